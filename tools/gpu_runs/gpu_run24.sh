cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kkt_gpu.py tests/test_solvers_gpu.py -m gpu -q -x 2>&1 | tail -3
CVXB_CHOL_PDL=1 timeout 900 python -m pytest tests/test_kkt_gpu.py tests/test_solvers_gpu.py -m gpu -q -x 2>&1 | tail -3
for P in 0 1; do
for G in 1 0; do
CVXB_CHOL_PDL=$P CVXB_GRAPH=$G timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ipm --no-i8 --no-driver 2>/dev/null | python -c "
import sys, json
b = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('PDL=$P GRAPH=$G ms_per_step', round(b['ms_per_step'], 3), b['breakdown_ms'])
"
done
done
CVXB_CHOL_PDL=1 timeout 600 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -x -k "l_cone_at_baseline or backward" 2>&1 | tail -3
