cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export CVXB_CHOL_TU=2
timeout 900 python -m pytest tests/test_kkt_gpu.py tests/test_solvers_gpu.py -m gpu -q -x 2>&1 | tail -4
timeout 900 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -x -k "l_cone_at_baseline" 2>&1 | tail -3
for TT in 6 3 12; do
CVXB_CHOL_TU_TILES=$TT timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ipm --no-i8 --no-driver 2>/dev/null | python -c "
import sys, json
b = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('TU=2 TILES=$TT ms_per_step', round(b['ms_per_step'], 3), b['breakdown_ms'])
"
done
CVXB_CHOL_TU=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ipm --no-i8 --no-driver 2>/dev/null | python -c "
import sys, json
b = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('TU=0 ms_per_step', round(b['ms_per_step'], 3), b['breakdown_ms'])
"
timeout 300 python tools/trace_potrf.py 8192 2>&1 | awk 'NR<=2 || (NR>=6 && NR<=10) || (NR>=40 && NR<=43)'
