cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kkt_gpu.py tests/test_solvers_gpu.py -m gpu -q -x 2>&1 | tail -4
timeout 900 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -x -k "l_cone_at_baseline or backward or potrf" 2>&1 | tail -3
for TT in 3 2 5 0; do
if [ $TT = 0 ]; then export CVXB_CHOL_TU=0; else export CVXB_CHOL_TU_TILES=$TT; fi
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ipm --no-i8 --no-driver 2>/dev/null | python -c "
import sys, json
b = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('TU_TILES=$TT ms_per_step', round(b['ms_per_step'], 3), b['breakdown_ms'])
"
done
unset CVXB_CHOL_TU
timeout 300 python tools/trace_potrf.py 8192 2>&1 | awk 'NR<=4 || (NR>=6 && NR<=12) || (NR>=40 && NR<=46)'
timeout 200 python tools/trace_potrf.py 4096 2>&1 | awk 'NR<=2'
timeout 200 python tools/trace_potrf.py 2048 2>&1 | awk 'NR<=2'
