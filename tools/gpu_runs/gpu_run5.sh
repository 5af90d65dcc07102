cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_kkt_gpu.py tests/test_solvers_gpu.py -q -x 2>&1 | tail -5
timeout 600 python -m pytest tests/test_fullsize_gpu.py -q -x -k "baseline_size and 8192 and 10000" 2>&1 | tail -5
for P in 1 0; do
  echo "== CVXB_CHOL_PAIR=$P"
  CVXB_CHOL_PAIR=$P timeout 600 python bench.py --no-cpu-baseline --no-ipm --no-i8 --steps 6 | python -c "
import sys, json
b = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms_per_step', b['ms_per_step'], 'breakdown', b['breakdown_ms'], 'value', b['value'])"
done
echo "== n=4096 potrf"
for P in 1 0; do
  CVXB_CHOL_PAIR=$P timeout 600 python bench.py --n 4096 --no-cpu-baseline --no-ipm --no-i8 --steps 6 | python -c "
import sys, json
b = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('pair=$P n=4096 ms_per_step', b['ms_per_step'], 'breakdown', b['breakdown_ms'])"
done
timeout 300 python tools/e2e_profile.py 2>&1 | grep -E "in-solver|idle gap|total"
} > gpurun_out/r02e_potrf_pair.txt 2>&1
cat gpurun_out/r02e_potrf_pair.txt
