cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests/test_conelp_device_gpu.py -q --durations=10 2>&1 | tail -30
for P in 0 1 24 32 40; do
  CVXB_CHOL_PAIR=$P timeout 600 python bench.py --no-cpu-baseline --no-ipm --no-i8 --steps 6 | python -c "
import sys, json
b = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('CVXB_CHOL_PAIR=$P ms_per_step', round(b['ms_per_step'],3), 'potrf', round(b['breakdown_ms']['potrf'],3), 'syrk', round(b['breakdown_ms']['syrk'],3))"
done
CVXB_CHOL_PAIR=1 timeout 300 python -m pytest tests/test_kkt_gpu.py -q -x -k "building_blocks or l_cones or equality" 2>&1 | tail -2
CVXB_CHOL_PAIR=32 timeout 300 python -m pytest tests/test_fullsize_gpu.py -q -x -k "baseline_size and 8192 and 10000" 2>&1 | tail -2
} > gpurun_out/r02l_conelp_pair.txt 2>&1
cat gpurun_out/r02l_conelp_pair.txt
