cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_batch_gpu.py tests/test_kkt_gpu.py tests/test_solvers_gpu.py tests/test_qr_ldl_gpu.py tests/test_tma_path_gpu.py -m gpu -q -x 2>&1 | tail -5
timeout 900 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -x -k "l_cone_at_baseline or config4 or config2" 2>&1 | tail -3
for C in 1 0; do
CVXB_BATCH_COMPACT=$C timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-i8 --no-driver 2>/dev/null | python -c "
import sys, json
b = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('COMPACT=$C ms_per_step', round(b['ms_per_step'], 3), b['breakdown_ms'])
bb = b['batch']; print({k: bb[k] for k in ('scatter_ms','solve_ms','gather_ms','ms','ipm_kernel_ms','ipm_iterations_total','all_optimal')})
print(b['ipm'])
"
done
