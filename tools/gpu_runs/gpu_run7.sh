cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_qr_ldl_gpu.py -q -m gpu 2>&1 | tail -40
timeout 900 python -m pytest tests/test_kkt_gpu.py tests/test_solvers_gpu.py tests/test_cvxprog_gpu.py -q -x -m gpu 2>&1 | tail -5
} > gpurun_out/r02g_qr_ldl.txt 2>&1
cat gpurun_out/r02g_qr_ldl.txt
