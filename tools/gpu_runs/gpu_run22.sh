cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kkt_gpu.py tests/test_scaling.py tests/test_misc_solvers_ext.py tests/test_conelp_device_gpu.py tests/test_wrappers_gpu.py -m gpu -q -x 2>&1 | tail -5
CVXB_JACOBI_COOP=0 timeout 900 python -m pytest tests/test_kkt_gpu.py tests/test_scaling.py -m gpu -q -x -k "max_step or scaling" 2>&1 | tail -3
timeout 600 python tools/conelp_profile.py > gpurun_out/r02t_conelp_profile.txt 2>&1
cat gpurun_out/r02t_conelp_profile.txt
