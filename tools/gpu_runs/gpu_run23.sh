cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_fullsize_gpu.py 2>&1 | tail -5
timeout 600 python tools/conelp_profile.py > gpurun_out/r02u_conelp_profile.txt 2>&1
cat gpurun_out/r02u_conelp_profile.txt
