cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests/test_conelp_device_gpu.py -q -x --durations=10 2>&1 | tail -40
timeout 300 python -m pytest tests/test_batch_gpu.py -q 2>&1 | tail -3
} > gpurun_out/r02k_conelp.txt 2>&1
cat gpurun_out/r02k_conelp.txt
