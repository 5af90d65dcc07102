cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r02h_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02h_pytest.txt
tail -15 gpurun_out/r02h_pytest.txt
timeout 600 python tools/batch_nsub_bench.py > gpurun_out/r02h_batch_nsub.txt 2>&1
cat gpurun_out/r02h_batch_nsub.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:oz_mma_kernel -c 1 -o gpurun_out/r02h_oz_mma tools/oz_probe perf 0 8192 16384 9 1 > gpurun_out/r02h_ncu.log 2>&1
tail -3 gpurun_out/r02h_ncu.log
