set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02a_smi.txt
./tools/int8_peak > gpurun_out/r02a_int8_peak.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=25 > gpurun_out/r02a_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_pytest.txt
timeout 900 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench rc=$?" >> gpurun_out/r02a_bench.err
tail -5 gpurun_out/r02a_pytest.txt
cat gpurun_out/r02a_int8_peak.txt
