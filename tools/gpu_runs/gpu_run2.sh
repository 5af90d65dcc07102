set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --durations=30 > gpurun_out/r02b_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02b_pytest.txt
timeout 300 python tools/e2e_profile.py > gpurun_out/r02b_e2e_profile.txt 2>&1
tail -40 gpurun_out/r02b_pytest.txt
