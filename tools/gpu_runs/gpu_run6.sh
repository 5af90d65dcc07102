cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_scaling.py -q -x -m gpu 2>&1 | tail -15
echo "== e2e profile, graphs on"
timeout 300 python tools/e2e_profile.py 2>&1 | grep -E "in-solver|total"
echo "== e2e profile, CVXB_GRAPH=0"
CVXB_GRAPH=0 timeout 300 python tools/e2e_profile.py 2>&1 | grep -E "in-solver|total"
} > gpurun_out/r02f_scaling_e2e.txt 2>&1
cat gpurun_out/r02f_scaling_e2e.txt
