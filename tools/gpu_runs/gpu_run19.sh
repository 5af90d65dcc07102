cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
echo "== k-steps-per-stage: correctness (default 4,1)"
timeout 60 tools/oz_probe full 0 300 200 9 | tail -1
timeout 60 tools/oz_probe full 0 640 40000 9 | tail -1
timeout 60 tools/oz_probe full 0 1100 700 9 | tail -1
timeout 60 tools/oz_probe full 0 517 333 8 | tail -1
CVXB_OZ_KB=6,2 timeout 60 tools/oz_probe full 0 1100 700 9 | tail -1
CVXB_OZ_KB=3,2 timeout 60 tools/oz_probe full 0 640 40000 9 | tail -1
for KB in 1,1 2,1 4,1 6,1 4,2 6,2; do
echo "== perf n=8192 KB=$KB"
CVXB_OZ_KB=$KB timeout 120 tools/oz_probe perf 0 8192 16384 9 2 | grep -E "rep|FAIL"
done
echo "== ablate 1 (no copies) KB=4,2"
CVXB_OZ_KB=4,2 CVXB_OZ_ABLATE=1 timeout 120 tools/oz_probe perf 0 8192 16384 9 2 | grep -E "rep"
echo "== groups 4,4,1 / 1,4,4 with KB 4,2"
CVXB_OZ_KB=4,2 CVXB_OZ_GROUPS=4,4,1 timeout 120 tools/oz_probe perf 0 8192 16384 9 2 | grep -E "rep|FAIL"
CVXB_OZ_KB=4,2 CVXB_OZ_GROUPS=1,4,4 timeout 120 tools/oz_probe perf 0 8192 16384 9 2 | grep -E "rep|FAIL"
CVXB_OZ_KB=4,2 CVXB_OZ_GROUPS=2,3,4 timeout 120 tools/oz_probe perf 0 8192 16384 9 2 | grep -E "rep|FAIL"
echo "== n=4096 m=8192"
timeout 120 tools/oz_probe perf 0 4096 8192 9 2 | grep -E "rep|FAIL"
} > gpurun_out/r02q_oz_kb.txt 2>&1
cat gpurun_out/r02q_oz_kb.txt
