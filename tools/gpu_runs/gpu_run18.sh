cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
echo "== batched-load epilogue: correctness"
timeout 60 tools/oz_probe full 0 300 200 9 | tail -1
timeout 60 tools/oz_probe full 0 640 40000 9 | tail -1
timeout 60 tools/oz_probe full 0 1100 700 9 | tail -1
CVXB_OZ_2SM=2 timeout 60 tools/oz_probe full 0 1100 700 9 | tail -1
CVXB_OZ_2SM=1 timeout 60 tools/oz_probe full 0 1100 700 9 | tail -1
for M in 0 2 1; do
echo "== perf n=8192 2SM=$M"
CVXB_OZ_2SM=$M timeout 120 tools/oz_probe perf 0 8192 16384 9 2 | grep -E "rep|PASS|FAIL"
for AB in 1 2 3; do
echo "-- ablate $AB (1: no copies, 2: no MMAs, 3: neither)"
CVXB_OZ_2SM=$M CVXB_OZ_ABLATE=$AB timeout 120 tools/oz_probe perf 0 8192 16384 9 2 | grep -E "rep"
done
done
} > gpurun_out/r02p_oz_ablate.txt 2>&1
cat gpurun_out/r02p_oz_ablate.txt
