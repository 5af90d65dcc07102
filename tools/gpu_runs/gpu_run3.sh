cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
echo "== 2SM kernel bring-up"
timeout 40 tools/oz_probe onehot 0
timeout 40 tools/oz_probe ints 0 256 64
timeout 60 tools/oz_probe full 0 300 200 9
timeout 60 tools/oz_probe full 0 640 40000 9
timeout 60 tools/oz_probe full 0 517 333 8
timeout 60 tools/oz_probe full 0 1100 700 9
echo "== perf 2SM n=8192"
timeout 120 tools/oz_probe perf 0 8192 16384 9 3
echo "== perf 1SM n=8192"
CVXB_OZ_2SM=0 timeout 120 tools/oz_probe perf 0 8192 16384 9 3 | grep -E "rep|PASS|FAIL"
echo "== perf 2SM n=4096 m=8192"
timeout 120 tools/oz_probe perf 0 4096 8192 9 2 | grep -E "rep|PASS|FAIL"
for G in 2,3,4 1,4,4 4,4,1 3,3,3; do
  echo "== 2SM groups $G"
  CVXB_OZ_GROUPS=$G timeout 120 tools/oz_probe perf 0 8192 16384 9 2 | grep -E "rep|FAIL"
done
} > gpurun_out/r02c_oz2sm.txt 2>&1
cat gpurun_out/r02c_oz2sm.txt
