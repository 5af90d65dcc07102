cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
echo "== 1SM kernel with N=256 paired MMAs"
timeout 40 tools/oz_probe onehot 0
timeout 40 tools/oz_probe ints 0 256 64
timeout 60 tools/oz_probe full 0 300 200 9
timeout 60 tools/oz_probe full 0 640 40000 9
timeout 60 tools/oz_probe full 0 517 333 8
echo "== perf n=8192 default groups"
timeout 120 tools/oz_probe perf 0 8192 16384 9 3
for G in 1,4,4 4,4,1 2,4,3 3,2,4; do
  echo "== groups $G"
  CVXB_OZ_GROUPS=$G timeout 120 tools/oz_probe perf 0 8192 16384 9 2 | grep -E "rep|FAIL"
done
echo "== perf n=4096 m=8192"
timeout 120 tools/oz_probe perf 0 4096 8192 9 2 | grep -E "rep|PASS|FAIL"
} > gpurun_out/r02d_oz_pair.txt 2>&1
cat gpurun_out/r02d_oz_pair.txt
timeout 300 python tools/e2e_profile.py 2>&1 | grep -E "idle gap|total" 
