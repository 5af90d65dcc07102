#!/bin/bash
# int8-slice SYRK: correctness at several shapes, then timing of the level-group variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 40 tools/oz_probe onehot 0
timeout 60 tools/oz_probe full 0 300 200 9
timeout 60 tools/oz_probe full 0 640 40000 9
timeout 60 tools/oz_probe full 0 517 333 8
echo "== default groups"
timeout 120 tools/oz_probe perf 0 8192 16384 9 3
for G in 2,3,4 2,4,3 4,4,1; do
  echo "== groups $G"
  CVXB_OZ_GROUPS=$G timeout 120 tools/oz_probe perf 0 8192 16384 9 2 | grep -E "rep|FAIL"
done
echo "== n=4096 m=8192"; timeout 120 tools/oz_probe perf 0 4096 8192 9 2 | grep -E "rep|PASS|FAIL"
