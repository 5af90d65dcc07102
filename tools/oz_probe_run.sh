#!/bin/bash
# bring-up sequence for the int8-slice SYRK: find the operand layout the MMA accepts, then accuracy + timing
cd "$(dirname "$0")/.."
OK=""
for L in 0 1 2; do
  echo "== layout $L"
  timeout 40 tools/oz_probe onehot $L; r1=$?
  timeout 40 tools/oz_probe ints $L 256 96; r2=$?
  echo "   rc onehot=$r1 ints=$r2"
  if [ $r1 -eq 0 ] && [ $r2 -eq 0 ] && [ -z "$OK" ]; then OK=$L; fi
done
echo "== chosen layout: '$OK'"
if [ -n "$OK" ]; then
  timeout 60 tools/oz_probe full $OK 300 200 9
  timeout 60 tools/oz_probe full $OK 1000 4100 9
  timeout 60 tools/oz_probe full $OK 640 40000 9
  timeout 120 tools/oz_probe perf $OK 8192 16384 9 3
  timeout 120 tools/oz_probe perf $OK 8192 16384 8 2
fi
