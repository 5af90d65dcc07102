cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
echo "== int8 peak incl. cta_group::2"
timeout 60 tools/int8_peak
echo "== WIDE 2SM kernel bring-up (CVXB_OZ_2SM=2)"
export CVXB_OZ_2SM=2
timeout 40 tools/oz_probe onehot 0
timeout 40 tools/oz_probe ints 0 256 64
timeout 40 tools/oz_probe ints 0 512 96
timeout 60 tools/oz_probe full 0 300 200 9
timeout 60 tools/oz_probe full 0 640 40000 9
timeout 60 tools/oz_probe full 0 517 333 8
timeout 60 tools/oz_probe full 0 1100 700 9
timeout 60 tools/oz_probe full 0 2500 900 9
echo "== perf WIDE n=8192"
timeout 120 tools/oz_probe perf 0 8192 16384 9 3
echo "== perf WIDE n=8192, no tail split"
CVXB_OZ_TAIL=0 timeout 120 tools/oz_probe perf 0 8192 16384 9 2 | grep -E "rep|PASS|FAIL"
echo "== perf 1SM n=8192"
CVXB_OZ_2SM=0 timeout 120 tools/oz_probe perf 0 8192 16384 9 2 | grep -E "rep|PASS|FAIL"
echo "== perf WIDE n=4096 m=8192 / n=4352 m=8704"
timeout 120 tools/oz_probe perf 0 4096 8192 9 2 | grep -E "rep|PASS|FAIL"
timeout 120 tools/oz_probe perf 0 4352 8704 9 2 | grep -E "rep|PASS|FAIL"
CVXB_OZ_2SM=0 timeout 120 tools/oz_probe perf 0 4096 8192 9 2 | grep -E "rep|PASS|FAIL"
for BD in 8 16 24; do
  echo "== WIDE band $BD"
  CVXB_OZ_BAND=$BD timeout 120 tools/oz_probe perf 0 8192 16384 9 2 | grep -E "rep|FAIL"
done
unset CVXB_OZ_2SM
} > gpurun_out/r02o_oz_wide.txt 2>&1
cat gpurun_out/r02o_oz_wide.txt
timeout 900 python -m pytest tests/test_batch_gpu.py tests/test_conelp_device_gpu.py -m gpu -q -x 2>&1 | tail -5
