cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
echo "== split-K tail of the int8-slice kernel"
timeout 60 tools/oz_probe full 0 640 40000 9
timeout 120 tools/oz_probe perf 0 8192 16384 9 3
echo "-- CVXB_OZ_TAIL=0"
CVXB_OZ_TAIL=0 timeout 120 tools/oz_probe perf 0 8192 16384 9 3 | grep -E "rep|PASS|FAIL"
echo "-- n=4224 (561 tiles = 3 x 148 + 117: no split)  and n=4352 (595 = 4 x 148 + 3)"
timeout 120 tools/oz_probe perf 0 4352 8704 9 2 | grep -E "rep|PASS|FAIL"
CVXB_OZ_TAIL=0 timeout 120 tools/oz_probe perf 0 4352 8704 9 2 | grep -E "rep|PASS|FAIL"
for P in 0 16 32 48; do
  CVXB_CHOL_PAIR=$P timeout 600 python bench.py --no-cpu-baseline --no-ipm --no-i8 --steps 6 | python -c "
import sys, json
b = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('CVXB_CHOL_PAIR=$P ms_per_step', round(b['ms_per_step'],3), 'potrf', round(b['breakdown_ms']['potrf'],3), 'syrk', round(b['breakdown_ms']['syrk'],3))"
done
timeout 300 python -m pytest tests/test_kkt_gpu.py tests/test_i8_syrk_gpu.py -q -x 2>&1 | tail -2
} > gpurun_out/r02i_tail_pair.txt 2>&1
cat gpurun_out/r02i_tail_pair.txt
timeout 1200 python bench.py > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err
echo "bench rc=$?"; tail -c 2500 gpurun_out/r02i_bench.json
