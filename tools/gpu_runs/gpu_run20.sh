cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
echo "== pass timeline of CTAs 0, 300, 1000 (us after CTA start: issue start, issue end, accumulators complete, epilogue end)"
for c in 0 300 1000; do
CVXB_OZ_TRACE_CTA=$c timeout 120 tools/oz_probe perf 0 8192 16384 9 1 | grep -E "trace|rep|FAIL"
done
echo "== perf after the slice-kernel rewrite"
timeout 120 tools/oz_probe perf 0 8192 16384 9 3 | grep -E "rep|PASS|FAIL"
timeout 60 tools/oz_probe full 0 1100 700 9 | tail -1
timeout 60 tools/oz_probe full 0 517 333 8 | tail -1
timeout 60 tools/oz_probe full 0 640 40000 9 | tail -1
} > gpurun_out/r02r_oz_trace.txt 2>&1
cat gpurun_out/r02r_oz_trace.txt
timeout 900 python -m pytest tests/test_i8_syrk_gpu.py tests/test_kkt_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x 2>&1 | tail -5
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02r_launches_probe.csv tools/oz_probe perf 0 8192 16384 9 1 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r02r_launches_probe.csv
