cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
echo "== 8 epilogue warps: correctness"
timeout 60 tools/oz_probe full 0 300 200 9 | tail -1
timeout 60 tools/oz_probe full 0 640 40000 9 | tail -1
timeout 60 tools/oz_probe full 0 1100 700 9 | tail -1
timeout 60 tools/oz_probe full 0 517 333 8 | tail -1
echo "== pass timeline of CTAs 300, 1000"
for c in 300 1000; do
CVXB_OZ_TRACE_CTA=$c timeout 120 tools/oz_probe perf 0 8192 16384 9 1 | grep -E "trace|rep|FAIL"
done
echo "== perf"
timeout 120 tools/oz_probe perf 0 8192 16384 9 3 | grep -E "rep|PASS|FAIL"
timeout 120 tools/oz_probe perf 0 4096 8192 9 2 | grep -E "rep|PASS|FAIL"
} > gpurun_out/r02s_oz_epi8.txt 2>&1
cat gpurun_out/r02s_oz_epi8.txt
timeout 600 python tools/conelp_profile.py > gpurun_out/r02s_conelp_profile.txt 2>&1
cat gpurun_out/r02s_conelp_profile.txt
