cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
timeout 60 tools/oz_probe full 0 300 200 9 | tail -1
timeout 60 tools/oz_probe full 0 640 40000 9 | tail -1
timeout 60 tools/oz_probe full 0 517 333 8 | tail -1
timeout 120 tools/oz_probe perf 0 8192 16384 9 3 | grep -E "rep|PASS|FAIL"
} > gpurun_out/r02z_oz_probe.txt 2>&1
cat gpurun_out/r02z_oz_probe.txt
timeout 1800 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02z_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02z_pytest.txt
tail -10 gpurun_out/r02z_pytest.txt
timeout 1500 python bench.py > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r02z_bench.json').read().strip().splitlines()[-1])
print({k:b[k] for k in ('value','ms_per_step','breakdown_ms','gpu_launches')}); print(b['roofline']['frac'], b['e2e'])
bb=b['batch']; print({k: bb[k] for k in ('scatter_ms','solve_ms','gather_ms','ms','ipm_kernel_ms','ipm_iterations_total','all_optimal')})
print({k:(v.get('device_conelp') or {}).get('seconds') for k,v in b['e2e_cones'].items()}); print(b.get('ipm')); print(b.get('clocks'))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02z_launches_n8192.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ipm --no-i8 --no-driver > gpurun_out/r02z_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/r02z_launches_n8192.csv | head -12
timeout 400 python bench.py --impl reference --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
