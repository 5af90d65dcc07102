cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r02w_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02w_pytest.txt
tail -12 gpurun_out/r02w_pytest.txt
timeout 1500 python bench.py > gpurun_out/r02w_bench.json 2> gpurun_out/r02w_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r02w_bench.json').read().strip().splitlines()[-1])
print({k:b[k] for k in ('value','ms_per_step','breakdown_ms','gpu_launches')}); print(b['roofline']['frac'], b['e2e'])
print(json.dumps(b.get('batch'))); print(json.dumps(b.get('e2e_cones'))); print(json.dumps(b.get('e2e_driver'))); print(b.get('ipm')); print(b.get('cpu_baseline')); print(b.get('clocks'))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oz_mma_kernel -c 1 -o gpurun_out/r02w_oz_mma tools/oz_probe perf 0 8192 16384 9 1 > gpurun_out/r02w_ncu.log 2>&1
ncu -i gpurun_out/r02w_oz_mma.ncu-rep --page raw --csv 2>/dev/null | python - <<'PY'
import sys, csv
rows = list(csv.reader(sys.stdin))
if len(rows) >= 3:
    hdr, units, vals = rows[0], rows[1], rows[2]
    want = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
            "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__block_size", "launch__grid_size")
    for h, u, v in zip(hdr, units, vals):
        if h in want or "utcimma" in h or "pipe_tensor" in h:
            print(h, u, v)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02w_launches_n8192.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ipm --no-i8 --no-driver > gpurun_out/r02w_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/r02w_launches_n8192.csv | head -24
