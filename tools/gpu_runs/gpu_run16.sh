cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02n_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02n_pytest.txt
tail -14 gpurun_out/r02n_pytest.txt
timeout 600 python tools/config_bench.py > gpurun_out/r02n_config_bench.json 2>&1
cat gpurun_out/r02n_config_bench.json
timeout 1500 python bench.py > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r02n_bench.json').read().strip().splitlines()[-1])
print({k:b[k] for k in ('value','ms_per_step','breakdown_ms','gpu_launches')}); print(b['roofline']['frac'], b['e2e'])
print(json.dumps(b.get('batch'))); print(json.dumps(b.get('e2e_cones'),indent=1)); print(json.dumps(b.get('e2e_driver'))); print(b.get('ipm')); print(b.get('cpu_baseline'))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02n_launches_n8192.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ipm --no-i8 > gpurun_out/r02n_ncu_bench.log 2>&1
echo "ncu rc=$?"
