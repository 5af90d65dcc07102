cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_abi.py tests/test_i8_syrk_gpu.py -q 2>&1 | tail -2
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ipm --no-i8 --no-driver > gpurun_out/r02z_bench_short.json 2>/dev/null
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r02z_bench_short.json').read().strip().splitlines()[-1])
print(b['ms_per_step'], b['breakdown_ms']); r=b['roofline']; print({k:r[k] for k in ('achieved','peak','frac','kernel_ms','frac_incl_slicing_kernels')})
PY
