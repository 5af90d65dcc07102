cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 --no-i8 --no-driver > gpurun_out/r02z_bench_n2.json 2> gpurun_out/r02z_bench_n2.err
echo "bench n2 rc=$?"; python - <<'PY'
import json
b=json.loads(open('gpurun_out/r02z_bench_n2.json').read().strip().splitlines()[-1])
print({k:b[k] for k in ('value','n_gpus','ms_per_step')}); print(json.dumps(b.get('batch'),indent=1))
PY
