cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -m pytest tests/test_multidevice_gpu.py -q 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_batch_check.py 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 4 --warmup 3 --no-i8 --no-driver > gpurun_out/r02z_bench_n2.json 2> gpurun_out/r02z_bench_n2.err
echo "bench n2 rc=$?"; tail -3 gpurun_out/r02z_bench_n2.err; python - <<'PY'
import json
b=json.loads(open('gpurun_out/r02z_bench_n2.json').read().strip().splitlines()[-1])
print({k:b[k] for k in ('value','n_gpus','ms_per_step')}); print(json.dumps(b.get('batch'),indent=1))
PY
