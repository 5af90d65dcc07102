cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
for P in 0 16 32 48; do
  CVXB_CHOL_PAIR=$P timeout 600 python bench.py --no-cpu-baseline --no-ipm --no-i8 --steps 6 | python -c "
import sys, json
b = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('CVXB_CHOL_PAIR=$P ms_per_step', round(b['ms_per_step'],3), 'potrf', round(b['breakdown_ms']['potrf'],3), 'syrk', round(b['breakdown_ms']['syrk'],3))"
done
timeout 300 python -m pytest tests/test_kkt_gpu.py -q -x -k "building_blocks or l_cones" 2>&1 | tail -2
} > gpurun_out/r02i_pair_hybrid.txt 2>&1
cat gpurun_out/r02i_pair_hybrid.txt
timeout 1200 python bench.py > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r02i_bench.json
