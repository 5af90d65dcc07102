cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
for MODE in "" "CVXB_CHOL_NOWAIT_EXPERIMENT=1" "CVXB_CHOL_PDL=1" "CVXB_CHOL_NOWAIT_EXPERIMENT=1 CVXB_CHOL_PDL=1"; do
echo "=== mode: $MODE"
env $MODE timeout 300 python tools/trace_potrf.py 8192 2>&1 | awk 'NR<=4 || (NR>=40 && NR<=52)'
done
} > gpurun_out/r02v_potrf_gap_experiment.txt 2>&1
cat gpurun_out/r02v_potrf_gap_experiment.txt
