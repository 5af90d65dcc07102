cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export CVXB_CHOL_TU=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chol_trailing_kernel -c 1 -o gpurun_out/r02y_tu python tools/trace_potrf.py 8192 > gpurun_out/r02y_ncu.log 2>&1
tail -3 gpurun_out/r02y_ncu.log
ncu -i gpurun_out/r02y_tu.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/r02y_tu_raw.csv
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02y_tu_raw.csv')))
hdr,units,vals=rows[0],rows[1],rows[2]
for h,u,v in zip(hdr,units,vals):
    if any(k in h for k in ("gpu__time_duration.sum","dmma","stalled","smsp__inst_executed.sum","l1tex__data_bank_conflicts","sm__cycles_active.avg","dram__bytes","warps_active","issue_active","pipe_fp64","lsu_mem_shared","shared_ld","shared_op_ld","xbar2l1tex_read_bytes")) and "pcsamp" not in h and ".max" not in h and ".min" not in h:
        print(h,u,v)
PY
