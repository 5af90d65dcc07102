cd $GRAFT_REPO_ROOT
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_kkt_gpu.py tests/test_batch_gpu.py tests/test_i8_syrk_gpu.py tests/test_scaling.py -m gpu -q 2>&1 | tail -3
