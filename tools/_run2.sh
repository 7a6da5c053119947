python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host_buffer_blocks or msm_vs_oracle or all_ops" 2>&1 | tail -2
