python -c "import __graft_entry__ as g; g.build()"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_fullsize.py -m gpu -x -q -k "engine or langevin or persistent or full_size or fullsize or 262144 or config1 or odd" 2>&1 | tail -4 > gpurun_out/r2_t14.log
tail -3 gpurun_out/r2_t14.log
( timeout 300 python profiles/tune_split.py; CHX_MD_CHUNK=4 timeout 300 python profiles/tune_split.py; CHX_MD_CHUNK=6 timeout 300 python profiles/tune_split.py; CHX_MD_CHUNK=12 timeout 300 python profiles/tune_split.py ) 2>&1 | grep -E "TUNE|rror" > gpurun_out/r2_tune14.log
cat gpurun_out/r2_tune14.log
