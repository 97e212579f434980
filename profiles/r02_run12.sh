python -c "import __graft_entry__ as g; g.build()"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_fullsize.py tests/test_gpu_statistics.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_t12.log
tail -4 gpurun_out/r2_t12.log
( CHX_MD_DEAL_KEY=0 timeout 300 python profiles/tune_split.py; timeout 300 python profiles/tune_split.py; NREP=8 CELLS=16,16,32 CHX_MD_DEAL_KEY=0 timeout 300 python profiles/tune_split.py; NREP=8 CELLS=16,16,32 timeout 300 python profiles/tune_split.py;) 2>&1 | grep -E "TUNE|rror" > gpurun_out/r2_tune12.log
cat gpurun_out/r2_tune12.log
