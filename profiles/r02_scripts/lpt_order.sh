# heaviest-block-first launch order of the step kernel (k_md_order): CHX_MD_LPT=0/1
python -c "import __graft_entry__ as g; g.build()"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_fullsize.py tests/test_gpu_multistate.py -m gpu -x -q 2>&1 | tail -4
( for P in 0 1 0 1; do CHX_MD_LPT=$P timeout 300 python profiles/tune_split.py | sed "s/^TUNE/TUNE lpt=$P/"; done
  for P in 0 1 0 1; do CHX_MD_LPT=$P NREP=8 CELLS=16,16,32 timeout 300 python profiles/tune_split.py | sed "s/^TUNE/TUNE lpt=$P/"; done
  for P in 0 1; do CHX_MD_LPT=$P NREP=64 CELLS=16,16,32 STEPS=300 timeout 300 python profiles/tune_split.py | sed "s/^TUNE/TUNE lpt=$P/"; done ) 2>&1 | grep -E "TUNE|rror" > gpurun_out/r2_tune_lpt.log
cat gpurun_out/r2_tune_lpt.log
