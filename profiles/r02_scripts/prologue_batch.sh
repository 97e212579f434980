# step kernel prologue: every load that depends on (replica, block) only -- table header, first tile words, own
# positions, control block -- issued as one batch before the halt check
python -c "import __graft_entry__ as g; g.build()"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_fullsize.py tests/test_gpu_multistate.py -m gpu -x -q 2>&1 | tail -4
( for P in 1 2; do timeout 300 python profiles/tune_split.py; done
  for P in 1 2; do NREP=8 CELLS=16,16,32 timeout 300 python profiles/tune_split.py; done
  NREP=64 CELLS=16,16,32 STEPS=300 timeout 300 python profiles/tune_split.py ) 2>&1 | grep -E "TUNE|rror" > gpurun_out/r2_tune_prologue.log
cat gpurun_out/r2_tune_prologue.log
