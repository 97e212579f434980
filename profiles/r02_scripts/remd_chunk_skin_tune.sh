python -c "import __graft_entry__ as g; g.build()"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t9.log
tail -4 gpurun_out/r2_t9.log
export NREP=8 CELLS=16,16,32 STEPS=600
( timeout 200 python profiles/tune_split.py
  CHX_FORCE_SPLIT=2 timeout 200 python profiles/tune_split.py
  CHX_MD_CHUNK=100 CHX_MD_SKIN=0.26 timeout 200 python profiles/tune_split.py
  CHX_MD_CHUNK=100 CHX_MD_SKIN=0.30 timeout 200 python profiles/tune_split.py
  CHX_MD_CHUNK=100 CHX_MD_SKIN=0.34 timeout 200 python profiles/tune_split.py
  CHX_MD_CHUNK=34 CHX_MD_SKIN=0.14 timeout 200 python profiles/tune_split.py
  CHX_MD_CHUNK=50 CHX_MD_SKIN=0.16 timeout 200 python profiles/tune_split.py
  CHX_MD_CHUNK=50 CHX_MD_SKIN=0.20 timeout 200 python profiles/tune_split.py
) 2>&1 | grep -E "TUNE|rror" > gpurun_out/r2_tune9.log
cat gpurun_out/r2_tune9.log
