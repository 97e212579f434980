python -c "import __graft_entry__ as g; g.build()"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t18.log
tail -5 gpurun_out/r2_t18.log
( timeout 300 python profiles/tune_split.py; NREP=8 CELLS=16,16,32 timeout 300 python profiles/tune_split.py; NREP=64 CELLS=16,16,32 STEPS=300 timeout 300 python profiles/tune_split.py ) 2>&1 | grep -E "TUNE|rror" > gpurun_out/r2_tune18.log
cat gpurun_out/r2_tune18.log
