# why is the step kernel at 8 x 8,192 particles 2x slower per particle than at 262,144?  same particle count, different geometry
python -c "import __graft_entry__ as g; g.build()"
( NREP=1 CELLS=32,32,64 timeout 300 python profiles/tune_split.py
  NREP=1 CELLS=40,40,40 timeout 300 python profiles/tune_split.py
  NREP=8 CELLS=16,16,32 timeout 300 python profiles/tune_split.py
  NREP=8 CELLS=20,20,20 timeout 300 python profiles/tune_split.py
  NREP=8 CELLS=16,16,32 CHX_FORCE_SPLIT=1 timeout 300 python profiles/tune_split.py
  NREP=1 CELLS=64,64,64 timeout 300 python profiles/tune_split.py ) 2>&1 | grep -E "TUNE|rror" > gpurun_out/r2_tune_geom.log
cat gpurun_out/r2_tune_geom.log
