# one GPU's share of config 5 at 8 GPUs (8 x 8,192) as 1 / 2 / 4 engines on their own streams and host threads
python -c "import __graft_entry__ as g; g.build()"
( NGROUPS=1 timeout 300 python profiles/tune_two_engines.py
  NGROUPS=2 timeout 300 python profiles/tune_two_engines.py
  NGROUPS=2 CHX_FORCE_SPLIT=2 timeout 300 python profiles/tune_two_engines.py
  NGROUPS=4 timeout 300 python profiles/tune_two_engines.py
  NGROUPS=4 CHX_FORCE_SPLIT=2 timeout 300 python profiles/tune_two_engines.py
  NGROUPS=1 NREP=16 timeout 300 python profiles/tune_two_engines.py
  NGROUPS=2 NREP=16 timeout 300 python profiles/tune_two_engines.py
  NGROUPS=1 timeout 300 python profiles/tune_two_engines.py ) 2>&1 | grep -E "TWO|rror|Trace" > gpurun_out/r2_two_engines.log
cat gpurun_out/r2_two_engines.log
