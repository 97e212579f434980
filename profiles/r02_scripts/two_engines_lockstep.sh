# _EngineGroups (engines joined after every 100-step run, like a REMD sweep) with and without staggered rebuilds
python -c "import __graft_entry__ as g; g.build()"
timeout 600 python -m pytest tests/test_gpu_multistate.py -m gpu -x -q 2>&1 | tail -3
( for N in 8 16; do
  LOCKSTEP=1 NGROUPS=1 NREP=$N timeout 300 python profiles/tune_two_engines.py
  LOCKSTEP=1 NGROUPS=2 NO_PHASE=1 NREP=$N timeout 300 python profiles/tune_two_engines.py
  LOCKSTEP=1 NGROUPS=2 NREP=$N timeout 300 python profiles/tune_two_engines.py
  LOCKSTEP=1 NGROUPS=4 NREP=$N timeout 300 python profiles/tune_two_engines.py
  done
  LOCKSTEP=1 NGROUPS=2 NREP=32 timeout 300 python profiles/tune_two_engines.py
  LOCKSTEP=1 NGROUPS=1 NREP=32 timeout 300 python profiles/tune_two_engines.py ) 2>&1 | grep -E "TWO|rror|Trace" > gpurun_out/r2_two_engines_lock.log
cat gpurun_out/r2_two_engines_lock.log
