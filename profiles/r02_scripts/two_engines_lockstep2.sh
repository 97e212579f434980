# _EngineGroups, engines joined after every 100-step run: 1 vs 2 engines at 8 / 16 / 32 / 64 replicas, no stagger, twice
python -c "import __graft_entry__ as g; g.build()"
( for rep in 1 2; do for N in 8 16 32 64; do for G in 1 2; do
  LOCKSTEP=1 NO_PHASE=1 NGROUPS=$G NREP=$N SWEEPS=10 timeout 300 python profiles/tune_two_engines.py
  done; done; done ) 2>&1 | grep -E "TWO|rror|Trace" > gpurun_out/r2_two_engines_lock2.log
cat gpurun_out/r2_two_engines_lock2.log
