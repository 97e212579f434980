# _EngineGroups: 2 vs 4 engines at 32 / 64 replicas
python -c "import __graft_entry__ as g; g.build()"
( for N in 32 64; do for G in 2 4; do
  LOCKSTEP=1 NO_PHASE=1 NGROUPS=$G NREP=$N SWEEPS=10 timeout 300 python profiles/tune_two_engines.py
  done; done ) 2>&1 | grep -E "TWO|rror|Trace" > gpurun_out/r2_two_engines_lock3.log
cat gpurun_out/r2_two_engines_lock3.log
