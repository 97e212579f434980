# _EngineGroups at 8 / 16 replicas with the sub-engines forced to one or two warps per block
python -c "import __graft_entry__ as g; g.build()"
( for N in 8 16; do for S in 1 2; do
  CHX_FORCE_SPLIT=$S LOCKSTEP=1 NO_PHASE=1 NGROUPS=2 NREP=$N SWEEPS=10 timeout 300 python profiles/tune_two_engines.py | sed "s/^TWO/TWO split=$S/"
  done; done ) 2>&1 | grep -E "TWO|rror|Trace" > gpurun_out/r2_two_engines_lock4.log
cat gpurun_out/r2_two_engines_lock4.log
