# _EngineGroups with the split chosen for the engine's share of the machine (chx_ljmd_set_gpu_share)
python -c "import __graft_entry__ as g; g.build()"
timeout 600 python -m pytest tests/test_gpu_multistate.py -m gpu -x -q 2>&1 | tail -2
( for N in 8 16 32; do for G in 1 2; do
  LOCKSTEP=1 NO_PHASE=1 NGROUPS=$G NREP=$N SWEEPS=10 timeout 300 python profiles/tune_two_engines.py
  done; done ) 2>&1 | grep -E "TWO|rror|Trace" > gpurun_out/r2_two_engines_lock5.log
cat gpurun_out/r2_two_engines_lock5.log
