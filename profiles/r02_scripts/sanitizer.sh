# compute-sanitizer memcheck / racecheck over the engine tests of the final tree
python -c "import __graft_entry__ as g; g.build()"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_engine or anisotropic or langevin_single" 2>&1 | tail -6 > gpurun_out/r2_memcheck_engine.log
tail -4 gpurun_out/r2_memcheck_engine.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_multistate.py tests/test_gpu_x64.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_memcheck_multistate_x64.log
tail -4 gpurun_out/r2_memcheck_multistate_x64.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batched_replicas and not step_loops or anisotropic" 2>&1 | tail -6 > gpurun_out/r2_racecheck_engine.log
tail -4 gpurun_out/r2_racecheck_engine.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mc_ or neighborlist or pairlist or mixture" 2>&1 | tail -6 > gpurun_out/r2_memcheck_mc_lists.log
tail -4 gpurun_out/r2_memcheck_mc_lists.log
