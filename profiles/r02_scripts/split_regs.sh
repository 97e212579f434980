# split (2 / 4 warps per block) variants of the step kernel at 72 registers (28 warps per SM) instead of 64 with spills
python -c "import __graft_entry__ as g; g.build()"
timeout 120 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multistate.py -m gpu -x -q -k "batched or engine_groups or prebuild or anisotropic or langevin_single" 2>&1 | tail -2
( NREP=8 CELLS=16,16,32 timeout 100 python profiles/tune_split.py ) 2>&1 | grep -E "TUNE|rror" | tee gpurun_out/r2_tune_split_regs.log
