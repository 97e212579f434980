# interior blocks skip the per-tile image resolution (md_tile_loop<..., IMG=false>)
python -c "import __graft_entry__ as g; g.build()"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_fullsize.py tests/test_gpu_multistate.py -m gpu -x -q 2>&1 | tail -4
( for P in 1 2; do timeout 300 python profiles/tune_split.py; done
  NREP=8 CELLS=16,16,32 timeout 300 python profiles/tune_split.py
  NREP=64 CELLS=16,16,32 STEPS=300 timeout 300 python profiles/tune_split.py ) 2>&1 | grep -E "TUNE|rror" > gpurun_out/r2_tune_interior.log
cat gpurun_out/r2_tune_interior.log
