set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "engine or langevin or persistent or full_size" 2>&1 | tail -15 > gpurun_out/r2_t3.log
tail -5 gpurun_out/r2_t3.log
( timeout 300 python profiles/tune_split.py; CHX_MD_PERSIST=0 timeout 300 python profiles/tune_split.py; NREP=8 CELLS=16,16,32 timeout 300 python profiles/tune_split.py;  NREP=8 CELLS=16,16,32 CHX_MD_PERSIST=0 timeout 300 python profiles/tune_split.py; NREP=64 CELLS=16,16,32 STEPS=300 timeout 300 python profiles/tune_split.py ) 2>&1 | grep -E "TUNE|Error|error" > gpurun_out/r2_tune1.log
cat gpurun_out/r2_tune1.log
