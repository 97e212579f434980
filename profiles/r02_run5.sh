python -c "import __graft_entry__ as g; g.build()"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_fullsize.py -m gpu -x -q -k "engine or langevin or persistent or full_size or fullsize or 262144 or config1 or odd" 2>&1 | tail -12 > gpurun_out/r2_t5.log
tail -4 gpurun_out/r2_t5.log
( timeout 300 python profiles/tune_split.py; CHX_MD_CHUNK=4 timeout 300 python profiles/tune_split.py; CHX_MD_CHUNK=16 timeout 300 python profiles/tune_split.py; CHX_MD_OLD_DEAL=1 timeout 300 python profiles/tune_split.py; NREP=8 CELLS=16,16,32 timeout 300 python profiles/tune_split.py; NREP=64 CELLS=16,16,32 STEPS=300 timeout 300 python profiles/tune_split.py ) 2>&1 | grep -E "TUNE|rror" > gpurun_out/r2_tune5.log
cat gpurun_out/r2_tune5.log
STEPS=200 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_v6.csv python profiles/tune_split.py > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/r2_launches_v6.csv 2>&1 | tail -30
