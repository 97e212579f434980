python -c "import __graft_entry__ as g; g.build()"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "engine or langevin or persistent or full_size" 2>&1 | tail -5 > gpurun_out/r2_t7.log
tail -3 gpurun_out/r2_t7.log
( timeout 300 python profiles/tune_split.py; NREP=8 CELLS=16,16,32 timeout 300 python profiles/tune_split.py; NREP=64 CELLS=16,16,32 STEPS=300 timeout 300 python profiles/tune_split.py ) 2>&1 | grep -E "TUNE|rror" > gpurun_out/r2_tune7.log
cat gpurun_out/r2_tune7.log
STEPS=200 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_v6.csv python profiles/tune_split.py > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/r2_launches_v6.csv 2>&1 | head -12
STEPS=100 timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_md_force -s 400 -c 1 -o gpurun_out/r2_force_v6 python profiles/tune_split.py > gpurun_out/r2_ncu7.log 2>&1
tail -2 gpurun_out/r2_ncu7.log
