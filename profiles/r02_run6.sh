python -c "import __graft_entry__ as g; g.build()"
( timeout 300 python profiles/tune_split.py; CHX_MD_OLD_DEAL=1 timeout 300 python profiles/tune_split.py; NREP=8 CELLS=16,16,32 timeout 300 python profiles/tune_split.py; NREP=8 CELLS=16,16,32 CHX_MD_OLD_DEAL=1 timeout 300 python profiles/tune_split.py; NREP=64 CELLS=16,16,32 STEPS=300 timeout 300 python profiles/tune_split.py ) 2>&1 | grep -E "TUNE|rror" > gpurun_out/r2_tune6.log
cat gpurun_out/r2_tune6.log
STEPS=200 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_v6.csv python profiles/tune_split.py > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/r2_launches_v6.csv 2>&1 | head -8
NREP=8 CELLS=16,16,32 STEPS=200 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_v6_8x8192.csv python profiles/tune_split.py > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/r2_launches_v6_8x8192.csv 2>&1 | head -12
