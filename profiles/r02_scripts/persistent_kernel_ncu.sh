python -c "import __graft_entry__ as g; g.build()"
timeout 300 python profiles/r02_debug_persist.py 2>&1 | grep -v "^$" | grep "n=60" > gpurun_out/r2_debug3.log
cat gpurun_out/r2_debug3.log
( timeout 300 python profiles/tune_split.py; CHX_MD_PERSIST=0 timeout 300 python profiles/tune_split.py; NREP=8 CELLS=16,16,32 timeout 300 python profiles/tune_split.py; NREP=8 CELLS=16,16,32 CHX_MD_PERSIST=0 timeout 300 python profiles/tune_split.py; NREP=64 CELLS=16,16,32 STEPS=300 timeout 300 python profiles/tune_split.py ) 2>&1 | grep -E "TUNE|rror" > gpurun_out/r2_tune4.log
cat gpurun_out/r2_tune4.log
STEPS=200 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_md_steps -s 12 -c 1 -o gpurun_out/r2_steps_v5c python profiles/tune_split.py > gpurun_out/r2_ncu3.log 2>&1
tail -2 gpurun_out/r2_ncu3.log
