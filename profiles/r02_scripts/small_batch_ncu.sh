# ncu --set full of the step kernel at 8 replicas x 8,192 particles (one rank's share of config 5 at 8 GPUs)
python -c "import __graft_entry__ as g; g.build()"
NREP=8 CELLS=16,16,32 STEPS=200 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_md_force -s 700 -c 1 -o gpurun_out/r2_step_8x8192 python profiles/tune_split.py > gpurun_out/r2_ncu_small.log 2>&1
tail -2 gpurun_out/r2_ncu_small.log
NREP=8 CELLS=16,16,32 STEPS=200 timeout 900 ncu --set full --cache-control none --clock-control none -k regex:k_md_force -s 700 -c 1 -o gpurun_out/r2_step_8x8192_warm python profiles/tune_split.py > gpurun_out/r2_ncu_small2.log 2>&1
tail -2 gpurun_out/r2_ncu_small2.log
