# does the step kernel pay for its 1.73 waves (8192 one-warp CTAs on 4736 warp slots)?  same density, 1.0 / 1.73 / 2.0 / 3.0 waves
python -c "import __graft_entry__ as g; g.build()"
( for C in 64,64,37 64,64,64 64,64,74 64,64,111 64,64,55 ; do NREP=1 CELLS=$C STEPS=400 timeout 300 python profiles/tune_split.py; done ) 2>&1 | grep -E "TUNE|rror" > gpurun_out/r2_waves.log
cat gpurun_out/r2_waves.log
