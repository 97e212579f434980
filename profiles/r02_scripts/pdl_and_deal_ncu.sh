# programmatic dependent launch of the step kernels (CHX_MD_PDL=0/1) + ncu of the deal kernel
python -c "import __graft_entry__ as g; g.build()"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_fullsize.py tests/test_gpu_multistate.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t_pdl.log
tail -5 gpurun_out/r2_t_pdl.log
( for P in 0 1 0 1; do CHX_MD_PDL=$P timeout 300 python profiles/tune_split.py | sed "s/^TUNE/TUNE pdl=$P/"; done
  for P in 0 1 0 1; do CHX_MD_PDL=$P NREP=8 CELLS=16,16,32 timeout 300 python profiles/tune_split.py | sed "s/^TUNE/TUNE pdl=$P/"; done
  for P in 0 1; do CHX_MD_PDL=$P NREP=64 CELLS=16,16,32 STEPS=300 timeout 300 python profiles/tune_split.py | sed "s/^TUNE/TUNE pdl=$P/"; done ) 2>&1 | grep -E "TUNE|rror" > gpurun_out/r2_tune_pdl.log
cat gpurun_out/r2_tune_pdl.log
STEPS=200 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_md_deal -s 3 -c 1 -o gpurun_out/r2_deal python profiles/tune_split.py > gpurun_out/r2_ncu_deal.log 2>&1
tail -2 gpurun_out/r2_ncu_deal.log
