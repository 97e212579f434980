# programmatic dependent launch, trigger after the tile loop
python -c "import __graft_entry__ as g; g.build()"
( for P in 0 1 0 1; do CHX_MD_PDL=$P timeout 300 python profiles/tune_split.py | sed "s/^TUNE/TUNE pdl=$P/"; done
  for P in 0 1 0 1; do CHX_MD_PDL=$P NREP=8 CELLS=16,16,32 timeout 300 python profiles/tune_split.py | sed "s/^TUNE/TUNE pdl=$P/"; done ) 2>&1 | grep -E "TUNE|rror" > gpurun_out/r2_tune_pdl2.log
cat gpurun_out/r2_tune_pdl2.log
