# internal skin of the engine tables (nm; default 0.35 sigma = 0.119) after the round-2 kernel and rebuild changes
python -c "import __graft_entry__ as g; g.build()"
( for S in 0.100 0.119 0.135 0.150 0.119 0.135; do CHX_MD_SKIN=$S timeout 300 python profiles/tune_split.py | sed "s/^TUNE/TUNE skin=$S/"; done ) 2>&1 | grep -E "TUNE|rror" > gpurun_out/r2_tune_skin2.log
cat gpurun_out/r2_tune_skin2.log
