# (historic) the lightest blocks of a step launch as pieces: apply profiles/r02_cut_pieces_variant.diff to
# chiron_b200/csrc/engine.cu first -- the variant was measured slower and is not in the tree
python -c "import __graft_entry__ as g; g.build()"
( CHX_MD_CUT_P=1 timeout 300 python profiles/tune_split.py | sed "s/^TUNE/TUNE P=1/"
  for P in 2 4 8; do for F in 0.73 0.5 0.3; do CHX_MD_CUT_P=$P CHX_MD_CUT_FRAC=$F timeout 300 python profiles/tune_split.py | sed "s/^TUNE/TUNE P=$P F=$F/"; done; done ) 2>&1 | grep -E "TUNE|rror" > gpurun_out/r2_tune_cut.log
cat gpurun_out/r2_tune_cut.log
