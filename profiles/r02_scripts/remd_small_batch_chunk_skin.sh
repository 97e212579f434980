# 8 replicas x 8,192 particles (a rank's share of config 5 at 8 GPUs): chunk length x internal skin
python -c "import __graft_entry__ as g; g.build()"
( for CS in "50 0.18" "100 0.22" "100 0.26" "100 0.30" "50 0.22" "34 0.15" "50 0.18"; do set -- $CS
  CHX_MD_CHUNK=$1 CHX_MD_SKIN=$2 LOCKSTEP=1 NO_PHASE=1 NGROUPS=1 NREP=8 SWEEPS=10 timeout 300 python profiles/tune_two_engines.py | sed "s/^TWO/TWO chunk=$1 skin=$2/"
  done ) 2>&1 | grep -E "TWO|rror|Trace" > gpurun_out/r2_remd_small_chunk_skin.log
cat gpurun_out/r2_remd_small_chunk_skin.log
