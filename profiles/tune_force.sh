#!/bin/bash
# Rebuild the library with different compile-time knobs on the GPU box and time the force kernel
# and the full step loop.  Usage (under gpurun): bash profiles/tune_force.sh "<extra flags 1>" "<extra flags 2>" ...
for flags in "$@"; do
  CHX_NVCC_EXTRA="$flags" python -m chiron_b200.build --force > /dev/null 2>&1
  grep -A2 "k_md_forceILb0" chiron_b200/lib/ptxas.log | grep -o "Used [0-9]* registers" | head -1
  TIME=1 TAG="$flags" python profiles/prof_force.py 2>/dev/null | grep -E "TIMING|lane_util"
done
python -m chiron_b200.build --force > /dev/null 2>&1
