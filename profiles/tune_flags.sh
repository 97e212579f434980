#!/bin/bash
# Rebuild the library on the GPU box with different compile-time knobs and time the force kernel and
# the step loop on the headline workload.  Usage (under gpurun): bash profiles/tune_flags.sh "<flags 1>" "<flags 2>" ...
for flags in "$@"; do
  CHX_NVCC_EXTRA="$flags" python -m chiron_b200.build --force > /dev/null 2>&1
  echo "== flags: $flags  $(grep -A2 'k_md_forceILb0ELi1ELb1' chiron_b200/lib/ptxas.log | grep -o 'Used [0-9]* registers' | head -1)"
  NREP=1 CELLS=64,64,64 python profiles/tune_split.py 2>/dev/null | grep TUNE
done
python -m chiron_b200.build --force > /dev/null 2>&1
