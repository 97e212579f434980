#!/bin/bash
# like tune_flags.sh, with the workload in the environment: "ENV=val ...|nvcc flags"
for spec in "$@"; do
  envs="${spec%%|*}"; flags="${spec#*|}"
  CHX_NVCC_EXTRA="$flags" python -m chiron_b200.build --force > /dev/null 2>&1
  echo "== env: $envs flags: $flags"
  env $envs python profiles/tune_split.py 2>/dev/null | grep TUNE
done
python -m chiron_b200.build --force > /dev/null 2>&1
