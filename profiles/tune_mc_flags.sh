#!/bin/bash
# Rebuild with different compile-time knobs and report the config-3 Monte Carlo rates.
for flags in "$@"; do
  CHX_NVCC_EXTRA="$flags" python -m chiron_b200.build --force > /dev/null 2>&1
  echo "== flags: $flags"
  python profiles/bench_mc.py 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cfg3 displacement %.0f  barostat %.0f  cfg2 full-energy %.0f' % (d['cfg3_displacement_moves_per_s'], d['cfg3_barostat_moves_per_s'], d['cfg2_single_particle_full_energy_moves_per_s']))"
done
python -m chiron_b200.build --force > /dev/null 2>&1
