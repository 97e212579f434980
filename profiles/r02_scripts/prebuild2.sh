# pre-build forked from the end-of-run event (not from the caller's stream position)
python -c "import __graft_entry__ as g; g.build()"
timeout 600 python -m pytest tests/test_gpu_multistate.py -m gpu -x -q 2>&1 | tail -3
for P in 2 1; do
  CHX_REMD_PREBUILD=$P timeout 600 python bench.py --steps 2 --warmup 1 --inner 200 --no-mc --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); r=d['remd']; p=r['phases_rank0']
print('PRE=$P sweeps_per_s=%.2f ms=%.3f propagate=%.3f exch=%.3f mix=%.3f sha=%s u_sum=%r' % (r['sweeps_per_s'], r['ms_per_sweep'], p['propagate_ms'], p['energies_and_exchange_ms'], p['mix_ms'], r['fingerprint']['state_indices_sha1'], r['fingerprint']['u_sum']))"
done | tee gpurun_out/r2_prebuild2.log
