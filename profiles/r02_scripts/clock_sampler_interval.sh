# does the in-region NVML clock sampler (10 ms) cost throughput?  10 ms vs 100 ms, alternating
python -c "import __graft_entry__ as g; g.build()"
for I in 0.01 0.1 0.01 0.1; do
  CHX_BENCH_CLOCK_S=$I timeout 600 python bench.py --no-remd --no-mc --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('CLK interval=$I value=%.1f ms_per_step=%.3f samples=%s' % (d['value'], d['ms_per_step'], d['clocks'].get('samples')))"
done | tee gpurun_out/r2_clock_interval.log
