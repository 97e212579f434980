python -c "import __graft_entry__ as g; g.build()"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_t13.log
tail -4 gpurun_out/r2_t13.log
timeout 1500 python bench.py > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
tail -2 gpurun_out/r2_bench2.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench2.json'))
for k in ('value','ms_per_step','clocks','e2e','gpu_launches','lane_utilisation','table_rebuilds_in_timed_region'): print(k, d.get(k))
print('roofline', {k:v for k,v in d['roofline'].items() if k in ('achieved','peak','frac','kernel_ms','traffic','step_frac')})
print('cpu', {k:v for k,v in (d.get('cpu_baseline') or {}).items() if k!='sample'})
r=d['remd']; print('remd', {k:r[k] for k in ('sweeps_per_s','ms_per_sweep','phases_rank0','cpu_baseline')})
PY
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
