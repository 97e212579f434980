python -c "import __graft_entry__ as g; g.build()"
timeout 1500 python bench.py --steps 6 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
tail -3 gpurun_out/r2_bench1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench1.json'))
for k in ('value','ms_per_step','clocks','e2e','gpu_launches'): print(k, d.get(k))
print('roofline', {k:v for k,v in d['roofline'].items() if k in ('achieved','peak','frac','kernel_ms','traffic','step_frac')})
print('cpu', {k:v for k,v in (d.get('cpu_baseline') or {}).items() if k!='sample'})
print('remd', json.dumps(d.get('remd'))[:1800])
print('mc', json.dumps(d.get('mc'))[:2500])
PY
( CHX_MD_SKIN=0.10 timeout 300 python profiles/tune_split.py; CHX_MD_SKIN=0.13 timeout 300 python profiles/tune_split.py; CHX_MD_SKIN=0.15 timeout 300 python profiles/tune_split.py; ) 2>&1 | grep -E "TUNE|rror" > gpurun_out/r2_tune8.log
cat gpurun_out/r2_tune8.log
