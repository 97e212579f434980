python -c "import __graft_entry__ as g; g.build()"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2_tests_final.log
tail -3 gpurun_out/r2_tests_final.log
timeout 1500 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
tail -1 gpurun_out/r2_bench_final.err | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --inner 200 --no-remd --no-mc --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_under_ncu.json 2>/dev/null
python profiles/summarize_launches.py gpurun_out/r2_launches_bench.csv 2>&1 | head -14
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_final.json')); r=json.load(open('gpurun_out/r2_bench_reference.json'))
print('ours', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['roofline']['kernel_ms'], d['clocks'])
print('ref', r['value'], r['ms_per_step'], r['steps'], {k:v for k,v in r['cpu_baseline'].items() if k!='sample'})
print('same_config', d['config']==r['config'], 'ratio', d['value']/r['value'], 'e2e ratio', d['e2e']['value']/r['value'])
print('remd', d['remd']['sweeps_per_s'], d['remd']['cpu_baseline'])
print('mc', {k:v for k,v in d['mc'].items() if 'moves_per_s' in k or 'steps_per_s' in k}, d['mc']['cpu_baseline'])
PY
