python -c "import __graft_entry__ as g; g.build()"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 6 --warmup 3 --no-mc > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err
tail -2 gpurun_out/r2_bench_8gpu.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_8gpu.json'))
for k in ('value','ms_per_step','e2e','n_gpus','clocks'): print(k, d.get(k))
r=d['remd']; print({k:r[k] for k in ('sweeps_per_s','ms_per_sweep','phases_rank0')}); print(r['fingerprint']['state_indices_sha1'], r['fingerprint']['u_sum'])
PY
