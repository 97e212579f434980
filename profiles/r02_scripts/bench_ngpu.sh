# usage: gpurun --gpus N -- bash profiles/r02_scripts/bench_ngpu.sh N
N=${1:-2}
python -c "import __graft_entry__ as g; g.build()"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 6 --warmup 3 --no-mc > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
tail -2 gpurun_out/r2_bench_${N}gpu.err | cut -c1-300
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_${N}gpu.json'))
for k in ('value','ms_per_step','n_gpus'): print(k, d.get(k))
print('e2e', d['e2e']['value'])
r=d['remd']; print({k:r[k] for k in ('sweeps_per_s','ms_per_sweep','phases_rank0')}); print(r['fingerprint']['state_indices_sha1'], r['fingerprint']['u_sum'])
PY
