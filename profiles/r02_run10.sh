python -c "import __graft_entry__ as g; g.build()"
timeout 900 python -m pytest tests/test_gpu_multigpu.py tests/test_gpu_multistate.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_t10.log
tail -4 gpurun_out/r2_t10.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 --no-mc > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
tail -2 gpurun_out/r2_bench_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_2gpu.json'))
for k in ('value','ms_per_step','e2e','n_gpus'): print(k, d.get(k))
r=d['remd']; print({k:r[k] for k in ('sweeps_per_s','ms_per_sweep','phases_rank0')}); print(r['fingerprint']['state_indices_sha1'], r['fingerprint']['u_sum'])
PY
