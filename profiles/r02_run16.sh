python -c "import __graft_entry__ as g; g.build()"
timeout 900 python -m pytest tests/test_gpu_multigpu.py tests/test_gpu_multistate.py tests/test_gpu_parity.py -m gpu -x -q -k "multi or remd or mixture or replica" 2>&1 | tail -5 > gpurun_out/r2_t16.log
tail -3 gpurun_out/r2_t16.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --no-mc --no-e2e > gpurun_out/r2_bench_2gpu_b.json 2> gpurun_out/r2_bench_2gpu_b.err
tail -2 gpurun_out/r2_bench_2gpu_b.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_2gpu_b.json'))
r=d['remd']; print({k:r[k] for k in ('sweeps_per_s','ms_per_sweep','phases_rank0')}); print(r['fingerprint']['state_indices_sha1'], r['fingerprint']['u_sum'])
PY
