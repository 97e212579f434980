# step kernel with tile words staged through shared memory by cp.async.bulk (CHX_TILE_BULK=1) vs the register pipeline.
# CHX_NVCC_EXTRA is exported for every command of a variant: tests/conftest.py rebuilds the library when the flags change.
{
export CHX_NVCC_EXTRA="-DCHX_TILE_BULK=1"
python -c "from chiron_b200 import build; build.build()" 2>&1 | grep -i error
cuobjdump -sass chiron_b200/lib/libchiron_b200.so | grep -c UBLKCP
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_fullsize.py -m gpu -x -q -k "engine or langevin or full_size or 262144 or odd" 2>&1 | tail -3
timeout 300 python profiles/tune_split.py 2>&1 | grep -E "TUNE|rror"
NREP=8 CELLS=16,16,32 timeout 300 python profiles/tune_split.py 2>&1 | grep -E "TUNE|rror"
STEPS=100 timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_md_force -s 400 -c 1 -o gpurun_out/r2_force_bulk python profiles/tune_split.py > /dev/null 2>&1
cuobjdump -sass chiron_b200/lib/libchiron_b200.so | grep -c UBLKCP
export CHX_NVCC_EXTRA=""
python -c "from chiron_b200 import build; build.build()" 2>&1 | grep -i error
timeout 300 python profiles/tune_split.py 2>&1 | grep -E "TUNE|rror"
} > gpurun_out/r2_bulk.log 2>&1
cat gpurun_out/r2_bulk.log
