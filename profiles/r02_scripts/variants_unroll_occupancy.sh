# step-kernel variants: trip-loop unroll x occupancy target, and the deal's first criterion
run() { echo "== $1 | $2"; CHX_NVCC_EXTRA="$1" python -c "from chiron_b200 import build; build.build()" 2>&1 | grep -i error; env $2 timeout 300 python profiles/tune_split.py 2>&1 | grep -E "TUNE|rror"; }
{
run "" "CHX_MD_DEAL_KEY=0"
run "" "CHX_MD_DEAL_KEY=1"
run "-DCHX_TRIP_UNROLL=2 -DCHX_FORCE_WARPS_PER_SM=32" "X=1"
run "-DCHX_TRIP_UNROLL=2 -DCHX_FORCE_WARPS_PER_SM=28" "X=1"
run "-DCHX_TRIP_UNROLL=1 -DCHX_FORCE_WARPS_PER_SM=28" "X=1"
run "-DCHX_TRIP_UNROLL=2 -DCHX_FORCE_WARPS_PER_SM=24" "X=1"
} > gpurun_out/r2_variants11.log 2>&1
cat gpurun_out/r2_variants11.log
