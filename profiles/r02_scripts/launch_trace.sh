# occupancy over time of one step-kernel launch (trace build, not the product library)
export CHX_NVCC_EXTRA="-DCHX_TRACE=1"
python -c "import __graft_entry__ as g; g.build()"
( echo "TRACE ---- N=262144"; timeout 300 python profiles/launch_trace.py
  echo "TRACE ---- 8 x 8192"; NREP=8 CELLS=16,16,32 timeout 300 python profiles/launch_trace.py
  echo "TRACE ---- one wave"; CELLS=64,64,37 timeout 300 python profiles/launch_trace.py ) 2>&1 | grep -E "TRACE|rror|Trace" > gpurun_out/r2_launch_trace.log
cat gpurun_out/r2_launch_trace.log
