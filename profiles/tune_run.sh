#!/bin/bash
# Run-time knobs of the engine (no rebuild of the library): chunk length and internal skin.
for ch in 16 32 48; do for sk in 0.10 0.119 0.14; do
  CHX_MD_CHUNK=$ch INTERNAL_SKIN=$sk TIME=1 TAG="chunk=$ch skin=$sk" python profiles/prof_force.py 2>/dev/null | grep TIMING
done; done
