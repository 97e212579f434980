#!/bin/bash
# REMD (config 5) sweep time against chunk length and internal skin, 64 replicas (1 GPU) and 8 (one GPU of 8)
for nrep in 64 8; do for ch in 32 50; do for sk in 0.119 0.15 0.18; do
  echo -n "nrep=$nrep chunk=$ch skin=$sk  "
  NREP=$nrep SWEEPS=10 CHX_MD_CHUNK=$ch CHX_MD_SKIN=$sk python profiles/prof_remd.py 2>&1 | grep "ms/sweep"
done; done; done
