#!/bin/bash
# Stage-removal experiments on the attention kernel (run on the B200 box through gpurun, from the repo root).
# Needs a library built with SFM_EXTRA_NVCC_FLAGS=-DSFM_ATT_EXPERIMENTS (python -m simplesfm_b200.build --force).
for d in 0 1 2 4 8 16 9 31; do
  echo -n "SFM_ATT_DBG=$d  "
  SFM_ATT_DBG=$d timeout 120 python tools/bench_kernels.py < /dev/null 2>&1 | tail -1
done
