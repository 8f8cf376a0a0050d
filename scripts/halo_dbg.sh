#!/bin/bash
# timing experiments on the halo kernel: which role paces it?
for m in 0 1 2 3 4 6 7; do
  echo "== DDPM_HALO_DBG=$m"
  DDPM_HALO_DBG=$m timeout 120 python scripts/bench_conv.py --gn --impls 3 2>&1 | head -8
done
