#!/bin/bash
# How much does the L1 share of the 256 KB SM memory matter?  (smem carve-out up => L1 down)
for cfg in "" "AFF_TUNE_CARVEOUT=100" "AFF_TUNE_SMEM_PAD=5000" "AFF_TUNE_SMEM_PAD=10000"; do
  echo "== $cfg"; env $cfg python tools/profile_run.py 2368 3 2>&1 | grep "launch 2"
done
