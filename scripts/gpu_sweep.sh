#!/bin/bash
# sweep streaming-kernel launch parameters (run on the GPU box)
for S in 4096 65536; do
for ct in 4 8; do
for occ in 1 2 3; do
  echo "== S=$S CT=$ct OCC=$occ"
  S=$S VP_STREAM_CT=$ct VP_STREAM_OCC=$occ timeout 120 python scripts/gpu_probe.py 2>&1 | grep -E "flush|fit again" | head -3
done; done; done
