#!/bin/bash
# Round 2, GPU call R (1 GPU): per-tile timing of the backward gather (measurement build).
O=gpurun_out
mkdir -p $O
VFMB_VARIANT=tt VFMB_NVCC_EXTRA=-DVFMB_TILE_TIMING timeout 600 python scripts/gather_tiles.py > $O/r2r_gather_tiles.txt 2>&1
echo done
