#!/bin/bash
# Round 2, GPU call T (1 GPU): hierarchical gather (block tiles, shared-memory combine): tests, timing, bench.
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sampled.py tests/test_gpu_bench_shapes.py tests/test_gpu_dp.py -m gpu -q -x > $O/r2t_tests.log 2>&1
echo "tests rc=$?" >> $O/r2t_tests.log
VFMB_VARIANT=tt VFMB_NVCC_EXTRA=-DVFMB_TILE_TIMING timeout 600 python scripts/gather_tiles.py > $O/r2t_gather_tiles.txt 2>&1
timeout 300 python bench.py --no-cpu --steps 300 --workload sideinfo > $O/r2t_bench_sideinfo.json 2> $O/r2t_bench_sideinfo.err
timeout 300 python bench.py --no-cpu --steps 1000 > $O/r2t_bench.json 2> $O/r2t_bench.err
python scripts/gather_variants.py > $O/r2t_variants_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:k_gather --csv \
    --log-file $O/r2t_gather_variants.csv python scripts/gather_variants.py > $O/r2t_ncu.log 2>&1
echo done
