#!/bin/bash
# Round 2, GPU call Q (1 GPU): gather knobs (dynamic tiles, fence flavour, keep-whole threshold) side by side.
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_sampled.py tests/test_gpu_closed.py tests/test_gpu_bench_shapes.py -m gpu -q -x > $O/r2q_tests.log 2>&1
echo "tests rc=$?" >> $O/r2q_tests.log
timeout 300 python bench.py --no-cpu --steps 300 --workload sideinfo > $O/r2q_bench_sideinfo.json 2> $O/r2q_bench_sideinfo.err
timeout 300 python bench.py --no-cpu --steps 1000 > $O/r2q_bench.json 2> $O/r2q_bench.err
python scripts/gather_variants.py > $O/r2q_variants_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:k_gather --csv \
    --log-file $O/r2q_gather_variants.csv python scripts/gather_variants.py > $O/r2q_ncu.log 2>&1
echo done
