#!/bin/bash
# Round 2, GPU call U (1 GPU): k_score restructured (32 samples per warp pass, rounds in flight, wide option).
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sampled.py tests/test_gpu_bench_shapes.py tests/test_gpu_dp.py -m gpu -q -x > $O/r2u_tests.log 2>&1
echo "tests rc=$?" >> $O/r2u_tests.log
timeout 300 python bench.py --no-cpu --steps 300 --workload sideinfo > $O/r2u_bench_sideinfo.json 2> $O/r2u_bench_sideinfo.err
timeout 300 python bench.py --no-cpu --steps 1000 > $O/r2u_bench.json 2> $O/r2u_bench.err
timeout 300 python bench.py --no-cpu --steps 1000 --tune score_wide=1 > $O/r2u_bench_wide.json 2> $O/r2u_bench_wide.err
timeout 300 python bench.py --no-cpu --steps 300 --workload sideinfo --tune score_wide=1 > $O/r2u_bench_sideinfo_wide.json 2> $O/r2u_bench_sideinfo_wide.err
python scripts/tune_variants.py score_wide=0 score_wide=1 > $O/r2u_variants_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"k_score|k_gather|k_stage|k_adam" --csv \
    --log-file $O/r2u_variants.csv python scripts/tune_variants.py score_wide=0 score_wide=1 > $O/r2u_ncu.log 2>&1
echo done
