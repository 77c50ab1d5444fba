#!/bin/bash
# Round 2, GPU call AQ (1 GPU): ncu launch list of bench.py itself (graph replays), after a plain run.
O=gpurun_out
mkdir -p $O
timeout 300 python bench.py --no-cpu --steps 3 --warmup 3 > $O/r2aq_bench_plain.json 2> $O/r2aq_bench_plain.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 600 --csv --log-file $O/r2aq_launches_bench.csv \
    python bench.py --no-cpu --steps 3 --warmup 3 > $O/r2aq_ncu.log 2>&1
echo done
