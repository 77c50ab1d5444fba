#!/bin/bash
# Round 2, GPU call AS (1 GPU): ncu --set full of one warm step (final kernels) + summary; gpu tests of the knob test.
O=gpurun_out
mkdir -p $O
python scripts/prof_step.py --reserve 1 > $O/r2as_prof_plain.log 2>&1 &&
timeout 900 ncu --set full --import-source on --clock-control none --cache-control none -s 104 -c 13 -f -o /tmp/r2as \
    python scripts/prof_step.py --reserve 1 > $O/r2as_ncu.log 2>&1
python scripts/ncu_summary.py /tmp/r2as.ncu-rep > $O/r2as_ncu_full_warm.jsonl 2> $O/r2as_summary.err
echo done
