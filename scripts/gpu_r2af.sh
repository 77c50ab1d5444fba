#!/bin/bash
# Round 2, GPU call AF (1 GPU): ncu --set full of one warm step (final kernels) + summary; gpu tests of the knob test.
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_sampled.py -m gpu -q -x -k "knobs" > $O/r2af_tests.log 2>&1
python scripts/prof_step.py --reserve 1 > $O/r2af_prof_plain.log 2>&1 &&
timeout 900 ncu --set full --import-source on --clock-control none --cache-control none -s 104 -c 13 -f -o /tmp/r2af \
    python scripts/prof_step.py --reserve 1 > $O/r2af_ncu.log 2>&1
python scripts/ncu_summary.py /tmp/r2af.ncu-rep > $O/r2af_ncu_full_warm.jsonl 2> $O/r2af_summary.err
echo done
