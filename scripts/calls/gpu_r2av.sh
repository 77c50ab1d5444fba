#!/bin/bash
# Round 2, GPU call AV (1 GPU): ragged-shape sweep of the step.
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_sampled.py -m gpu -q -x -k "ragged_shapes" > $O/r2av_tests.log 2>&1
echo "tests rc=$?" >> $O/r2av_tests.log
echo done
