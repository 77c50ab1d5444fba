#!/bin/bash
# Round 2, GPU call AT (1 GPU): epoch-level metric parity test.
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_sampled.py -m gpu -q -x -k "epoch_level or training_loop" > $O/r2at_tests.log 2>&1
echo "tests rc=$?" >> $O/r2at_tests.log
echo done
