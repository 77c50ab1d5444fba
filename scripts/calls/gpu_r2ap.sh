#!/bin/bash
# Round 2, GPU call AP (1 GPU): the added benchmark-shape case (second level of the cut-row finisher).
O=gpurun_out
mkdir -p $O
rm -f $O/parity_bench_shapes.jsonl
timeout 900 python -m pytest tests/test_gpu_bench_shapes.py -m gpu -q -x > $O/r2ap_tests.log 2>&1
echo "tests rc=$?" >> $O/r2ap_tests.log
echo done
