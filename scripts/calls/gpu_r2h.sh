#!/bin/bash
# Round 2, GPU call H: full suite incl. the fused mode-B path on emulated ranks, guard-band test.
O=gpurun_out
mkdir -p $O
rm -f $O/parity_bench_shapes.jsonl
timeout 1500 python -m pytest tests -m gpu -q > $O/r2h_tests.log 2>&1
echo "tests rc=$?" >> $O/r2h_tests.log
timeout 300 python bench.py --no-cpu --steps 1000 > $O/r2h_bench.json 2> $O/r2h_bench.err
echo done
