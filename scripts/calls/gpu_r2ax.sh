#!/bin/bash
# Round 2, GPU call AX (1 GPU): last verification of the committed tree -- suite, smoke, both bench arms.
O=gpurun_out
mkdir -p $O
rm -f $O/parity_bench_shapes.jsonl
python -c "from vae_b200 import _lib; print('stale', _lib._stale())" > $O/r2ax_stale.txt 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r2ax_tests.log 2>&1
echo "tests rc=$?" >> $O/r2ax_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2ax_smoke.log 2>&1
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2ax_bench_reference.json 2> $O/r2ax_bench_reference.err
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2ax_bench_driver_args.json 2> $O/r2ax_bench_driver_args.err
echo done
