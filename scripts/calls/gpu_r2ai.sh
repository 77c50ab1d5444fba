#!/bin/bash
# Round 2, GPU call AI (1 GPU): what the driver runs at round end -- suite, smoke, bench with its arguments.
O=gpurun_out
mkdir -p $O
rm -f $O/parity_bench_shapes.jsonl
python -c "from vae_b200 import _lib; print('stale', _lib._stale())" > $O/r2ai_stale.txt 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r2ai_tests.log 2>&1
echo "tests rc=$?" >> $O/r2ai_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2ai_smoke.log 2>&1
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2ai_bench_reference.json 2> $O/r2ai_bench_reference.err
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2ai_bench_driver_args.json 2> $O/r2ai_bench_driver_args.err
timeout 600 python bench.py > $O/r2ai_bench_default.json 2> $O/r2ai_bench_default.err
timeout 300 python bench.py --no-cpu --steps 300 --workload sideinfo > $O/r2ai_bench_sideinfo.json 2> $O/r2ai_bench_sideinfo.err
echo done
