#!/bin/bash
# Round 2, GPU call AO (1 GPU): final verification -- suite, smoke, bench lines of every single-GPU workload.
O=gpurun_out
mkdir -p $O
rm -f $O/parity_bench_shapes.jsonl
python -c "from vae_b200 import _lib; print('stale', _lib._stale())" > $O/r2ao_stale.txt 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r2ao_tests.log 2>&1
echo "tests rc=$?" >> $O/r2ao_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2ao_smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2ao_bench_driver_args.json 2> $O/r2ao_bench_driver_args.err
timeout 600 python bench.py --no-cpu > $O/r2ao_bench_default.json 2> $O/r2ao_bench_default.err
for w in sideinfo ml100k fraction; do
  timeout 300 python bench.py --steps 300 --workload $w > $O/r2ao_bench_$w.json 2> $O/r2ao_bench_$w.err
done
echo done
