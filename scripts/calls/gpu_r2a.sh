#!/bin/bash
# Round 2, GPU call A: full GPU test suite, bandwidth ceiling at the row update's size, k_adam_rows
# schedule variants, fused step with the Adam moments prefetched into L2 by the stage kernel.
O=gpurun_out
mkdir -p $O
rm -f $O/parity_bench_shapes.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2a_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2a_tests.log 2>&1
echo "tests rc=$?" >> $O/r2a_tests.log
./build/stream_ceiling 59336 > $O/r2a_ceiling.txt 2>&1
for v in "" bal balnopf nopf; do
  VFMB_VARIANT=$v timeout 300 python scripts/adam_micro.py >> $O/r2a_adam_micro.txt 2>&1
done
for t in 0 1 2 3; do
  timeout 300 python bench.py --no-cpu --steps 1000 --tune prefetch_mv=$t > $O/r2a_bench_pf$t.json 2> $O/r2a_bench_pf$t.err
done
for v in bal balnopf nopf; do
  VFMB_VARIANT=$v timeout 300 python bench.py --no-cpu --steps 1000 > $O/r2a_bench_$v.json 2> $O/r2a_bench_$v.err
  VFMB_VARIANT=$v timeout 300 python bench.py --no-cpu --steps 1000 --tune prefetch_mv=3 > $O/r2a_bench_${v}_pf3.json 2> $O/r2a_bench_${v}_pf3.err
done
timeout 400 python bench.py --steps 20 --warmup 5 > $O/r2a_bench_driver.json 2> $O/r2a_bench_driver.err
echo done
