#!/bin/bash
# Round 2, GPU call G: full GPU suite (predict_proba, closed autograd bridge ...), default bench lines,
# compute-sanitizer memcheck over smoke().
O=gpurun_out
mkdir -p $O
rm -f $O/parity_bench_shapes.jsonl
timeout 1500 python -m pytest tests -m gpu -q > $O/r2g_tests.log 2>&1
echo "tests rc=$?" >> $O/r2g_tests.log
timeout 400 python bench.py --steps 20 --warmup 5 > $O/r2g_bench_driver.json 2> $O/r2g_bench_driver.err
timeout 300 python bench.py --no-cpu --steps 1000 > $O/r2g_bench.json 2> $O/r2g_bench.err
timeout 300 python bench.py --no-cpu --steps 1000 --tune adam_pipe=0 > $O/r2g_bench_nopipe.json 2> $O/r2g_bench_nopipe.err
timeout 300 python bench.py --no-cpu --steps 1000 > $O/r2g_bench_again.json 2> $O/r2g_bench_again.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2g_smoke_plain.log 2>&1 &&
timeout 900 compute-sanitizer --tool memcheck --leak-check no python -c "import __graft_entry__ as g; g.smoke()" > $O/r2g_memcheck.log 2>&1
echo "memcheck rc=$?" >> $O/r2g_memcheck.log
echo done
