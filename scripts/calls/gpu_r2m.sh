#!/bin/bash
# Round 2, GPU call M (1 GPU): full suite, sideinfo step breakdown, short-run (driver arguments) behaviour.
O=gpurun_out
mkdir -p $O
rm -f $O/parity_bench_shapes.jsonl
timeout 1500 python -m pytest tests -m gpu -q > $O/r2m_tests.log 2>&1
echo "tests rc=$?" >> $O/r2m_tests.log
timeout 300 python bench.py --no-cpu --steps 20 --warmup 5 > $O/r2m_bench_k20_w5.json 2> $O/r2m_bench_k20_w5.err
timeout 300 python bench.py --no-cpu --steps 20 --warmup 300 > $O/r2m_bench_k20_w300.json 2> $O/r2m_bench_k20_w300.err
timeout 300 python bench.py --no-cpu --steps 1000 > $O/r2m_bench.json 2> $O/r2m_bench.err
timeout 300 python bench.py --no-cpu --steps 300 --workload sideinfo > $O/r2m_bench_sideinfo.json 2> $O/r2m_bench_sideinfo.err
python scripts/prof_step.py --reserve 1 --workload sideinfo --rows 1000000 > $O/r2m_prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 78 -c 52 --csv --log-file $O/r2m_launches_sideinfo_warm.csv python scripts/prof_step.py --reserve 1 --workload sideinfo --rows 1000000 > $O/r2m_ncu.log 2>&1
echo done
