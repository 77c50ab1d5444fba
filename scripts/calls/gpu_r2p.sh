#!/bin/bash
# Round 2, GPU call P (1 GPU): gather tiles snapped to row boundaries; suite + sideinfo/ml20m + ncu of k_gather.
O=gpurun_out
mkdir -p $O
rm -f $O/parity_bench_shapes.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2p_tests.log 2>&1
echo "tests rc=$?" >> $O/r2p_tests.log
timeout 300 python bench.py --no-cpu --steps 300 --workload sideinfo > $O/r2p_bench_sideinfo.json 2> $O/r2p_bench_sideinfo.err
timeout 300 python bench.py --no-cpu --steps 1000 > $O/r2p_bench.json 2> $O/r2p_bench.err
timeout 300 python bench.py --no-cpu --steps 300 --workload ml100k > $O/r2p_bench_ml100k.json 2> $O/r2p_bench_ml100k.err
python scripts/prof_step.py --reserve 1 --workload sideinfo --rows 1000000 > $O/r2p_prof_plain.log 2>&1 &&
timeout 600 ncu --set full --import-source on --clock-control none --cache-control none -k regex:k_gather -s 6 -c 1 -f -o /tmp/r2p \
    python scripts/prof_step.py --reserve 1 --workload sideinfo --rows 1000000 > $O/r2p_ncu.log 2>&1
ncu -i /tmp/r2p.ncu-rep --page raw --csv > $O/r2p_gather_raw.csv 2>/dev/null
ncu -i /tmp/r2p.ncu-rep --page details --csv > $O/r2p_gather_details.csv 2>/dev/null
ncu -i /tmp/r2p.ncu-rep --page source --csv > $O/r2p_gather_source.csv 2>/dev/null
echo done
