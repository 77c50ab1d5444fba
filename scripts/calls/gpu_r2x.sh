#!/bin/bash
# Round 2, GPU call X (1 GPU): k_stage restructured; wide knobs; adam_reserve default.
O=gpurun_out
mkdir -p $O
rm -f $O/parity_bench_shapes.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2x_tests.log 2>&1
echo "tests rc=$?" >> $O/r2x_tests.log
run() { name=$1; shift; timeout 300 "$@" > $O/r2x_$name.json 2> $O/r2x_$name.err; }
run ml20m_s0 python bench.py --no-cpu --steps 1000 --tune score_wide=1
run ml20m_s1 python bench.py --no-cpu --steps 1000 --tune score_wide=1 --tune stage_wide=1
run ml20m_00 python bench.py --no-cpu --steps 1000
run side_s0 python bench.py --no-cpu --steps 300 --workload sideinfo --tune score_wide=1
run side_s1 python bench.py --no-cpu --steps 300 --workload sideinfo --tune score_wide=1 --tune stage_wide=1
run side_s1_ar0 python bench.py --no-cpu --steps 300 --workload sideinfo --tune score_wide=1 --tune stage_wide=1 --tune adam_reserve=0
python scripts/tune_variants.py score_wide=1,stage_wide=0 score_wide=1,stage_wide=1 > $O/r2x_variants_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"k_score|k_gather|k_stage|k_adam" --csv \
    --log-file $O/r2x_variants.csv python scripts/tune_variants.py score_wide=1,stage_wide=0 score_wide=1,stage_wide=1 > $O/r2x_ncu.log 2>&1
echo done
