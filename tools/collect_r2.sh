#!/bin/bash
# Final evidence run of round 2 on one B200 (gpurun): tests, smoke, bench lines, config sweep, ncu.
set -u
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > $O/r2_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2_pytest.log; tail -3 $O/r2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/r2_smoke.log 2>&1; tail -2 $O/r2_smoke.log
timeout 600 python bench.py > $O/r2_bench.json 2> $O/r2_bench.err; tail -c 300 $O/r2_bench.json; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference_arm.json 2> $O/r2_bench_reference_arm.err; tail -c 200 $O/r2_bench_reference_arm.json; echo
[ -n "${QUICK-}" ] || { bash tools/ab_frontends.sh $O/r2_frontends_ab.jsonl > /dev/null; }
bash profiles/run_configs.sh > /dev/null 2>&1; cp $O/configs.jsonl $O/r2_configs.jsonl; grep -c '^{' $O/r2_configs.jsonl
CMD="python bench.py --steps 2 --warmup 3 --min-seconds 0 --no-cpu-baseline --no-extras --e2e-steps 2"
[ -n "${QUICK-}" ] || { $CMD > $O/r2_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_ncu.csv $CMD > $O/r2_ncu_list.log 2>&1; }
$CMD --blur-scale 20 > $O/r2_plain_k97.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_ncu_k97.csv $CMD --blur-scale 20 > $O/r2_ncu_list_k97.log 2>&1
[ -n "${QUICK-}" ] || { $CMD --mode default > $O/r2_plain_default.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_ncu_default.csv $CMD --mode default > $O/r2_ncu_list_default.log 2>&1; }
[ -n "${QUICK-}" ] || { $CMD > $O/r2_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 1 -o $O/r2_fused $CMD > $O/r2_ncu_full.log 2>&1; }
$CMD --blur-scale 20 > $O/r2_plain_k97.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_wide -s 4 -c 2 -o $O/r2_wide $CMD --blur-scale 20 > $O/r2_ncu_full_k97.log 2>&1
ls -la $O/r2_wide.ncu-rep
