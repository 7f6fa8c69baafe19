set -u
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > $O/r2_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2_pytest.log; tail -3 $O/r2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/r2_smoke.log 2>&1; tail -2 $O/r2_smoke.log
timeout 600 python bench.py > $O/r2_bench.json 2> $O/r2_bench.err; tail -c 300 $O/r2_bench.json; echo
CMD="python bench.py --steps 2 --warmup 3 --min-seconds 0 --no-cpu-baseline --no-extras --e2e-steps 2"
$CMD --mode default > $O/r2_plain_default.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_ncu_default.csv $CMD --mode default > $O/r2_ncu_list_default.log 2>&1
