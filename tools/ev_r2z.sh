set -x
O=gpurun_out/r2z
mkdir -p $O
timeout 300 python tools/fuzz_gpu.py 90 11 ssw > $O/fuzz_ssw.log 2>&1; echo "rc=$?" >> $O/fuzz_ssw.log
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
PROF_SW_TASKS=200000 timeout 500 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:sw_ --csv --log-file $O/sw_kernels.csv python tools/prof.py > $O/prof.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --seed-reads 0 --no-traffic-probe > $O/bench_c1.json 2> $O/bench_c1.err; echo "rc=$?" >> $O/bench_c1.err
QUICK=1 GENOME=20000000 READS=1000000 timeout 900 python tools/dropin_speed.py > $O/dropin_speed_quick.json 2> $O/dropin_speed_quick.err
