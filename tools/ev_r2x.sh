set -x
O=gpurun_out/r2x
mkdir -p $O
(timeout 1500 python tools/seed_bench.py --genome 300000000 --reads 1000000 > $O/seed_bench_300M.json 2> $O/seed_bench_300M.err) &
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
PROF_SW_TASKS=200000 timeout 500 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:sw_ --csv --log-file $O/sw_kernels.csv python tools/prof.py > $O/prof.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --seed-reads 0 --no-traffic-probe > $O/bench_c1.json 2> $O/bench_c1.err; echo "rc=$?" >> $O/bench_c1.err
wait
