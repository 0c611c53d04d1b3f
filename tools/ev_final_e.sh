# last evidence pass of the round on the final code: fuzz, all GPU tests, smoke, SSW timings, sweep, both bench lines
set -x
O=gpurun_out/ff
mkdir -p $O
timeout 300 python tools/fuzz_gpu.py 60 51 ssw > $O/fuzz_ssw.log 2>&1; echo "rc=$?" >> $O/fuzz_ssw.log
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
PROF_SW_TASKS=200000 timeout 500 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active --clock-control none -k regex:sw_ --csv --log-file $O/sw_kernels.csv python tools/prof.py > $O/prof.log 2>&1
timeout 600 python tools/sweep.py > $O/sweep.json 2> $O/sweep.err
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_c1.json 2> $O/bench_c1.err; echo "rc=$?" >> $O/bench_c1.err
timeout 900 python bench.py --steps 10 --warmup 3 --genome 3100000000 --read-len 150 --snp-rate 0.0047 --pe-pairs 1000000 --seed-reads 0 --cpu-sample 100000 > $O/bench_c2.json 2> $O/bench_c2.err; echo "rc=$?" >> $O/bench_c2.err
