# SSW traceback stage check: fuzz, tests, kernel timings, sweep (run through gpurun)
set -x
O=gpurun_out/fe
mkdir -p $O
timeout 300 python tools/fuzz_gpu.py 60 41 ssw > $O/fuzz_ssw.log 2>&1; echo "rc=$?" >> $O/fuzz_ssw.log
timeout 900 python -m pytest tests -m gpu -q -k "ssw or pair or narrow or dropin" > $O/pytest_gpu_subset.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_subset.log
PROF_SW_TASKS=200000 timeout 500 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:sw_ --csv --log-file $O/sw_kernels.csv python tools/prof.py > $O/prof.log 2>&1
timeout 600 python tools/sweep.py > $O/sweep.json 2> $O/sweep.err
