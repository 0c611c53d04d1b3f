set -x
O=gpurun_out/r2y
mkdir -p $O
timeout 300 python tools/fuzz_gpu.py 75 7 ssw > $O/fuzz_ssw.log 2>&1; echo "rc=$?" >> $O/fuzz_ssw.log
timeout 600 python -m pytest tests -m gpu -q -k "ssw or pair or dropin or narrow or chunk" > $O/pytest_gpu_subset.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_subset.log
PROF_SW_TASKS=200000 timeout 500 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:sw_ --csv --log-file $O/sw_kernels.csv python tools/prof.py > $O/prof.log 2>&1
timeout 900 python tools/dropin_scaling.py > $O/dropin_scaling.json 2> $O/dropin_scaling.err
