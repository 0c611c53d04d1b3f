# final single-GPU evidence, part A: tests, bench lines, ncu launch list and full captures (run through gpurun)
set -x
O=gpurun_out/fa
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_c1.json 2> $O/bench_c1.err; echo "rc=$?" >> $O/bench_c1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 900 python bench.py --steps 10 --warmup 3 --genome 3100000000 --read-len 150 --snp-rate 0.0047 --pe-pairs 1000000 --seed-reads 0 --cpu-sample 100000 > $O/bench_c2.json 2> $O/bench_c2.err; echo "rc=$?" >> $O/bench_c2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-traffic-probe --pe-pairs 0 --seed-reads 0 > $O/ncu_launch.log 2>&1
PROF_SW_TASKS=200000 timeout 500 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:sw_ --csv --log-file $O/sw_kernels.csv python tools/prof.py > $O/prof.log 2>&1
PROF_READS=2000000 PROF_SW_TASKS=200000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'nogap_fused|lv_filter|lv_tpp|scan_gap|lv_cigar' --launch-skip 5 -c 7 -o $O/verify python tools/prof.py > $O/ncu_verify.log 2>&1
python tools/ncu_summary.py $O/verify.ncu-rep > $O/ncu_full_verify.txt 2>> $O/ncu_verify.log
PROF_READS=200000 PROF_SW_TASKS=200000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:sw_ --launch-skip 8 -c 8 -o $O/ssw python tools/prof.py > $O/ncu_ssw.log 2>&1
python tools/ncu_summary.py $O/ssw.ncu-rep > $O/ncu_full_ssw.txt 2>> $O/ncu_ssw.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'seed_kernel|locate_kernel|sort_long' --launch-skip 6 -c 3 -o $O/seed python tools/seed_bench.py --genome 20000000 --reads 400000 --cpu-sample 20000 > $O/ncu_seed.log 2>&1
python tools/ncu_summary.py $O/seed.ncu-rep > $O/ncu_full_seed.txt 2>> $O/ncu_seed.log
rm -f $O/*.ncu-rep
