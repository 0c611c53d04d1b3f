# refresh of the ncu artefacts that name SSW kernels, on the final code
set -x
O=gpurun_out/fg
mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-traffic-probe --pe-pairs 0 --seed-reads 0 > $O/ncu_launch.log 2>&1
PROF_READS=200000 PROF_SW_TASKS=200000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:sw_ --launch-skip 9 -c 9 -o $O/ssw python tools/prof.py > $O/ncu_ssw.log 2>&1
python tools/ncu_summary.py $O/ssw.ncu-rep > $O/ncu_full_ssw.txt 2>> $O/ncu_ssw.log
rm -f $O/*.ncu-rep
