set -x
mkdir -p gpurun_out/r2u
O=gpurun_out/r2u
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_c1.json 2> $O/bench_c1.err; echo "rc=$?" >> $O/bench_c1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 900 python bench.py --steps 10 --warmup 3 --genome 3100000000 --read-len 150 --snp-rate 0.0047 --pe-pairs 1000000 --seed-reads 0 --cpu-sample 100000 > $O/bench_c2.json 2> $O/bench_c2.err; echo "rc=$?" >> $O/bench_c2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-traffic-probe --pe-pairs 0 --seed-reads 0 > $O/ncu_launch.log 2>&1
