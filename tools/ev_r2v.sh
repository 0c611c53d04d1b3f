set -x
O=gpurun_out/r2v
mkdir -p $O
(timeout 1200 python tools/seed_bench.py --genome 100000000 --reads 1000000 > $O/seed_bench_100M.json 2> $O/seed_bench_100M.err) &
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --seed-reads 0 --no-traffic-probe > $O/bench_c1.json 2> $O/bench_c1.err; echo "rc=$?" >> $O/bench_c1.err
wait
QUICK=1 GENOME=20000000 READS=1000000 timeout 900 python tools/dropin_speed.py > $O/dropin_speed_quick.json 2> $O/dropin_speed_quick.err
