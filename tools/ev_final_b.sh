# final single-GPU evidence, part B: whole-program timing, kernel sweep, seeding leg (run through gpurun)
set -x
O=gpurun_out/fb
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -k "two_driver or chunk_pair" > $O/pytest_two_driver.log 2>&1; echo "pytest rc=$?" >> $O/pytest_two_driver.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_c1.json 2> $O/bench_c1.err; echo "rc=$?" >> $O/bench_c1.err
QUICK=1 GENOME=20000000 READS=2000000 timeout 1500 python tools/dropin_speed.py > $O/dropin_speed_2M_t16.json 2> $O/dropin_speed_2M_t16.err
GENOME=20000000 READS=400000 timeout 1500 python tools/dropin_speed.py > $O/dropin_speed_400k.json 2> $O/dropin_speed_400k.err
timeout 900 python tools/sweep.py > $O/sweep.json 2> $O/sweep.err
timeout 900 python tools/seed_bench.py --genome 20000000 --reads 1000000 > $O/seed_bench.json 2> $O/seed_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'seed|locate|sort_long|gather|scan3' -c 300 --csv --log-file $O/seed_launches.csv python tools/seed_bench.py --genome 20000000 --reads 400000 --cpu-sample 20000 > $O/ncu_seed_launch.log 2>&1
