# both bench lines once more (the cpu_baseline leg now also compares its sample with the e2e results)
set -x
O=gpurun_out/fh
mkdir -p $O
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_c1.json 2> $O/bench_c1.err; echo "rc=$?" >> $O/bench_c1.err
timeout 900 python bench.py --steps 10 --warmup 3 --genome 3100000000 --read-len 150 --snp-rate 0.0047 --pe-pairs 1000000 --seed-reads 0 --cpu-sample 100000 > $O/bench_c2.json 2> $O/bench_c2.err; echo "rc=$?" >> $O/bench_c2.err
