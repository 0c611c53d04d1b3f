set -x
O=gpurun_out/r2u
mkdir -p $O
for N in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err
done
timeout 600 python tools/multi_bench.py --chunk 100000 > $O/multi_bench_100k.json 2> $O/multi_bench_100k.err
timeout 600 python tools/multi_bench.py --chunk 400000 > $O/multi_bench_400k.json 2> $O/multi_bench_400k.err
