mkdir -p gpurun_out
N=${NGPU:-2}
T=${TAG:-r03d}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err; echo "rc=$?" >> gpurun_out/${T}_bench_n$N.err
