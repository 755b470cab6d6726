set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_gather.py > gpurun_out/r02f_gather.log 2>&1
tail -3 gpurun_out/r02f_gather.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02f_bench_n2.json 2> gpurun_out/r02f_bench_n2.err
tail -5 gpurun_out/r02f_bench_n2.err; head -c 3000 gpurun_out/r02f_bench_n2.json
