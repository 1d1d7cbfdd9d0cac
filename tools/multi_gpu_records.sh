# bench.py --gpus N and tools/dist_search.py under torchrun on N GPUs of one box: bash tools/multi_gpu_records.sh N  (records under gpurun_out/)
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2G_bench_n$N.json 2> gpurun_out/r2G_bench_n$N.err; echo "bench rc $?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/dist_search.py s150 10 > gpurun_out/r2G_dist_s150_n$N.json 2> gpurun_out/r2G_dist_n$N.err; echo "dist s150 rc $?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/dist_search.py s150d3 5 > gpurun_out/r2G_dist_s150d3_n$N.json 2>> gpurun_out/r2G_dist_n$N.err; echo "dist s150d3 rc $?"
tail -1 gpurun_out/r2G_dist_s150_n$N.json; tail -1 gpurun_out/r2G_dist_s150d3_n$N.json
