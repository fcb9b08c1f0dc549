mkdir -p gpurun_out/r2p
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 -n 368 --steps 10 --warmup 3 --no-e2e > gpurun_out/r2p/bench_n8_weak_n368.json 2> gpurun_out/r2p/bench_n8_weak.err
tail -c 1200 gpurun_out/r2p/bench_n8_weak_n368.json; tail -3 gpurun_out/r2p/bench_n8_weak.err
