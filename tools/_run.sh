set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/r6_bench_8gpu_c.json 2> gpurun_out/r6_bench_8gpu_c.err; echo rc=$?
tail -c 300 gpurun_out/r6_bench_8gpu_c.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/r6_bench_4gpu_c.json 2> gpurun_out/r6_bench_4gpu_c.err; echo rc=$?
