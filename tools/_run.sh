set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r8_bench_2gpu_full.json 2> gpurun_out/r8_bench_2gpu_full.err; echo rc=$?
grep -v "NCCL\|^$\|OMP_NUM\|\*\*\*" gpurun_out/r8_bench_2gpu_full.err | tail -n 8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r8_bench_2gpu_ref.json 2> gpurun_out/r8_bench_2gpu_ref.err; echo rc=$?
