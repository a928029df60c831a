set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/bench_config4.py --n 8 --chunk 512 > gpurun_out/r7_config4_2gpu_small.jsonl 2> gpurun_out/r7_config4_2gpu_small.err; echo rc=$?
grep -v "NCCL\|^$\|OMP_NUM\|\*\*\*" gpurun_out/r7_config4_2gpu_small.err | tail -n 12
cut -c1-900 gpurun_out/r7_config4_2gpu_small.jsonl
