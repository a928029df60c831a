set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/bench_config4.py --box 44 --chunk 2048 > gpurun_out/r7_config4_8gpu.jsonl 2> gpurun_out/r7_config4_8gpu.err; echo rc=$?
grep -v "NCCL\|^$\|OMP_NUM\|\*\*\*" gpurun_out/r7_config4_8gpu.err | tail -n 12
cut -c1-700 gpurun_out/r7_config4_8gpu.jsonl
