set -x
python -m pytest tests/test_gpu_sharding.py -x -q > gpurun_out/r6_tests_sh.log 2>&1; echo "tests rc=$?"
tail -n 3 gpurun_out/r6_tests_sh.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/r6_bench_2gpu_c.json 2> gpurun_out/r6_bench_2gpu_c.err; echo rc=$?
tail -c 300 gpurun_out/r6_bench_2gpu_c.err
