set -x
python -m pytest tests/test_gpu_hdiv.py tests/test_gpu_tensor_forms.py tests/test_gpu_golden.py tests/test_gpu_streamed.py -x -q > gpurun_out/r8_tests_bdm3.log 2>&1; echo "tests rc=$?"
tail -n 4 gpurun_out/r8_tests_bdm3.log
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"local_h1_load_stream" -s 1 -c 1 -o gpurun_out/r8_load -f python bench.py --no-cpu --no-e2e --no-varying --steps 2 --warmup 1 --no-graph > gpurun_out/r8_ncu_load.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:"SegPtrs|p2tet|local_h1_load|pair_gather|load_wt|node_gather" -c 30 --csv --log-file gpurun_out/r8_launches.csv python bench.py --no-cpu --no-e2e --no-varying --steps 3 --warmup 1 --no-graph > gpurun_out/r8_ncu_ll.log 2>&1
