set -x
python -m pytest tests/test_gpu_tensor_forms.py tests/test_gpu_streamed.py tests/test_gpu_golden.py -x -q > gpurun_out/r9_tests_tensor.log 2>&1; echo "tests rc=$?"
tail -n 4 gpurun_out/r9_tests_tensor.log
timeout 1500 python tools/bench_config4.py --box 23 --chunk 2048 > gpurun_out/r9_config4.jsonl 2> gpurun_out/r9_config4.err; echo rc=$?
tail -n 5 gpurun_out/r9_config4.err
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"local_hdiv_tensor" -s 1 -c 1 -o gpurun_out/r9_tensor -f python tools/bench_config4.py --box 8 --chunk 1024 > gpurun_out/r9_ncu_tensor.log 2>&1
