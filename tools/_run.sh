set -x
python -m pytest tests/test_gpu_streamed.py -x -q > gpurun_out/r7_tests_stream.log 2>&1; echo "tests rc=$?"
tail -n 25 gpurun_out/r7_tests_stream.log
