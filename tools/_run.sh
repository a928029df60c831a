set -x
python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r5_tests2.log 2>&1; echo "tests rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:"local_h1_load|pair_gather" -c 4 --csv --log-file gpurun_out/r5_l1.csv python bench.py --no-cpu --no-e2e --no-varying --steps 2 --warmup 1 --no-graph > /dev/null 2>&1
NFA_DEBUG_NONODE=1 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:"local_h1_load|pair_gather" -c 4 --csv --log-file gpurun_out/r5_l2.csv python bench.py --no-cpu --no-e2e --no-varying --steps 2 --warmup 1 --no-graph > /dev/null 2>&1
grep -v "^==" gpurun_out/r5_l1.csv | cut -d, -f5,15- | head -6
grep -v "^==" gpurun_out/r5_l2.csv | cut -d, -f5,15- | head -6
