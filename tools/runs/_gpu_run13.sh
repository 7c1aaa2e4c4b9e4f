mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_fit.py -x -q 2>&1 | tail -15) > gpurun_out/s13_pytest_fit.log; tail -6 gpurun_out/s13_pytest_fit.log
timeout 400 python tools/fit_bench.py --out gpurun_out/s13_fit_bench.json 2>&1 | cut -c1-900
ncu --set full --clock-control none --import-source on -k regex:fit_count_kernel -s 4 -c 1 -o gpurun_out/s13_fit_count_dmv -f python tools/fit_bench.py --reps 3 > gpurun_out/s13_ncu.log 2>&1; tail -2 gpurun_out/s13_ncu.log
