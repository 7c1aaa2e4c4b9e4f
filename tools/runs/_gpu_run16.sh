mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/s16_pytest.log; tail -3 gpurun_out/s16_pytest.log
timeout 300 python bench.py > gpurun_out/s16_bench.json 2> gpurun_out/s16_bench.err; tail -c 600 gpurun_out/s16_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s16_bench_ref.json 2>> gpurun_out/s16_bench.err; tail -c 300 gpurun_out/s16_bench_ref.json
timeout 400 python tools/fit_bench.py --out gpurun_out/s16_fit_bench.json 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
