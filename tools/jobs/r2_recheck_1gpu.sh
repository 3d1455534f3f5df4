# Re-check of HEAD on one B200: GPU tests, smoke, bench (both arms), config-5 stage times.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/r2b_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2b_bench_reference.json 2> gpurun_out/r2b_bench_reference.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_1gpu.json 2> gpurun_out/r2b_bench_1gpu.err
timeout 300 python tools/prof_selfplay.py --stages > gpurun_out/r2b_selfplay_stages.json 2> gpurun_out/r2b_selfplay_stages.err
cat gpurun_out/r2b_pytest_gpu.log; tail -2 gpurun_out/r2b_smoke.log; tail -c 600 gpurun_out/r2b_bench_1gpu.json; tail -3 gpurun_out/r2b_bench_1gpu.err
