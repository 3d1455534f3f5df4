# Final round-2 validation on one B200 (gpurun --timeout 2400 -- bash tools/jobs/r2_final_1gpu.sh): GPU tests, smoke, bench (both arms),
# config-5 stage times and launch list, the bench launch list, compute-sanitizer over every kernel family.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/r2g_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2g_bench_reference.json 2> gpurun_out/r2g_bench_reference.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench_1gpu.json 2> gpurun_out/r2g_bench_1gpu.err
timeout 300 python tools/prof_selfplay.py --stages > gpurun_out/r2g_selfplay_stages_k7_k8.json 2> gpurun_out/r2g_selfplay_stages.err
timeout 300 python tools/prof_selfplay.py --stages --glue 0 --fused 0 > gpurun_out/r2g_selfplay_stages_library_only.json 2>> gpurun_out/r2g_selfplay_stages.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_selfplay_launches.csv python tools/prof_selfplay.py --eager 3 > gpurun_out/r2g_selfplay_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 60 --csv --log-file gpurun_out/r2g_launches_bench.csv python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-configs > gpurun_out/r2g_bench_under_ncu.log 2>&1
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/r2g_sanitizer_$tool.log 2>&1
done
cat gpurun_out/r2g_pytest_gpu.log; tail -2 gpurun_out/r2g_smoke.log; tail -c 400 gpurun_out/r2g_bench_1gpu.json; tail -3 gpurun_out/r2g_bench_1gpu.err; tail -n 3 gpurun_out/r2g_sanitizer_*.log
