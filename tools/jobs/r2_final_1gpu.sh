# Final round-2 validation on one B200 (gpurun --timeout 2400 -- bash tools/jobs/r2_final_1gpu.sh): GPU tests, smoke, bench (both arms),
# config-5 stage times and launch list, the bench launch list, ncu --set full of K9, compute-sanitizer over every kernel family.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/r2k_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2k_bench_reference.json 2> gpurun_out/r2k_bench_reference.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench_1gpu.json 2> gpurun_out/r2k_bench_1gpu.err
timeout 300 python tools/prof_selfplay.py --stages > gpurun_out/r2k_selfplay_stages_k7_k9_k8.json 2> gpurun_out/r2k_selfplay_stages.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2k_selfplay_launches.csv python tools/prof_selfplay.py --eager 3 > gpurun_out/r2k_selfplay_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 60 --csv --log-file gpurun_out/r2k_launches_bench.csv python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-configs > gpurun_out/r2k_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wide_layers_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_k9 python tools/prof_kernels.py --which k8 --n 32768 > gpurun_out/r2_ncu_k9.log 2>&1
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/r2k_sanitizer_$tool.log 2>&1
done
cat gpurun_out/r2k_pytest_gpu.log; tail -2 gpurun_out/r2k_smoke.log; tail -c 300 gpurun_out/r2k_bench_1gpu.json; tail -3 gpurun_out/r2k_bench_1gpu.err; tail -n 3 gpurun_out/r2k_sanitizer_*.log
