# The round-2 validation run on one B200 (gpurun --timeout 3000 -- bash tools/jobs/r2_validate_1gpu.sh): GPU tests, smoke, bench (both arms),
# K5 sweep, ncu captures of K5 at four launch shapes, the bench launch list, compute-sanitizer.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
timeout 600 python tools/k5sweep.py --sizes 65536,131072,262144,1048576 --layouts cramped_room --tiles 64 > gpurun_out/r2_k5sweep_final.jsonl 2>&1
timeout 300 python tools/k5sweep.py --sizes 131072 --layouts asymmetric_advantages --tiles 64 >> gpurun_out/r2_k5sweep_final.jsonl 2>&1
timeout 300 python tools/k5sweep.py --sizes 262144 --layouts cramped_room,asymmetric_advantages,coordination_ring,forced_coordination,counter_circuit --tiles 128 >> gpurun_out/r2_k5sweep_final.jsonl 2>&1
timeout 300 python tools/k5sweep.py --sizes 65536 --layouts cramped_room --tiles 64 --formats codes >> gpurun_out/r2_k5sweep_final.jsonl 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 2 -c 1 -o gpurun_out/r2_prof_k5_config2 python tools/prof_kernels.py --which k5 > gpurun_out/r2_ncu_k5_config2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 2 -c 1 -o gpurun_out/r2_prof_k5_config4 python tools/prof_kernels.py --which k5 --n 131072 --layouts asymmetric_advantages > gpurun_out/r2_ncu_k5_config4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 2 -c 1 -o gpurun_out/r2_prof_k5_config3 python tools/prof_kernels.py --which k5 --n 262144 --layouts cramped_room,asymmetric_advantages,coordination_ring,forced_coordination,counter_circuit > gpurun_out/r2_ncu_k5_config3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 2 -c 1 -o gpurun_out/r2_prof_k5_target python tools/prof_kernels.py --which k5 --n 131072 > gpurun_out/r2_ncu_k5_target.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 60 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-configs > gpurun_out/r2_bench_under_ncu.log 2>&1
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/r2_sanitizer_$tool.log 2>&1
done
