mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 60 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-configs > gpurun_out/r2_bench_under_ncu.log 2>&1
for m in 0 1 2; do
OVC_EXPAND_NT=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-configs --no-cpu > gpurun_out/r2_bench_nt$m.json 2> gpurun_out/r2_bench_nt$m.err
done
OVC_E2E_CHUNK=50 timeout 300 python bench.py --steps 10 --warmup 3 --no-configs --no-cpu > gpurun_out/r2_bench_chunk50.json 2>/dev/null
OVC_E2E_CHUNK=200 timeout 300 python bench.py --steps 10 --warmup 3 --no-configs --no-cpu > gpurun_out/r2_bench_chunk200.json 2>/dev/null
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1; lscpu | head -30 >> gpurun_out/r2_topo.txt; nproc >> gpurun_out/r2_topo.txt
