mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/r2_pytest_gpu.log
for m in 0 1; do
OVC_ENC_TEMPLATE=$m timeout 400 python tools/kbench.py --what k2 --sizes 32768,65536,262144,1048576 --ios 1 > gpurun_out/r2_kbench_k2_tmpl$m.jsonl 2>&1
done
OVC_ENC_TEMPLATE=0 timeout 300 python bench.py --workload config5 --steps 3 --warmup 3 > gpurun_out/r2_bench_config5_tmpl0.json 2>/dev/null
OVC_ENC_TEMPLATE=1 timeout 300 python bench.py --workload config5 --steps 3 --warmup 3 > gpurun_out/r2_bench_config5_tmpl1.json 2>/dev/null
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
