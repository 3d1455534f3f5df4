mkdir -p gpurun_out
timeout 600 python tools/k5_backtoback.py --tiles 32,64 --libs exitfirst > gpurun_out/r2_k5_backtoback.jsonl 2>&1
timeout 300 python tools/k5sweep.py --sizes 65536 --layouts cramped_room --tiles 32 --libs exitfirst > gpurun_out/r2_k5sweep_exitfirst.jsonl 2>&1
timeout 300 python tools/k5sweep.py --sizes 131072 --layouts asymmetric_advantages --tiles 64 --libs exitfirst >> gpurun_out/r2_k5sweep_exitfirst.jsonl 2>&1
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-configs --no-cpu > gpurun_out/r2_bench_e2e_$i.json 2> gpurun_out/r2_bench_e2e_$i.err
done
