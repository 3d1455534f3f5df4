# Multi-GPU run (gpurun --gpus N -- bash tools/jobs/r2_bench_multigpu.sh N): the GPU tests that need >= 2 GPUs, bench.py and its reference arm under torchrun.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "two_ranks or long_rollouts or sparse_event or golden_trajectories" 2>&1 | tail -15 > gpurun_out/r2_pytest_gpu_2gpu.log
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference_${N}gpu.json 2> gpurun_out/r2_bench_reference_${N}gpu.err
