# ncu --set full of K7 (encode_linear) and K8 (policy_tail) inside eager self-play transitions at 32768 envs (config 5's shard).
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_linear_kernel -s 2 -c 1 -o gpurun_out/r2_prof_k7 python tools/prof_kernels.py --which k7 --n 32768 > gpurun_out/r2_ncu_k7.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:policy_tail_kernel -s 2 -c 1 -o gpurun_out/r2_prof_k8 python tools/prof_kernels.py --which k8 --n 32768 > gpurun_out/r2_ncu_k8.log 2>&1
tail -3 gpurun_out/r2_ncu_k7.log gpurun_out/r2_ncu_k8.log; ls -la gpurun_out/*.ncu-rep
