# ncu --set full of K9 (wide_layers_kernel, tcgen05) inside eager self-play transitions at 32768 envs.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wide_layers_kernel -s 2 -c 1 -o gpurun_out/r2_prof_k9 python tools/prof_kernels.py --which k8 --n 32768 > gpurun_out/r2_ncu_k9.log 2>&1
tail -n 3 gpurun_out/r2_ncu_k9.log; ls -la gpurun_out/r2_prof_k9.ncu-rep
