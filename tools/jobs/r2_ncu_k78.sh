# ncu --set full of config 5's own kernels — K7 (encode_linear), K9 (wide_layers, tcgen05), K8 (policy_tail) — inside eager
# self-play transitions at 32768 envs (gpurun --timeout 1500 -- bash tools/jobs/r2_ncu_k78.sh); read here with
# python tools/ncu_summarize.py --tag r2 config5_k7=gpurun_out/r2_prof_k7.ncu-rep:32768x1 config5_k9=gpurun_out/r2_prof_k9.ncu-rep:65536x1 config5_k8=gpurun_out/r2_prof_k8.ncu-rep:65536x1
mkdir -p gpurun_out
for k in k7:encode_linear_kernel k9:wide_layers_kernel k8:policy_tail_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:${k#*:} -s 2 -c 1 -f -o gpurun_out/r2_prof_${k%%:*} python tools/prof_kernels.py --which k8 --n 32768 > gpurun_out/r2_ncu_${k%%:*}.log 2>&1
done
tail -n 3 gpurun_out/r2_ncu_k7.log gpurun_out/r2_ncu_k9.log gpurun_out/r2_ncu_k8.log; ls -la gpurun_out/*.ncu-rep
