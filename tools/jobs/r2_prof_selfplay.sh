# Config 5 kernel-by-kernel: the self-play GPU test, stage times (CUDA events), the ncu launch list of three eager transitions.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k selfplay 2>&1 | tail -5 > gpurun_out/r2_selfplay_test.log
timeout 300 python tools/prof_selfplay.py --stages > gpurun_out/r2_selfplay_stages.json 2> gpurun_out/r2_selfplay_stages.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_selfplay_launches.csv python tools/prof_selfplay.py --eager 3 > gpurun_out/r2_selfplay_ncu.log 2>&1
cat gpurun_out/r2_selfplay_test.log; tail -3 gpurun_out/r2_selfplay_stages.json; tail -5 gpurun_out/r2_selfplay_stages.err
