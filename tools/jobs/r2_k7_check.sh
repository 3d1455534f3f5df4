# K7 (ovc_encode_linear) + the sampling / return kernels on one B200: their tests, the config-5 stage times with and without them.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 300 -x -k "k7 or selfplay or sample_actions or accumulate" 2>&1 | tail -25 > gpurun_out/r2d_pytest_k7.log
timeout 300 python tools/prof_selfplay.py --stages > gpurun_out/r2d_selfplay_stages_fused_glue.json 2> gpurun_out/r2d_selfplay_stages.err
timeout 300 python tools/prof_selfplay.py --stages --glue 0 > gpurun_out/r2d_selfplay_stages_fused.json 2>> gpurun_out/r2d_selfplay_stages.err
timeout 300 python tools/prof_selfplay.py --stages --glue 0 --fused 0 > gpurun_out/r2d_selfplay_stages_before.json 2>> gpurun_out/r2d_selfplay_stages.err
cat gpurun_out/r2d_pytest_k7.log; tail -n 2 gpurun_out/r2d_selfplay_stages_*.json; tail -3 gpurun_out/r2d_selfplay_stages.err
