# K7 / K8 + the sampling / return kernels on one B200: their tests, the config-5 stage times with and without them.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 300 -x -k "k7 or k8 or selfplay or sample_actions or accumulate" 2>&1 | tail -25 > gpurun_out/r2f_pytest_k78.log
timeout 300 python tools/prof_selfplay.py --stages > gpurun_out/r2f_selfplay_stages_k7_k8.json 2> gpurun_out/r2f_selfplay_stages.err
timeout 300 python tools/prof_selfplay.py --stages --tail 0 > gpurun_out/r2f_selfplay_stages_k7_glue.json 2>> gpurun_out/r2f_selfplay_stages.err
cat gpurun_out/r2f_pytest_k78.log; tail -n 2 gpurun_out/r2f_selfplay_stages_*.json; tail -3 gpurun_out/r2f_selfplay_stages.err
