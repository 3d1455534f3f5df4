# K9 (ovc_wide_layers, tcgen05) on one B200: its tests under a short timeout, then the config-5 stage times with it.
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -q --timeout 100 -x -k "k9" 2>&1 | tail -30 > gpurun_out/r2j_pytest_k9.log
cat gpurun_out/r2j_pytest_k9.log
if grep -q "passed" gpurun_out/r2j_pytest_k9.log && ! grep -q "failed\|error" gpurun_out/r2j_pytest_k9.log; then
  timeout 300 python -m pytest tests/test_gpu_parity.py -q --timeout 200 -x -k "selfplay or k7 or k8" 2>&1 | tail -15 > gpurun_out/r2j_pytest_selfplay.log
  timeout 300 python tools/prof_selfplay.py --stages > gpurun_out/r2j_selfplay_stages_k7_k9_k8.json 2> gpurun_out/r2j_selfplay_stages.err
  cat gpurun_out/r2j_pytest_selfplay.log; tail -n 2 gpurun_out/r2j_selfplay_stages_k7_k9_k8.json; tail -3 gpurun_out/r2j_selfplay_stages.err
fi
