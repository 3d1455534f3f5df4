# compute-sanitizer over every kernel family (tools/sanitize_smoke.py), three tools.
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/r2g_sanitizer_$tool.log 2>&1
done
tail -n 4 gpurun_out/r2g_sanitizer_*.log
