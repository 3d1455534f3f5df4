mkdir -p gpurun_out
export PSN=$PWD/overcooked_ai_b200/csrc/libovc_b200_psn.so
OVC_B200_LIB=$PSN timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/r2_pytest_gpu_psn.log
timeout 300 python tools/k5sweep.py --sizes 65536,131072,1048576 --layouts cramped_room --tiles 64 --libs psn > gpurun_out/r2_k5sweep_psn.jsonl 2>&1
timeout 300 python tools/k5sweep.py --sizes 131072 --layouts asymmetric_advantages --tiles 64 --libs psn >> gpurun_out/r2_k5sweep_psn.jsonl 2>&1
timeout 300 python tools/k5sweep.py --sizes 262144 --layouts cramped_room,asymmetric_advantages,coordination_ring,forced_coordination,counter_circuit --tiles 128 --libs psn >> gpurun_out/r2_k5sweep_psn.jsonl 2>&1
timeout 300 python tools/k5sweep.py --sizes 65536 --layouts cramped_room --tiles 64 --formats codes --libs psn >> gpurun_out/r2_k5sweep_psn.jsonl 2>&1
OVC_B200_LIB=$PSN timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/r2_sanitizer_memcheck_psn.log 2>&1
OVC_B200_LIB=$PSN timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_psn.log 2>&1
