mkdir -p gpurun_out
timeout 300 python tools/k5sweep.py --sizes 65536,131072 --layouts cramped_room --tiles 64 --libs psn,hoist,hoist2 > gpurun_out/r2_k5sweep_hoist.jsonl 2>&1
timeout 300 python tools/k5sweep.py --sizes 131072 --layouts asymmetric_advantages --tiles 64 --libs psn,hoist,hoist2 >> gpurun_out/r2_k5sweep_hoist.jsonl 2>&1
OVC_B200_LIB=$PWD/overcooked_ai_b200/csrc/libovc_b200_hoist2.so timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/r2_pytest_gpu_hoist2.log
