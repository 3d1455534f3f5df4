mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest_gpu.log
python tools/k5sweep.py --sizes 65536,262144,1048576 --layouts cramped_room > gpurun_out/r2_k5sweep_cramped.jsonl 2>&1
python tools/k5sweep.py --sizes 65536 --layouts cramped_room --formats codes > gpurun_out/r2_k5sweep_codes.jsonl 2>&1
python tools/k5sweep.py --sizes 262144 --layouts cramped_room,asymmetric_advantages,coordination_ring,forced_coordination,counter_circuit --tiles 64,128 > gpurun_out/r2_k5sweep_mixed5.jsonl 2>&1
python tools/k5sweep.py --sizes 131072 --layouts asymmetric_advantages > gpurun_out/r2_k5sweep_asym.jsonl 2>&1
