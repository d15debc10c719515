#!/bin/bash
# compute-sanitizer over the CUDA paths (memcheck on everything quick, racecheck on the shared-memory heavy kernels)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 --print-limit 20 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_zoo.py -q -x -k "not 4096 and not scorer" > gpurun_out/sanitize_memcheck_mlp.log 2>&1; echo "memcheck mlp+zoo rc=$?"; tail -4 gpurun_out/sanitize_memcheck_mlp.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -k "c1_1h or c4_e16 or empty or out_of_range" > gpurun_out/sanitize_memcheck_parity.log 2>&1; echo "memcheck parity rc=$?"; tail -4 gpurun_out/sanitize_memcheck_parity.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 --print-limit 20 python -m pytest tests/test_gpu_mlp.py -q -x -k "fast_path and 640" > gpurun_out/sanitize_racecheck_tail.log 2>&1; echo "racecheck tail rc=$?"; tail -4 gpurun_out/sanitize_racecheck_tail.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 --print-limit 20 python -m pytest tests/test_gpu_train.py -q -x -k "bn or flat_adam_matches" > gpurun_out/sanitize_memcheck_train.log 2>&1; echo "memcheck train rc=$?"; tail -4 gpurun_out/sanitize_memcheck_train.log
