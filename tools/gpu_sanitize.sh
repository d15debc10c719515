#!/bin/bash
# compute-sanitizer over the CUDA paths added in round 2 (memcheck; racecheck on the mbarrier / TMEM pipeline) plus the
# round-1 selection (tools/README.md).  Results: profiles/r2_v4_sanitizer.md
mkdir -p gpurun_out
S="compute-sanitizer --error-exitcode 1 --print-limit 20"
timeout 1200 $S --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "tensor_memory_kernel_matches_oracle or row_chunks or lazy_check" > gpurun_out/sanitize_memcheck_tmem.log 2>&1; echo "memcheck tmem rc=$?"; tail -3 gpurun_out/sanitize_memcheck_tmem.log
timeout 900 $S --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_train.py -q -x -k "c1_1h or c4_e16 or empty or out_of_range or bn or flat_adam_matches or fused_backward" > gpurun_out/sanitize_memcheck_parity.log 2>&1; echo "memcheck parity+train rc=$?"; tail -3 gpurun_out/sanitize_memcheck_parity.log
timeout 900 $S --tool racecheck python -m pytest tests/test_gpu_parity.py -q -x -k "tensor_memory_kernel_matches_oracle and (1.7-39-10-4-128-1.0 or 2.0-39)" > gpurun_out/sanitize_racecheck_tmem.log 2>&1; echo "racecheck tmem rc=$?"; tail -12 gpurun_out/sanitize_racecheck_tmem.log
timeout 600 $S --tool memcheck python -m pytest tests/test_gpu_mlp.py tests/test_gpu_zoo.py -q -x -k "not 4096-5120 and not scorer and not 4100" > gpurun_out/sanitize_memcheck_mlp.log 2>&1; echo "memcheck mlp+zoo rc=$?"; tail -3 gpurun_out/sanitize_memcheck_mlp.log
