#!/bin/bash
# compute-sanitizer over the small parity cases (gpurun): memcheck + racecheck of every kernel family.
set -u
out=gpurun_out/sanitizer.txt
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
{
echo "compute-sanitizer on a B200 (gpurun); $(date -u +%F)"
echo "== memcheck: ranking (golden, ragged, exact path, wide spans, hash lengths, two-class bins), pack push"
timeout 1500 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "golden or ragged or all_codes_equal or wide or hash_lengths or push" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|error" | tail -5
echo "== memcheck: real-valued mode"
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_real_valued.py -q -x -k "dyadic or edge or golden" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|error" | tail -5
echo "== memcheck: encoder (fp32 and implicit-GEMM tensor-core convolution, fc GEMM)"
timeout 1500 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_encoder.py -q -x -k "tensor_core_convolution or gemm_tf32" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|error" | tail -5
echo "== racecheck: ranking (golden), real-valued mode (edge cases)"
timeout 1500 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "golden" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|error" | tail -5
timeout 900 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_real_valued.py -q -x -k "edge or golden" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|error" | tail -5
} > $out 2>&1
cat $out
