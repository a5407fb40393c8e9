#!/bin/bash
# compute-sanitizer over the small parity cases (gpurun): memcheck + racecheck of every kernel family.
set -u
out=gpurun_out/sanitizer.txt
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
F='grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|error"'
{
echo "compute-sanitizer on a B200 (gpurun); $(date -u +%F); round-2 tree"
echo "== memcheck: ranking (golden, ragged, exact path, wide spans, hash lengths, short / long codes on the tensor cores, queued select forced on small / ragged / dense shapes, dense walk, precision/recall, pack push, host path)"
timeout 1500 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "golden or ragged or all_codes_equal or wide or hash_lengths or push or short_and_long or dense or precision or class_sorted or host_entry or queued" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|error" | tail -5
echo "== memcheck: real-valued mode (candidate pass + full-row path)"
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_real_valued.py -q -x -k "dyadic or edge or golden or chunks or pm1" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|error" | tail -5
echo "== memcheck: real-valued mode, tensor-core contraction (HG_REAL_TC=1)"
HG_REAL_TC=1 timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_real_valued.py -q -x -k "edge or golden or chunks" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|error" | tail -5
echo "== memcheck: encoder (fused first stage at 32 and 64 pixels, separate kernels, implicit-GEMM convolution, fc GEMM, fused pool + LRN)"
timeout 1500 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_encoder.py -q -x -k "tensor_core_convolution or gemm_tf32 or fused_first_stage or other_image_sizes" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|error" | tail -5
echo "== racecheck: ranking (golden, dense walk, queued select), real-valued mode (edge cases), fused first encoder stage, encoder at other image sizes (warp pool + LRN kernel)"
timeout 1500 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "golden or dense or queued" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|error" | tail -5
timeout 900 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_real_valued.py -q -x -k "edge or golden" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|error" | tail -5
timeout 1500 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_encoder.py -q -x -k "(fused_first_stage and 32-True) or (other_image_sizes and 16-tf32x3)" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|error" | tail -5
} > $out 2>&1
cat $out
