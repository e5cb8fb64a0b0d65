#!/usr/bin/env bash
# compute-sanitizer passes over the kernel tests (the reference has no race / memory checking of
# its own; the decode chain relies on hand-rolled sentinels and software grid barriers, the training kernels on shared-
# memory atomics). Run on a GPU box:  tools/sanitize.sh [memcheck|racecheck|synccheck|initcheck] [pytest -k expression]
# Reports land in gpurun_out/sanitize_<tool>.log. The persistent decode-chain kernel spins on flags written by other
# CTAs; under the sanitizer's serialisation that can take minutes per launch, so it is excluded by default.
set -u
TOOL="${1:-memcheck}"
EXPR="${2:-not dlinear and not generate and not fulldepth and not cfg}"
mkdir -p gpurun_out
compute-sanitizer --tool "$TOOL" --print-limit 50 --error-exitcode 9 \
  python -m pytest tests/test_train_gpu.py tests/test_ops_gpu.py tests/test_gemm_gpu.py -x -q -m gpu -k "$EXPR" \
  > "gpurun_out/sanitize_${TOOL}.log" 2>&1
rc=$?
tail -15 "gpurun_out/sanitize_${TOOL}.log"
echo "compute-sanitizer --tool $TOOL exit code $rc"
exit $rc
