#!/bin/bash
# compute-sanitizer passes over the GPU parity tests (run under gpurun).  Logs under gpurun_out/.
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests -q -m gpu -x > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$? $(tail -1 gpurun_out/sanitizer_memcheck.log)"
timeout 300 compute-sanitizer --tool synccheck --print-limit 4 python -m pytest tests/test_kernels_gpu.py -q -m gpu > gpurun_out/sanitizer_synccheck.log 2>&1
echo "synccheck rc=$? $(tail -1 gpurun_out/sanitizer_synccheck.log)"
if [ "${RACECHECK:-0}" == "1" ]; then   # ~4 minutes
timeout 600 compute-sanitizer --tool racecheck --print-limit 8 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention or linear or layernorm or round or posterior" > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$? $(tail -1 gpurun_out/sanitizer_racecheck.log)"
fi
