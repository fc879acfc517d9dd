#!/bin/bash
# One gpurun call: attention unit tests + A/B timing; the full GPU suite and the bench only if the unit tests pass.
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_ops_gpu.py -x -q -k attention > gpurun_out/attn_tests.log 2>&1
rc=$?
tail -15 gpurun_out/attn_tests.log
echo "attention tests rc=$rc"
timeout -s KILL 300 python tools/attn_bench.py > gpurun_out/attn_bench.log 2>&1
echo "attn_bench rc=$?"
cat gpurun_out/attn_bench.log
if [ $rc -eq 0 ]; then
  timeout -s KILL 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1
  echo "gpu tests rc=$?"
  tail -5 gpurun_out/gpu_tests.log
  timeout -s KILL 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench rc=$?"
  cat gpurun_out/bench.json
fi
