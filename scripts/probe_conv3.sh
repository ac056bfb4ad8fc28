#!/bin/bash
# correctness of the 3x3 halo kernel under both descriptor modes, then per-layer timing
for mode in 0 1; do
  echo "=== YMS_CONV3_DESC=$mode"
  YMS_CONV3_DESC=$mode timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "3x3 and not s2" 2>&1 | tail -4
done
echo "=== streaming-weights path forced (mode 0)"
YMS_CONV3_STREAM=1 timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "3x3 and not s2" 2>&1 | tail -3
