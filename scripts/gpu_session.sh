#!/bin/bash
# One GPU-box session: parity tests, then the bench lines, logs into gpurun_out/ (merged back by gpurun).
# usage: scripts/gpu_session.sh [pytest -k expression]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
K="${1:-}"
if [ -n "$K" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q -k "$K" --timeout 300 > gpurun_out/pytest.log 2>&1
else
  timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest.log 2>&1
fi
echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
