#!/bin/bash
# compute-sanitizer over one small launch of every kernel family (scripts/sanitize_target.py) and, with >= 2 GPUs, memcheck
# over the 2-rank exchange (scripts/check_multigpu.py).  Summaries go to gpurun_out/sanitize_*.log (copied to profiles/).
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  timeout 1500 $CS --tool $tool --print-limit 20 python scripts/sanitize_target.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target ok|Error|hazard" gpurun_out/sanitize_$tool.log | head -8
done
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  timeout 1500 $CS --tool memcheck --target-processes all --print-limit 20 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
      --master-addr 127.0.0.1 --master-port 29577 scripts/check_multigpu.py > gpurun_out/sanitize_memcheck_2rank.log 2>&1
  echo "== memcheck 2-rank rc=$?"; grep -E "ERROR SUMMARY|multigpu ok|fused peer" gpurun_out/sanitize_memcheck_2rank.log | head -8
fi
