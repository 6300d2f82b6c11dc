#!/bin/bash
export PHYSS_RT_VERBOSE=1
for d in 8 16 32; do
  python bench.py --state-dim $d --series 256 --sub-batch 256 --T 50 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline 2>&1 >/dev/null | grep "physs rt" | sort -u
done
