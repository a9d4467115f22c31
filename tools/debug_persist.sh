#!/bin/bash
# Runs tools/debug_persist.py configuration by configuration, each in its own process under a timeout.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in 1 2 3 4 6 7 8 9 11 0; do
  echo "=== stop $k"
  timeout 180 python tools/debug_persist.py --stop $k 2>&1 | grep -v "^$" | tail -12
done
echo "=== stop 0, two layers"
timeout 180 python tools/debug_persist.py --stop 0 --layers 2 --batch 64 2>&1 | tail -4
echo "=== full"
timeout 300 python tools/debug_persist.py --full 2>&1 | tail -6
