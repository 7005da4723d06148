#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/tc_trace.txt
while read -r a; do
  timeout 120 python scripts/tc_trace.py $a >> gpurun_out/tc_trace.txt 2>&1 || echo "FAIL $a" >> gpurun_out/tc_trace.txt
done <<'EOL'
4096 4096 4096 256 0 3
4096 4096 4096 128 0 3
8192 768 768 256 0 3
EOL
head -40 gpurun_out/tc_trace.txt
