#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/tc_trace3.txt
for it in 3 40; do
while read -r a; do
  TC_TRACE_ITERS=$it DLGRAD_TC_TRACE=3 timeout 120 python scripts/tc_trace.py $a >> gpurun_out/tc_trace3.txt 2>&1 || echo "FAIL $a" >> gpurun_out/tc_trace3.txt
done <<'EOL'
4096 4096 4096 128 0 4
4096 4096 4096 192 0 4
8192 768 768 192 0 4
8192 768 768 128 0 4
1024 64 1024 64 0 4
EOL
done
cat gpurun_out/tc_trace3.txt
