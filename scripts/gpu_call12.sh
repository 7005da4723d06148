#!/bin/bash
mkdir -p gpurun_out
for part in v4 perf4; do
  timeout 300 python scripts/tc_probe.py $part > gpurun_out/tc_probe_$part.log 2>&1; echo "exit $?" >> gpurun_out/tc_probe_$part.log
  grep -v "bn64 \|bn256 " gpurun_out/tc_probe_$part.log | tail -40
done
: > gpurun_out/tc_trace.txt
while read -r a; do
  timeout 120 python scripts/tc_trace.py $a >> gpurun_out/tc_trace.txt 2>&1 || echo "FAIL $a" >> gpurun_out/tc_trace.txt
done <<'EOL'
4096 4096 4096 128 0 4
1024 64 1024 64 0 4
EOL
