#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/tc_trace.txt
for e in "" mma2 mma1 nobsplit; do
while read -r a; do
  echo "## experiment='$e'" >> gpurun_out/tc_trace.txt
  DLGRAD_TC_EXPERIMENT=$e timeout 120 python scripts/tc_trace.py $a 2>&1 | awk 'NR<=2 || NR>=14' >> gpurun_out/tc_trace.txt || echo "FAIL $a" >> gpurun_out/tc_trace.txt
done <<'EOL'
4096 4096 4096 128 0 4
1024 64 1024 64 0 4
EOL
done
grep -c . gpurun_out/tc_trace.txt
