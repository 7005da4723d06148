#!/bin/bash
# ncu --set full on dp_adam / pack_multi / uniform_philox / split-K GEMM (after the same command ran clean)
mkdir -p gpurun_out
CMD="python scripts/ncu_new_kernels.py"
timeout 200 $CMD || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dp_adam|pack_multi|uniform_philox_9f230c9ab7f3|mm_tcts_k" -c 12 -f -o gpurun_out/prof_new $CMD > gpurun_out/ncu_new.log 2>&1
tail -2 gpurun_out/ncu_new.log; ls -la gpurun_out/prof_new.ncu-rep
