#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cuda_parity.py -x -q -m gpu -k "attention_block or fused_attention" > gpurun_out/attn_test.log 2>&1; echo "exit $?"; tail -25 gpurun_out/attn_test.log | cut -c1-220
