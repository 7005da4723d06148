#!/bin/bash
# round 2 evidence run (one B200): bench (+reference arm), same-box ablations, step profile, ncu launch list, ncu --set full of the
# dominant kernels, the reference's small configs, pipeline traces.  Every ncu pass runs only after its command exited 0 without ncu.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,power.limit,memory.total --format=csv > $O/r02_gpu_box_env.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)" >> $O/r02_gpu_box_env.txt
if [ "$1" = "tests" ]; then
  DLGRAD_PARITY_REPORT=$O/r02_parity_report.jsonl timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > $O/r02_pytest_gpu.txt 2>&1
  echo "pytest exit $?"; tail -4 $O/r02_pytest_gpu.txt
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.txt 2>&1; echo "smoke exit $?"; tail -1 $O/r02_smoke.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.err; echo "bench exit $?"
[ "$2" = "quick" ] || { timeout 500 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_reference_arm.json 2> $O/r02_bench_reference_arm.err; echo "ref exit $?"; }
B="python bench.py --steps 10 --warmup 3 --no-ops --no-small --no-cpu"
{
  echo "same-box ablations (bench.py --steps 10 --warmup 3, ms/step):"
  for v in "" "DLGRAD_GROUP_LINEAR=0" "DLGRAD_FLASH_ATTN=0" "DLGRAD_HSPLIT_DEFER=0" "DLGRAD_LAZY_LINEAR=0" "DLGRAD_GROUP_LINEAR=0 DLGRAD_FLASH_ATTN=0 DLGRAD_HSPLIT_DEFER=0 DLGRAD_LAZY_LINEAR=0" ""; do
    r=$(env $v timeout 300 $B 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['value'],2), d['gpu_launches']//d['steps'])")
    echo "  [${v:-all on (default)}]  ms/step, steps/s, launches/step = $r"
  done
} > $O/r02_ablation.txt 2>&1; cat $O/r02_ablation.txt
timeout 300 python bench.py --profile > $O/r02_step_profile_events.txt 2>&1; echo "profile exit $?"
S="python bench.py --steps 2 --warmup 1 --no-ops --no-small --no-cpu"
timeout 300 $S > /dev/null 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r02_ncu_launches.csv $S > $O/r02_ncu_launches.log 2>&1; echo "ncu list exit $?"
# (the report is exported to CSV here and deleted: gpurun copies back at most 64 MiB)
timeout 900 ncu --set full --clock-control none -k regex:"flash_|mm_tcts" -s 190 -c 48 -f -o /tmp/r02_ncu_full_step $S > $O/r02_ncu_full_step.log 2>&1; echo "ncu full exit $?"; tail -2 $O/r02_ncu_full_step.log
ncu -i /tmp/r02_ncu_full_step.ncu-rep --page raw --csv > $O/r02_ncu_full_step_raw.csv 2>/dev/null; ls -la $O/r02_ncu_full_step_raw.csv
timeout 600 python scripts/small_configs.py > $O/r02_small_configs.jsonl 2> $O/r02_small_configs.err; echo "small exit $?"
[ "$2" = "quick" ] || { for m in fwd dv ds; do DLGRAD_FLASH_TRACE=$m timeout 120 python scripts/flash_trace.py > $O/r02_flash_trace_$m.txt 2>&1; done; echo traces done; }
ls -la $O | head -40
