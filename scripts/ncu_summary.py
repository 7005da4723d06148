"""Summarise an `ncu --set full` report (exported with `ncu -i X.ncu-rep --page raw --csv`) as a markdown table.

    ncu -i gpurun_out/prof_targets.ncu-rep --page raw --csv > /tmp/raw.csv
    python scripts/ncu_summary.py /tmp/raw.csv [traffic.json kernel-name-prefix]
"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[0], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def col(r, name, default="0"):
    i = ix.get(name)
    return r[i].replace(",", "") if i is not None and r[i] != "" else default


units = dict(zip(hdr, rows[1]))
print("| # | kernel | grid | block | time us | DRAM read MB | DRAM write MB | DRAM active % | tensor pipe % | SM % | regs | smem KB/blk | warps active % |")
print("|---|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
out = []
for n, r in enumerate(data):
    t = float(col(r, "gpu__time_duration.sum"))
    t_us = t / 1e3 if units.get("gpu__time_duration.sum", "ns") in ("ns", "nsecond") else t
    rd, wr = float(col(r, "dram__bytes_read.sum")), float(col(r, "dram__bytes_write.sum"))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    rd *= scale.get(units.get("dram__bytes_read.sum", "byte"), 1.0)
    wr *= scale.get(units.get("dram__bytes_write.sum", "byte"), 1.0)
    rec = dict(kernel=col(r, "Kernel Name"), grid=col(r, "launch__grid_size"), block=col(r, "launch__block_size"), time_us=t_us,
               dram_read=rd, dram_write=wr,
               dram_pct=float(col(r, "dram__cycles_active.avg.pct_of_peak_sustained_elapsed")),
               tensor_pct=float(col(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                                    col(r, "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"))),
               sm_pct=float(col(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed")),
               regs=col(r, "launch__registers_per_thread"),
               smem=float(col(r, "launch__shared_mem_per_block_dynamic", "0")) / 1024.0 if units.get("launch__shared_mem_per_block_dynamic") == "byte" else float(col(r, "launch__shared_mem_per_block_dynamic", "0")),
               warps=float(col(r, "sm__warps_active.avg.pct_of_peak_sustained_active")))
    out.append(rec)
    print(f"| {n} | {rec['kernel']} | {rec['grid']} | {rec['block']} | {rec['time_us']:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | "
          f"{rec['dram_pct']:.1f} | {rec['tensor_pct']:.1f} | {rec['sm_pct']:.1f} | {rec['regs']} | {rec['smem']:.0f} | {rec['warps']:.1f} |")
if len(sys.argv) > 3:
    hits = [r for r in out if r["kernel"].startswith(sys.argv[3])]
    if hits:
        h = hits[-1]
        json.dump({"kernel": h["kernel"], "dram_bytes_read": h["dram_read"], "dram_bytes_write": h["dram_write"],
                   "bytes": h["dram_read"] + h["dram_write"], "time_us_under_ncu": h["time_us"]}, open(sys.argv[2], "w"), indent=1)
