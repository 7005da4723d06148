"""Markdown summary of a `DLGRAD_PARITY_REPORT=<path> pytest tests -m gpu` log (tests/runner.rel_check / gemm_check).

    python scripts/parity_report_summary.py gpurun_out/r02_parity_report.jsonl > profiles/r02_parity_report_summary.md
"""
import json
import sys

rows = [json.loads(ln) for ln in open(sys.argv[1]) if ln.strip()]
asserted = [r for r in rows if r.get("elementwise")]
reported = [r for r in rows if not r.get("elementwise")]
print("# Per-element parity figures of the GPU suite (round 2)\n")
print(f"`DLGRAD_PARITY_REPORT=... pytest tests -m gpu` on one B200: {len(rows)} comparisons logged by `tests/runner.rel_check` / `gemm_check`.")
print("For every comparison: the worst per-element relative error over the entries with |expected| > 1e-3 x max, and the worst "
      "error of the remaining entries in ulps of the largest magnitude.\n")
print(f"## asserted per element ({len(asserted)} comparisons: pointwise / softmax forward values at the stated tolerance; GEMMs "
      "against tol x (|A||B|)_ij)\n")
print("| comparison | tolerance | worst per-element error / tolerance | small entries: worst error (ulp of max) |")
print("|---|---:|---:|---:|")
for r in sorted(asserted, key=lambda r: -r["worst_rel"] / r["tol"])[:40]:
    print(f"| {r['what']} | {r['tol']:g} | {r['worst_rel'] / r['tol']:.3f} | {r['rest_ulp']:.1f} |")
if len(asserted) > 40:
    print(f"| ... {len(asserted) - 40} more, all below | | | |")
print(f"\n## reported only ({len(reported)} comparisons: sums that cancel — gradients of whole models, reductions; the max-norm "
      "bound is asserted)\n")
print("| comparison | tolerance (of max) | worst per-element relative error | small entries: worst error (ulp of max) |")
print("|---|---:|---:|---:|")
for r in sorted(reported, key=lambda r: -r["worst_rel"])[:50]:
    print(f"| {r['what']} | {r['tol']:g} | {r['worst_rel']:.2e} | {r['rest_ulp']:.1f} |")
if len(reported) > 50:
    print(f"| ... {len(reported) - 50} more, all below | | | |")
