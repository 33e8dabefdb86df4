"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share of the summed device time per kernel.
usage: python scripts/summarize_launches.py launches.csv[.gz] "<command that was profiled>" > summary.md"""
import csv, gzip, io, sys
from collections import defaultdict

path = sys.argv[1]
cmd = sys.argv[2] if len(sys.argv) > 2 else "?"
raw = (gzip.open(path, "rt") if path.endswith(".gz") else open(path)).read()
start = raw.find('"ID"')
rows = list(csv.DictReader(io.StringIO(raw[start:])))
tot = defaultdict(lambda: [0.0, 0])
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v_ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
    k = r["Kernel Name"]
    tot[k][0] += v_ms
    tot[k][1] += 1
total = sum(v[0] for v in tot.values())
n = sum(v[1] for v in tot.values())
ours = sum(v[0] for k, v in tot.items() if "mc::" in k or k.startswith("mc::") or "temporal_attn" in k or "cross_attn" in k
           or "groupnorm" in k or "geglu" in k or "layernorm" in k or "cfg_ddim" in k or "motion_loss" in k
           or "bias_residual" in k or "top1" in k or "add_noise" in k or "gelu_lut" in k or "self_attn_short" in k)
print(f"# ncu launch list of `{cmd}` (profiling run, not a bench value)\n")
print(f"`ncu --metrics gpu__time_duration.sum --clock-control none`; {n} launches, {total:.1f} ms summed device time "
      "(serialised, cold cache: compare SHARES).")
print(f"Kernels of this package: {100 * ours / total:.1f} % of the summed time.\n")
print("| share | time (ms) | launches | kernel |\n|---|---|---|---|")
for k, (ms, c) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"| {100 * ms / total:.1f} % | {ms:.2f} | {c} | `{k[:100]}` |")
