"""ncu CSV (--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:<kernel>) ->
profiles/<tag>_temporal_traffic.json: DRAM bytes per launch of the roofline kernel, averaged over every launch of the
capture (one guided + one plain DDIM step at the bench shapes). bench.py's `roofline.traffic` reads it."""
import csv
import gzip
import json
import sys

src, dst = sys.argv[1], sys.argv[2]
op = gzip.open if src.endswith(".gz") else open
rows = {}
with op(src, "rt") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for r in csv.DictReader(lines):
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except Exception:
        continue
    unit = r.get("Metric Unit", "")
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6}.get(unit, 1.0)
    rows.setdefault(r["ID"], {"kernel": r["Kernel Name"]})[r["Metric Name"]] = v * mult
n = len(rows)
rd = sum(x.get("dram__bytes_read.sum", 0.0) for x in rows.values())
wr = sum(x.get("dram__bytes_write.sum", 0.0) for x in rows.values())
ns = sum(x.get("gpu__time_duration.sum", 0.0) for x in rows.values())
out = {"kernel": "temporal_attn_fwd_kernel", "launches": n, "dram_bytes_per_launch": (rd + wr) / max(n, 1),
       "dram_read_bytes_per_launch": rd / max(n, 1), "dram_write_bytes_per_launch": wr / max(n, 1),
       "avg_launch_us_under_ncu": ns / max(n, 1) / 1e3,
       "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum averaged over the {n} temporal_attn_fwd launches of one "
                 f"guided + one plain DDIM step at 16x512x512 ({src.split('/')[-1]}; cold cache, serialised replay)"}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out))
