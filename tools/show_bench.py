"""prints the headline fields of a bench.py JSON line (file argument)"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d["value"], 2), d["unit"], "ms/step", round(d["ms_per_step"], 4), "stages", {k: round(v, 4) for k, v in d.get("stage_ms", {}).items()})
e = d.get("e2e", {})
print("e2e", round(e.get("value", 0), 2), "h2d", e.get("h2d_bytes_per_step"), "d2h", e.get("d2h_bytes_per_step"), "ms", e.get("ms_per_step"), "rle==grid", e.get("rle_stream_equals_grid"),
      "| full-grid", round(e.get("full_grid", {}).get("value", 0), 2), "serial", round(e.get("full_grid", {}).get("serial_value", 0), 2))
print("roofline", {k: d["roofline"][k] for k in ("kernel", "frac", "achieved")}, "step frac", d.get("roofline_step", {}).get("frac"))
print("parity", d.get("parity_checked"), (d.get("parity") or {}).get("mismatching_cells"), "launches", d.get("gpu_launches"), "clocks", d.get("clocks"))
for k in ("slab", "batch", "vessel"):
    if k in d: print(k, {kk: d[k][kk] for kk in list(d[k])[:8] if kk != "workload"})
