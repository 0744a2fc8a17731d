"""Turns what tools/collect_profiles.sh left in gpurun_out/ into the tracked summaries under profiles/ (run here, needs `ncu` to read the reports)."""
import collections, csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"

# ---- bench lines
for src, dst in (("bench_cfg3.json", f"{tag}_bench.json"), ("bench_reference.json", f"{tag}_bench_reference.json"), ("bench_batch.json", f"{tag}_bench_batch_1gpu.json"),
                 ("bench_batch_1job.json", f"{tag}_bench_batch_1gpu_1job.json"), ("bench_dataset.json", f"{tag}_bench_dataset_1gpu.json"),
                 ("flood_timings.txt", f"{tag}_flood_timings.txt"), ("voxelizer_timings.txt", f"{tag}_voxelizer_timings.txt")):
    if os.path.exists(os.path.join(G, src)):
        shutil.copyfile(os.path.join(G, src), os.path.join(P, dst))

# ---- launch list
rows = [r for r in csv.reader(open(os.path.join(G, "launches.csv"))) if len(r) > 10 and r[0].isdigit()]
shutil.copyfile(os.path.join(G, "launches.csv"), os.path.join(P, f"{tag}_launches.csv"))
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    agg.setdefault(name, []).append(float(r[-1].replace(",", "")) / 1e3)
tot = sum(sum(v) for v in agg.values())
bench = json.load(open(os.path.join(G, "bench_cfg3.json")))
with open(os.path.join(P, f"{tag}_launches_summary.md"), "w") as f:
    f.write(f"# {tag} — ncu launch list of `python bench.py --steps 3 --warmup 3 --no-cpu-baseline`\n\n")
    f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 "
            "--warmup 3 --no-cpu-baseline` on one B200 (tools/collect_profiles.sh). Per-launch times are cold-cache and serialised: compare SHARES with "
            f"bench.py's `stage_ms`, not absolutes. Raw list: `profiles/{tag}_launches.csv`.\n\n| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"| `{k}` | {len(v)} | {sum(v):.1f} | {100 * sum(v) / tot:.1f}% | {sum(v) / len(v):.1f} |\n")
    sm = bench["stage_ms"]
    st = sum(sm.values())
    f.write("\nbench.py `stage_ms` of the same build (CUDA events, warm): " + ", ".join(f"{k} {v:.3f} ms ({100 * v / st:.1f}%)" for k, v in sm.items()) + "\n\n")
    grp = {"naive": ["naive_brick"], "remove_isolated": ["ccl_", "zero_kernel"], "erode": ["stencil_"], "histogram_undo_mask": ["histogram", "pointwise"]}
    f.write("Launch-list shares grouped the same way: " + ", ".join(
        f"{g} {100 * sum(sum(v) for k, v in agg.items() if any(p in k for p in pats)) / tot:.1f}%" for g, pats in grp.items()) + "\n")

# ---- ncu --set full
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]
STALL = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALLS = ["short_scoreboard", "long_scoreboard", "barrier", "wait", "branch_resolving", "not_selected", "math_pipe_throttle", "mio_throttle", "lg_throttle", "no_instruction"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


with open(os.path.join(P, f"{tag}_ncu_full_summary.md"), "w") as f:
    f.write(f"# {tag} — ncu --set full summaries (one B200)\n\nCommands (tools/collect_profiles.sh): `ncu --set full --clock-control none --import-source on -k regex:\"naive_brick|"
            "stencil_fast|ccl_tile|ccl_border|ccl_select|histogram|pointwise\" -s 14 -c 14 python tools/prof_stage.py 512 naive,c1,erode,hist 2` (the cfg3 stages on the dense "
            "512^3 grid), `-k regex:flood_round -s 6 -c 2 python tools/prof_flood1.py 1` (cfg2 vessel) and `-k regex:\"solid_scatter|solid_expand|tetra_setup|voxelize_brick\" "
            "-s 8 -c 4 python tools/prof_solid.py 256` (V1 / V2 on the 19.8k-triangle vessel at 176x256x176). One row per distinct kernel (first captured launch); "
            "read with `ncu -i ... --page raw --csv`.\n")
    traffic = {}
    for rep in ("cfg3_full.ncu-rep", "flood_full.ncu-rep", "voxelizer_full.ncu-rep"):
        path = os.path.join(G, rep)
        if not os.path.exists(path):
            continue
        hdr, units, body = raw(path)
        seen = set()
        for r in body:
            name = r[hdr.index("Kernel Name")]
            short = name.split("(")[0].replace("void ", "").replace("<unnamed>::", "")
            if short in seen:
                continue
            seen.add(short)
            f.write(f"\n## {short}\n\n| metric | value |\n|---|---|\n")
            for w in WANT:
                if w in hdr:
                    f.write(f"| {w} | {r[hdr.index(w)]} {units[hdr.index(w)]} |\n")
            st = sorted(((float(r[hdr.index(STALL % s)]), s) for s in STALLS if STALL % s in hdr), reverse=True)[:4]
            f.write("| top stalls (warps per issue) | " + ", ".join(f"{s} {v:.2f}" for v, s in st) + " |\n")
            if "naive_brick" in short:
                rd, wr = float(r[hdr.index("dram__bytes_read.sum")]), float(r[hdr.index("dram__bytes_write.sum")])
                scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
                traffic["512"] = rd * scale[units[hdr.index("dram__bytes_read.sum")]] + wr * scale[units[hdr.index("dram__bytes_write.sum")]]
    if traffic:
        json.dump(traffic, open(os.path.join(P, "naive_traffic.json"), "w"))
        f.write(f"\nnaive_brick_kernel DRAM traffic per launch (read + write): {traffic['512'] / 1e6:.1f} MB against 536.9 MB algorithmic (profiles/naive_traffic.json feeds bench.py's `roofline.traffic`).\n")
print("profiles written with tag", tag)
