"""Turns what tools/collect_r2.sh left in gpurun_out/ (tag given on the command line) into the tracked summaries under profiles/:
   <tag>_bench.json, <tag>_bench_reference.json, <tag>_gpu_tests.log, <tag>_flood_timings.txt, <tag>_vessel_stage_timings.txt,
   <tag>_launches.csv + <tag>_launches_summary.md, <tag>_ncu_full_summary.md, kernel_traffic.json (DRAM bytes per launch of the dominant kernels).
Run here after the gpurun call (reads the raw-page CSV dumps the GPU box wrote with `ncu -i ... --page raw --csv`)."""
import collections, csv, json, os, shutil, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1]

for suffix in ("bench.json", "bench_reference.json", "gpu_tests.log", "flood_timings.txt", "vessel_stage_timings.txt", "voxelize_timings.txt", "launches.csv"):
    src = os.path.join(G, f"{tag}_{suffix}")
    if os.path.exists(src):
        shutil.copyfile(src, os.path.join(P, f"{tag}_{suffix}"))


def short(name):
    return name.split("(")[0].replace("void ", "").replace("<unnamed>::", "").strip()


# ---- launch list
rows = [r for r in csv.reader(open(os.path.join(G, f"{tag}_launches.csv"))) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    agg.setdefault(short(r[4]), []).append(float(r[-1].replace(",", "")) / 1e3)
tot = sum(sum(v) for v in agg.values())
bench = json.loads(open(os.path.join(G, f"{tag}_bench.json")).read().strip().splitlines()[-1])
with open(os.path.join(P, f"{tag}_launches_summary.md"), "w") as f:
    f.write(f"# {tag} — ncu launch list of `python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-batch --no-vessel`\n\n")
    f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 600 --csv --log-file gpurun_out/" + tag + "_launches.csv python bench.py "
            "--steps 3 --warmup 3 --no-cpu-baseline --no-batch --no-vessel` on one B200 (tools/collect_r2.sh). Per-launch times are cold-cache and serialised: "
            f"compare SHARES with bench.py's `stage_ms`, not absolutes. Raw list: `profiles/{tag}_launches.csv` ({len(rows)} launches). The list covers the timed "
            "steps, the per-stage pass, the F1-alone loop and the end-to-end passes (upload expansion and `.rle` encoder kernels included).\n\n"
            "| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"| `{k}` | {len(v)} | {sum(v):.1f} | {100 * sum(v) / tot:.1f}% | {sum(v) / len(v):.1f} |\n")
    sm = bench["stage_ms"]
    st = sum(sm.values())
    f.write("\nbench.py `stage_ms` of the same build (CUDA events, warm): " + ", ".join(f"{k} {v:.3f} ms ({100 * v / st:.1f}%)" for k, v in sm.items()) + "\n\n")
    grp = {"naive": ["naive_brick"], "remove_isolated": ["vfc1::", "ccl_"], "erode": ["stencil_", "erode_sparse", "sweep_"], "histogram_undo_mask": ["histogram", "pointwise"]}
    per_step = {g: sum(sum(v) / len(v) * n for k, v in agg.items() for p, n in pats.items() if p in k) for g, pats in
                {"naive": {"naive_brick": 1}, "remove_isolated": {"fill_kernel": 1, "plant_kernel": 1, "certificate": 1, "resolve": 1, "publish": 1},
                 "erode": {"stencil_fast_kernel<0": 1, "stencil_fast_kernel<1": 1, "stencil_fast_kernel<2": 0, "erode_sparse": 2, "sweep_sparse": 1, "sweep_apply": 1, "zero_kernel": 1}, "histogram_undo_mask": {"histogram": 1}}.items()}
    ps = sum(per_step.values())
    f.write("One step's kernels from the list's per-kernel averages (launch counts of one step): " + ", ".join(f"{g} {v:.0f} us ({100 * v / ps:.1f}%)" for g, v in per_step.items()) + "\n")

# ---- ncu --set full
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size"]
SCALE = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
traffic = {}
with open(os.path.join(P, f"{tag}_ncu_full_summary.md"), "w") as f:
    f.write(f"# {tag} — ncu --set full summaries (one B200)\n\nCommands (tools/collect_r2.sh): `ncu --set full --clock-control none --import-source on -k regex:\"naive_brick|"
            "certificate|resolve_kernel|stencil_fast|erode_sparse|sweep_sparse|sweep_apply|histogram\" -s 9 -c 9 python tools/prof_stage.py 512 naive,c1,erode,hist 2` (the cfg3 stages on the dense "
            "512^3 grid, second repetition) and `-k regex:flood_round -c 1 python tools/prof_flood1.py 1` (the cfg2 vessel: one cooperative launch = the whole flood phase) and `-k regex:voxelize_brick -s 4 -c 1 python tools/prof_voxelize.py 512` (V2 on the 352x512x352 vessel grid, a warm call). "
            "One section per distinct kernel (first captured launch); read with `ncu -i ... --page raw --csv`.\n")
    for dump in (f"{tag}_cfg3_full_raw.csv", f"{tag}_flood_full_raw.csv", f"{tag}_vox_full_raw.csv"):
        path = os.path.join(G, dump)
        if not os.path.exists(path):
            continue
        rws = list(csv.reader(open(path)))
        hdr, units, body = rws[0], rws[1], rws[2:]
        seen = set()
        for r in body:
            name = short(r[hdr.index("Kernel Name")])
            full = r[hdr.index("Kernel Name")]
            key = name
            if key in seen:
                continue
            seen.add(key)
            f.write(f"\n## {key}\n\n| metric | value |\n|---|---|\n")
            for w in WANT:
                if w in hdr:
                    f.write(f"| {w} | {r[hdr.index(w)]} {units[hdr.index(w)]} |\n")
            st = sorted(((float(r[i].replace(",", "")), h) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and r[i]), reverse=True)[:4]
            f.write("| top stalls (warps per issue) | " + ", ".join(f"{h.split('issue_stalled_')[1].split('_per')[0]} {v:.2f}" for v, h in st) + " |\n")
            rd = float(r[hdr.index("dram__bytes_read.sum")].replace(",", "")) * SCALE[units[hdr.index("dram__bytes_read.sum")]]
            wr = float(r[hdr.index("dram__bytes_write.sum")].replace(",", "")) * SCALE[units[hdr.index("dram__bytes_write.sum")]]
            traffic[key] = rd + wr
    f.write("\n## DRAM traffic per launch (read + write)\n\n| kernel | MB |\n|---|---|\n")
    for k, v in traffic.items():
        f.write(f"| `{k}` | {v / 1e6:.1f} |\n")
if traffic:
    json.dump(traffic, open(os.path.join(P, "kernel_traffic.json"), "w"), indent=1)
    nb = [v for k, v in traffic.items() if "naive_brick" in k]
    if nb:
        json.dump({"512": nb[0]}, open(os.path.join(P, "naive_traffic.json"), "w"))
print("profiles written with tag", tag)
