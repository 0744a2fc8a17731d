"""Reads one `ncu --page raw --csv` dump (tools/ncu_one.sh) and prints the metrics the profiles/ summaries quote; with a sass csv, the hottest instructions."""
import csv, sys
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "?")[:90])
    for k in WANT:
        if k in d: print(f"  {k:70s} {d[k]} {units[hdr.index(k)]}")
    st = sorted(((float(d[k].replace(",", "")), k) for k in hdr if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and d[k]), reverse=True)[:6]
    print("  stalls:", ", ".join(f"{k.split('issue_stalled_')[1].split('_per')[0]} {v:.2f}" for v, k in st))
if len(sys.argv) > 2:
    rows = list(csv.reader(open(sys.argv[2])))
    while rows and "Source" not in rows[0]: rows = rows[1:]
    h = rows[0]
    ci = {n: i for i, n in enumerate(h)}
    src = ci.get("Source"); ex = ci.get("Instructions Executed"); sm = ci.get("# Samples") if "# Samples" in ci else ci.get("Samples")
    body = [r for r in rows[1:] if len(r) == len(h)]
    tot_ex = sum(float(r[ex] or 0) for r in body); tot_sm = sum(float(r[sm] or 0) for r in body)
    print(f"sass: {len(body)} instr, {tot_ex/1e6:.1f} M warp-instr, {tot_sm:.0f} samples")
    top = sorted(enumerate(body), key=lambda t: -float(t[1][sm] or 0))[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]
    if len(sys.argv) > 4:  # full listing with executed counts
        for i, r in enumerate(body): print(f"  {i:5d} {float(r[sm] or 0)/tot_sm*100:5.1f}% smp {float(r[ex] or 0)/1e3:9.0f}k ex  {r[src][:110]}")
        top = []
    for i, r in sorted(top):
        print(f"  {i:5d} {float(r[sm] or 0)/tot_sm*100:5.1f}% smp {float(r[ex] or 0)/tot_ex*100:5.1f}% ex  {r[src][:110]}")
