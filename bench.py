#!/usr/bin/env python
"""bench.py — measures the hot path on B200 (contract: see DESIGN.md §measurement).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the CPU oracle on the host cores (bounded sample)

Workload (N=1 default) = BASELINE.json configs[2], the configuration the metric "Gvoxels/s fragmented at 512^3" is quoted on:
512^3 grid, NAIVE EUCLIDEAN nearest-seed fragmentation with 64 seeds, boundary noise (erode ELLIPSE 3, 3 iterations, p .5,
thr .5), connected-to-seed small-fragment removal and the per-fragment histogram.  The grid is the DENSE variant (every cell
occupied: worst case, algorithmic bytes 4N for the naive kernel); the sparse synthetic-vessel variant is reported alongside.
A step = one pass of that pipeline over one grid; with N GPUs every rank runs its own grid (weak scaling, no collective on
the data path).  The input grid is rewritten before every timed call, so between the rewrite and the timed region a 256 MiB
scratch buffer (larger than the 126 MB L2) is overwritten on the same stream: the timed kernels start on a grid that is not L2-resident.
The last step's result is compared with the oracle's on the same seeds and noise (`parity_checked`).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


_JSON_OUT = None


def claim_stdout():
    """The driver reads ONE JSON line from stdout.  Libraries write there too (NCCL prints its version banner with printf at communicator
    creation, whatever NCCL_DEBUG_FILE says), so fd 1 is pointed at stderr for the whole run and the JSON line goes to a saved copy of it."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit_json(line):
    out = _JSON_OUT or sys.stdout
    out.write(line + "\n")
    out.flush()


def load_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.first = index, [], None, 0

    def mark(self):
        """samples from here on are the ones reported (the timed region starts)"""
        self.first = max(0, len(self.rows) - 1)

    def start(self):
        if self.index < 0:  # ranks other than 0 do not sample: eight nvidia-smi pollers on one box disturb what they measure
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["not sampled on this rank" if self.index < 0 else "nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[self.first:]:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ workload definition
CFG3 = dict(nseeds=64, dfunc=0, erosion=(1, 3, 3, 0.5, 0.5), rng_seed=80, nnoise=1000000)


def synth_seeds_dense(n: int, nseeds: int, rng_raw):
    """Seeder::uniform restated for a fully occupied grid with location BOTH (a dense grid has no OUTER cell): rejection only
    on duplicates, three draws per attempt, std::set order, labels 2.. — identical on the CUDA and the oracle side because
    both receive this array."""
    seen = set()
    while len(seen) < nseeds:
        x, y, z = (int(np.float32(0.0) + np.float32(n - 1) * rng_raw()) for _ in range(3))
        seen.add((x, y, z))
    pts = sorted(seen)
    return np.array([[x, y, z, 2 + i] for i, (x, y, z) in enumerate(pts)], dtype=np.uint32)


def rng_uniform_stream(seed):
    """The reference float recipe on numpy's MT19937 (init_genrand seeding == std::mt19937(seed))."""
    rs = np.random.RandomState(seed)

    def draw():
        raw = int(rs.randint(0, 2**32, dtype=np.uint64))
        u = np.float32(raw) * np.float32(2.3283064365386963e-10)
        return np.float32(0.99999994) if u >= np.float32(1.0) else u

    return draw


def noise_table(seed, n):
    rs = np.random.RandomState(seed)
    raw = rs.randint(0, 2**32, size=n, dtype=np.uint64).astype(np.float32)
    u = raw * np.float32(2.3283064365386963e-10)
    u[u >= 1.0] = np.float32(0.99999994)
    return u.astype(np.float32)


# ------------------------------------------------------------------------------------------------ CUDA arm
def run_cuda(args):
    import torch

    import voxelfragmentml_b200 as vf

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    ctx = vf.Context(local_rank)
    n = args.size
    N = n**3
    dims = (n, n, n)
    peak, peak_src = load_peak()

    seeds = synth_seeds_dense(n, CFG3["nseeds"], rng_uniform_stream(CFG3["rng_seed"] + rank))
    noise = noise_table(CFG3["rng_seed"] + 1000 + rank, CFG3["nnoise"])
    pristine = torch.ones(N, dtype=torch.int16, device="cuda")  # dense variant: every cell FREE (1)
    work = torch.empty_like(pristine)
    grid = vf.RegularGrid(ctx, dims, device_ptr=work.data_ptr())
    ctx.reserve(dims)
    naive = vf.NaiveFracturer()
    naive.setDistanceFunction(CFG3["dfunc"])
    stream = torch.cuda.ExternalStream(ctx.stream)
    stages_run = []
    stage_launches = {}

    flush_buf = torch.empty(max(N, 128 << 20), dtype=torch.int16, device="cuda")  # >= 256 MiB: twice the L2

    def restore():
        """rewrite the input, then push it out of L2 (the rewrite left up to 126 MB of the grid L2-resident)"""
        with torch.cuda.stream(stream):
            work.copy_(pristine)
            flush_buf.fill_(1)

    def pipeline(record=None):
        """one step of cfg3 on the device-resident grid; returns the histogram when the stage exists"""
        t = {}

        def stage(name, fn):
            try:
                if record is not None:
                    l0 = ctx.kernel_launches
                    ctx.timer_start()
                fn()
                if record is not None:
                    t[name] = ctx.timer_stop()
                    stage_launches[name] = ctx.kernel_launches - l0
                if name not in stages_run:
                    stages_run.append(name)
            except vf.VoxFragError as e:
                if e.status != 7:  # VF_ERR_UNSUPPORTED: stage not built yet
                    raise

        # reference call order (CADScene::fractureModel, CADScene.cpp:657-688, then prepareScene :791-813):
        # NaiveFracturer::build = naive + removeIsolatedRegions, then erode, then countValues, then undoMask
        stage("naive", lambda: naive.build(grid, seeds))
        stage("remove_isolated", lambda: vf.NaiveFracturer.removeIsolatedRegions(grid, seeds))
        et, es, ei, ep, eth = CFG3["erosion"]
        stage("erode", lambda: grid.erode(et, es, ei, ep, eth, noise=noise))
        # prepareScene's grid side: countValues (toTriangleMesh) then undoMask — one fused pass (vf_histogram_undo_mask)
        stage("histogram_undo_mask", lambda: grid.countValuesUndoMask())
        if record is not None:
            record.append(t)

    # ---- warm-up, then K timed steps (CUDA events on the context's stream; the input restore is outside the events).
    # The clock sampler (nvidia-smi -lms 20 on rank 0's GPU) is started BEFORE the warm-up: the tool needs about a second to come up, and the
    # timed region of this workload lasts tens of milliseconds; it keeps sampling through the timed steps and the per-kernel timing below.
    sampler = ClockSampler(local_rank if rank == 0 else -1)
    sampler.start()
    t_warm = time.perf_counter()
    for _ in range(args.warmup):
        restore()
        pipeline()
    while rank == 0 and not sampler.rows and time.perf_counter() - t_warm < 3.0:  # first sample in: the sampler is live under load
        restore()
        pipeline()
    ctx.synchronize()
    launches0 = ctx.kernel_launches
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.mark()
    step_ms = []
    for _ in range(args.steps):
        restore()
        ctx.synchronize()
        ctx.timer_start()
        pipeline()
        step_ms.append(ctx.timer_stop())
    torch.cuda.synchronize()
    launches = ctx.kernel_launches - launches0
    total_ms = float(sum(step_ms))
    if world > 1:
        tt = torch.tensor([total_ms], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
        dist.barrier()

    # ---- per-stage breakdown (separate pass: event pairs per stage add sync points, so they are not part of the timed steps)
    breakdown = []
    for _ in range(3):
        restore()
        ctx.synchronize()
        pipeline(record=breakdown)
    stage_ms = {k: float(np.median([b[k] for b in breakdown if k in b])) for k in stages_run}

    # ---- roofline of the dominant kernel, timed alone with CUDA events on its launch stream
    naive_ms = []
    for _ in range(max(10, args.steps)):
        restore()
        ctx.synchronize()
        ctx.timer_start()
        naive.build(grid, seeds)
        naive_ms.append(ctx.timer_stop())
    clocks = sampler.stop()
    naive_t = float(np.median(naive_ms)) * 1e-3
    algo_bytes = 4.0 * N  # dense: read 2N + write 2N
    achieved = algo_bytes / naive_t / 1e9
    # per-stage achieved GB/s on algorithmic bytes (B = 2N read + 2 N_w written; dense grid => N_w = N for the stages
    # that rewrite labels, 0 for the histogram; erode = detect + 3 x erode + sweep counted as 5 passes of one read each and
    # 4 passes that rewrite every label — the reference's second and third detectBoundaries calls cannot change the grid and are not
    # launched (csrc/stencil.cu vf_erode), so their bytes are not counted as useful either)
    passes = {"naive": (1, 1), "remove_isolated": (1, 0), "erode": (5, 4), "histogram": (1, 0), "undo_mask": (1, 1), "histogram_undo_mask": (1, 1)}
    stage_roofline = {k: {"algorithmic_bytes": 2.0 * N * (passes[k][0] + passes[k][1]), "achieved_gbs": 2.0 * N * (passes[k][0] + passes[k][1]) / (v * 1e-3) / 1e9,
                          "frac": 2.0 * N * (passes[k][0] + passes[k][1]) / (v * 1e-3) / 1e9 / peak, "share_of_step": v / sum(stage_ms.values())}
                      for k, v in stage_ms.items() if k in passes}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "naive_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(str(n))
        except Exception:
            traffic = None
    # `roofline` = the dominant kernel of the timed step (largest share of the step's device time); the F1 kernel the north star's 60 % target
    # names stays beside it as `roofline_f1`, the whole step as `roofline_step`.  Algorithmic bytes per launch = the stage's algorithmic bytes
    #  / the launches the library issued for it; launch duration = the stage's CUDA-event time / the same count.
    dom = max(stage_ms, key=stage_ms.get)
    dom_launches = max(1, int(stage_launches.get(dom, 1)))
    dom_bytes = stage_roofline[dom]["algorithmic_bytes"] / dom_launches
    dom_ms = stage_ms[dom] / dom_launches
    dom_kernels = {"naive": "naive_brick_kernel<EUCLIDEAN,8>", "remove_isolated": "vfc1::certificate_kernel (+ plant / resolve list work)",
                   "erode": "stencil_fast_kernel<DETECT|ERODE3,TMA> (2 full passes) + erode_sparse_kernel x2 + sweep_sparse_kernel + sweep_apply_kernel",
                   "histogram_undo_mask": "histogram_kernel<UNMASK>"}
    # measured DRAM bytes of the dominant stage's kernels (one ncu --set full capture, profiles/kernel_traffic.json), per launch like `achieved`
    dom_traffic = None
    try:
        kt = json.load(open(os.path.join(ROOT, "profiles", "kernel_traffic.json")))
        pick = {"naive": ["naive_brick"], "remove_isolated": ["vfc1::"], "erode": ["stencil_fast", "erode_sparse", "erode_sparse", "sweep_sparse", "sweep_apply"], "histogram_undo_mask": ["histogram"]}[dom]
        tot_b = sum(v for pat in pick for k, v in kt.items() if pat in k)
        dom_traffic = tot_b / dom_launches if (n == 512 and tot_b > 0) else None
    except Exception:
        dom_traffic = None
    step_bytes = float(sum(v["algorithmic_bytes"] for v in stage_roofline.values()))

    # ---- the sparse variant BASELINE.md lists as cfg3's primary input: the ~20k-triangle synthetic vessel voxelized at 512-max (352 x 512 x 352),
    #      64 OUTER seeds from RNG seed 80, same pipeline.  Reported beside the dense headline (parity: tests/test_fullsize_gpu.py).
    vessel = None
    if not args.no_vessel and n == 512:
        from voxelfragmentml_b200 import synth

        vv, vfaces = synth.vessel_mesh(0)
        vmn, vmx = synth.mesh_aabb(vv)
        vdims = np.zeros(3, np.uint32)
        vf._capi.load().vf_dims_rule(vmn.ctypes.data, vmx.ctypes.data, 512, vdims.ctypes.data)
        vdims = tuple(int(d) for d in vdims)
        vN = int(np.prod(vdims))
        vwork = torch.empty(vN, dtype=torch.int16, device="cuda")
        vgrid = vf.RegularGrid(ctx, vdims, device_ptr=vwork.data_ptr())
        vgrid.setAABB(vmn, vmx, vdims)
        vox_ms = []
        for _ in range(4):
            ctx.synchronize()
            ctx.timer_start()
            vgrid.fill(vv, vfaces)
            vox_ms.append(ctx.timer_stop())
        ctx.initSeed(CFG3["rng_seed"])
        vseeds = vf.Seeder.uniform(vgrid, CFG3["nseeds"])
        vpristine = vwork.clone()
        vstage, vstep = [], []
        et, es, ei, ep, eth = CFG3["erosion"]
        for it in range(3 + max(5, min(args.steps, 10))):
            with torch.cuda.stream(stream):
                vwork.copy_(vpristine)
                flush_buf.fill_(1)
            ctx.synchronize()
            t = {}
            ctx.timer_start()
            naive.build(vgrid, vseeds)
            vf.NaiveFracturer.removeIsolatedRegions(vgrid, vseeds)
            vgrid.erode(et, es, ei, ep, eth, noise=noise)
            _, vocc = vgrid.countValuesUndoMask()
            if it >= 3:
                vstep.append(ctx.timer_stop())
        vms = float(np.median(vstep))
        occupied_in = int((vpristine != 0).sum().item())
        vessel = {"workload": f"cfg3-vessel: synthetic vessel ({len(vfaces)} triangles) SAT-voxelized at {vdims[0]}x{vdims[1]}x{vdims[2]}, {CFG3['nseeds']} OUTER seeds, "
                              "same pipeline as the headline", "grid": list(vdims), "occupied_voxels": occupied_in, "labelled_after_cleanup": int(vocc),
                  "ms_per_step": vms, "value": vN / (vms * 1e-3) / 1e9, "unit": "Gvoxels/s (grid cells)", "occupied_gvoxels_per_s": occupied_in / (vms * 1e-3) / 1e9,
                  "voxelize_ms": float(np.median(vox_ms)),
                  "algorithmic_bytes": 2.0 * vN * 9 + 2.0 * occupied_in * 6, "roofline_frac": (2.0 * vN * 9 + 2.0 * occupied_in * 6) / (vms * 1e-3) / 1e9 / peak}
        vgrid.close()
        del vwork, vpristine

    # ---- end to end through the public API with HOST buffers: pinned upload -> pipeline -> download, every step.
    # (a) serial: one grid, each step waits for its own download (latency of one call sequence);
    # (b) pipelined: three grids, each on its own context/stream, rotate through upload / pipeline / download so that the copies
    #     of neighbouring steps overlap the kernels of the current one (PCIe is full duplex); every step still uploads its own
    #     input and downloads its own result inside the timed region.  (b) is the throughput a batch producer gets and is `value`.
    h_in = torch.ones(N, dtype=torch.int16).pin_memory()
    h_out = torch.empty(N, dtype=torch.int16).pin_memory()
    noise_pinned = torch.from_numpy(noise).pin_memory()  # the noise table is part of every step's upload: pinned like the grid
    noise = noise_pinned.numpy()
    e2e_steps = max(2, min(args.steps, 5))
    for it in range(1 + e2e_steps):
        if it == 1:
            ctx.synchronize()
            t0 = time.perf_counter()
        grid.upload_async(h_in)
        pipeline()
        grid.download_async(h_out)
        ctx.synchronize()
    serial_s = (time.perf_counter() - t0) / e2e_steps

    # producer threads of the compact end-to-end path (the full-grid rotation uses the first three slots).  Producers with few cores per GPU
    # (the 8-GPU box: 32 hardware threads for 8 ranks) sleep on blocking events instead of spinning, as batch generation does
    # (vf_ctx_set_blocking_sync); a sleeping producer wakes up late, so twice as many of them keep the GPU fed (one GPU, tools/e2e_modes.sh:
    # spinning x4 0.99 ms per step, blocking x4 1.19, blocking x8 0.99)
    e2e_block = args.e2e_blocking == "on" or (args.e2e_blocking == "auto" and host_cores() // max(1, world) < 8)
    NPROD = max(3, args.e2e_producers if args.e2e_producers > 0 else (8 if e2e_block else 4))
    slots = []
    for k in range(NPROD):
        c = ctx if k == 0 else vf.Context(local_rank)
        g = grid if k == 0 else vf.RegularGrid(c, dims)
        c.reserve(dims)
        slots.append((c, g, h_out if k == 0 else (torch.empty(N, dtype=torch.int16).pin_memory() if k < 3 else None)))

    def run_pipelined(nsteps):
        # On this platform a transfer submitted while another one is in flight waits for it, whatever the stream or direction;
        # two transfers submitted together share the link (H2D + D2H = 97 GB/s, tools/pcie_probe*.py).  So each step submits the
        # next step's upload and the previous step's download back to back, then runs its kernels; unchanged seeds / noise are
        # not re-sent by the library, and the step's only small transfer (the histogram read-back) closes the step.
        et, es, ei, ep, eth = CFG3["erosion"]
        slots[0][1].upload_async(h_in)
        for i in range(nsteps):
            c, g, ho = slots[i % 3]
            if i + 1 < nsteps:
                slots[(i + 1) % 3][1].upload_async(h_in)               # step i+1's input
            if i > 0:
                slots[(i - 1) % 3][1].download_async(slots[(i - 1) % 3][2])  # step i-1's result
            naive.build(g, seeds)
            vf.NaiveFracturer.removeIsolatedRegions(g, seeds)
            g.erode(et, es, ei, ep, eth, noise=noise)
            g.countValuesUndoMask()
        last = (nsteps - 1) % 3
        slots[last][1].download_async(slots[last][2])
        for c, _, _ in slots:
            c.synchronize()

    pipe_steps = max(6, 2 * args.steps)
    run_pipelined(3)
    if world > 1:
        dist.barrier()
    e2e_runs = []  # three passes of pipe_steps steps each; the median pass is reported (host-side DMA scheduling varies from pass to pass)
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run_pipelined(pipe_steps)
        e2e_runs.append((time.perf_counter() - t0) / pipe_steps)
    e2e_s = sorted(e2e_runs)[1]
    # ---- the same steps with COMPACT host buffers: the input travels as one occupancy bit per cell (vf_grid_upload_bits: everything a
    # fragmentation starts from is EMPTY / FREE) and the result as the `.rle` stream the reference exports (runs found on the device,
    # vf_grid_encode_rle), so a step moves tens of megabytes instead of 2 x 2N bytes.  Three host threads, one context each, run whole steps
    # (ctypes releases the GIL inside a call): copies of one step overlap the kernels of another, as in a batch producer.
    import threading

    h_bits = torch.from_numpy(np.packbits(np.ones(N, np.uint8), bitorder="little")).pin_memory()
    rle_bytes = [0] * NPROD
    rle_out = [torch.empty(max(1 << 20, N // 4), dtype=torch.uint8).pin_memory() for _ in range(NPROD)]  # room for N / 24 runs (cfg3: ~1.6 M of 134 M cells)

    def compact_steps(k, nsteps):
        c, g, _ = slots[k]
        et, es, ei, ep, eth = CFG3["erosion"]
        for _ in range(nsteps):
            g.upload_bits(h_bits)
            naive.build(g, seeds)
            vf.NaiveFracturer.removeIsolatedRegions(g, seeds)
            g.erode(et, es, ei, ep, eth, noise=noise)
            g.countValuesUndoMask()
            rle_bytes[k] = g.encodeRLE_into(rle_out[k])  # one call: runs found, packed and copied into the producer's pinned buffer
            assert rle_bytes[k] > 0

    def run_compact(per_thread):
        ths = [threading.Thread(target=compact_steps, args=(k, per_thread)) for k in range(NPROD)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()

    for c, _, _ in slots:
        c.setBlockingSync(e2e_block)
    run_compact(1)
    if world > 1:
        dist.barrier()  # all ranks measure the same phase at the same time (the full-grid passes of a late rank would share the host's DMA with them)
    compact_runs = []
    # a pass is long enough for its ramp-up / drain (NPROD steps in flight) not to weigh on the figure: at least 10 steps per producer at full length
    per_thread = max(2, (2 * args.steps + NPROD - 1) // NPROD, min(10, args.steps))
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run_compact(per_thread)
        compact_runs.append((time.perf_counter() - t0) / (NPROD * per_thread))
    compact_s = sorted(compact_runs)[1]
    for c, _, _ in slots:
        c.setBlockingSync(False)
    # parity of the compact path: the decoded stream equals the full-grid download of the same step
    compact_ok = None
    if rank == 0:
        import oracle as orc_chk

        c0, g0, _ = slots[0]
        rle_stream = g0.encodeRLE()
        compact_ok = bool(np.array_equal(orc_chk.decode_rle(rle_stream).reshape(-1), g0.updateGrid().reshape(-1)))
    if world > 1:
        tt = torch.tensor([e2e_s, serial_s, compact_s], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s, serial_s, compact_s = (float(v) for v in tt.tolist())

    out = {
        "metric": "Gvoxels/s fragmented at 512^3", "value": world * N * args.steps / (total_ms * 1e-3) / 1e9, "unit": "Gvoxels/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u16 labels / int32 distance keys", "data": "synthetic",
        "config": cfg3_config(n, stages_run, world),
        "stage_ms": stage_ms,
        "stage_roofline": stage_roofline,
        "fragmentation_only": {"value": world * N / naive_t / 1e9, "unit": "Gvoxels/s", "note": "F1 operator alone (NaiveFracturer::build without cleanup), CUDA events"},
        "roofline": {"kernel": dom_kernels.get(dom, dom), "stage": dom, "bound": "hbm", "achieved": dom_bytes / (dom_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": dom_bytes / (dom_ms * 1e-3) / 1e9 / peak, "traffic": dom_traffic, "peak_source": peak_src, "algorithmic_bytes": dom_bytes,
                     "kernel_ms": dom_ms, "launches_per_step": dom_launches, "share_of_step": stage_ms[dom] / sum(stage_ms.values())},
        "roofline_f1": {"kernel": "naive_brick_kernel<EUCLIDEAN,8>", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes": algo_bytes,
                        "kernel_ms": naive_t * 1e3, "note": "the operator BASELINE.json's >= 60 % target names, timed alone"},
        "roofline_step": {"bound": "hbm", "algorithmic_bytes": step_bytes, "achieved": step_bytes / (total_ms / args.steps * 1e-3) / 1e9, "peak": peak,
                          "unit": "GB/s", "frac": step_bytes / (total_ms / args.steps * 1e-3) / 1e9 / peak,
                          "note": "all stages of the timed step: sum of the per-stage algorithmic bytes / ms_per_step"},
        # headline end-to-end figure: compact host buffers (what a producer that keeps occupancy as bits and consumes `.rle` moves);
        # "full_grid" is the same step with 16-bit grids both ways, as RegularGrid::updateSSBO / updateGrid move them
        "e2e": {"value": world * N / compact_s / 1e9, "unit": "Gvoxels/s", "h2d_bytes_per_step": int(h_bits.numel()),
                "d2h_bytes_per_step": int(max(rle_bytes)) + 16 + 4 * 1024, "ms_per_step": compact_s * 1e3, "steps": NPROD * per_thread,
                "mode": "compact: input = 1 occupancy bit per cell from pinned memory (vf_grid_upload_bits, expanded on the device), result = the `.rle` "
                        "byte stream (runs found on the device, vf_grid_encode_rle) + the histogram; `producer_threads` host threads with one context each run whole "
                        "steps, so the copies of one step overlap the kernels of another; seeds and noise table are sent once (unchanged tables are skipped)",
                "passes_ms_per_step": [t * 1e3 for t in compact_runs], "rle_stream_equals_grid": compact_ok, "producer_threads": NPROD,
                "host_waits": "blocking events" if e2e_block else "spinning", "host_cores_per_rank": host_cores() // max(1, world),
                "full_grid": {"value": world * N / e2e_s / 1e9, "unit": "Gvoxels/s", "h2d_bytes_per_step": 2 * N, "d2h_bytes_per_step": 2 * N + 4 * 32768,
                              "ms_per_step": e2e_s * 1e3, "steps": pipe_steps,
                              "mode": "3 grids in rotation: upload / kernels / download of consecutive steps overlap; every step copies its own 16-bit input "
                                      "grid and result grid", "passes_ms_per_step": [t * 1e3 for t in e2e_runs],
                              "serial_ms_per_step": serial_s * 1e3, "serial_value": world * N / serial_s / 1e9}},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if vessel is not None:
        out["vessel"] = vessel
    if not args.no_batch:
        # BASELINE.json's second metric (cfg4, models/s), measured beside the headline: 128 meshes per GPU, every mesh its own shape (the full
        # config is 1024 meshes over the box = 128 per GPU at N = 8)
        jobs = args.jobs or default_jobs(world)
        bm = 128 * world
        bdt, _, npool = batch_measure(rank, local_rank, world, dist if world > 1 else None, bm, 2, jobs, 0, passes=3)
        out["batch"] = {"metric": "models/s batch voxelize+fragment", "value": bm / bdt, "unit": "models/s", "meshes": bm, "jobs_per_gpu": jobs,
                        "passes_models_per_s_rank0": [bm / t for t in LAST_BATCH_INFO.get("pass_seconds", [])], "reported_pass": "median of 3",
                        "fragmentations_per_s": bm * 10 / bdt, "scaling": "weak", "host_cores": host_cores(),
                        "host_cpu_s_per_model_rank0": LAST_BATCH_INFO.get("host_cpu_s_per_model"),
                        "host_waits_per_fragmentation": LAST_BATCH_INFO.get("host_waits_per_fragmentation"),
                        "launches_per_fragmentation": LAST_BATCH_INFO.get("launches_per_fragmentation"),
                        "workload": f"cfg4-batch: {bm} synthetic vessels ({npool} distinct shapes on this rank) x 10 fragmentations at 256-max, FLOOD CHEBYSHEV, "
                                    "nf 2..10, 2*nf extra seeds, mesh m -> rank m mod N, RNG seed 80 + m"}
    if world >= 2 and not args.no_slab:
        # BASELINE config 5 on the driver's record: the slab-partitioned flood over all ranks — 2048^3 on 8 GPUs, 1024^3 on 2 or 4 (the key field of a
        # slab is 4 B per cell: 2048^3 / 8 ranks = 4.3 GiB of keys + 2.1 GiB of labels per GPU)
        torch.cuda.empty_cache()
        sn = args.slab_size or (2048 if world >= 8 else 1024)
        out["slab"] = slab_measure(rank, local_rank, world, dist, sn, 256, 3, 1, native=True)
        # the operators that shard with a one-cell halo (naive + erode) on a 1024^3 grid cut the same way (the integer-exact Euclidean kernel
        # covers extents up to 1182)
        torch.cuda.empty_cache()
        out["slab"]["labels"] = slab_labels_measure(rank, local_rank, world, dist, 1024, 64, 3, 1)
    if rank == 0 and not args.no_cpu_baseline:
        cn = args.cpu_size or n
        out["cpu_baseline"], want = cpu_baseline(cn, stages_run)
        # parity of exactly what was timed: one more step on the device from the same input, its final grid and histogram against the oracle's
        # (same seeds, same noise table; oracle = checker only).  A mismatch is reported, never hidden: the line still prints, with the count.
        if cn == n and stages_run == ["naive", "remove_isolated", "erode", "histogram_undo_mask"]:
            restore()
            naive.build(grid, seeds)
            vf.NaiveFracturer.removeIsolatedRegions(grid, seeds)
            et, es, ei, ep, eth = CFG3["erosion"]
            grid.erode(et, es, ei, ep, eth, noise=noise)
            counts, occ = grid.countValuesUndoMask()
            got = grid.updateGrid()
            bad = int((got.reshape(-1) != want["grid"].reshape(-1)).sum())
            hist_ok = bool(np.array_equal(np.asarray(counts), want["counts"]) and int(occ) == int(want["occupied"]))
            out["parity_checked"] = bad == 0 and hist_ok
            out["parity"] = {"against": "oracle (OpenMP restatement, pinned by the reference's C++ compiled in place and by the reference's compute shaders compiled as C++: tests/test_oracle_vs_ref.py, tests/test_oracle_vs_glsl.py)",
                             "cells": int(N), "mismatching_cells": bad, "histogram_equal": hist_ok, "occupied": int(occ)}
        else:
            out["parity_checked"] = False
            out["parity"] = {"skipped": f"cpu sample {cn}^3 differs from the timed {n}^3 grid or a stage is missing"}
    if rank == 0:
        emit_json(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ CPU arms (oracle)
def oracle_pipeline(orc, grid, seeds, noise, stages):
    """cfg3 with the oracle's operators, in place; returns the histogram (counts, occupied) taken before undoMask"""
    hist = (None, None)
    if "naive" in stages:
        orc.naive(grid, seeds, CFG3["dfunc"])
    if "remove_isolated" in stages:
        orc.remove_isolated_regions_cpu(grid, seeds)
    if "erode" in stages:
        et, es, ei, ep, eth = CFG3["erosion"]
        orc.erode(grid, noise, et, es, ei, ep, eth)
    if "histogram_undo_mask" in stages:
        hist = orc.count_values(grid)
        orc.undo_mask(grid, 15, False)
    if "histogram" in stages:
        hist = orc.count_values(grid)
    if "undo_mask" in stages:
        orc.undo_mask(grid, 15, False)
    return hist


def cpu_baseline(n, stages, steps=1):
    """-> (the cpu_baseline block, {"grid", "counts", "occupied"} of the oracle's result for the parity check)"""
    import oracle as orc

    orc.use_all_cores()
    seeds = synth_seeds_dense(n, CFG3["nseeds"], rng_uniform_stream(CFG3["rng_seed"]))
    noise = noise_table(CFG3["rng_seed"] + 1000, CFG3["nnoise"])
    best = None
    for _ in range(steps):
        g = np.ones((n, n, n), np.uint16)
        t0 = time.perf_counter()
        counts, occ = oracle_pipeline(orc, g, seeds, noise, stages)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    info = {"value": n**3 / best / 1e9, "unit": "Gvoxels/s", "cores": orc.num_threads(), "kind": "port",
            "sample": f"{n}^3 dense grid, same pipeline and seed count (stages {stages}), {best:.2f} s", "seconds": best}
    return info, {"grid": g, "counts": counts, "occupied": occ}


ALL_STAGES = ["naive", "remove_isolated", "erode", "histogram", "undo_mask"]


def load_ref_lib():
    """oracle/_ref/libvf_ref.so: the reference's own NaiveFracturer.cpp (buildCPU, removeIsolatedRegionsCPU) compiled in place in the
    build container (oracle/ref_shim/Makefile); it travels to the GPU box with the snapshot.  None when it was not built."""
    import ctypes as C

    path = os.path.join(ROOT, "oracle", "_ref", "libvf_ref.so")
    if not os.path.exists(path):
        return None
    try:
        L = C.CDLL(path)
    except OSError:
        return None
    u16 = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
    u32 = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
    L.ref_naive_build_cpu.argtypes = [u16, u32, u32, C.c_uint32, C.c_int]
    L.ref_remove_isolated_regions_cpu.argtypes = [u16, u32, u32, C.c_uint32]
    return L


def load_glsl_lib():
    """oracle/_ref/libvf_ref_glsl.so: the reference's own compute shaders compiled as C++ from the text under /root/reference
    (oracle/ref_shim/glsl2cpp.py + ref_glsl.cpp, OpenMP over the invocations); built in the build container, travels with the snapshot."""
    import ctypes as C

    path = os.path.join(ROOT, "oracle", "_ref", "libvf_ref_glsl.so")
    if not os.path.exists(path):
        return None
    try:
        L = C.CDLL(path)
    except OSError:
        return None
    u16 = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
    u32 = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
    f32 = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    L.glsl_naive.argtypes = [u16, u32, u32, C.c_uint32, C.c_int]
    L.glsl_erode.argtypes = [u16, u32, C.c_int, C.c_uint32, C.c_uint32, C.c_float, C.c_float, f32, C.c_uint32, C.c_int]
    L.glsl_undo_mask.argtypes = [u16, u32, C.c_uint32, C.c_int]
    return L


def cfg3_config(n, stages, world):
    """the `config` block of the cfg3 line: the CUDA arm and the reference arm describe the same workload"""
    return {"workload": f"cfg3-dense: {n}^3 all-occupied grid, NAIVE EUCLIDEAN {CFG3['nseeds']} seeds + connected-to-seed cleanup"
                        " + erode(ELLIPSE,3,3it,p.5,thr.5) + histogram + undoMask", "stages": stages, "grid": [n, n, n],
            "l2": "input rewritten, then a 256 MiB scratch buffer (2x the 126 MB L2) overwritten on the same stream before every timed call",
            "parallelism": f"replicas x{world} (one grid per GPU)"}


def reference_pipeline(ref, glsl, orc, grid, seeds, noise):
    """cfg3 on the host cores with the REFERENCE'S OWN code: naiveFracturer-comp.glsl (NaiveFracturer::buildGPU's dispatch), then
    NaiveFracturer::removeIsolatedRegionsCPU (NaiveFracturer.cpp:111-150, serial as written; oracle/_ref/libvf_ref.so), then RegularGrid::erode's
    loop over detectBoundaries / erodeGrid / copyGrid / removeIsolatedRegionsGrid-comp.glsl, then undoMask-comp.glsl.  Only countValues, serial
    host code inside RegularGrid.cpp (which needs OpenGL to compile), runs the oracle's restatement."""
    dims = np.asarray(grid.shape, np.uint32)
    s32 = np.ascontiguousarray(seeds, np.uint32)
    glsl.glsl_naive(grid, dims, s32, len(s32), CFG3["dfunc"])
    ref.ref_remove_isolated_regions_cpu(grid, dims, s32, len(s32))
    et, es, ei, ep, eth = CFG3["erosion"]
    glsl.glsl_erode(grid, dims, et, es, ei, ep, eth, noise, len(noise), 2)  # the final sweep concurrently in place, as the GPU runs it
    orc.count_values(grid)
    glsl.glsl_undo_mask(grid, dims, 15, 0)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as orc

    orc.use_all_cores()
    ref = None if args.ref_impl == "port" else load_ref_lib()
    glsl = None if args.ref_impl == "port" else load_glsl_lib()
    have_ref = ref is not None and glsl is not None
    if args.ref_impl == "ref" and not have_ref:
        emit_json(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libvf_ref.so / libvf_ref_glsl.so were not built (need /root/reference at build time)"}))
        return
    full = args.size
    # bounded sample: a step of the reference's shaders on the host cores takes ~10 s at 256^3 on 8 cores; the grid edge is chosen below so that
    # the timed steps end within about four minutes.  Throughput per voxel barely depends on the edge (slightly better when smaller).
    n = args.cpu_size or 256
    seeds = synth_seeds_dense(n, CFG3["nseeds"], rng_uniform_stream(CFG3["rng_seed"]))
    noise = np.ascontiguousarray(noise_table(CFG3["rng_seed"] + 1000, CFG3["nnoise"]), np.float32)

    def step(g):
        if have_ref:
            reference_pipeline(ref, glsl, orc, g, seeds, noise)
        else:
            oracle_pipeline(orc, g, seeds, noise, ALL_STAGES)

    # one untimed probe on a small grid sizes the sample
    if not args.cpu_size:
        pn = 128
        pseeds = synth_seeds_dense(pn, CFG3["nseeds"], rng_uniform_stream(CFG3["rng_seed"]))
        seeds_keep, seeds = seeds, pseeds
        t0 = time.perf_counter()
        step(np.ones((pn, pn, pn), np.uint16))
        per_voxel = (time.perf_counter() - t0) / pn**3
        seeds = seeds_keep
        budget = 240.0 / max(1, args.steps + 1)  # the whole arm ends within about four minutes
        n = int(min(full, max(64, (budget / per_voxel) ** (1.0 / 3.0))) // 32 * 32)
        seeds = synth_seeds_dense(n, CFG3["nseeds"], rng_uniform_stream(CFG3["rng_seed"]))
    step(np.ones((n, n, n), np.uint16))  # warm-up at the sample size
    times = []
    for _ in range(args.steps):
        g = np.ones((n, n, n), np.uint16)
        t0 = time.perf_counter()
        step(g)
        times.append(time.perf_counter() - t0)
    total = float(sum(times))
    v = n**3 * args.steps / total / 1e9
    cores = orc.num_threads()
    if have_ref:
        kind = "reference"
        what = (f"bounded sample: the same pipeline and seed count on a {n}^3 all-occupied grid per step.  Reference shaders (naiveFracturer, detectBoundaries, "
                f"erodeGrid, copyGrid, removeIsolatedRegionsGrid, undoMask -comp.glsl compiled as C++ from the reference's text) on {cores} host threads; "
                "removeIsolatedRegionsCPU is the reference's C++ (serial as written); countValues is the oracle's restatement")
    else:
        kind = "port"
        what = (f"bounded sample: the same pipeline and seed count on a {n}^3 all-occupied grid per step (oracle/_ref not built: the CPU oracle, an OpenMP "
                f"restatement, {cores} threads)")
    emit_json(json.dumps({
        "impl": "reference", "metric": "Gvoxels/s fragmented at 512^3", "value": v, "unit": "Gvoxels/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u16 labels / f32 distances", "data": "synthetic",
        "config": cfg3_config(full, ["naive", "remove_isolated", "erode", "histogram_undo_mask"], int(os.environ.get("WORLD_SIZE", "1"))),
        "sample_grid": [n, n, n], "same_size_as_config": n == full,
        "cpu_baseline": {"value": v, "unit": "Gvoxels/s", "cores": cores, "kind": kind, "sample": what},
        "e2e": {"value": v, "unit": "Gvoxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ secondary workloads (not the driver's default)
def _dist_setup():
    import torch

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    return rank, local_rank, world, dist


def slab_measure(rank, local_rank, world, dist, n, nseeds, steps, warmup, native=True, args_breakdown=False):
    """BASELINE config 5: one n^3 analytic solid, FLOOD MANHATTAN, `nseeds` seeds (15-bit ids), x-slabs over the ranks with key-plane exchange
    over NCCL.  native: the exchange loop runs inside libvoxfrag (C++: grouped ncclSend / ncclRecv + ncclAllReduce on the context's stream,
    vf_flood_slab_run); otherwise the Python loop over torch.distributed.  Labels are bit-exact against the single-address-space oracle
    (tests/test_slab_gpu.py, tests/test_fullsize_gpu.py).  Timed: key-field init + seeds, exchange loop, finalize; wall clock, max over ranks."""
    import torch

    import voxelfragmentml_b200 as vf
    from voxelfragmentml_b200 import slab, synth

    ctx = vf.Context(local_rank)
    params = synth.solid_vessel_params(0)
    seeds = synth.solid_vessel_seeds(n, nseeds, params, 80)
    x0, x1 = slab.partition(n, world)[rank]
    lib = vf._capi.load()
    comm = slab.nccl_comm(ctx, rank, world, dist) if (native and world > 1) else None

    def fill(grid):
        vf._capi.check(lib.vf_synth_solid_vessel(grid._h, x0 - 1, n, *params))

    times, iters, moved, maxd = [], 0, 0, 0
    for it in range(warmup + steps):
        s = slab.GpuSlab(ctx, fill, seeds, x0, x1, n, 1, shape=(x1 - x0 + 2, n, n), defer_init=True)  # allocation + synthetic input
        ctx.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s.start()
        if args_breakdown:
            ctx.synchronize()
        t1 = time.perf_counter()
        if native:
            iters, moved = s.run_native(comm, rank, world)
        elif world > 1:
            iters, moved = slab.run_distributed(s, rank, world, dist)
        else:
            iters, moved = slab.run_local([s])
        t2 = time.perf_counter()
        s.finalize(download=False)
        ctx.synchronize()
        dt = time.perf_counter() - t0
        phases = [(t1 - t0) * 1e3, (t2 - t1) * 1e3, (time.perf_counter() - t2) * 1e3]
        maxd = s.max_dist
        s.close()
        if it >= warmup:
            times.append(dt)
    if comm is not None:
        lib.vf_nccl_comm_destroy(comm)
    ctx.close()
    total = float(sum(times))
    if dist is not None:
        tt = torch.tensor([total, float(moved), float(maxd)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total, moved, maxd = float(tt[0]), float(tt[1]), float(tt[2])
    N = n**3
    ms = total / steps * 1e3
    peak, _ = load_peak()
    return {"metric": "Gvoxels/s flood-fragmented, one grid over N GPUs", "value": N * steps / total / 1e9, "unit": "Gvoxels/s", "n_gpus": world,
            "ms_per_step": ms, "scaling": "strong", "grid": [n, n, n], "seeds": nseeds,
            "workload": f"cfg5-slab: {n}^3 analytic solid vessel, FLOOD MANHATTAN, {nseeds} seeds (15-bit ids), {world} x-slabs, "
                        + ("C++ exchange loop over NCCL inside libvoxfrag (vf_flood_slab_run)" if native else "Python exchange loop over torch.distributed"),
            "exchange_iterations": iters, "halo_bytes_per_rank": moved, "max_geodesic_distance": maxd,
            "phases_ms_rank0_last_step": {"init_keys_and_seeds": phases[0], "exchange_loop": phases[1], "finalize": phases[2]} if args_breakdown else None,
            # F2's algorithmic bytes: read every label once, write every label once, against the aggregate HBM roofline of the N GPUs
            "roofline_frac_aggregate": 4.0 * N / (ms * 1e-3) / 1e9 / (peak * world)}


def slab_labels_measure(rank, local_rank, world, dist, n, nseeds, steps, warmup):
    """the operators that shard with a one-cell halo and no iteration — nearest-seed fragmentation (halo planes computed) and
    RegularGrid::erode (detectBoundaries, three erosion passes, the 3^3 sweep; label planes exchanged over NCCL after every pass) on one
    n^3 analytic solid cut into x-slabs.  Bit-exact against the oracle on the whole grid in tests/test_slab_gpu.py (small grids, NCCL and
    in-process exchange).  Wall clock per step, max over ranks."""
    import torch

    import voxelfragmentml_b200 as vf
    from voxelfragmentml_b200 import slab, synth

    ctx = vf.Context(local_rank)
    params = synth.solid_vessel_params(0)
    seeds = synth.solid_vessel_seeds(n, nseeds, params, 80)
    noise = noise_table(1080, 1000000)
    x0, x1 = slab.partition(n, world)[rank]
    lib = vf._capi.load()
    first = x0 - int(x0 > 0)
    planes = (x1 - x0) + int(x0 > 0) + int(x1 < n)

    def fill(grid):
        vf._capi.check(lib.vf_synth_solid_vessel(grid._h, first, n, *params))

    s = slab.LabelSlab(ctx, fill, x0, x1, n, shape=(planes, n, n))
    ex = (lambda: slab.exchange_labels(s, rank, world, dist)) if world > 1 else (lambda: None)
    et, es, ei, ep, eth = CFG3["erosion"]
    times, t_naive = [], 0.0
    for it in range(warmup + steps):
        fill(s.grid)
        ctx.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s.naive(seeds, CFG3["dfunc"])
        ctx.synchronize()
        t1 = time.perf_counter()
        slab.erode_slabs([s], et, es, ei, ep, eth, noise, 0, exchange=ex)
        ctx.synchronize()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
            t_naive = t1 - t0
    s.close()
    ctx.close()
    total = float(sum(times))
    if dist is not None:
        tt = torch.tensor([total], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total = float(tt[0])
    N = n**3
    ms = total / steps * 1e3
    peak, _ = load_peak()
    return {"metric": "Gvoxels/s nearest-seed fragmented + eroded, one grid over N GPUs", "value": N * steps / total / 1e9, "unit": "Gvoxels/s", "n_gpus": world,
            "ms_per_step": ms, "naive_ms_rank0_last_step": t_naive * 1e3, "scaling": "strong", "grid": [n, n, n], "seeds": nseeds,
            "workload": f"{n}^3 analytic solid vessel in {world} x-slabs: NAIVE EUCLIDEAN {nseeds} seeds (halo planes computed) + erode(ELLIPSE,3,3it,p.5,thr.5) = "
                        "detectBoundaries, 3 erosion passes, 3^3 sweep with a label-plane exchange over NCCL after each of the 5 passes",
            "halo_bytes_per_rank": 5 * 2 * 2 * n * n * (1 if world > 1 else 0),
            # 4 B per voxel for F1 + 18 B per voxel for the erode stage, against the aggregate HBM roofline of the N GPUs
            "roofline_frac_aggregate": 22.0 * N / (ms * 1e-3) / 1e9 / (peak * world)}


def run_slab(args):
    rank, local_rank, world, dist = _dist_setup()
    blk = slab_measure(rank, local_rank, world, dist, args.size, args.seeds, args.steps, args.warmup, native=not args.python_exchange, args_breakdown=True)
    if rank == 0:
        emit_json(json.dumps({
            "metric": blk["metric"], "value": blk["value"], "unit": blk["unit"], "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": blk["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32 keys (dist<<15|order)",
            "data": "synthetic", "config": {k: blk[k] for k in ("workload", "grid", "seeds", "exchange_iterations", "halo_bytes_per_rank", "max_geodesic_distance",
                                                                  "phases_ms_rank0_last_step")},
            "roofline_frac_aggregate": blk["roofline_frac_aggregate"]}))
    if dist is not None:
        dist.destroy_process_group()


def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def default_jobs(world):
    """concurrent contexts per GPU for the batch workload: enough to fill the SMs.  With a core per job the contexts spin (16 jobs); when the
    ranks of a node have fewer host cores than that, the contexts wait with blocking events (vf_ctx_set_blocking_sync) so that a waiting job
    does not hold a core, and more jobs hide the wake-up latency (measured on one GPU restricted to 4 cores: 8 / 16 / 24 / 32 jobs ->
    145 / 202 / 213 / 221 models/s)."""
    return 16 if host_cores() // max(1, world) >= 16 else 32


BATCH_FLOOD_MODE = 0
BATCH_FLOOD_FRONT = -1
LAST_BATCH_INFO = {}  # side information of the last batch_measure call (rank-local)


def batch_measure(rank, local_rank, world, dist, meshes, warmup, jobs, mesh_pool, blocking="auto", passes=1):
    """BASELINE config 4: dataset generation — per mesh: SAT voxelization at 256-max, then 10 fragmentations with the reference's
    dataset defaults (FLOOD + CHEBYSHEV, numSeeds = nf cycling 2..10, numExtraSeeds = 2 nf, detectBoundaries, histogram, undoMask;
    CADScene.cpp:294-332, FragmentationProcedure.h:12-13).  Mesh m goes to rank m mod N with RNG seed 80 + m, so results do not
    depend on N.  No collective on the data path.  A flood round of a 256-max shell keeps ~170 tiles busy — a fraction of one B200 —
    so every rank drives `jobs` contexts (one CUDA stream and one host thread each; ctypes releases the GIL inside a call): the
    latency-bound rounds of different meshes overlap on the SMs.  Results do not depend on `jobs` either (per-mesh RNG stream).
    `passes` > 1: the timed pass over all meshes is repeated and the MEDIAN pass is reported (a pass of 128 meshes lasts about half a second: one
    burst of foreign load on the host cores — the producers are host threads — would otherwise decide the figure; every pass is listed in
    LAST_BATCH_INFO["pass_seconds"]).
    Returns (seconds for all meshes: max over ranks, checksum of this rank, number of distinct shapes)."""
    import threading

    import torch

    import voxelfragmentml_b200 as vf
    from voxelfragmentml_b200 import synth

    lib = vf._capi.load()
    my = [m for m in range(meshes) if m % world == rank]
    # distinct shapes, generated once outside the timed region; mesh_pool <= 0: every mesh its own shape (only this rank's are generated)
    if mesh_pool <= 0:
        pool = {m: synth.vessel_mesh(m) for m in my}
    else:
        pool = {i: synth.vessel_mesh(i) for i in range(mesh_pool)}
    checksums = [0] * jobs
    workers = []
    block = blocking == "on" or (blocking == "auto" and jobs * world > host_cores())  # measured: 4 cores per rank, 16 jobs: 120 (spin) vs 164 models/s
    if blocking == "yield":
        block = 2
    for j in range(jobs):
        ctx = vf.Context(local_rank)
        ctx.setBlockingSync(block)
        if jobs > 1:
            ctx.setFloodLevels(8)  # throughput setting: with several jobs per GPU total tile work matters, not the latency of one flood
            # tiles only (0) unless --flood-front says otherwise: the thin-front solver is the latency path of ONE flood; this loop is bound by its host
            # threads and gains nothing from it (profiles/r2j_front_batch.txt), and a 16-CTA cluster per phase and job is the wrong grain when jobs share the SMs
            ctx.setFloodFront(BATCH_FLOOD_FRONT if BATCH_FLOOD_FRONT >= 0 else 0)
            ctx.setFloodMode(BATCH_FLOOD_MODE)  # CTAs per SM of a job's cooperative round loop (0: one launch per round), so that jobs share the SMs
        grid = vf.RegularGrid(ctx, (256, 256, 256))  # allocated once at the clamp size, re-dimensioned per model (CADScene.cpp:529-543)
        ctx.reserve((256, 256, 256))
        workers.append((ctx, grid))

    def one_mesh(j, m):
        ctx, grid = workers[j]
        v, f = pool[m] if mesh_pool <= 0 else pool[m % len(pool)]
        mn, mx = synth.mesh_aabb(v)
        dims = np.zeros(3, np.uint32)
        lib.vf_dims_rule(mn.ctypes.data, mx.ctypes.data, 256, dims.ctypes.data)
        grid.setAABB(mn, mx, tuple(int(d) for d in dims))
        grid.fill(v, f)
        ctx.initSeed(80 + m)
        for k in range(10):
            nf = 2 + (k % 9)
            p = vf.FractureParameters(_numSeeds=nf, _numExtraSeeds=2 * nf)
            grid.homogenize()  # resetFilling/homogenize between fragmentations (CADScene::rebuildGrid, FloodFracturer.cpp:99)
            vf.fracture_model(grid, p)
            counts, occ = grid.countValuesUndoMask()
            checksums[j] += int(occ)

    def run(todo):
        def work(j):
            for m in todo[j::jobs]:
                one_mesh(j, m)
            workers[j][0].synchronize()

        ts = [threading.Thread(target=work, args=(j,)) for j in range(jobs)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()

    run(my[: max(warmup, jobs)])
    waits0, launches0 = sum(w[0].host_waits for w in workers), sum(w[0].kernel_launches for w in workers)
    passes = max(1, int(passes))
    pass_s, pass_cpu = [], []
    for _ in range(passes):
        checksums[:] = [0] * jobs  # the checksum covers one timed pass only: it must not depend on `jobs` or N
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t0, c0 = time.perf_counter(), time.process_time()
        run(my)
        pass_s.append(time.perf_counter() - t0)
        pass_cpu.append(time.process_time() - c0)
    dt = sorted(pass_s)[(passes - 1) // 2]  # the median pass (the only one when passes == 1)
    # host CPU time of this rank (all threads) per model of the reported pass: what a batch producer pays in host cores
    LAST_BATCH_INFO["host_cpu_s_per_model"] = pass_cpu[pass_s.index(dt)] / max(1, len(my))
    LAST_BATCH_INFO["host_waits_per_fragmentation"] = (sum(w[0].host_waits for w in workers) - waits0) / max(1, 10 * len(my) * passes)
    LAST_BATCH_INFO["launches_per_fragmentation"] = (sum(w[0].kernel_launches for w in workers) - launches0) / max(1, 10 * len(my) * passes)
    LAST_BATCH_INFO["pass_seconds"] = list(pass_s)
    if dist is not None:
        tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt[0])
    for ctx, grid in workers:
        grid.close()
        ctx.close()
    return dt, int(sum(checksums)), len(pool)


def run_batch(args):
    rank, local_rank, world, dist = _dist_setup()
    jobs = args.jobs or default_jobs(world)
    dt, checksum, npool = batch_measure(rank, local_rank, world, dist, args.meshes, args.warmup, jobs, args.mesh_pool, args.blocking_sync)
    if rank == 0:
        emit_json(json.dumps({
            "metric": "models/s batch voxelize+fragment", "value": args.meshes / dt, "unit": "models/s", "n_gpus": world, "steps": args.meshes,
            "warmup": args.warmup, "ms_per_step": dt / args.meshes * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u16 labels", "data": "synthetic",
            "config": {"workload": f"cfg4-batch: {args.meshes} synthetic vessels (pool of {npool} shapes) x 10 fragmentations at 256-max, FLOOD CHEBYSHEV, "
                                   "nf 2..10, 2*nf extra seeds, per-mesh RNG seed 80+m, mesh m -> rank m mod N", "jobs_per_gpu": jobs,
                       "host_cores": host_cores(), "blocking_sync": args.blocking_sync,
                       "host_cpu_s_per_model_rank0": LAST_BATCH_INFO.get("host_cpu_s_per_model"),
                       "fragmentations_per_s": args.meshes * 10 / dt, "checksum_rank0": checksum},
        }))
    if dist is not None:
        dist.destroy_process_group()


def run_dataset(args):
    """The batch workload through the NATIVE dataset driver (csrc/dataset.cpp = CADScene::generateDataset's voxel path): per mesh
    the dims rule at clamp 256, SAT voxelization, 10 fragmentations (numFragments 2..11, one iteration each, FLOOD CHEBYSHEV, 2n extra
    seeds), histogram, undoMask, and — unlike `--workload batch` — the `.rle` export of every grid (runs found on the device, files
    written by the driver's writer threads to --out).  Mesh m -> rank m mod N, RNG seed 80 + m; --jobs contexts per GPU."""
    import shutil
    import tempfile
    import threading

    import torch

    import voxelfragmentml_b200 as vf
    from voxelfragmentml_b200 import dataset, synth

    rank, local_rank, world, dist = _dist_setup()
    jobs = args.jobs or default_jobs(world)
    pool = [synth.vessel_mesh(i) for i in range(args.mesh_pool)]
    proc = vf.FragmentationProcedure(_fragmentInterval=(2, 11), _iterationInterval=(1, 1), _maxFragmentsModel=1 << 40)
    proc._fractureParameters._clampVoxelMetricUnit = 256
    proc._fractureParameters._voxelPerMetricUnit = 256
    if args.no_export:  # the cfg4 batch loop driven natively (one C call per model): what `--workload batch` does from Python threads
        proc._exportGrid = False
    workers = []
    for j in range(jobs):
        ctx = vf.Context(local_rank)
        ctx.setBlockingSync(jobs * world > host_cores())
        if jobs > 1:
            ctx.setFloodLevels(8)
            ctx.setFloodFront(BATCH_FLOOD_FRONT if BATCH_FLOOD_FRONT >= 0 else 0)
            ctx.setFloodMode(BATCH_FLOOD_MODE)
        workers.append((ctx, dataset.dataset_grid(ctx, proc), vf._capi.VfDatasetStats()))
    out = tempfile.mkdtemp(prefix=f"vf_dataset_r{rank}_", dir=args.out or None)
    my = [m for m in range(args.meshes) if m % world == rank]

    def run(todo, timed):
        def work(j):
            ctx, grid, st = workers[j]
            for m in todo[j::jobs]:
                v, f = pool[m % len(pool)]
                ctx.initSeed(80 + m)
                dataset.generate_model(grid, proc, f"VS_{m:04d}", v, f, out + "/", st if timed else vf._capi.VfDatasetStats())
            ctx.synchronize()

        ts = [threading.Thread(target=work, args=(j,)) for j in range(jobs)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()

    run(my[: max(args.warmup, jobs)], False)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0, c0 = time.perf_counter(), time.process_time()
    run(my, True)
    dt = time.perf_counter() - t0
    host_cpu = (time.process_time() - c0) / max(1, len(my))
    shutil.rmtree(out, ignore_errors=True)
    if dist is not None:
        tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt[0])
    if rank == 0:
        tot = {k: sum(getattr(w[2], k) for w in workers) for k, _ in workers[0][2]._fields_}
        emit_json(json.dumps({
            "metric": "models/s batch voxelize+fragment+export", "value": args.meshes / dt, "unit": "models/s", "n_gpus": world, "steps": args.meshes,
            "warmup": args.warmup, "ms_per_step": dt / args.meshes * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u16 labels", "data": "synthetic",
            "config": {"workload": f"cfg4-dataset: {args.meshes} synthetic vessels (pool of {len(pool)} shapes) x 10 fragmentations at clamp 256 through the native "
                                   "driver, FLOOD CHEBYSHEV, n = 2..11 seeds + 2n extra, .rle export of every grid (device run detection, async writers)",
                       "jobs_per_gpu": jobs, "rank0": tot, "fragmentations_per_s": args.meshes * 10 / dt, "grid_export": not args.no_export,
                       "host_cores": host_cores(), "host_cpu_s_per_model_rank0": host_cpu},
        }))
    if dist is not None:
        dist.destroy_process_group()


def main():
    global BATCH_FLOOD_MODE, BATCH_FLOOD_FRONT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--cpu-size", type=int, default=0, help="edge of the CPU sample grid (default: 512 once for cpu_baseline, 384 per step for --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-impl", default="auto", choices=["auto", "ref", "port"],
                    help="--impl reference: ref = the reference's own buildCPU / removeIsolatedRegionsCPU from oracle/_ref (+ oracle for the GLSL-only stages), "
                         "port = the OpenMP oracle throughout, auto = ref when oracle/_ref was built")
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "slab", "batch", "dataset"],
                    help="cfg3 = the driver's default; slab = cfg5; batch = cfg4; dataset = cfg4 through the native driver with .rle export")
    ap.add_argument("--out", default="", help="dataset workload: parent directory of the (temporary) output folder")
    ap.add_argument("--no-export", action="store_true", help="dataset workload: no grid files (the batch loop through the native driver, metadata files only)")
    ap.add_argument("--flood-mode", type=int, default=0, help="batch workload: vf_ctx_set_flood_mode of the job contexts (0 = one launch per round, 1..4 = CTAs per SM of the cooperative loop)")
    ap.add_argument("--flood-front", type=int, default=-1, help="batch / dataset workloads: vf_ctx_set_flood_front of the job contexts when several share a GPU (-1 = 0 = tiles only)")
    ap.add_argument("--seeds", type=int, default=256, help="slab workload: number of seeds")
    ap.add_argument("--python-exchange", action="store_true", help="slab workload: the Python exchange loop over torch.distributed instead of the C++ loop over NCCL")
    ap.add_argument("--slab-size", type=int, default=0, help="default workload at N >= 2: edge of the cfg5 grid (default 2048 at 8 GPUs, else 1024)")
    ap.add_argument("--no-slab", action="store_true", help="default workload at N >= 2: skip the cfg5 slab measurement reported under \"slab\"")
    ap.add_argument("--meshes", type=int, default=64, help="batch workload: number of meshes")
    ap.add_argument("--jobs", type=int, default=0, help="batch workload: concurrent contexts (streams + host threads) per GPU; 0 = 16 with a core per job, else 32 with blocking waits")
    ap.add_argument("--blocking-sync", default="auto", choices=["auto", "on", "off", "yield"],
                    help="batch workload: contexts wait on blocking events instead of spinning (auto: when jobs x ranks exceed the host cores)")
    ap.add_argument("--e2e-producers", type=int, default=0, help="default workload: host threads (one context each) of the compact end-to-end path (0: 4 when they spin, 8 when they wait on blocking events)")
    ap.add_argument("--e2e-blocking", default="auto", choices=["auto", "on", "off"], help="default workload: the end-to-end producer threads wait on blocking events (auto: fewer than 8 host cores per rank)")
    ap.add_argument("--no-vessel", action="store_true", help="default workload: skip the sparse cfg3-vessel sub-line reported under \"vessel\"")
    ap.add_argument("--no-batch", action="store_true", help="default workload: skip the short cfg4 batch measurement reported under \"batch\"")
    ap.add_argument("--mesh-pool", type=int, default=8, help="batch workload: distinct synthetic shapes generated up front")
    args = ap.parse_args()
    BATCH_FLOOD_MODE = args.flood_mode
    BATCH_FLOOD_FRONT = args.flood_front
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "slab":
        run_slab(args)
    elif args.workload == "dataset":
        run_dataset(args)
    elif args.workload == "batch":
        run_batch(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_cuda(args)


if __name__ == "__main__":
    main()
