#!/usr/bin/env python
"""bench.py — BASELINE.json metric on configs[1]: synthetic 1M-triangle random soup, BLAS build + 1M random-direction
closest-hit rays.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference ...                      (the reference's own CPU code from oracle/_ref, host cores)

One JSON line on stdout (rank 0). `value` = incoherent closest-hit Mrays/s over all ranks with the scene and the rays
resident in HBM; `build` = BVH build Mtris/s on one GPU; `e2e` = the same trace through the C ABI with pinned HOST ray
buffers (H2D + trace + D2H inside the timed region). A step is one trace of the ray batch; the BVH build is timed in its
own loop of the same K steps. N > 1 is weak scaling: the BVH is replicated (every rank builds it), every rank traces its
own 1M rays, and one NCCL all-gather of the 16-byte hit records per step is inside the timed region.
"""
import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TRIS = 1_000_000
SETTLE_S = 0.4        # seconds of untimed load before a timed region (single-GPU loops)
SETTLE_STEPS = 300    # ... and the step count that replaces it where all ranks must agree
N_RAYS = 1_000_000
WORKLOAD = "C2: synthetic 1M-triangle random soup (seed 1234): BLAS build + 1M random-direction closest-hit rays (seed 5678)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.lines = []
        self.proc = None
        self.device = device

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(rank):
    from atlas_engine_b200 import workloads as W
    tris = W.soup(N_TRIS, seed=1234)
    boxes = W.tri_boxes(tris)
    lo, hi = boxes[:, :3].min(0), boxes[:, 3:].max(0)
    rays = W.random_rays(N_RAYS, lo, hi, seed=5678 + rank)
    root = np.concatenate([lo, hi])[None].astype(np.float32)
    return tris, boxes, root, rays


def reference_arm(args, rank, out):
    """The reference's own CPU implementation on the host cores: Atlas::Volume::BVH built by the unmodified
    src/engine/volume/BVH.cpp (oracle/_ref) and BVH::GetIntersection over it, rays split over all hardware threads."""
    if rank != 0:
        return
    from oracle import pyoracle
    pyoracle.build()
    tris, boxes, root, rays = make_inputs(0)
    cores = os.cpu_count() or 1
    sample = 200_000
    r8 = np.concatenate([rays[:sample, 0:3], rays[:sample, 4:7], np.zeros((sample, 1), np.float32),
                         np.full((sample, 1), 1e12, np.float32)], axis=1)
    if pyoracle.Ref.available():
        ref = pyoracle.Ref()
        kind = "reference"
        t0 = time.perf_counter()
        bvh = ref.build_blas(boxes, tris, parallel=True, keep=True)
        build_s = time.perf_counter() - t0

        def step():
            ref.intersect_closest(bvh, r8, cores)
    else:   # the reference sources are not on this box and no prebuilt _ref travelled: time the port instead
        orc = pyoracle.Oracle()
        kind = "port"
        from atlas_engine_b200 import workloads as W
        t0 = time.perf_counter()
        ob = orc.build_blas(boxes, tris)
        build_s = time.perf_counter() - t0
        ot = orc.build_tlas(root)
        sc = pyoracle.Scene(ot.gpu_nodes(), W.identity_instance(), [ob.gpu_nodes()], [W.pack_bvh_triangles(tris, ob.order, ob.end_of_node)])

        def step():
            orc.trace(sc, rays[:sample], nthreads=cores)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = sample / dt / 1e6
    line = {
        "impl": "reference", "metric": "closest_hit_incoherent", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"first {sample} of the 1M rays per step"},
        "build": {"metric": "bvh_build", "value": N_TRIS / build_s / 1e6, "unit": "Mtris/s", "ms_per_build": build_s * 1e3,
                  "note": "Atlas::Volume::BVH(aabbs, data, parallelBuild=true), constructor in to constructor out, one run"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": kind,
                         "sample": f"BVH::GetIntersection over the reference-built BLAS, first {sample} rays, {cores} threads"},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=out, flush=True)


def _claim_stdout():
    """Keep stdout for the ONE JSON line: everything else any library prints to fd 1 (NCCL's version banner, make) is
    sent to stderr. Returns a file object for the real stdout."""
    real = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    return real


def main():
    real_stdout = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true",
                    help="for runs under ncu: no settling warm-up, no lead iterations, no sampler keep-alive loop (numbers are not bench values)")
    args = ap.parse_args()
    global SETTLE_S, SETTLE_STEPS
    if args.profile:
        SETTLE_S, SETTLE_STEPS = 0.0, 0
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank, real_stdout)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as graft
    if local_rank == 0:
        graft.build()
    from atlas_engine_b200 import capi, sharding

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
    capi.lib()

    # a real (non-default) torch stream: the library launches on it and torch.cuda.Event times it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = capi.Context(local_rank, stream.cuda_stream)
    tris, boxes, root, rays = make_inputs(rank)

    # ---- resident inputs
    d_boxes = torch.from_numpy(boxes).to(dev)
    d_tris = torch.from_numpy(tris).to(dev)
    d_rays = torch.from_numpy(rays).to(dev)
    d_out = torch.empty_like(d_rays)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    h_rays = torch.from_numpy(rays).pin_memory()
    h_out = torch.empty_like(h_rays).pin_memory()
    gathered = torch.empty((world * N_RAYS, 4), dtype=torch.float32, device=dev) if world > 1 else None

    def build_once():
        blas = ctx.build_blas(d_boxes, d_tris, N_TRIS, flags=capi.ASYNC)
        return blas

    # scene used for tracing (built by the CUDA builder itself)
    blas = build_once()
    tlas = ctx.build_tlas(root)
    mesh = ctx.pack_mesh(blas, d_tris, N_TRIS)
    from atlas_engine_b200 import workloads as W
    scene = ctx.create_scene([mesh], W.identity_instance(), tlas)
    torch.cuda.synchronize()

    def trace_step():
        ctx.trace(scene, d_rays, N_RAYS, out=d_out, flags=capi.ASYNC)

    # N > 1: the all-gather of step i's hit records runs on a side stream while step i+1 is traced (double-buffered
    # outputs); all gathers are joined before the closing event, so every step's gather is inside the timed region.
    side = torch.cuda.Stream(device=dev) if world > 1 else None
    outs = [d_out, torch.empty_like(d_out)] if world > 1 else [d_out]
    gaths = [gathered, torch.empty_like(gathered)] if world > 1 else [None]

    def timed_pipelined(steps, warmup):
        def run(k, done):
            buf, dst = outs[k % 2], gaths[k % 2]
            if done[k % 2] is not None:
                stream.wait_event(done[k % 2])          # the gather that last read this buffer has finished
            ctx.trace(scene, d_rays, N_RAYS, out=buf, flags=capi.ASYNC)
            traced = torch.cuda.Event()
            traced.record(stream)
            with torch.cuda.stream(side):
                side.wait_event(traced)
                sharding.gather_hits(buf, dst)
                ev = torch.cuda.Event()
                ev.record(side)
            done[k % 2] = ev
        done = [None, None]
        # warm-up: W steps at least, and (same count on every rank, the steps contain a collective) enough of them to keep
        # the GPU under load for a few hundred ms: the first ~100 ms after an idle spell run at ramping clocks
        for k in range(max(warmup, SETTLE_STEPS)):
            run(k, done)
        stream.wait_stream(side)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for k in range(steps):
            run(k, done)
        stream.wait_stream(side)
        b.record(stream)
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps

    # the cyclic garbage collector stays off while timing: a full collection in the middle of a build (whose host code
    # waits for the device twice) shows up as device time
    gc.collect()
    gc.disable()

    def timed(fn, steps, warmup):
        # warm-up: W steps at least, continued until the GPU has been under load for SETTLE_S: after an idle spell (scene
        # set-up on the host, the clock sampler's teardown) the first ~100 ms run at ramping clocks with stalls of
        # milliseconds (seen: single 1.8 ms builds taking 9, 33 and 120 ms among the first four timed ones)
        t_settle = time.perf_counter() + SETTLE_S
        n = 0
        while n < warmup or time.perf_counter() < t_settle:
            flush.zero_()
            fn()
            torch.cuda.synchronize()
            n += 1
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        # the loop below runs LEAD + K iterations of exactly the same shape and keeps the last K: the first few iterations
        # of a loop without device-wide syncs still differ (seen: the 3rd and 5th build 2-4 ms slower, identically on both
        # GPUs of a 2-rank run - the stream-ordered allocator settling into its reuse pattern)
        LEAD = 0 if args.profile else 8
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(LEAD + steps)]
        for a, b in evs:             # torch creates the CUDA event at its first record(): do that outside the timed region
            a.record(stream)
            b.record(stream)
        torch.cuda.synchronize()
        for a, b in evs:
            flush.zero_()            # L2 flush between timed iterations (outside the event bracket)
            a.record(stream)
            fn()
            b.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        timed.samples = [a.elapsed_time(b) for a, b in evs[LEAD:]]
        total_ms = sum(timed.samples)
        if world > 1:
            t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms / steps

    # The build loop is timed FIRST, with no nvidia-smi process anywhere near it: the sampler's start-up, its queries and
    # above all its teardown stall the driver for milliseconds (a build's read-backs then wait on it: single builds of
    # 9 ms and once 120 ms were seen right after the sampler stopped, against a steady 1.8-2.3 ms).
    built = []

    def build_step():
        for b in built:
            b.free()
        built.clear()
        built.append(build_once())
    build_ms = timed(build_step, args.steps, args.warmup)
    print("build samples (ms):", " ".join(f"{x:.2f}" for x in timed.samples), file=sys.stderr)
    build_samples = sorted(timed.samples)
    for b in built:
        b.free()

    # clocks / throttle reasons are sampled with nvidia-smi while the headline (trace) region is timed
    with ClockSampler(local_rank) as clk:
        trace_ms = timed(trace_step, args.steps, args.warmup) if world == 1 else timed_pipelined(args.steps, args.warmup)
        launches0 = ctx.launches()
        trace_step()                                   # kernels of ONE step (the library counts its own launches) x K
        torch.cuda.synchronize()
        launches = (ctx.launches() - launches0) * args.steps
        if args.steps * trace_ms < 600.0 and not args.profile:   # keep the sampler alive for at least three 200 ms samples under load
            extra = int(600.0 / max(trace_ms, 1e-3)) - args.steps
            for _ in range(max(0, extra)):
                trace_step()
            torch.cuda.synchronize()
    clocks = clk.summary()

    # ---- end to end with HOST ray buffers (pinned): H2D + trace + (gather) + D2H inside the timed region.
    # N == 1: one atlas_rt_trace_closest call with host pointers (the library stages, pipelines and copies). N > 1: the
    # same call with host rays in and ATLAS_RT_DEVICE_OUTPUT, so the all-gather runs on the device-resident hits; every
    # rank then reads back its own 16 B/ray share of the gathered hit records.
    h_hits = torch.empty((N_RAYS, 4), dtype=torch.float32).pin_memory() if world > 1 else None

    def e2e_step():
        if world == 1:
            ctx.check(ctx.L.atlas_rt_trace_closest(ctx.h, scene.h, h_rays.data_ptr(), N_RAYS, capi.MASK_ALL, 0.0, capi.INF, h_out.data_ptr(), 0))
        else:
            # host rays in, hits left on the device for the gather: the library overlaps the upload with the trace
            ctx.check(ctx.L.atlas_rt_trace_closest(ctx.h, scene.h, h_rays.data_ptr(), N_RAYS, capi.MASK_ALL, 0.0, capi.INF, d_out.data_ptr(),
                                                   capi.DEVICE_OUTPUT | capi.ASYNC))
            sharding.gather_hits(d_out, gathered)
            h_hits.copy_(gathered[rank * N_RAYS:(rank + 1) * N_RAYS], non_blocking=True)   # this rank's share of the gathered hits
            stream.synchronize()
    for _ in range(args.warmup):
        e2e_step()
    torch.cuda.synchronize()
    # untimed settling as in timed(): the sampler's teardown has left the GPU idle for a few hundred ms
    if world == 1:
        t_settle = time.perf_counter() + SETTLE_S
        while time.perf_counter() < t_settle:
            e2e_step()
    else:
        for _ in range(SETTLE_STEPS // 2):   # same count on every rank: the step contains a collective
            e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())

    # ---- build end to end: pinned host boxes/triangles in, host nodes/order/flags out (Volume::BVH constructor shape)
    hb = torch.from_numpy(boxes).pin_memory()
    ht = torch.from_numpy(tris).pin_memory()
    reps = max(3, args.steps // 4)
    hn = torch.empty((2 * N_TRIS, 14), dtype=torch.int32).pin_memory().numpy().view(np.uint32)   # room for spatial-split duplicates
    ho = torch.empty(2 * N_TRIS, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    hf = torch.empty(2 * N_TRIS, dtype=torch.uint8).pin_memory().numpy()
    ctx.build_blas(hb.numpy(), ht.numpy()).free()
    t0 = time.perf_counter()
    for _ in range(reps):
        b = ctx.build_blas(hb.numpy(), ht.numpy())
        nodes, order, eon = b.download(hn, ho, hf)
        b.free()
    build_e2e_ms = (time.perf_counter() - t0) * 1e3 / reps

    # ---- algorithmic bytes per launch from the traversal's own visit counters
    ctx.trace(scene, d_rays, N_RAYS, out=d_out, flags=capi.COUNTERS)
    ct = ctx.trace_counters()
    bytes_per_launch = 96 * N_RAYS + 64 * (ct["tlas_nodes"] + ct["blas_nodes"] + ct["instances"]) + 48 * ct["triangles"]
    pk, pk_src = peaks()
    peak = float(pk["hbm_gbs"])
    achieved = bytes_per_launch / (trace_ms * 1e-3) / 1e9
    if world > 1:   # roofline is per GPU: gather time is inside trace_ms, which only makes this conservative
        pass
    traffic = None
    tp = os.path.join(ROOT, "profiles", "trace_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")

    # build roofline: SURVEY.md §8d B_build, with the leaf-depth sum taken from the tree the GPU just built
    nodes_n, refs_n = blas.counts()
    build_stats = blas.stats()

    line = {
        "metric": "closest_hit_incoherent", "value": world * N_RAYS / trace_ms / 1e3, "unit": "Mrays/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": trace_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_gpu": N_RAYS, "triangles": N_TRIS, "bvh": "replicated per GPU",
                   "l2": "flushed between timed iterations (256 MiB memset)" if world == 1 else "per-GPU working set (112 MB tree + 96 MB rays) exceeds the 126 MB L2; no flush",
                   "gather": "nccl all_gather of 16 B hit records on a side stream, overlapped with the next step's trace, all joined inside the timed region" if world > 1 else "none"},
        "clocks": clocks,
        "e2e": {"value": world * N_RAYS / e2e_ms / 1e3, "unit": "Mrays/s", "h2d_bytes_per_step": 48 * N_RAYS, "d2h_bytes_per_step": 48 * N_RAYS if world == 1 else 16 * N_RAYS,
                "ms_per_step": e2e_ms},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "trace_kernel<closest>", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": pk_src,
                     "algorithmic_bytes_per_launch": bytes_per_launch, "visits": ct},
        "build": {"metric": "bvh_build", "value": N_TRIS / build_ms / 1e3, "unit": "Mtris/s", "ms_per_build": build_ms, "ms_per_build_median": statistics.median(build_samples),
                  "ms_per_build_min_max": [build_samples[0], build_samples[-1]], "n_gpus": 1,
                  "e2e": {"value": N_TRIS / build_e2e_ms / 1e3, "unit": "Mtris/s", "ms": build_e2e_ms,
                          "h2d_bytes": 60 * N_TRIS, "d2h_bytes": 56 * nodes_n + 5 * refs_n},
                  "nodes": nodes_n, "refs": refs_n, "stats": build_stats},
    }

    if rank == 0 and not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(tris, boxes, root, rays)
        # build roofline needs the tree's leaf-depth sum, which the oracle reports
        sld = line["cpu_baseline"].get("sum_leaf_depth")
        if sld:
            b_build = 24 * N_TRIS + 32 * N_TRIS + 96 * sld + 64 * nodes_n + 48 * refs_n + 36 * N_TRIS
            ach = b_build / (build_ms * 1e-3) / 1e9
            line["build"]["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                         "algorithmic_bytes": b_build, "traffic": None}
    if rank == 0:
        print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(tris, boxes, root, rays):
    """Bounded CPU legs on the host cores: (1) the traversal restatement (oracle port, all threads) over the oracle-built
    scene on the first 500k rays; (2) the reference builder itself (oracle/_ref, parallelBuild=true) on the full mesh."""
    from oracle import pyoracle
    from atlas_engine_b200 import workloads as W
    pyoracle.build()
    cores = os.cpu_count() or 1
    orc = pyoracle.Oracle()
    t0 = time.perf_counter()
    ob = orc.build_blas(boxes, tris)
    port_build_s = time.perf_counter() - t0
    ot = orc.build_tlas(root)
    sc = pyoracle.Scene(ot.gpu_nodes(), W.identity_instance(), [ob.gpu_nodes()], [W.pack_bvh_triangles(tris, ob.order, ob.end_of_node)])
    sample = 500_000
    orc.trace(sc, rays[:20000], nthreads=cores)
    t0 = time.perf_counter()
    orc.trace(sc, rays[:sample], nthreads=cores)
    dt = time.perf_counter() - t0
    out = {"value": sample / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
           "sample": f"GLSL-order traversal restatement (oracle/atlas_oracle.cpp) over the first {sample} rays, {cores} threads",
           "sum_leaf_depth": ob.stats["sum_leaf_depth"]}
    if pyoracle.Ref.available():
        ref = pyoracle.Ref()
        dtb = ref.build_blas_timed(boxes, tris, parallel=True)
        out["build"] = {"value": N_TRIS / dtb / 1e6, "unit": "Mtris/s", "cores": cores, "kind": "reference",
                        "sample": "Atlas::Volume::BVH(aabbs, data, parallelBuild=true) on the full 1M soup, one run",
                        "ms": dtb * 1e3}
    else:
        out["build"] = {"value": N_TRIS / port_build_s / 1e6, "unit": "Mtris/s", "cores": 1, "kind": "port",
                        "sample": "oracle restatement, serial, full 1M soup", "ms": port_build_s * 1e3}
    return out


if __name__ == "__main__":
    main()
