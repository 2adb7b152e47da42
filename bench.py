#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on B200.

    python bench.py --gpus N --steps K --warmup W              (N > 1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference ...                        (the reference's own CPU code from oracle/_ref, host cores)
    python bench.py --workload c5 ...                           (BASELINE configs[4]: path tracer, STRONG scaling)
    python bench.py --metric build ...                          (the BVH-build half of the metric as its own JSON line)

Default workload = BASELINE configs[1] (C2): synthetic 1M-triangle random soup, BLAS build + 1M random-direction closest-hit
rays per GPU (weak scaling). One JSON line on stdout (rank 0):
  value   incoherent closest-hit Mrays/s over all ranks, scene and rays resident in HBM, every rank writing compact 16-byte
          hit records; N > 1: plus ONE NCCL gather of the records to rank 0 per step (atlas_rt_trace_sharded, C ABI), the
          gather of step k overlapping the trace of step k + 1, everything joined inside the timed region.
  e2e     the same trace through atlas_rt_trace_closest with pinned HOST rays in and HOST hit records out on every rank
          (H2D 48 B/ray + trace + D2H 16 B/ray inside the timed region; no collective: each rank returns its own hits).
  build   BVH build Mtris/s on one GPU (device-resident input), its end-to-end variant, and C4's 64-mesh batch.
L2 policy, identical at every N: no flush kernel; consecutive steps alternate between two replicas of the scene and two ray
buffers, so a step's working set (112 MB tree + 48 MB rays + 16 MB hits) was last touched two steps (352 MB) ago — beyond the
126 MB L2."""
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
N_RAYS = 1_000_000
SETTLE_S = 0.4        # seconds of untimed load before a timed region: the first ~100 ms after an idle spell run at ramping clocks
WORKLOAD = "C2: synthetic 1M-triangle random soup (seed 1234): BLAS build + 1M random-direction closest-hit rays per GPU (seed 5678 + rank)"
WORKLOAD_C5 = ("C5: path tracer, 3840x2160 x 16 spp, 4 bounces (+ shadow rays), one 16-spp frame per step, on the C4 scene (64 BLASes, 10k instances, "
               "1.02M triangles); blocks of 64 rayGen tiles dealt round-robin to the GPUs, rays generated on the device, one gather of the image")


PEER = os.environ.get("ATLAS_BENCH_GATHER", "nccl") == "peer"


def config_c2(world):
    return {"workload": WORKLOAD, "rays_per_gpu": N_RAYS, "triangles": N_TRIS, "bvh": "replicated per GPU",
            "l2": "no flush kernel: steps alternate between two scene replicas and two ray buffers (352 MB between reuses > 126 MB L2), same at every N",
            "output": "16-byte hit records (t, hitID, hitInstanceID, v) per ray",
            "gather": ("fused into the traversal: every rank's kernel stores its hit records straight into rank 0's memory over NVLink (CUDA IPC window, "
                       "atlas_rt_trace_sharded with ATLAS_RT_PEER_OUTPUT), completion counters polled with cuStreamWaitValue32, joined inside the timed region; N = 1: none")
            if PEER else
                      ("one NCCL gather of the hit records to rank 0 per step through atlas_rt_trace_sharded (C ABI), overlapped with the next step's trace, "
                       "joined inside the timed region; N = 1: none")}


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
    src/engine/volume/BVH.cpp (oracle/_ref) and BVH::GetIntersection over it, ALL rank-0 rays of the workload per step, split
    over all hardware threads."""
    if rank != 0:
        return
    from oracle import pyoracle
    pyoracle.build()
    tris, boxes, root, rays = make_inputs(0)
    cores = os.cpu_count() or 1
    r8 = np.concatenate([rays[:, 0:3], rays[:, 4:7], np.zeros((N_RAYS, 1), np.float32), np.full((N_RAYS, 1), 1e12, np.float32)], axis=1)
    if pyoracle.Ref.available():
        ref = pyoracle.Ref()
        kind = "reference"
        t0 = time.perf_counter()
        bvh = ref.build_blas(boxes, tris, parallel=True, keep=True)
        build_s = time.perf_counter() - t0

        def step():
            ref.intersect_closest(bvh, r8, cores)
        what = "BVH::GetIntersection over the reference-built BLAS"
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
            orc.trace(sc, rays, nthreads=cores)
        what = "GLSL-order traversal restatement (oracle port)"
    if args.metric == "build":
        reps = max(1, min(args.steps, 5))
        ts = []
        for _ in range(reps):
            ts.append(ref.build_blas_timed(boxes, tris, parallel=True) if kind == "reference" else build_s)
        dt = statistics.median(ts)
        value = N_TRIS / dt / 1e6
        line = {"impl": "reference", "metric": "bvh_build", "value": value, "unit": "Mtris/s", "n_gpus": args.gpus, "steps": reps, "warmup": 0,
                "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_c2(args.gpus),
                "cpu_baseline": {"value": value, "unit": "Mtris/s", "cores": cores, "kind": kind,
                                 "sample": "Atlas::Volume::BVH(aabbs, data, parallelBuild=true) on the full 1M soup, constructor in to constructor out"},
                "e2e": {"value": value, "unit": "Mtris/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line), file=out, flush=True)
        return
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = N_RAYS / dt / 1e6
    line = {
        "impl": "reference", "metric": "closest_hit_incoherent", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_c2(args.gpus),
        "build": {"metric": "bvh_build", "value": N_TRIS / build_s / 1e6, "unit": "Mtris/s", "ms_per_build": build_s * 1e3,
                  "note": "Atlas::Volume::BVH(aabbs, data, parallelBuild=true), constructor in to constructor out, one run"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": kind,
                         "sample": f"{what}, all {N_RAYS} rays of rank 0 per step, {cores} threads"},
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
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"])
    ap.add_argument("--metric", default="trace", choices=["trace", "build"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the out-of-L2 terrain figure and the batch-build figure")
    ap.add_argument("--profile", action="store_true",
                    help="for runs under ncu: no settling warm-up, no sampler keep-alive loop (numbers are not bench values)")
    args = ap.parse_args()
    global SETTLE_S
    if args.profile:
        SETTLE_S = 0.0
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if args.workload != "c2":
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "the reference has no CPU path tracer (its bounce loop exists only as GLSL)"}),
                      file=real_stdout, flush=True)
            return
        reference_arm(args, rank, real_stdout)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as graft
    if local_rank == 0:
        graft.build()
    from atlas_engine_b200 import capi, sharding, workloads as W

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
    comm = sharding.init_comm(ctx) if world > 1 else None
    gc.collect()
    gc.disable()   # a collection in the middle of a build (whose host code waits for the device) shows up as device time

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_loop(step, join, steps, warmup):
        """W warm-up steps (continued until the GPU has been under load for SETTLE_S; the same count on every rank: it is
        agreed on by rank 0), then EXACTLY K steps between two CUDA events on the launching stream, bracketed by a barrier +
        synchronize on both sides; returns ms per step, max over ranks."""
        n = 0
        t_end = time.perf_counter() + SETTLE_S
        while True:
            for _ in range(warmup if n == 0 else 8):
                step(n)
                n += 1
            join()
            torch.cuda.synchronize()
            more = torch.tensor([1 if time.perf_counter() < t_end else 0], device=dev)
            if world > 1:
                dist.broadcast(more, src=0)
            if not int(more.item()):
                break
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for k in range(steps):
            step(n + k)
        join()
        b.record(stream)
        barrier()
        return max_over_ranks(a.elapsed_time(b) / steps)

    if args.workload == "c5":
        run_c5(args, ctx, comm, dev, stream, rank, world, local_rank, timed_loop, barrier, max_over_ranks, real_stdout)
        if world > 1:
            comm.close()
            dist.barrier()
            dist.destroy_process_group()
        return

    tris, boxes, root, rays = make_inputs(rank)

    # ---- resident inputs: two replicas of the scene and two ray buffers (L2 policy, see the module docstring)
    d_boxes = torch.from_numpy(boxes).to(dev)
    d_tris = torch.from_numpy(tris).to(dev)
    d_rays = [torch.from_numpy(rays).to(dev), torch.from_numpy(rays).to(dev)]
    d_hits = [torch.empty((N_RAYS, 4), dtype=torch.float32, device=dev) for _ in range(2)]
    h_rays = torch.from_numpy(rays).pin_memory()
    h_hits = torch.empty((N_RAYS, 4), dtype=torch.float32).pin_memory()
    gathered = [torch.empty((world * N_RAYS, 4), dtype=torch.float32, device=dev) for _ in range(2)] if (world > 1 and rank == 0) else [None, None]

    def build_once():
        return ctx.build_blas(d_boxes, d_tris, N_TRIS, flags=capi.ASYNC)

    scenes, keep = [], []
    for _ in range(2):
        blas = build_once()
        tlas = ctx.build_tlas(root)
        mesh = ctx.pack_mesh(blas, d_tris, N_TRIS)
        scenes.append(ctx.create_scene([mesh], W.identity_instance(), tlas))
        keep.append((blas, tlas, mesh))
    blas = keep[0][0]
    torch.cuda.synchronize()

    # ---- the BVH build, timed FIRST with no nvidia-smi process anywhere near it (its start-up and teardown stall the driver for
    # milliseconds; a build's read-backs then wait on it). Per-build CUDA events, 256 MiB memset between builds to flush L2.
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    built = []

    def build_step():
        for b in built:
            b.free()
        built.clear()
        built.append(build_once())

    def timed_builds(steps, warmup):
        t_settle = time.perf_counter() + SETTLE_S
        n = 0
        while n < warmup or time.perf_counter() < t_settle:
            flush.zero_()
            build_step()
            torch.cuda.synchronize()
            n += 1
        lead = 0 if args.profile else 8
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(lead + steps)]
        for a, b in evs:
            a.record(stream)
            b.record(stream)
        torch.cuda.synchronize()
        for a, b in evs:
            flush.zero_()
            a.record(stream)
            build_step()
            b.record(stream)
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs[lead:]]
    build_samples = sorted(timed_builds(args.steps, args.warmup))
    build_ms = sum(build_samples) / len(build_samples)
    print("build samples (ms):", " ".join(f"{x:.2f}" for x in build_samples), file=sys.stderr)
    for b in built:
        b.free()
    built.clear()
    del flush

    # ---- the headline: closest-hit trace, device resident
    def trace_step(k):
        s = k & 1
        if world == 1:
            ctx.trace(scenes[s], d_rays[s], N_RAYS, out=d_hits[s], flags=capi.ASYNC | capi.HITS_ONLY)
        else:
            if PEER:   # every rank's traversal kernel stores its records straight into rank 0's window over NVLink: no gather step
                comm.trace_sharded(scenes[s], d_rays[s], world * N_RAYS, hits_out=None, flags=capi.ASYNC | capi.PEER_OUTPUT | capi.DEVICE_OUTPUT)
            else:
                comm.trace_sharded(scenes[s], d_rays[s], world * N_RAYS, hits_out=gathered[s], flags=capi.ASYNC)

    def trace_join():
        if world > 1:
            comm.synchronize()

    with ClockSampler(local_rank) as clk:
        trace_ms = timed_loop(trace_step, trace_join, args.steps, args.warmup)
        launches0 = ctx.launches()
        trace_step(0)                                   # kernels of ONE step (the library counts its own launches) x K
        trace_join()
        torch.cuda.synchronize()
        launches = (ctx.launches() - launches0) * args.steps
        if args.steps * trace_ms < 600.0 and not args.profile:   # keep the sampler alive for at least three 200 ms samples under load
            extra = int(600.0 / max(trace_ms, 1e-3)) - args.steps
            for k in range(max(0, extra)):
                trace_step(k)
            trace_join()
            torch.cuda.synchronize()
    clocks = clk.summary()

    # ---- parity of what the timed loop produced (outside the timed region): the gathered records on rank 0 must equal a
    # single-GPU trace of every rank's rays, and a scene assembled with the BLAS builds dealt across the GPUs must equal the local one
    parity = None
    if world > 1:
        ok_hits = ok_scene = True
        if PEER:   # the records the kernels of all ranks put into rank 0's window, handed on to `gathered` by the library
            if rank == 0:
                gathered[0].zero_()
            comm.trace_sharded(scenes[0], d_rays[0], world * N_RAYS, hits_out=gathered[0], flags=capi.PEER_OUTPUT)
            comm.synchronize()
        else:
            comm.trace_sharded(scenes[0], d_rays[0], world * N_RAYS, hits_out=gathered[0])
        if rank == 0:
            got = gathered[0].cpu().numpy().view(np.uint32)
            lo, hi = root[0, :3], root[0, 3:]
            for r in range(world):
                rr = W.random_rays(N_RAYS, lo, hi, seed=5678 + r)
                mine = ctx.trace(scenes[0], rr, flags=capi.HITS_ONLY)
                ok_hits &= bool(np.array_equal(mine.view(np.uint32), got[r * N_RAYS:(r + 1) * N_RAYS]))
        small = [W.soup(20000 + 3000 * k, seed=40 + k, extent=0.05) for k in range(6)]
        mb = [np.concatenate([W.tri_boxes(t)[:, :3].min(0), W.tri_boxes(t)[:, 3:].max(0)]) for t in small]
        ib, ir = W.random_instances(500, mb, seed=6, extent=(20.0, 5.0, 20.0))
        sharded = comm.build_scene_sharded(small, ib, ir)
        lb = [ctx.build_blas(W.tri_boxes(t), t) for t in small]
        lm = [ctx.pack_mesh(b, t) for b, t in zip(lb, small)]
        lt = ctx.build_tlas(ib)
        ls = ctx.create_scene(lm, ir, lt)
        nn, ni = lt.counts()
        ia, na = sharded.download(ni, nn)
        ib2, nb2 = ls.download()
        pr = W.random_rays(100000, ib[:, :3].min(0), ib[:, 3:].max(0), seed=9)
        ok_scene = bool(np.array_equal(ia, ib2) and np.array_equal(na.view(np.uint32), nb2.view(np.uint32)) and
                        np.array_equal(ctx.trace(sharded, pr).view(np.uint32), ctx.trace(ls, pr).view(np.uint32)))
        flag = torch.tensor([int(ok_hits and ok_scene)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        parity = {"gathered_hits_equal_single_gpu_trace": ok_hits, "sharded_scene_build_equals_local_on_every_rank": bool(flag.item()) and ok_scene,
                  "ok": bool(flag.item())}
        for o in [ls, lt] + lm + lb + [sharded]:
            o.free()

    # ---- end to end with HOST ray buffers (pinned): H2D 48 B/ray + trace + D2H 16 B/ray inside the timed region, on every rank.
    # Two figures: the LATENCY of one synchronous call, and the THROUGHPUT of double-buffered asynchronous calls (ATLAS_RT_ASYNC |
    # ATLAS_RT_PIPELINED: the upload of step k + 1 overlaps the last chunks of step k; every step still uploads its 48 MB and
    # downloads its 16 MB inside the timed region, into alternating host result buffers) - the way a renderer feeds batches.
    h_hits2 = [h_hits, torch.empty_like(h_hits).pin_memory()]

    def e2e_sync_step(k):
        ctx.check(ctx.L.atlas_rt_trace_closest(ctx.h, scenes[k & 1].h, h_rays.data_ptr(), N_RAYS, capi.MASK_ALL, 0.0, capi.INF, h_hits.data_ptr(),
                                               capi.HITS_ONLY))

    def e2e_step(k):
        ctx.check(ctx.L.atlas_rt_trace_closest(ctx.h, scenes[k & 1].h, h_rays.data_ptr(), N_RAYS, capi.MASK_ALL, 0.0, capi.INF, h_hits2[k & 1].data_ptr(),
                                               capi.HITS_ONLY | capi.ASYNC | capi.PIPELINED))

    def e2e_join():
        ctx.trace_join()
        torch.cuda.synchronize()
    for k in range(args.warmup):
        e2e_sync_step(k)
    t_settle = time.perf_counter() + SETTLE_S
    while time.perf_counter() < t_settle:
        e2e_sync_step(0)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        e2e_sync_step(k)
    torch.cuda.synchronize()
    e2e_latency_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
    for k in range(max(args.warmup, 4)):
        e2e_step(k)
    e2e_join()
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        e2e_step(k)
    e2e_join()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
    # what the host received (both result buffers of the pipelined calls) is what the device-resident path produced
    ctx.trace(scenes[1], d_rays[1], N_RAYS, out=d_hits[1], flags=capi.HITS_ONLY)
    want = d_hits[1].cpu().numpy().view(np.uint32)
    e2e_equal = bool(np.array_equal(h_hits2[0].numpy().view(np.uint32), want) and np.array_equal(h_hits2[1].numpy().view(np.uint32), want))

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
        b.download(hn, ho, hf)
        b.free()
    build_e2e_ms = (time.perf_counter() - t0) * 1e3 / reps

    # ---- algorithmic bytes per launch from the traversal's own visit counters: 48 B ray in + 16 B hit record out
    full = torch.empty_like(d_rays[0])
    ctx.trace(scenes[0], d_rays[0], N_RAYS, out=full, flags=capi.COUNTERS)
    ct = ctx.trace_counters()
    bytes_per_launch = 64 * N_RAYS + 64 * (ct["tlas_nodes"] + ct["blas_nodes"] + ct["instances"]) + 48 * ct["triangles"]
    pk, pk_src = peaks()
    peak = float(pk["hbm_gbs"])
    achieved = bytes_per_launch / (trace_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "r2_trace_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            tj = json.load(f)
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")

    nodes_n, refs_n = blas.counts()
    line = {
        "metric": "closest_hit_incoherent", "value": world * N_RAYS / trace_ms / 1e3, "unit": "Mrays/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": trace_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_c2(world),
        "clocks": clocks,
        "e2e": {"value": world * N_RAYS / e2e_ms / 1e3, "unit": "Mrays/s", "h2d_bytes_per_step": 48 * N_RAYS, "d2h_bytes_per_step": 16 * N_RAYS,
                "ms_per_step": e2e_ms, "latency_ms_one_synchronous_call": e2e_latency_ms,
                "calls": "asynchronous, double buffered (ATLAS_RT_ASYNC | ATLAS_RT_PIPELINED, two host result buffers), joined inside the timed region",
                "per": "rank (every rank uploads its own rays and downloads its own hit records; no collective)",
                "host_records_equal_device_path": e2e_equal},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "trace_kernel<closest>", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": pk_src,
                     "algorithmic_bytes_per_launch": bytes_per_launch, "visits": ct,
                     "note": "per GPU; ms_per_step (CUDA events, includes the 3 small ray-ordering kernels and at N > 1 the overlapped gather) is used as the launch duration"},
        "build": {"metric": "bvh_build", "value": N_TRIS / build_ms / 1e3, "unit": "Mtris/s", "ms_per_build": build_ms, "ms_per_build_median": statistics.median(build_samples),
                  "ms_per_build_min_max": [build_samples[0], build_samples[-1]], "n_gpus": 1, "l2": "256 MiB memset between builds",
                  "e2e": {"value": N_TRIS / build_e2e_ms / 1e3, "unit": "Mtris/s", "ms": build_e2e_ms,
                          "h2d_bytes": 60 * N_TRIS, "d2h_bytes": 56 * nodes_n + 5 * refs_n},
                  "nodes": nodes_n, "refs": refs_n, "stats": blas.stats()},
    }
    if parity is not None:
        line["parity"] = parity

    if rank == 0 and world == 1 and not args.no_extras:
        line["out_of_l2"] = out_of_l2_figure(ctx, dev, stream, peak)
        line["build"]["batch"] = batch_build_figure(ctx, dev, stream)
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(tris, boxes, root, rays)
        sld = line["cpu_baseline"].get("sum_leaf_depth")
        if sld:   # build roofline needs the tree's leaf-depth sum, which the oracle reports
            b_build = 24 * N_TRIS + 32 * N_TRIS + 96 * sld + 64 * nodes_n + 48 * refs_n + 36 * N_TRIS
            ach = b_build / (build_ms * 1e-3) / 1e9
            line["build"]["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                         "algorithmic_bytes": b_build, "traffic": None}
    if args.metric == "build":   # the build half of BASELINE's metric as the headline of its own line (same run, same numbers)
        bl = line["build"]
        line = {"metric": "bvh_build", "value": bl["value"], "unit": "Mtris/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": bl["ms_per_build"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_c2(world), "clocks": clocks,
                "e2e": {"value": bl["e2e"]["value"], "unit": "Mtris/s", "h2d_bytes_per_step": bl["e2e"]["h2d_bytes"], "d2h_bytes_per_step": bl["e2e"]["d2h_bytes"]},
                "gpu_launches": None, "roofline": bl.get("roofline"), "cpu_baseline": (line.get("cpu_baseline") or {}).get("build"), "build": bl}
    if rank == 0:
        print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        comm.close()
        dist.barrier()
        dist.destroy_process_group()


def out_of_l2_figure(ctx, dev, stream, peak):
    """The regime in which 'HBM roofline' means HBM: incoherent rays through the 8M-triangle terrain (512 MB of nodes + 384 MB
    of triangles, 7x the L2). Closest hit, 4M random rays, device resident, L2 flushed between iterations."""
    import torch
    from atlas_engine_b200 import capi, workloads as W
    tris = W.heightfield(2000, 2000)
    boxes = W.tri_boxes(tris)
    d_t, d_b = torch.from_numpy(tris).to(dev), torch.from_numpy(boxes).to(dev)
    blas = ctx.build_blas(d_b, d_t, len(tris))
    lo, hi = boxes[:, :3].min(0), boxes[:, 3:].max(0)
    root = np.concatenate([lo, hi])[None].astype(np.float32)
    tlas = ctx.build_tlas(root)
    mesh = ctx.pack_mesh(blas, d_t, len(tris))
    scene = ctx.create_scene([mesh], W.identity_instance(), tlas)
    n = 4_000_000
    lo2, hi2 = lo.copy(), hi.copy()
    hi2[1] += 40.0                                      # origins in a slab above and inside the relief
    rays = W.random_rays(n, lo2, hi2, seed=77)
    d_r = torch.from_numpy(rays).to(dev)
    d_h = torch.empty((n, 4), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for k in range(8):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        ctx.trace(scene, d_r, n, out=d_h, flags=capi.ASYNC | capi.HITS_ONLY)
        b.record(stream)
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = statistics.median(ts[3:])
    full = torch.empty_like(d_r)
    ctx.trace(scene, d_r, n, out=full, flags=capi.COUNTERS)
    ct = ctx.trace_counters()
    nbytes = 64 * n + 64 * (ct["tlas_nodes"] + ct["blas_nodes"] + ct["instances"]) + 48 * ct["triangles"]
    hit_rate = float((d_h[:, 1].view(torch.int32) >= 0).float().mean().item())
    out = {"workload": "8M-triangle terrain (896 MB tree), 4M random-origin random-direction closest-hit rays, device resident, L2 flushed",
           "value": n / ms / 1e3, "unit": "Mrays/s", "ms": ms, "hit_rate": hit_rate,
           "roofline": {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": nbytes / (ms * 1e-3) / 1e9 / peak,
                        "algorithmic_bytes_per_launch": nbytes, "traffic": None}}
    tp = os.path.join(ROOT, "profiles", "r2_trace_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            tj = json.load(f)
        out["roofline"]["traffic"] = tj.get("out_of_l2_dram_bytes_per_launch")
        out["roofline"]["traffic_source"] = tj.get("out_of_l2_source")
    for o in (scene, mesh, tlas, blas):
        o.free()
    return out


def batch_build_figure(ctx, dev, stream):
    """C4's 64 BLASes (1k-100k triangles) as ONE atlas_rt_build_blas_batch call, device-resident inputs, CUDA events."""
    import torch
    from atlas_engine_b200 import capi, workloads as W
    rng = np.random.default_rng(64)
    meshes = []
    for k in range(64):
        n = int(np.exp(rng.uniform(np.log(1000), np.log(100000))))
        if k % 2 == 0:
            seg = max(8, int(np.sqrt(n / 2)))
            meshes.append(W.uv_sphere(seg, max(4, seg // 2), radius=1.0 + 0.1 * k))
        else:
            side = max(4, int(np.sqrt(n / 2)))
            meshes.append(W.heightfield(side, side, spacing=20.0 / side) * np.float32(0.2))
    d_t = [torch.from_numpy(t).to(dev) for t in meshes]
    d_b = [torch.from_numpy(W.tri_boxes(t)).to(dev) for t in meshes]
    counts = [len(t) for t in meshes]
    ts = []
    for _ in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        out = ctx.build_blas_batch(d_b, d_t, counts, flags=capi.ASYNC)
        b.record(stream)
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
        for o in out:
            o.free()
    ms = statistics.median(ts[3:])
    return {"workload": "C4's 64 BLASes (1k-100k triangles each) in one atlas_rt_build_blas_batch call", "meshes": 64, "triangles": int(sum(counts)),
            "ms": ms, "value": sum(counts) / ms / 1e3, "unit": "Mtris/s"}


def run_c5(args, ctx, comm, dev, stream, rank, world, local_rank, timed_loop, barrier, max_over_ranks, real_stdout):
    """BASELINE configs[4], STRONG scaling: the frame's rayGen slots are cut into blocks of 64 tiles that are dealt round-robin
    to the ranks (a contiguous split gives one GPU the sky and another all the bounces: measured 1.34x on 2 GPUs), every rank
    runs the whole bounce loop for its blocks on the device (atlas_rt_pathtrace_bounces_interleaved, no host rays) into a
    compact buffer, and one NCCL gather brings the buffers to rank 0 (atlas_rt_image_from_shards puts the pixels in place).
    A step = one frame of the configuration: 16 sample passes (16 spp) of the full image, issued as ONE call (the passes run on
    two lanes side by side inside the library, see pathtrace.cu) + that gather."""
    import torch
    import torch.distributed as dist
    from atlas_engine_b200 import capi, sharding, workloads as W
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_configs import c4_scene
    meshes, ib, ir = c4_scene()
    blas = ctx.build_blas_batch([W.tri_boxes(t) for t in meshes], meshes)
    gm = []
    for b, t in zip(blas, meshes):
        m = ctx.pack_mesh(b, t)
        m.pack_shading(t, payload11=ctx.pack_shading_words(t, W.smooth_normals(t)))
        gm.append(m)
    tlas = ctx.build_tlas(ib)
    scene = ctx.create_scene(gm, ir, tlas)
    scene.set_materials(capi.make_materials(1))
    w, h, bounces = 3840, 2160, 4
    spp = int(os.environ.get("ATLAS_BENCH_C5_SPP", "16"))
    cam = W.camera_frame((1000.0, 260.0, -300.0), (1000.0, 60.0, 1000.0), aspect=w / h)
    ld = np.array([0.3, 0.9, -0.3]) / np.linalg.norm([0.3, 0.9, -0.3])
    prm = capi.pt_params(ld, (3.0, 3.0, 2.5), (0.4, 0.5, 0.8), max_bounces=bounces)
    # interleaved shards: blocks of 64 tiles (4096 pixels) dealt round-robin, so every GPU gets the same mix of sky and geometry
    block = 4096
    sizes_px = [ctx.pathtrace_bounces_interleaved(scene, cam, w, h, prm, 1, 0, np.zeros(bounces + 1, np.float32), r, world, block)[0] for r in range(world)]
    part = torch.zeros((sizes_px[rank], 4), dtype=torch.float32, device=dev)
    gathered = torch.zeros((w * h, 4), dtype=torch.float32, device=dev) if rank == 0 else None
    image = torch.zeros((w * h, 4), dtype=torch.float32, device=dev) if rank == 0 else None
    sizes = [n * 16 for n in sizes_px]
    offsets = [sum(sizes[:r]) for r in range(world)]
    flags = capi.ASYNC | (capi.RAY_BINNING if os.environ.get("ATLAS_BENCH_BINNING") else 0)

    def seeds_of(k):
        return (np.arange(spp * (bounces + 1), dtype=np.float32) + np.float32(k * spp * (bounces + 1))) * np.float32(0.754878) + np.float32(0.5)

    def step(k):
        ctx.pathtrace_bounces_interleaved(scene, cam, w, h, prm, spp, k * spp, seeds_of(k), rank, world, block, accum_local=part, flags=flags, count_rays=False)
        if world > 1:
            comm.gather(part, sizes[rank], gathered, sizes, offsets, flags=capi.ASYNC)

    def join():
        if world > 1:
            comm.synchronize()

    with ClockSampler(local_rank) as clk:
        ms = timed_loop(step, join, args.steps, args.warmup)
        launches0 = ctx.launches()
        step(0)
        join()
        torch.cuda.synchronize()
        launches = (ctx.launches() - launches0) * args.steps
    clocks = clk.summary()
    # rays per step: count them once (same seeds as step 0) outside the timed region
    part.zero_()
    seeds0 = seeds_of(0)
    _, traced = ctx.pathtrace_bounces_interleaved(scene, cam, w, h, prm, spp, 0, seeds0, rank, world, block, accum_local=part)
    total = torch.tensor([traced], dtype=torch.float64, device=dev)
    parity = None
    if world > 1:
        dist.all_reduce(total)
    if world > 1:
        comm.gather(part, sizes[rank], gathered, sizes, offsets)
    else:
        gathered = part
    if rank == 0:   # the gathered, re-assembled image of the sharded frame == the frame rendered whole on one GPU
        ctx.image_from_shards(gathered, w, h, world, block, image)
        whole = torch.zeros((w * h, 4), dtype=torch.float32, device=dev)
        ctx.pathtrace_bounces(scene, cam, w, h, prm, spp, 0, seeds0, whole)
        a, b = image.cpu().numpy(), whole.cpu().numpy()
        ok = bool(np.array_equal(a[:, 3], b[:, 3]) and np.allclose(a[:, :3], b[:, :3], rtol=1e-5, atol=1e-6))
        parity = {"assembled_image_equals_single_gpu_frame": ok, "ok": ok}
    rays_per_step = float(total.item())
    line = {"metric": "pathtrace_closest_hit", "value": rays_per_step / ms / 1e3, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_C5, "width": w, "height": h, "bounces": bounces, "spp_per_step": spp, "closest_hit_rays_per_step": rays_per_step,
                       "lanes": int(os.environ.get("ATLAS_RT_PT_LANES", "4")),
                       "note": "Mrays/s counts closest-hit rays only; every lit hit also casts one shadow (any-hit, opacity-aware) ray",
                       "l2": "per-step working set (rays 8.3M x 48 B x 2 + scene 112 MB) exceeds the 126 MB L2; no flush",
                       "binning": bool(flags & capi.RAY_BINNING)},
            "clocks": clocks,
            "e2e": {"value": rays_per_step / ms / 1e3, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "note": "rays are generated on the device and the image stays on rank 0's device: the reference's path tracer has no host transfers either"},
            "gpu_launches": int(launches)}
    if parity is not None:
        line["parity"] = parity
    if rank == 0:
        print(json.dumps(line), file=real_stdout, flush=True)


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
