"""GPU test of the path tracer's bounce loop (ray gen -> closest hit -> shadow any-hit -> diffuse bounce -> compaction).
Per-bounce parity as SURVEY.md §8d prescribes: the rays the GPU traced are re-traced by the oracle bit for bit, and the
shading / sampling / Russian-roulette / compaction step is compared with a numpy statement of the same recipe within
float tolerance (GPU sin/cos differ from libm in the last ulp, so later bounces are checked on the GPU's own rays)."""
import numpy as np
import pytest

from atlas_engine_b200 import capi, workloads as W
from oracle.pyoracle import Scene as OScene

pytestmark = pytest.mark.gpu
f32 = np.float32


def hash1(x):
    x = x.astype(np.uint32)
    x = x + (x << np.uint32(10)); x ^= (x >> np.uint32(6)); x = x + (x << np.uint32(3)); x ^= (x >> np.uint32(11)); x = x + (x << np.uint32(15))
    return x


def random2(x, seed):
    m = hash1(x.view(np.uint32) ^ hash1(np.full_like(x, seed).view(np.uint32)))
    return ((m & np.uint32(0x007FFFFF)) | np.uint32(0x3F800000)).view(f32) - f32(1.0)


def numpy_bounce(rays_hit, payload, shadow_hit, inst, tris_by_mesh, prm):
    """numpy statement of shade_prepare / shade_finish for the rays in rays_hit (hits already traced)."""
    n = len(rays_hit)
    ids = rays_hit[:, 3].view(np.int32)
    hid = rays_hit[:, 9].view(np.int32)
    hin = rays_hit[:, 10].view(np.int32)
    o, d, t = rays_hit[:, 0:3].astype(np.float64), rays_hit[:, 4:7].astype(np.float64), rays_hit[:, 8].astype(np.float64)
    rad = payload[:, 0:3].astype(np.float64).copy()
    thr = payload[:, 4:7].astype(np.float64).copy()
    hit = hid >= 0
    P = o + t[:, None] * d
    N = np.zeros((n, 3))
    for i in np.nonzero(hit)[0]:
        I = inst[hin[i]]
        M = I[:12].view(f32).reshape(3, 4).astype(np.float64)
        T = tris_by_mesh[int(I[12])][hid[i]].astype(np.float64)
        a, b, c = T[0:3], T[4:7], T[8:11]
        nrm = np.cross(a - b, a - c)
        g = M[:, :3].T @ nrm
        g /= np.linalg.norm(g)
        if g @ d[i] > 0:
            g = -g
        N[i] = g
    L = np.asarray(prm["light_dir"], dtype=np.float64)
    ndl = N @ L
    lit = hit & (ndl > 0) & (~shadow_hit)
    direct = np.where(lit[:, None], thr * (np.asarray(prm["albedo"]) / np.pi) * np.asarray(prm["light_radiance"]) * ndl[:, None], 0.0)
    if prm["bounce"] > 0:
        mx = np.maximum(direct.max(axis=1), 10.0)
        direct *= (10.0 / mx)[:, None]
    rad = rad + np.where(hit[:, None], direct, np.minimum(np.asarray(prm["sky"]) * thr, 10.0))
    seedf = ids.astype(f32)
    s = f32(prm["seed"])
    random2(seedf, s); s += f32(1)
    random2(seedf, s); s += f32(1)
    u0 = random2(seedf, s).astype(np.float64); s += f32(1)
    u1 = random2(seedf, s).astype(np.float64); s += f32(1)
    rr = random2(seedf, s).astype(np.float64)
    r, phi = np.sqrt(u0), 2 * np.pi * u1
    lx, ly, lz = r * np.cos(phi), r * np.sin(phi), np.sqrt(1 - u0)
    up = np.where((np.abs(N[:, 2]) < 0.999)[:, None], np.array([0.0, 0.0, 1.0]), np.array([1.0, 0.0, 0.0]))
    tg = np.cross(up, N)
    tg /= np.maximum(np.linalg.norm(tg, axis=1, keepdims=True), 1e-30)
    bt = np.cross(N, tg)
    nd = tg * lx[:, None] + bt * ly[:, None] + N * lz[:, None]
    nd /= np.maximum(np.linalg.norm(nd, axis=1, keepdims=True), 1e-30)
    no = P - d * 0.1
    thr2 = thr * np.asarray(prm["albedo"])
    prob = np.clip(thr2.max(axis=1), 0.01, 0.99)
    if prm["bounce"] < 3:
        prob = np.minimum(3 * prob, 1.0)
    dead = (rr > prob) | ((nd * N).sum(axis=1) <= 0)
    thr2 = np.where(dead[:, None], 0.0, thr2 / prob[:, None])
    thr_out = np.where(hit[:, None], thr2, 0.0)
    alive = (thr_out.sum(axis=1) != 0) & (prm["bounce"] != prm["max_bounces"]) & (ids >= 0)
    return dict(alive=alive, origin=no, direction=nd, radiance=rad, throughput=thr_out, N=N, P=P, ndl=ndl, hit=hit, margin=np.abs(rr - prob))


def test_bounce_loop_against_oracle_and_numpy(ctx, oracle):
    import torch
    meshes = [W.uv_sphere(24, 12), W.heightfield(40, 40)]
    mb = [np.concatenate([W.tri_boxes(t)[:, :3].min(0), W.tri_boxes(t)[:, 3:].max(0)]) for t in meshes]
    ib, ir = W.random_instances(60, mb, seed=3, extent=(60.0, 10.0, 60.0), scale=(1.0, 4.0))
    blas = [ctx.build_blas(W.tri_boxes(t), t) for t in meshes]
    gm = [ctx.pack_mesh(b, t) for b, t in zip(blas, meshes)]
    tl = ctx.build_tlas(ib)
    scene = ctx.create_scene(gm, ir, tl)
    inst, tnodes = scene.download()
    gpu_nodes = [m.download() for m in gm]
    osc = OScene(tnodes, inst, [g[0] for g in gpu_nodes], [g[1] for g in gpu_nodes])
    tris_by_mesh = [g[1] for g in gpu_nodes]

    w, h, spp = 96, 64, 2
    eye, origin, right, bottom = W.camera_frame((30.0, 40.0, -20.0), (30.0, 0.0, 30.0), aspect=w / h)
    rays0 = ctx.generate_primary_rays(eye, origin, right, bottom, w, h, spp, jitter=np.array([[0.5, 0.5], [0.25, 0.75]], np.float32))
    n = w * h * spp
    assert np.array_equal(np.sort(rays0[:, 3].view(np.int32)), np.arange(n))
    dev = torch.device("cuda", 0)
    d_in = torch.from_numpy(rays0).to(dev)
    d_out = torch.empty_like(d_in)
    p_in = torch.zeros((n, 8), dtype=torch.float32, device=dev)
    p_out = torch.zeros_like(p_in)
    accum = torch.zeros((w * h, 4), dtype=torch.float32, device=dev)
    prm = dict(light_dir=np.array([0.3, 0.9, -0.3]) / np.linalg.norm([0.3, 0.9, -0.3]), light_radiance=[3.0, 3.0, 2.5],
               albedo=[0.7, 0.6, 0.5], sky=[0.4, 0.5, 0.8], max_bounces=3)
    count, total_finished, ref_accum = n, 0, np.zeros((w * h, 4))
    for bounce in range(4):
        traced_input = d_in[:count].cpu().numpy().copy()
        pay_in = p_in[:count].cpu().numpy().copy()
        bp = capi.BounceParams((capi._f32 * 3)(*prm["light_dir"]), (capi._f32 * 3)(*prm["light_radiance"]), (capi._f32 * 3)(*prm["albedo"]),
                               (capi._f32 * 3)(*prm["sky"]), 17.0 + bounce, bounce, prm["max_bounces"], spp)
        survivors = ctx.pathtrace_bounce(scene, bp, d_in, p_in, count, d_out, p_out, accum)
        # (1) the closest hits written in place equal the oracle's on exactly these rays
        hits = d_in[:count].cpu().numpy()
        ohits, _ = oracle.trace(osc, traced_input, nthreads=4)
        assert np.array_equal(hits.view(np.uint32), ohits.view(np.uint32)), f"bounce {bounce}: closest hits differ"
        # (2) shading / bounce against the numpy recipe, shadow visibility from the oracle's any-hit
        if bounce == 0:
            pay_in[:, 4:7] = 1.0
        step = numpy_bounce(hits, pay_in, np.zeros(count, bool), inst, tris_by_mesh, dict(prm, bounce=bounce, seed=17.0 + bounce))
        sh = np.zeros((count, 12), dtype=f32)
        need = step["hit"] & (step["ndl"] > 0)
        sh[:, 0:3] = (step["P"] + step["N"] * 0.1).astype(f32)
        sh[:, 4:7] = prm["light_dir"].astype(f32)
        sh[:, 3] = np.where(need, hits[:, 3].view(np.int32), -1).astype(np.int32).view(f32)
        sh[:, 8] = f32(1e12)
        occl, _ = oracle.trace(osc, sh, any_hit=True, per_ray_tmax=True, cull_mask=W.MASK_SHADOW, nthreads=4)
        step = numpy_bounce(hits, pay_in, occl[:, 9].view(np.int32) >= 0, inst, tris_by_mesh, dict(prm, bounce=bounce, seed=17.0 + bounce))
        out_rays = d_out[:survivors].cpu().numpy()
        out_pay = p_out[:survivors].cpu().numpy()
        ids_in = hits[:, 3].view(np.int32)
        ids_out = out_rays[:, 3].view(np.int32)
        expect_ids = set(ids_in[step["alive"]].tolist())
        got_ids = set(ids_out.tolist())
        borderline = set(ids_in[(step["margin"] < 1e-5)].tolist())
        assert (expect_ids ^ got_ids) <= borderline, f"bounce {bounce}: survivor sets differ"
        assert len(got_ids) == survivors
        if bounce < prm["max_bounces"]:
            assert survivors > 0.2 * count
        else:
            assert survivors == 0
        pos = {int(i): k for k, i in enumerate(ids_in)}
        sel = np.array([pos[int(i)] for i in ids_out if int(i) in expect_ids], dtype=np.int64)
        keep = np.array([int(i) in expect_ids for i in ids_out], dtype=bool)
        assert np.allclose(out_rays[keep, 0:3], step["origin"][sel], rtol=1e-4, atol=1e-3)
        assert np.allclose(out_rays[keep, 4:7], step["direction"][sel], rtol=0, atol=2e-3)
        assert np.allclose(out_pay[keep, 0:3], step["radiance"][sel], rtol=1e-3, atol=1e-4)
        assert np.allclose(out_pay[keep, 4:7], step["throughput"][sel], rtol=1e-3, atol=1e-4)
        fin = (~step["alive"]) & (ids_in >= 0)
        np.add.at(ref_accum, ids_in[fin] // spp, np.concatenate([step["radiance"][fin], np.ones((fin.sum(), 1))], axis=1))
        total_finished += int(fin.sum())
        count = survivors
        d_in, d_out = d_out, d_in
        p_in, p_out = p_out, p_in
        if count == 0:
            break
    acc = accum.cpu().numpy()
    assert abs(acc[:, 3].sum() - n) <= 4            # every path finished exactly once (max_bounces = 3)
    assert np.allclose(acc[:, :3].sum(axis=0), ref_accum[:, :3].sum(axis=0), rtol=2e-2)
    assert np.allclose(acc[:, :3], ref_accum[:, :3], rtol=1e-2, atol=0.5)
