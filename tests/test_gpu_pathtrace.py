"""GPU tests of the path tracer (SURVEY.md 8a rows rayGen / rayHit, BASELINE configs[4]) against the oracle's line-cited
restatement of pathtracer/rayGen.csh, pathtracer/rayHit.csh, brdf/*.hsh, common/random.hsh (oracle/atlas_oracle_shade.cpp).

Per-bounce parity as SURVEY.md 8d prescribes: the rays the GPU traced are re-traced by the oracle bit for bit; the hit
shader is compared ray by ray — RNG words, lobe choice, Russian roulette and therefore the SURVIVOR SET exactly (rays whose
roulette draw lies within 1e-6 of the probability are the only ones allowed to differ), float terms within FLOAT_RTOL.
GPU sin / cos / pow differ from libm in the last ulp, so later bounces start from the GPU's own rays."""
import numpy as np
import pytest

from atlas_engine_b200 import capi, workloads as W
from oracle.pyoracle import Scene as OScene

pytestmark = pytest.mark.gpu
f32 = np.float32
FLOAT_RTOL = 2e-5      # shading terms (transcendentals + a 3x3 inverse differ in the last ulps between GPU and libm)
SHADOW_TMAX = float(f32(1e12) - f32(0.2))


def pt_scene(ctx, oracle, with_texture=True, n_inst=60, seed=3):
    """Instanced scene with smooth normals, uvs, five materials (diffuse, rough metal, emissive, translucent, textured
    opacity) and per-instance material offsets; returns the GPU scene and the oracle's view of the same arrays."""
    sphere = W.uv_sphere(24, 12)
    field = W.heightfield(40, 40)
    meshes = [sphere, field]
    mb = [np.concatenate([W.tri_boxes(t)[:, :3].min(0), W.tri_boxes(t)[:, 3:].max(0)]) for t in meshes]
    ib, ir = W.random_instances(n_inst, mb, seed=seed, extent=(60.0, 10.0, 60.0), scale=(1.0, 4.0))
    ir[:, 13] = (np.arange(n_inst) % 3).astype(np.uint32)              # materialOffset 0..2
    mats = capi.make_materials(5)
    mats[1]["metalness"], mats[1]["roughness"], mats[1]["baseR"] = 0.9, 0.35, 0.95
    mats[2]["emissR"], mats[2]["emissG"], mats[2]["emissB"] = 2.0, 1.5, 0.5
    mats[3]["opacity"] = 0.6
    mats[4]["opacityTexture"] = 0 if with_texture else -1
    mats[4]["roughness"] = 0.6
    rng = np.random.default_rng(5)
    tex = [(rng.random((32, 48)) > 0.45).astype(np.uint8) * 255, (rng.random((8, 8)) * 255).astype(np.uint8)]
    gm, blas, t96s, obls = [], [], [], []
    for k, tris in enumerate(meshes):
        boxes = W.tri_boxes(tris)
        n = len(tris)
        midx = (np.arange(n) % 3).astype(np.int32)                      # materialIndex 0..2 (+ offset 0..2 -> materials 0..4)
        op = np.where(midx == 2, f32(-1.0), f32(1.0)).astype(f32) if with_texture else np.ones(n, f32)   # index 2: textured where it maps to material 4
        words = ctx.pack_shading_words(tris, W.smooth_normals(tris), W.planar_uvs(tris, 0.13), None)
        assert np.array_equal(words, oracle.pack_shading_words(tris, W.smooth_normals(tris), W.planar_uvs(tris, 0.13), None))
        b = ctx.build_blas(boxes, tris)
        m = ctx.pack_mesh(b, tris, material_idx=midx, opacity=op)
        m.pack_shading(tris, material_idx=midx, opacity=op, payload11=words)
        blas.append(b); gm.append(m); t96s.append(m.download_shading())
    tl = ctx.build_tlas(ib)
    scene = ctx.create_scene(gm, ir, tl)
    scene.set_materials(mats, tex if with_texture else ())
    inst, tnodes = scene.download()
    dl = [m.download() for m in gm]
    osc = OScene(tnodes, inst, [d[0] for d in dl], [d[1] for d in dl], t96s)
    osc.set_materials(mats, tex if with_texture else ())
    return scene, osc, ib, (blas, gm, tl)


def test_primary_rays_equal_oracle_raygen(ctx, oracle):
    """rayGen.csh: IDs, storage order (8x8 tiles + ragged borders) and the hash jitter exactly; directions bit for bit
    (normalize is a division-free product of IEEE operations in both)."""
    eye, origin, right, bottom = W.camera_frame((3.0, 2.0, 1.0), (0.0, 0.5, 0.0))
    for (w, h, spf, sc) in ((64, 40, 1, 0), (70, 37, 2, 5), (13, 9, 1, 123)):
        jit = capi.sample_jitter(sc)
        rays = ctx.generate_primary_rays(eye, origin, right, bottom, w, h, spf, jitter=np.tile(jit, (spf, 1)))
        ref = oracle.raygen(eye, origin, right, bottom, w, h, spf, sc)
        assert np.array_equal(rays[:, 3].view(np.int32), ref[:, 3].view(np.int32))
        assert np.array_equal(rays.view(np.uint32), ref.view(np.uint32))
    assert 0.0 <= capi.sample_jitter(7)[0] < 1.0 and not np.array_equal(capi.sample_jitter(7), capi.sample_jitter(8))


def test_ray_binning_matches_the_octahedral_bins(ctx, oracle):
    import torch
    rng = np.random.default_rng(9)
    n = 50_001
    d = rng.normal(size=(n, 3)).astype(f32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[:7] = [[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0], [0, -1, 0], [0, 0, -1], [0.5, 0.0, 0.5]]     # axis directions: saturating coordinates
    rays = W.pack_rays(rng.random((n, 3)).astype(f32), d)
    pay = rng.integers(0, 2**32, size=(n, 4), dtype=np.uint64).astype(np.uint32)
    bins = oracle.ray_bins(rays)
    assert bins.max() <= 72
    order = np.argsort(bins, kind="stable")
    dev = torch.device("cuda", 0)
    d_in, p_in = torch.from_numpy(rays).to(dev), torch.from_numpy(pay.view(np.int32)).to(dev)
    d_out, p_out = torch.empty_like(d_in), torch.empty_like(p_in)
    ctx.bin_rays(d_in, p_in, n, d_out, p_out)
    assert np.array_equal(d_out.cpu().numpy().view(np.uint32), rays[order].view(np.uint32))
    assert np.array_equal(p_out.cpu().numpy().view(np.uint32), pay[order])


def test_textured_opacity_in_traversal(ctx, oracle):
    """HitClosestTransparency / HitAnyTransparency with GetOpacity (surface.hsh:147-160): triangles whose opacity is < 0 are
    resolved through the material's opacity texture at the hit's interpolated texture coordinates."""
    scene, osc, ib, keep = pt_scene(ctx, oracle)
    rays = W.random_rays(150_000, ib[:, :3].min(0), ib[:, 3:].max(0), seed=21)
    out = ctx.trace(scene, rays, flags=capi.OPACITY)
    ref, _ = oracle.trace(osc, rays, opacity=True, nthreads=8)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    plain = ctx.trace(scene, rays)
    assert (plain[:, 9].view(np.int32) != out[:, 9].view(np.int32)).sum() > 20       # holes in the texture let rays through
    for mask in (W.MASK_ALL, W.MASK_SHADOW):
        out = ctx.trace(scene, rays, any_hit=True, cull_mask=mask, t_max=150.0, flags=capi.OPACITY)
        ref, _ = oracle.trace(osc, rays, any_hit=True, cull_mask=mask, t_max=150.0, opacity=True, nthreads=8)
        assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    tr = out[:, 7]
    assert ((tr > 0.0) & (tr < 1.0)).sum() > 20                                       # partial transparency (opacity 0.6, bilinear edges)
    hits = ctx.trace(scene, rays, flags=capi.OPACITY | capi.HITS_ONLY)               # compact 16-byte records
    full = ctx.trace(scene, rays, flags=capi.OPACITY)
    assert hits.shape == (len(rays), 4) and np.array_equal(hits.view(np.uint32), full[:, 8:12].view(np.uint32))


def unpack_payload(p):
    h = p.view(np.uint32).reshape(-1, 4)
    lo = lambda w: (w & 0xffff).astype(np.uint16).view(np.float16).astype(np.float64)
    hi = lambda w: (w >> 16).astype(np.uint16).view(np.float16).astype(np.float64)
    return np.stack([lo(h[:, 0]), hi(h[:, 0]), lo(h[:, 2])], 1), np.stack([lo(h[:, 1]), hi(h[:, 1]), hi(h[:, 2])], 1)


@pytest.mark.parametrize("with_texture", [True, False])
def test_bounce_loop_against_the_oracle(ctx, oracle, with_texture):
    """with_texture=False: every triangle is fully opaque and every instance carries the shadow bit, so the library runs the
    OPACITY_CHECK traces as plain 48-byte ones — the oracle still runs HitClosestTransparency / HitAnyTransparency, and the
    results must stay bit-identical."""
    import torch
    scene, osc, ib, keep = pt_scene(ctx, oracle, with_texture=with_texture)
    w, h, spf = 96, 64, 2
    cam = W.camera_frame((30.0, 40.0, -20.0), (30.0, 0.0, 30.0), aspect=w / h)
    n = w * h * spf
    jit = capi.sample_jitter(3)
    rays0 = ctx.generate_primary_rays(*cam, w, h, spf, jitter=np.tile(jit, (spf, 1)))
    dev = torch.device("cuda", 0)
    d_in = torch.from_numpy(rays0).to(dev)
    d_out = torch.empty_like(d_in)
    p_in = torch.zeros((n, 4), dtype=torch.float32, device=dev)
    p_out = torch.zeros_like(p_in)
    accum = torch.zeros((w * h, 4), dtype=torch.float32, device=dev)
    ld = np.array([0.3, 0.9, -0.3]) / np.linalg.norm([0.3, 0.9, -0.3])
    prm = capi.pt_params(ld, (3.0, 3.0, 2.5), (0.4, 0.5, 0.8), max_bounces=3, samples_per_frame=spf)
    count, ref_accum, finished_paths = n, np.zeros((w * h, 4)), 0
    for bounce in range(4):
        seed = 17.25 + 3 * bounce
        traced_input = d_in[:count].cpu().numpy().copy()
        pay_in = p_in[:count].cpu().numpy().copy()
        survivors = ctx.pathtrace_bounce(scene, prm, seed, bounce, d_in, p_in, count, d_out, p_out, accum, w, h)
        # (1) HitClosestTransparency in place: equal to the oracle's on exactly these rays
        hits = d_in[:count].cpu().numpy()
        ohits, _ = oracle.trace(osc, traced_input, opacity=True, nthreads=8)
        assert np.array_equal(hits.view(np.uint32), ohits.view(np.uint32)), f"bounce {bounce}: closest hits differ"
        # (2) the oracle's shadow rays + HitAnyTransparency, then its hit shader
        sh = oracle.pt_shadow_rays(osc, hits, prm)
        vis, _ = oracle.trace(osc, sh, any_hit=True, cull_mask=W.MASK_SHADOW, t_max=SHADOW_TMAX, opacity=True, nthreads=8)
        step = oracle.pt_shade(osc, hits, pay_in, vis[:, 7], prm, seed, bounce)
        out_rays = d_out[:survivors].cpu().numpy()
        out_pay = p_out[:survivors].cpu().numpy()
        ids_in = hits[:, 3].view(np.int32)
        ids_out = out_rays[:, 3].view(np.int32)
        assert len(set(ids_out.tolist())) == survivors
        drew = step["rr"][:, 1] > 0.0                                     # rays that reached the Russian roulette
        borderline = set(ids_in[drew & (np.abs(step["rr"][:, 0] - step["rr"][:, 1]) < 1e-6)].tolist())
        expect = set(ids_in[step["alive"]].tolist())
        assert (expect ^ set(ids_out.tolist())) <= borderline, f"bounce {bounce}: survivor sets differ"
        assert len(borderline) < 5
        if bounce < 3:
            assert survivors > 0.15 * count
        else:
            assert survivors == 0
        pos = {int(i): k for k, i in enumerate(ids_in)}
        keep_o = np.array([int(i) in expect for i in ids_out], dtype=bool)
        sel = np.array([pos[int(i)] for i in ids_out[keep_o]], dtype=np.int64)
        if len(sel):
            assert np.allclose(out_rays[keep_o, 0:3], step["rays"][sel, 0:3], rtol=FLOAT_RTOL, atol=1e-4)
            assert np.allclose(out_rays[keep_o, 4:7], step["rays"][sel, 4:7], rtol=0, atol=2e-5)
            assert np.array_equal(out_rays[keep_o, 8:11].view(np.uint32), step["rays"][sel, 8:11].view(np.uint32))
            gr, gt = unpack_payload(out_pay[keep_o])
            orr, ot = unpack_payload(step["payload"][sel])
            assert np.allclose(gr, orr, rtol=2e-3, atol=1e-4) and np.allclose(gt, ot, rtol=2e-3, atol=1e-4)     # one half-precision ulp
            exact = (out_pay[keep_o].view(np.uint32) == step["payload"][sel]).all(axis=1).mean()
            assert exact > 0.98                                                                             # and nearly always the same halves
        fin = (~step["alive"]) & (ids_in >= 0)
        np.add.at(ref_accum, ids_in[fin] // spf, np.concatenate([step["finished"][fin].astype(np.float64), np.ones((fin.sum(), 1))], axis=1))
        finished_paths += int(fin.sum())
        count = survivors
        d_in, d_out = d_out, d_in
        p_in, p_out = p_out, p_in
        if count == 0:
            break
    acc = accum.cpu().numpy()
    assert abs(acc[:, 3].sum() - n) <= len(borderline) + 2 and finished_paths >= n - 8
    assert np.allclose(acc[:, :3].sum(axis=0), ref_accum[:, :3].sum(axis=0), rtol=1e-3)
    close = np.isclose(acc[:, :3], ref_accum[:, :3], rtol=1e-3, atol=1e-3).all(axis=1)
    assert close.mean() > 0.999          # pixels touched by a borderline roulette ray may differ


def test_device_side_bounce_loop_and_slot_sharding(ctx, oracle):
    """atlas_rt_pathtrace_bounces (no host round trip per bounce) == the same frames driven bounce by bounce through
    atlas_rt_pathtrace_bounce; and two slot ranges rendered separately (what two GPUs would do) add up to the whole image."""
    import torch
    scene, osc, ib, keep = pt_scene(ctx, oracle, n_inst=40, seed=8)
    w, h, spf, bounces, frames = 104, 60, 1, 4, 3
    cam = W.camera_frame((30.0, 40.0, -20.0), (30.0, 0.0, 30.0), aspect=w / h)
    ld = np.array([0.2, 0.8, 0.4]) / np.linalg.norm([0.2, 0.8, 0.4])
    prm = capi.pt_params(ld, (3.0, 3.0, 2.5), (0.4, 0.5, 0.8), max_bounces=bounces, samples_per_frame=spf)
    seeds = (np.arange(frames * (bounces + 1), dtype=np.float32) * np.float32(1.618) + np.float32(0.5))
    dev = torch.device("cuda", 0)
    n = w * h * spf
    # reference: bounce by bounce
    accum_ref = torch.zeros((w * h, 4), dtype=torch.float32, device=dev)
    traced_ref = 0
    d_a, d_b = torch.empty((n, 12), dtype=torch.float32, device=dev), torch.empty((n, 12), dtype=torch.float32, device=dev)
    p_a, p_b = torch.zeros((n, 4), dtype=torch.float32, device=dev), torch.zeros((n, 4), dtype=torch.float32, device=dev)
    for f in range(frames):
        ctx.generate_primary_rays(*cam, w, h, spf, jitter=np.tile(capi.sample_jitter(10 + f), (spf, 1)), out=d_a)
        count, ri, ro, pi, po = n, d_a, d_b, p_a, p_b
        for b in range(bounces + 1):
            traced_ref += count
            count = ctx.pathtrace_bounce(scene, prm, float(seeds[f * (bounces + 1) + b]), b, ri, pi, count, ro, po, accum_ref, w, h)
            ri, ro, pi, po = ro, ri, po, pi
            if count == 0:
                break
    ref = accum_ref.cpu().numpy()
    assert abs(ref[:, 3].sum() - frames * n) < 0.5
    for flags in (0, capi.RAY_BINNING):
        accum = torch.zeros((w * h, 4), dtype=torch.float32, device=dev)
        traced = ctx.pathtrace_bounces(scene, cam, w, h, prm, frames, 10, seeds, accum, flags=flags)
        got = accum.cpu().numpy()
        assert traced == traced_ref
        assert np.array_equal(got[:, 3], ref[:, 3])
        assert np.allclose(got[:, :3], ref[:, :3], rtol=1e-5, atol=1e-6)          # same paths; only the order of the atomic adds differs
    # two shards in tile order
    total_slots = n
    cut = (total_slots // 2 // 64) * 64
    tiles = torch.zeros((w * h, 4), dtype=torch.float32, device=dev)
    t0 = ctx.pathtrace_bounces(scene, cam, w, h, prm, frames, 10, seeds, tiles, slot_begin=0, slot_end=cut, flags=capi.ACCUM_TILE_ORDER)
    first = tiles.cpu().numpy().copy()
    assert np.all(first[cut // spf:, 3] == 0) and np.all(first[:cut // spf, 3] == frames)     # a shard owns a contiguous slice
    t1 = ctx.pathtrace_bounces(scene, cam, w, h, prm, frames, 10, seeds, tiles, slot_begin=cut, slot_end=total_slots, flags=capi.ACCUM_TILE_ORDER)
    assert t0 + t1 == traced_ref
    # tile order -> pixel order through the ray IDs of a primary batch
    prim = ctx.generate_primary_rays(*cam, w, h, 1)
    pixel_of_tile_index = prim[:, 3].view(np.int32)
    both = tiles.cpu().numpy()
    untiled = np.zeros_like(both)
    untiled[pixel_of_tile_index] = both
    assert np.array_equal(untiled[:, 3], ref[:, 3])
    assert np.allclose(untiled[:, :3], ref[:, :3], rtol=1e-5, atol=1e-6)


def test_interleaved_shards_assemble_the_whole_frame(ctx, oracle):
    """Three interleaved shards (what three GPUs would render) gathered one after the other and put back by
    atlas_rt_image_from_shards == the frame rendered whole; resolution with ragged tile borders, 2 samples per frame."""
    import torch
    scene, osc, ib, keep = pt_scene(ctx, oracle, n_inst=40, seed=8)
    w, h, spf, bounces, frames, parts, block = 108, 61, 2, 3, 2, 3, 128
    cam = W.camera_frame((30.0, 40.0, -20.0), (30.0, 0.0, 30.0), aspect=w / h)
    ld = np.array([0.2, 0.8, 0.4]) / np.linalg.norm([0.2, 0.8, 0.4])
    prm = capi.pt_params(ld, (3.0, 3.0, 2.5), (0.4, 0.5, 0.8), max_bounces=bounces, samples_per_frame=spf)
    seeds = np.arange(frames * (bounces + 1), dtype=np.float32) * np.float32(1.25) + np.float32(0.75)
    dev = torch.device("cuda", 0)
    whole = torch.zeros((w * h, 4), dtype=torch.float32, device=dev)
    traced_whole = ctx.pathtrace_bounces(scene, cam, w, h, prm, frames, 4, seeds, whole)
    sizes = [ctx.pathtrace_bounces_interleaved(scene, cam, w, h, prm, frames, 4, seeds, p, parts, block)[0] for p in range(parts)]
    assert sum(sizes) == w * h and max(sizes) - min(sizes) <= block
    gathered = torch.zeros((w * h, 4), dtype=torch.float32, device=dev)
    traced, off = 0, 0
    for p in range(parts):
        n, t = ctx.pathtrace_bounces_interleaved(scene, cam, w, h, prm, frames, 4, seeds, p, parts, block, accum_local=gathered[off:off + sizes[p]])
        assert n == sizes[p]
        traced += t
        off += n
    assert traced == traced_whole
    image = torch.empty_like(whole)
    ctx.image_from_shards(gathered, w, h, parts, block, image)
    a, b = image.cpu().numpy(), whole.cpu().numpy()
    assert np.array_equal(a[:, 3], b[:, 3]) and np.all(b[:, 3] == frames * spf)
    assert np.allclose(a[:, :3], b[:, :3], rtol=1e-5, atol=1e-6)


def test_c5_sample_pass_at_full_resolution(ctx, oracle):
    """BASELINE configs[4]: one 3840x2160 sample pass of the 4-bounce path tracer on the instanced scene; the first bounce's
    8.3M closest hits are compared with the oracle on a 100k-ray sample, the image invariants over all pixels."""
    import torch
    from test_gpu_configs import c4_scene
    meshes, ib, ir = c4_scene()
    mats = capi.make_materials(2)
    mats[1]["metalness"], mats[1]["roughness"] = 0.8, 0.4
    ir[:, 13] = (np.arange(len(ir)) % 2).astype(np.uint32)
    blas = ctx.build_blas_batch([W.tri_boxes(t) for t in meshes], meshes)
    gm, t96 = [], []
    for b, t in zip(blas, meshes):
        m = ctx.pack_mesh(b, t)
        m.pack_shading(t, payload11=ctx.pack_shading_words(t, W.smooth_normals(t)))
        gm.append(m)
    tlas = ctx.build_tlas(ib)
    scene = ctx.create_scene(gm, ir, tlas)
    scene.set_materials(mats)
    w, h, bounces = 3840, 2160, 4
    cam = W.camera_frame((1000.0, 260.0, -300.0), (1000.0, 60.0, 1000.0), aspect=w / h)
    ld = np.array([0.3, 0.9, -0.3]) / np.linalg.norm([0.3, 0.9, -0.3])
    prm = capi.pt_params(ld, (3.0, 3.0, 2.5), (0.4, 0.5, 0.8), max_bounces=bounces)
    dev = torch.device("cuda", 0)
    accum = torch.zeros((w * h, 4), dtype=torch.float32, device=dev)
    seeds = np.arange(bounces + 1, dtype=np.float32) + np.float32(0.25)
    traced = ctx.pathtrace_bounces(scene, cam, w, h, prm, 1, 0, seeds, accum)
    acc = accum.cpu().numpy()
    assert np.all(acc[:, 3] == 1.0)                     # every pixel's path finished exactly once
    assert np.isfinite(acc).all() and acc[:, :3].min() >= 0.0
    assert w * h < traced < (bounces + 1) * w * h
    # first bounce against the oracle on a sample
    rays = ctx.generate_primary_rays(*cam, w, h, 1, jitter=capi.sample_jitter(0)[None])
    idx = np.r_[0:40000, 4_000_000:4_040_000, len(rays) - 20000:len(rays)]
    out = ctx.trace(scene, rays[idx], flags=capi.OPACITY)
    inst, tnodes = scene.download()
    dl = [m.download() for m in gm]
    osc = OScene(tnodes, inst, [d[0] for d in dl], [d[1] for d in dl], [m.download_shading() for m in gm])
    osc.set_materials(mats)
    ref, _ = oracle.trace(osc, rays[idx], opacity=True, nthreads=16)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    miss = out[:, 9].view(np.int32) < 0
    sky = np.minimum(np.array([0.4, 0.5, 0.8]), 10.0)
    pix = rays[idx][miss][:, 3].view(np.int32)
    assert np.allclose(acc[pix, :3], sky, rtol=1e-6)    # primary rays that miss see the sky


def test_sample_pass_lanes_give_the_same_image(ctx, oracle):
    """The sample passes of one atlas_rt_pathtrace_bounces call run on up to ATLAS_RT_PT_LANES lanes side by side (own stream and
    buffers each, images added in lane order). One lane (strictly sequential passes), the default four and eight must trace
    the same rays and give the same image up to the order of the per-pixel float additions; a repeated call is bit-identical."""
    import os
    import torch
    dev = torch.device("cuda", 0)
    w, h, spf, bounces, frames = 120, 72, 1, 4, 7
    cam = W.camera_frame((30.0, 40.0, -20.0), (30.0, 0.0, 30.0), aspect=w / h)
    ld = np.array([0.2, 0.8, 0.4]) / np.linalg.norm([0.2, 0.8, 0.4])
    prm = capi.pt_params(ld, (3.0, 3.0, 2.5), (0.4, 0.5, 0.8), max_bounces=bounces, samples_per_frame=spf)
    seeds = np.arange(frames * (bounces + 1), dtype=np.float32) * np.float32(1.37) + np.float32(0.25)
    images, traced = {}, {}
    old = os.environ.get("ATLAS_RT_PT_LANES")
    try:
        for lanes in (1, 4, 8):
            os.environ["ATLAS_RT_PT_LANES"] = str(lanes)
            c = capi.Context(0)
            try:
                scene, osc, ib, keep = pt_scene(c, oracle, n_inst=40, seed=8)
                accum = torch.zeros((w * h, 4), dtype=torch.float32, device=dev)
                traced[lanes] = c.pathtrace_bounces(scene, cam, w, h, prm, frames, 3, seeds, accum)
                images[lanes] = accum.cpu().numpy().copy()
                if lanes == 4:
                    accum.zero_()
                    c.pathtrace_bounces(scene, cam, w, h, prm, frames, 3, seeds, accum)
                    assert np.array_equal(accum.cpu().numpy().view(np.uint32), images[4].view(np.uint32))   # deterministic
            finally:
                c.close()
    finally:
        if old is None:
            os.environ.pop("ATLAS_RT_PT_LANES", None)
        else:
            os.environ["ATLAS_RT_PT_LANES"] = old
    assert traced[1] == traced[4] == traced[8] > frames * w * h
    for lanes in (4, 8):
        assert np.array_equal(images[lanes][:, 3], images[1][:, 3]) and np.all(images[1][:, 3] == frames * spf)
        assert np.allclose(images[lanes][:, :3], images[1][:, :3], rtol=1e-5, atol=1e-6)
