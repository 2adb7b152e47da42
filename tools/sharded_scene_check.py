"""Multi-GPU check (torchrun, N ranks) of the C-ABI multi-GPU entry points (csrc/comm.cu, NCCL):
  1. atlas_rt_build_scene_sharded (BLAS builds dealt round-robin, trees broadcast, TLAS from rank 0) gives, on every rank, a
     scene that is byte-identical to one built entirely on that rank;
  2. atlas_rt_trace_sharded: every rank traces its share, the root's gathered 16-byte hit records equal a single-GPU trace
     (NCCL gather, and the peer-memory window where the traversal kernels store straight into the root's memory);
  3. atlas_rt_scene_replicate: rank 0's complete scene (96-byte triangles, materials, textures) on every rank traces
     identically (closest + opacity-aware any-hit);
  4. the path tracer sharded by rayGen slot ranges + atlas_rt_comm_gather of the tile-ordered image == the single-GPU image.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/sharded_scene_check.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from atlas_engine_b200 import capi, sharding, workloads as W
rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
dist.init_process_group("nccl", device_id=dev)
ctx = capi.Context(local_rank)
comm = sharding.init_comm(ctx)
meshes = [W.uv_sphere(12 + 2 * k, 6 + k) if k % 2 else W.heightfield(10 + 3 * k, 10 + 2 * k) * np.float32(0.1) for k in range(9)]
mb = [np.concatenate([W.tri_boxes(t)[:, :3].min(0), W.tri_boxes(t)[:, 3:].max(0)]) for t in meshes]
ib, ir = W.random_instances(3000, mb, seed=4, extent=(120.0, 30.0, 120.0))
ir[:, 13] = (np.arange(len(ir)) % 2).astype(np.uint32)
results = {}

# ---- 1. sharded scene build vs a local build
scene = comm.build_scene_sharded(meshes, ib, ir)
blas = [ctx.build_blas(W.tri_boxes(t), t) for t in meshes]
gm = [ctx.pack_mesh(b, t) for b, t in zip(blas, meshes)]
tl = ctx.build_tlas(ib)
local_scene = ctx.create_scene(gm, ir, tl)
n_nodes, n_inst = tl.counts()
inst_a, nodes_a = scene.download(n_inst, n_nodes)
inst_b, nodes_b = local_scene.download()
results["sharded_scene_equals_local"] = bool(np.array_equal(inst_a, inst_b) and np.array_equal(nodes_a.view(np.uint32), nodes_b.view(np.uint32)))

# ---- 2. sharded trace + gather vs single GPU
count = 400_037
rays = W.random_rays(count, ib[:, :3].min(0), ib[:, 3:].max(0), seed=5)
b, e = sharding.shard_bounds(count, rank, world)
full = ctx.trace(local_scene, rays)
ok = True
for any_hit in (False, True):
    ref = ctx.trace(local_scene, rays, any_hit=any_hit, t_max=90.0)
    hits = comm.trace_sharded(scene, rays[b:e], count, any_hit=any_hit, t_max=90.0)                       # host rays, host hits on root
    d_hits = torch.empty((count, 4), dtype=torch.float32, device=dev) if rank == 0 else None
    comm.trace_sharded(scene, torch.from_numpy(rays[b:e]).to(dev), count, hits_out=d_hits, any_hit=any_hit, t_max=90.0)   # device in / out
    if rank == 0:
        ok &= np.array_equal(hits.view(np.uint32), ref[:, 8:12].view(np.uint32))
        ok &= np.array_equal(d_hits.cpu().numpy().view(np.uint32), ref[:, 8:12].view(np.uint32))
results["gathered_hits_equal_single_gpu"] = bool(ok)

# ---- 2b. the same with the gather fused into the traversal (ATLAS_RT_PEER_OUTPUT): every rank's kernel stores its records in
# rank 0's window over NVLink; several calls in a row exercise both window slots and the flow-control counters
ok = True
for rep in range(5):
    any_hit = bool(rep & 1)
    ref = ctx.trace(local_scene, rays, any_hit=any_hit, t_max=90.0)
    d_hits = torch.zeros((count, 4), dtype=torch.float32, device=dev) if rank == 0 else None
    if rep < 3:
        comm.trace_sharded(scene, torch.from_numpy(rays[b:e]).to(dev), count, hits_out=d_hits, any_hit=any_hit, t_max=90.0, flags=capi.PEER_OUTPUT)
        got = d_hits.cpu().numpy() if rank == 0 else None
    else:   # host rays in, host records out on the root
        got = comm.trace_sharded(scene, rays[b:e], count, hits_out=np.zeros((count, 4), np.float32) if rank == 0 else None, any_hit=any_hit, t_max=90.0,
                                 flags=capi.PEER_OUTPUT)
    comm.synchronize()
    if rank == 0:
        ok &= np.array_equal(got.view(np.uint32), ref[:, 8:12].view(np.uint32))
results["peer_window_hits_equal_single_gpu"] = bool(ok)

# ---- 3. replicate rank 0's complete scene (shading triangles, materials, an opacity texture)
mats = capi.make_materials(3)
mats[1]["opacity"] = 0.5
mats[2]["opacityTexture"] = 0
tex = [(np.random.default_rng(3).random((16, 16)) > 0.5).astype(np.uint8) * 255]
full_scene = None
if rank == 0:
    for m, t in zip(gm, meshes):
        n = len(t)
        midx = (np.arange(n) % 2).astype(np.int32)
        op = np.where(midx == 1, np.float32(-1.0), np.float32(1.0)).astype(np.float32)
        m.pack_shading(t, material_idx=midx, opacity=op, payload11=ctx.pack_shading_words(t, W.smooth_normals(t), W.planar_uvs(t, 0.2)))
    full_scene = ctx.create_scene(gm, ir, tl)
    full_scene.set_materials(mats, tex)
replica = comm.replicate_scene(full_scene, 0)
small = rays[:100_000]
mine = {k: ctx.trace(replica, small, any_hit=k, t_max=120.0, flags=capi.OPACITY) for k in (False, True)}
gathered = [torch.empty((len(small), 12), dtype=torch.float32, device=dev) for _ in range(world)]
ok = True
for k in (False, True):
    dist.all_gather(gathered, torch.from_numpy(mine[k]).to(dev))
    ok &= all(torch.equal(g.view(torch.int32), gathered[0].view(torch.int32)) for g in gathered)
results["replicated_scene_traces_identically"] = bool(ok)

# ---- 4. path tracer sharded by slot ranges + gather of the tile-ordered image
w, h, bounces, frames = 200, 120, 3, 2
cam = W.camera_frame((60.0, 50.0, -30.0), (60.0, 0.0, 60.0), aspect=w / h)
prm = capi.pt_params(np.array([0.3, 0.9, -0.3]) / np.linalg.norm([0.3, 0.9, -0.3]), (3.0, 3.0, 2.5), (0.4, 0.5, 0.8), max_bounces=bounces)
seeds = np.arange(frames * (bounces + 1), dtype=np.float32) + np.float32(0.5)
sb, se = sharding.shard_bounds(w * h, rank, world)
part = torch.zeros((w * h, 4), dtype=torch.float32, device=dev)
ctx.pathtrace_bounces(replica, cam, w, h, prm, frames, 0, seeds, part, slot_begin=sb, slot_end=se, flags=capi.ACCUM_TILE_ORDER)
image = torch.zeros((w * h, 4), dtype=torch.float32, device=dev) if rank == 0 else None
bounds = [sharding.shard_bounds(w * h, r, world) for r in range(world)]
comm.gather(part[sb:se], (se - sb) * 16, image, [(e_ - b_) * 16 for b_, e_ in bounds], [b_ * 16 for b_, _ in bounds])
if rank == 0:
    whole = torch.zeros((w * h, 4), dtype=torch.float32, device=dev)
    ctx.pathtrace_bounces(replica, cam, w, h, prm, frames, 0, seeds, whole, flags=capi.ACCUM_TILE_ORDER)
    a, b_ = image.cpu().numpy(), whole.cpu().numpy()
    results["sharded_pathtrace_equals_single_gpu"] = bool(np.array_equal(a[:, 3], b_[:, 3]) and np.allclose(a[:, :3], b_[:, :3], rtol=1e-5, atol=1e-6))
else:
    results["sharded_pathtrace_equals_single_gpu"] = True

flag = torch.tensor([int(all(results.values()))], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world={world} " + "; ".join(f"{k}: {v}" for k, v in results.items()) + f"; all ranks ok: {bool(flag.item())}")
comm.close()
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
