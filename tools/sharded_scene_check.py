"""Multi-GPU check (torchrun, N ranks): BLASes built round-robin across GPUs + exchanged over NCCL give, on every rank,
a scene whose trace output is bit-identical to a scene built entirely on that rank; rays are sharded and gathered.
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
rng = np.random.default_rng(7)
meshes = [W.uv_sphere(12 + 2 * k, 6 + k) if k % 2 else W.heightfield(10 + 3 * k, 10 + 2 * k) * np.float32(0.1) for k in range(9)]
mb = [np.concatenate([W.tri_boxes(t)[:, :3].min(0), W.tri_boxes(t)[:, 3:].max(0)]) for t in meshes]
ib, ir = W.random_instances(3000, mb, seed=4, extent=(120.0, 30.0, 120.0))
scene, keep = sharding.build_scene_sharded(ctx, meshes, ib, ir)
# reference: everything built locally
blas = [ctx.build_blas(W.tri_boxes(t), t) for t in meshes]
gm = [ctx.pack_mesh(b, t) for b, t in zip(blas, meshes)]
tl = ctx.build_tlas(ib)
local_scene = ctx.create_scene(gm, ir, tl)
count = 400_000
rays = W.random_rays(count, ib[:, :3].min(0), ib[:, 3:].max(0), seed=5)
b, e = sharding.shard_bounds(count, rank, world)
mine = torch.from_numpy(rays[b:e]).to(dev)
out = torch.empty_like(mine)
ctx.trace(scene, mine, e - b, out=out)
gathered = sharding.gather_hits_ragged(out, count)
full = ctx.trace(local_scene, rays)
same = np.array_equal(gathered.cpu().numpy().view(np.uint32), full[:, 8:12].view(np.uint32))
inst_a, nodes_a = scene.download(); inst_b, nodes_b = local_scene.download()
same_scene = np.array_equal(inst_a, inst_b) and np.array_equal(nodes_a.view(np.uint32), nodes_b.view(np.uint32))
flag = torch.tensor([int(same and same_scene)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world={world} sharded-build scene == local scene: {same_scene}; gathered sharded hits == single-GPU hits: {same}; all ranks ok: {bool(flag.item())}")
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
