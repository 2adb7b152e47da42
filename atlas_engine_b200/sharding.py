"""Multi-GPU sharding of the traversal: one process per GPU, scene replicated, ray batch split into
contiguous ranges, ONE gather of 16-byte hit records per batch. Builds do not shard (a single BVH is top-down dependent);
independent BLASes are dealt round-robin with `blas_owner`.

The product path is in the C ABI (csrc/comm.cu: atlas_rt_comm_init / atlas_rt_build_scene_sharded / atlas_rt_scene_replicate /
atlas_rt_trace_sharded over NCCL); `init_comm` below only hands the NCCL unique id around with torch.distributed, which is
plumbing. The torch.distributed statements of the same exchanges further down (gather_hits*, exchange_flat_trees) are what
the CPU tests run over gloo (tests/test_sharding_gloo.py) to check the host-side arithmetic: shares, offsets, ownership.
The reference has no multi-GPU code at all; this is the B200-native addition.
"""
import torch
import torch.distributed as dist


def init_comm(ctx):
    """capi.Comm over all ranks of the torch.distributed job: rank 0 makes the NCCL unique id, everyone receives it."""
    from . import capi
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    box = [capi.comm_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    return capi.Comm(ctx, box[0], rank, world)


def shard_bounds(count, rank, world, align=64):
    """Contiguous [begin, end) of `count` rays for `rank`; same arithmetic as atlas_rt_shard_range (C ABI): whole
    `align`-ray units (64 = one 8x8 rayGen tile) are dealt as evenly as possible, earlier ranks get the extra unit."""
    units = (count + align - 1) // align
    base, extra = divmod(units, world)
    b = (base * rank + min(rank, extra)) * align
    e = b + (base + (1 if rank < extra else 0)) * align
    return min(b, count), min(e, count)


def blas_owner(mesh_index, world):
    """Which rank builds BLAS `mesh_index` when a scene's meshes are built in parallel across GPUs."""
    return mesh_index % world


def hit_records(rays_out):
    """(n, 12) PackedRay tensor -> contiguous (n, 4) hit records (t, bits(hitID), bits(instanceID), v)."""
    return rays_out.view(-1, 3, 4)[:, 2, :].contiguous()


def gather_hits(rays_out, gathered=None):
    """All-gather the 16-byte hit records of every rank's (equal-sized) share; returns the (world*n, 4) tensor."""
    hits = hit_records(rays_out)
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return hits
    if gathered is None:
        gathered = torch.empty((world * hits.shape[0], 4), dtype=hits.dtype, device=hits.device)
    dist.all_gather_into_tensor(gathered, hits)
    return gathered


def gather_hits_ragged(rays_out, count, align=64):
    """All-gather for shares made by shard_bounds (sizes may differ by one unit): pads to the largest share, gathers,
    and strips the padding. Returns the (count, 4) hit records in global ray order on every rank."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    hits = hit_records(rays_out)
    if world == 1:
        return hits
    sizes = [shard_bounds(count, r, world, align) for r in range(world)]
    biggest = max(e - b for b, e in sizes)
    padded = torch.zeros((biggest, 4), dtype=hits.dtype, device=hits.device)
    padded[: hits.shape[0]] = hits
    out = torch.empty((world * biggest, 4), dtype=hits.dtype, device=hits.device)
    dist.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * biggest: r * biggest + (e - b)] for r, (b, e) in enumerate(sizes)], dim=0)


# --------------------------------------------------------------------------------------- instance sets across GPUs
def exchange_flat_trees(local, count, device):
    """Every rank ends up with all `count` flattened trees. `local` maps tree index -> (nodes (n,14) int32,
    order (m,) int32, end_of_node (m,) uint8) tensors on `device` for the trees this rank built (blas_owner). Sizes are
    exchanged first (one all_reduce), then each tree is broadcast from its owner. Backend-agnostic (NCCL / gloo)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return [local[k] for k in range(count)]
    sizes = torch.zeros((count, 2), dtype=torch.int64, device=device)
    for k, (nodes, order, _) in local.items():
        sizes[k, 0], sizes[k, 1] = nodes.shape[0], order.shape[0]
    dist.all_reduce(sizes)
    sizes = sizes.cpu()
    out = []
    for k in range(count):
        n, m = int(sizes[k, 0]), int(sizes[k, 1])
        if k in local:
            nodes, order, eon = local[k]
        else:
            nodes = torch.empty((n, 14), dtype=torch.int32, device=device)
            order = torch.empty(m, dtype=torch.int32, device=device)
            eon = torch.empty(m, dtype=torch.uint8, device=device)
        src = blas_owner(k, world)
        for t in (nodes, order, eon):
            if t.numel():
                dist.broadcast(t, src=src)
        out.append((nodes, order, eon))
    return out


def build_scene_sharded(ctx, mesh_tris, inst_boxes, inst_records):
    """the BLASes of an instanced scene are independent, so rank r builds meshes r, r+world, ...;
    the flattened trees are exchanged over NCCL and imported on the other ranks (atlas_rt_bvh_import); the TLAS is built
    on rank 0 and broadcast. Every rank returns an identical, complete scene. Returns (scene, keepalive objects)."""
    from . import workloads as W
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    device = torch.device("cuda", ctx.device)
    built, local = {}, {}
    for k, tris in enumerate(mesh_tris):
        if blas_owner(k, world) == rank:
            built[k] = ctx.build_blas(W.tri_boxes(tris), tris)
            local[k] = built[k].download_device()
    trees = exchange_flat_trees(local, len(mesh_tris), device)
    torch.cuda.current_stream().synchronize()   # the broadcasts ran on torch's / NCCL's streams; the imports use the context's
    blas = [built[k] if k in built else ctx.import_bvh_device(*trees[k]) for k in range(len(mesh_tris))]
    meshes = [ctx.pack_mesh(b, t) for b, t in zip(blas, mesh_tris)]
    tl_local = {}
    tlas = None
    if rank == 0:
        tlas = ctx.build_tlas(inst_boxes)
        tl_local[0] = tlas.download_device()
    tl_tree = exchange_flat_trees(tl_local, 1, device)[0]
    torch.cuda.current_stream().synchronize()
    if tlas is None:
        tlas = ctx.import_bvh_device(*tl_tree)
    scene = ctx.create_scene(meshes, inst_records, tlas)
    return scene, (blas, meshes, tlas)
