"""Multi-GPU sharding of the traversal: one process per GPU, scene replicated, ray batch split into
contiguous ranges, ONE all-gather of hit records per batch. Builds do not shard (a single BVH is top-down dependent);
independent BLASes are dealt round-robin with `blas_owner`.

The reference has no multi-GPU code at all; this is the B200-native addition. torch.distributed (NCCL over
NVLink/NVSwitch on GPU, gloo in the CPU tests) is plumbing only — no collective sits inside the traversal itself.
"""
import torch
import torch.distributed as dist


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
