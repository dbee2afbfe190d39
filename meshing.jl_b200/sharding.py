"""x-slab sharding of Marching Cubes across ranks.

Output order is x-outermost (src/marching_cubes.jl:40), so rank r's voxel x-planes [xa_r, xb_r) produce a
contiguous range of the global vertex and face arrays.  A slab needs one halo sample plane at high x and ONE
exchange: the all-gather of (nverts_r, nfaces_r) -- 16 bytes per rank -- whose exclusive prefix is the
vertex base added to the slab's face indices (b200iso_generate's vertex_base) and the offset of the slab in
the stitched arrays.  Concatenating the slabs in rank order is byte-identical to the unsharded call.
MT shards the same way with one extra GHOST voxel row per slab (a slab's first own row references vertices owned by
the previous slab's last row): the ghost row is counted locally, not emitted, and excluded from the totals.
"""
import numpy as np


def slab_bounds(nx, world, rank, ghost=False):
    """Sample range [xa, xb) of rank `rank`: voxel planes [xa, xb-1); one halo plane at high x.
    Voxel planes are split as evenly as possible (nx-1 planes over `world` ranks).
    ghost=True (Marching Tetrahedra): slabs that do not start at x = 0 additionally carry the previous slab's
    last voxel row (one more sample plane at low x) -- b200iso_params.x_ghost."""
    nvx = max(nx - 1, 0)
    va = nvx * rank // world
    vb = nvx * (rank + 1) // world
    if vb <= va:
        return va, va
    return (va - 1 if ghost and va > 0 else va), vb + 1


def exclusive_bases(counts, rank):
    """counts: (world, 2) gathered (nverts, nfaces) -> (vertex_base, face_base) of `rank`."""
    c = np.asarray(counts).reshape(-1, 2)
    return int(c[:rank, 0].sum()), int(c[:rank, 1].sum())


def allgather_counts(nverts, nfaces, group=None):
    """The sharded path's only collective: all-gather of this rank's (nverts, nfaces).  Works on any
    torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.tensor([nverts, nfaces], dtype=torch.int64, device=dev)
    out = torch.zeros(world * 2, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(out, mine, group=group)
    return out.view(world, 2).cpu().numpy()


def stitch(parts):
    """Concatenate per-rank (vertices, faces) in rank order (faces already carry their vertex base)."""
    v = np.concatenate([p[0] for p in parts]) if parts else np.empty((0, 3), np.float32)
    f = np.concatenate([p[1] for p in parts]) if parts else np.empty((0, 3), np.int64)
    return v, f
