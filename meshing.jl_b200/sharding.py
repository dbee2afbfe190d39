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


class PeerExchange:
    """The same exchange without NCCL in the data path: every rank's 16 bytes go to every rank as peer-memory stores
    over NVLink, enqueued on the handle's own stream (b200iso_set_peer_exchange / b200iso_exchange_async).
    torch's symmetric memory does the plumbing: it allocates one PEER_BYTES buffer per rank and maps all of them
    into every process.  Raises if the ranks cannot map each other's memory (callers then keep the all-gather)."""

    def __init__(self, handle, group=None, device=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        from . import capi

        group = group or dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > capi.PEER_MAX:
            raise ValueError(f"at most {capi.PEER_MAX} ranks")
        device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.buf = symm.empty(capi.PEER_BYTES // 8, dtype=torch.int64, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, group)
        torch.cuda.synchronize(device)
        dist.barrier(group)  # every buffer is zero before anyone publishes
        self.handle = handle
        handle.set_peer_exchange(self.rank, self.world, [int(p) for p in self.hdl.buffer_ptrs])
        self.bases = torch.zeros(4, dtype=torch.int64, device=device)       # vertex base, face base, total nverts, total nfaces
        self.all = torch.zeros((self.world, 2), dtype=torch.int64, device=device)

    def exchange_async(self):
        """After handle.count_async: publish + gather on the handle's stream; returns the device pointer to pass as
        vertex_base_dev to generate_async."""
        self.handle.exchange_async(self.bases.data_ptr(), self.all.data_ptr())
        return self.bases.data_ptr()

    def close(self):
        self.handle.set_peer_exchange(0, 0, None)
