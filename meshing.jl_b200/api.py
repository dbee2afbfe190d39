"""Host-side mirror of the reference's public interface for the hot path.

Same names, argument meaning and result layout as Meshing.jl v0.7.0:

    isosurface(sdf, method, X, Y, Z) -> (vertices, faces)        src/marching_cubes.jl:27,
                                                                 src/marching_tetrahedra.jl:129
    isosurface(sdf, X, Y, Z) == isosurface(sdf, MarchingCubes(), X, Y, Z)   src/isosurface.jl:30-32
    MarchingCubes(iso=0.0), MarchingTetrahedra(iso=0.0, eps=1e-3)           src/algorithmtypes.jl:23-40

Julia carries the types of `iso`, `eps` and of the range endpoints in the method/range objects; here they
are carried by the Python scalar type: `numpy.float32` <-> Float32, Python `float`/`numpy.float64` <->
Float64, Python `int` <-> Int.  `Float32` / `Float64` are exported as aliases, so the Float32-vertex call of
the benchmark reads `isosurface(sdf, MarchingCubes(iso=Float32(0)))`.

Everything is computed by libb200iso.so (CUDA, sm_100a).  Inputs the accelerated path does not cover
(fields that are not Float32/Float64) raise TypeError -- there is no CPU fallback.
"""
from dataclasses import dataclass
from typing import Any

import numpy as np

from . import capi

Float32 = np.float32
Float64 = np.float64


@dataclass(frozen=True)
class MarchingCubes:
    """MarchingCubes(iso=0.0)  (src/algorithmtypes.jl:23-25)"""
    iso: Any = 0.0


@dataclass(frozen=True)
class MarchingTetrahedra:
    """MarchingTetrahedra(iso=0.0, eps=1e-3)  (src/algorithmtypes.jl:37-40)"""
    iso: Any = 0.0
    eps: Any = 1e-3


def _scalar_kind(v, what):
    """-> (value as float, is_f32).  An Int level takes the type of the other operands (promote_type(Int64,
    Float32) == Float32; next to a Float64 field the Float32 flag is harmless: everything promotes to Float64),
    provided it is exact in Float32."""
    if isinstance(v, (np.float32,)):
        return float(v), True
    if isinstance(v, (bool, np.bool_)):
        raise TypeError(f"{what} must be a real number")
    if isinstance(v, (int, np.integer)):
        if float(np.float32(v)) != float(v):
            raise TypeError(f"integer {what}={v} is not exactly representable in Float32")
        return float(v), True
    if isinstance(v, (float, np.float64)):
        return float(v), False
    raise TypeError(f"unsupported type for {what}: {type(v).__name__} (Float32, Float64 or Int)")


def _range_kind(r, name):
    """first/last element and element kind of a range-like (tuple, list, range, ndarray).  Only the
    endpoints are used, like LinRange(first(X), last(X), n) in src/marching_cubes.jl:36-38."""
    if isinstance(r, range):
        if len(r) == 0:
            raise ValueError(f"{name} is empty")
        return float(r[0]), float(r[-1]), capi.RANGE_INT
    a = r[0], r[-1]
    kinds = []
    for v in a:
        if isinstance(v, (int, np.integer)) and not isinstance(v, (bool, np.bool_)):
            kinds.append(capi.RANGE_INT)
        elif isinstance(v, np.float32):
            kinds.append(capi.RANGE_F32)
        elif isinstance(v, (float, np.float64)):
            kinds.append(capi.RANGE_F64)
        else:
            raise TypeError(f"unsupported endpoint type in {name}: {type(v).__name__}")
    # The reference takes the vertex type from eltype(first(X)) only (src/marching_cubes.jl:31) but the LinRange
    # element type from BOTH endpoints (LinRange(first(X), last(X), n), :36-38).  In terms of b200iso_params:
    #   first is Float64                      -> RANGE_F64 (Float64 coordinates, Float64 vertices)
    #   else promote(first, last) is Float32  -> RANGE_F32 (Float32 coordinates)
    #   else                                  -> RANGE_INT (Float64 coordinates, no promotion of the vertex type):
    #                                            (Int, Int), (Int, Float64), (Float32, Float64)
    if kinds[0] == capi.RANGE_F64:
        kind = capi.RANGE_F64
    elif capi.RANGE_F64 not in kinds and capi.RANGE_F32 in kinds:
        kind = capi.RANGE_F32
    else:
        kind = capi.RANGE_INT
    return float(a[0]), float(a[1]), kind


def make_params(method, X=(-1, 1), Y=(-1, 1), Z=(-1, 1)):
    """Flatten (method, X, Y, Z) of the reference call into struct b200iso_params."""
    p = capi.Params()
    if isinstance(method, MarchingCubes):
        p.algo = capi.MC
        p.iso, isf = _scalar_kind(method.iso, "iso")
        p.eps, epf = 1e-3, True
    elif isinstance(method, MarchingTetrahedra):
        p.algo = capi.MT
        p.iso, isf = _scalar_kind(method.iso, "iso")
        p.eps, epf = _scalar_kind(method.eps, "eps")
    else:
        raise TypeError("method must be MarchingCubes(...) or MarchingTetrahedra(...)")
    p.iso_is_f32, p.eps_is_f32 = int(isf), int(epf)
    x0, x1, kx = _range_kind(X, "X")
    y0, y1, ky = _range_kind(Y, "Y")
    z0, z1, kz = _range_kind(Z, "Z")
    if len({kx, ky, kz}) != 1:
        raise TypeError("X, Y, Z must have the same element type (Int, Float32 or Float64)")
    p.range_kind = kx
    p.x0, p.x1, p.y0, p.y1, p.z0, p.z1 = x0, x1, y0, y1, z0, z1
    return p


_handles = {}


def get_handle(device=0):
    h = _handles.get(device)
    if h is None:
        h = _handles[device] = capi.Handle(device)
    return h


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def isosurface(sdf, *args, device=None, capacity=None):
    """isosurface(sdf[, method][, X, Y, Z]) -> (vertices, faces)

    sdf      : 3-D Float32 (fast path) or Float64 array, `sdf[x, y, z]`.  numpy array (host; any strides, made x-contiguous like a
               Julia Array) or a CUDA torch tensor whose x stride is 1 (device-resident: outputs are CUDA
               tensors too and nothing crosses PCIe).
    method   : MarchingCubes(iso=...) (default) or MarchingTetrahedra(iso=..., eps=...)
    X, Y, Z  : ranges whose first/last elements give the extent; default -1:1 on every axis
    capacity : optional (max_vertices, max_faces) for a host field instead of the built-in guess (see below)
    returns  : vertices (nverts, 3) Float32|Float64 by the reference's promotion rule, faces (nfaces, 3)
               int64, 1-based, in the reference's order (x-outermost, z-innermost voxel scan).

    A host field takes the one-shot slab pipeline (b200iso_extract_host): the mesh streams out of the GPU while the
    field still streams in.  The output arrays must exist before the mesh size is known, so they are allocated from
    a guess -- the totals of the previous call with the same shape and method, else a surface-area estimate -- and
    trimmed afterwards (untouched pages of the over-allocation are never backed by memory).  If the guess was short
    the mesh is fetched into arrays of the exact size from the slabs still resident on the device
    (b200iso_extract_host_resident): the field is uploaded once in every case.
    """
    if args and isinstance(args[0], (MarchingCubes, MarchingTetrahedra)):
        method, rest = args[0], args[1:]
    else:
        method, rest = MarchingCubes(), args  # src/isosurface.jl:30-32
    if len(rest) not in (0, 3):
        raise TypeError("isosurface(sdf[, method][, X, Y, Z])")
    params = make_params(method, *rest)

    if _is_torch(sdf):
        return _isosurface_torch(sdf, params, device)
    a = np.asarray(sdf)
    if a.ndim != 3:
        raise TypeError("sdf must be a 3-D array")
    if a.dtype not in (np.float32, np.float64):
        raise TypeError(f"the B200 path accepts Float32 and Float64 fields (got {a.dtype}); there is no CPU fallback")
    a = np.asfortranarray(a)
    params.field_is_f64 = int(a.dtype == np.float64)
    nx, ny, nz = a.shape
    h = get_handle(0 if device is None else device)
    f64 = _vert_is_f64(params)
    key = (h.device, a.shape, a.dtype.str, params.algo)
    vcap, fcap = (int(capacity[0]), int(capacity[1])) if capacity is not None else _capacity_guess(key, a.shape, params.algo)
    verts = np.empty((vcap, 3), dtype=np.float64 if f64 else np.float32)
    faces = np.empty((fcap, 3), dtype=np.int64)
    nv, nf, f64, fits = h.extract_host(params, a.ctypes.data, nx, ny, nz, nx, verts.ctypes.data, vcap, faces.ctypes.data, fcap)
    if not fits:  # the guess was short: exact arrays, filled from the slabs that are still on the device
        verts = np.empty((nv, 3), dtype=np.float64 if f64 else np.float32)
        faces = np.empty((nf, 3), dtype=np.int64)
        nv2, nf2, fits = h.extract_host_resident(verts.ctypes.data, nv, faces.ctypes.data, nf)
        if not fits or (nv2, nf2) != (nv, nf):
            raise RuntimeError("b200iso_extract_host_resident: totals changed between the two passes")
    _capacity_memo[key] = (nv, nf)
    return verts[:nv], faces[:nf]


def isosurface_two_phase(sdf, *args, device=None):
    """The same result through the two-phase pair of the C ABI -- b200iso_count (upload, classify, count: the caller
    learns the sizes) then b200iso_generate into exact arrays -- i.e. without overlap between the upload and the
    download.  Kept as a public form (a caller that cannot over-allocate) and as the cross-check of the one-shot path."""
    if args and isinstance(args[0], (MarchingCubes, MarchingTetrahedra)):
        method, rest = args[0], args[1:]
    else:
        method, rest = MarchingCubes(), args
    params = make_params(method, *rest)
    a = np.asfortranarray(np.asarray(sdf))
    if a.ndim != 3 or a.dtype not in (np.float32, np.float64):
        raise TypeError("3-D Float32/Float64 field expected")
    params.field_is_f64 = int(a.dtype == np.float64)
    nx, ny, nz = a.shape
    h = get_handle(0 if device is None else device)
    nv, nf, f64 = h.count(params, a.ctypes.data, capi.HOST, nx, ny, nz, nx)
    verts = np.empty((nv, 3), dtype=np.float64 if f64 else np.float32)
    faces = np.empty((nf, 3), dtype=np.int64)
    h.generate(verts.ctypes.data, faces.ctypes.data, capi.HOST, 0)
    return verts, faces


_capacity_memo = {}


def _capacity_guess(key, shape, algo):
    """(max_vertices, max_faces) for the one-shot call before the mesh size is known: the previous totals for this
    shape and method plus 1/8, else an estimate from the surface area an isosurface of this grid typically has
    (the 1024^3 gyroid of the benchmark: 38.7 n^2 MC vertices, 19.3 n^2 faces; MT 27.7 n^2 / 55.2 n^2)."""
    memo = _capacity_memo.get(key)
    if memo is not None:
        return memo[0] + memo[0] // 8 + 1024, memo[1] + memo[1] // 8 + 1024
    nx, ny, nz = (max(int(d) - 1, 0) for d in shape)
    area = (nx * ny + ny * nz + nx * nz) / 3.0
    nvox = nx * ny * nz
    if algo == capi.MC:
        return int(min(64 * area, 12 * nvox)) + 1024, int(min(32 * area, 5 * nvox)) + 1024
    return int(min(48 * area, 7 * nvox)) + 1024, int(min(96 * area, 12 * nvox)) + 1024


def _vert_is_f64(params):
    """float(promote_type(eltype(X), eltype(Y), eltype(Z), T_sdf, typeof(iso)[, typeof(eps)])) == Float64
    (src/marching_cubes.jl:31, src/marching_tetrahedra.jl:131)."""
    return bool(params.field_is_f64 or not params.iso_is_f32 or params.range_kind == 2
                or (params.algo == 1 and not params.eps_is_f32))


def isosurface_into(sdf, verts_out, faces_out, *args, device=None):
    """isosurface() of a host field into caller-owned host arrays (e.g. pinned, re-used every time step) through the
    one-shot slab pipeline b200iso_extract_host.  verts_out: (vcap, 3) float32|float64 C-contiguous, faces_out:
    (fcap, 3) int64.  Returns (nverts, nfaces); raises B200IsoError(ECAPACITY) if the mesh does not fit."""
    if args and isinstance(args[0], (MarchingCubes, MarchingTetrahedra)):
        method, rest = args[0], args[1:]
    else:
        method, rest = MarchingCubes(), args
    params = make_params(method, *rest)
    a = np.asarray(sdf)
    if a.ndim != 3 or a.dtype not in (np.float32, np.float64):
        raise TypeError("3-D Float32/Float64 field expected")
    if not a.flags.f_contiguous:
        raise TypeError("isosurface_into takes a Fortran-ordered (x-contiguous) field: it does not copy")
    params.field_is_f64 = int(a.dtype == np.float64)
    want = np.float64 if _vert_is_f64(params) else np.float32
    if verts_out.dtype != want or faces_out.dtype != np.int64 or not verts_out.flags.c_contiguous or not faces_out.flags.c_contiguous:
        raise TypeError(f"verts_out must be C-contiguous {np.dtype(want)} (n, 3), faces_out C-contiguous int64 (n, 3)")
    nx, ny, nz = a.shape
    h = get_handle(0 if device is None else device)
    nv, nf, _, fits = h.extract_host(params, a.ctypes.data, nx, ny, nz, nx, verts_out.ctypes.data, verts_out.shape[0],
                                     faces_out.ctypes.data, faces_out.shape[0])
    if not fits:
        raise capi.B200IsoError(capi.ECAPACITY, f"mesh has {nv} vertices / {nf} faces; capacity {verts_out.shape[0]} / {faces_out.shape[0]}")
    return nv, nf


def _isosurface_torch(t, params, device):
    import torch

    if t.dim() != 3 or t.dtype not in (torch.float32, torch.float64) or not t.is_cuda:
        raise TypeError("torch input must be a 3-D float32/float64 CUDA tensor")
    params.field_is_f64 = int(t.dtype == torch.float64)
    nx, ny, nz = t.shape
    sx, sy, sz = t.stride()
    if nx > 1 and sx != 1:
        raise TypeError("torch input must be x-contiguous (stride 1 on dim 0), i.e. Julia/Fortran order")
    ldx = sy if ny > 1 else max(nx, 1)
    if ny > 1 and nz > 1 and sz != ldx * ny:
        raise TypeError("torch input must have strides (1, ldx, ldx*ny)")
    dev = t.device.index
    h = get_handle(dev)
    with torch.cuda.device(dev):
        h.set_stream(torch.cuda.current_stream().cuda_stream)
        try:
            nv, nf, f64 = h.count(params, t.data_ptr(), capi.DEVICE, nx, ny, nz, ldx)
            verts = torch.empty((nv, 3), dtype=torch.float64 if f64 else torch.float32, device=t.device)
            faces = torch.empty((nf, 3), dtype=torch.int64, device=t.device)
            h.generate(verts.data_ptr(), faces.data_ptr(), capi.DEVICE, 0)
        finally:
            h.use_own_stream()  # the cached handle must not stay bound to torch's stream if anything raised
    return verts, faces


def isosurface_slab(sdf_slab, method, x_offset, nx_global, vertex_base, *ranges, device=0):
    """One rank's share of a sharded call (MC, or MT with its ghost row): `sdf_slab` holds samples [x_offset, x_offset+nx) of
    a volume with nx_global samples along x; `ranges` are X, Y, Z of the WHOLE volume.  Returns the slab's
    (vertices, faces) with `vertex_base` already added to the face indices.  Two-phase use
    (count -> all-gather -> generate) is `slab_count` + `slab_generate`."""
    h, nv, nf, f64 = slab_count(sdf_slab, method, x_offset, nx_global, *ranges, device=device)
    return slab_generate(h, nv, nf, f64, vertex_base)


def slab_count(sdf_slab, method, x_offset, nx_global, *ranges, device=0):
    """Phase 1 of a slab: classify + count.  numpy slab (host) or x-contiguous CUDA torch slab (device-resident)."""
    params = make_params(method, *ranges)
    params.x_offset, params.nx_global = x_offset, nx_global
    # Marching Tetrahedra: every slab that does not start at x = 0 carries the previous slab's last voxel row
    params.x_ghost = int(isinstance(method, MarchingTetrahedra) and x_offset > 0)
    if _is_torch(sdf_slab):
        import torch
        t = sdf_slab
        if t.dim() != 3 or t.dtype != torch.float32 or not t.is_cuda or (t.shape[0] > 1 and t.stride(0) != 1):
            raise TypeError("torch slab must be a 3-D float32 CUDA tensor with x stride 1")
        nx, ny, nz = t.shape
        h = get_handle(t.device.index)
        torch.cuda.current_stream(t.device).synchronize()  # the handle runs on its own stream
        nv, nf, f64 = h.count(params, t.data_ptr(), capi.DEVICE, nx, ny, nz, t.stride(1) if ny > 1 else nx)
        h._slab_device = t.device
        return h, nv, nf, f64
    a = np.asfortranarray(np.asarray(sdf_slab))
    if a.dtype not in (np.float32, np.float64) or a.ndim != 3:
        raise TypeError("3-D Float32/Float64 field expected")
    nx, ny, nz = a.shape
    h = get_handle(device)
    h._slab_device = None
    params.field_is_f64 = int(a.dtype == np.float64)
    nv, nf, f64 = h.count(params, a.ctypes.data, capi.HOST, nx, ny, nz, nx)
    return h, nv, nf, f64


def slab_generate(h, nv, nf, f64, vertex_base):
    """Phase 2 of a slab: generate with the slab's global vertex base added to the face indices."""
    dev = getattr(h, "_slab_device", None)
    if dev is not None:
        import torch
        verts = torch.empty((nv, 3), dtype=torch.float64 if f64 else torch.float32, device=dev)
        faces = torch.empty((nf, 3), dtype=torch.int64, device=dev)
        h.generate(verts.data_ptr(), faces.data_ptr(), capi.DEVICE, vertex_base)
        return verts, faces
    verts = np.empty((nv, 3), dtype=np.float64 if f64 else np.float32)
    faces = np.empty((nf, 3), dtype=np.int64)
    h.generate(verts.ctypes.data, faces.ctypes.data, capi.HOST, vertex_base)
    return verts, faces


def case_indices(sdf, method=None, device=0):
    """Per-voxel case index (_get_cubeindex, src/common.jl:10-20) in scan-rank order, computed on the GPU
    from the classify kernel's bit-field (parity check hook)."""
    method = method or MarchingCubes()
    a = np.asfortranarray(np.asarray(sdf))
    if a.dtype not in (np.float32, np.float64) or a.ndim != 3:
        raise TypeError("3-D Float32/Float64 field expected")
    nx, ny, nz = a.shape
    h = get_handle(device)
    params = make_params(method)
    params.field_is_f64 = int(a.dtype == np.float64)
    h.count(params, a.ctypes.data, capi.HOST, nx, ny, nz, nx)
    out = np.empty(max(nx - 1, 0) * max(ny - 1, 0) * max(nz - 1, 0), dtype=np.uint8)
    h.case_indices(out.ctypes.data, capi.HOST)
    return out
