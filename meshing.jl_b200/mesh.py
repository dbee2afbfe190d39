"""Mesh consumers: what callers run next on the extracted mesh (not part of Meshing.jl v0.7.0).

    vertex_normals(sdf, vertices[, X, Y, Z])       unit normals from the gradient of the field (GPU)
    isosurface_welded(sdf, MarchingCubes(...)...)  the indexed form of the Marching Cubes mesh: every grid-edge vertex once,
                                                   in first-occurrence order of the reference's scan (GPU)
    write_ply / write_stl                          binary PLY (optional normals) / binary STL

Everything is computed by libb200iso.so; torch is used only to hold device memory.
"""
import numpy as np

from . import api, capi


def _to_device_field(sdf):
    """-> (torch CUDA tensor with logical shape (nx, ny, nz) and x stride 1, ldx)"""
    import torch
    if api._is_torch(sdf):
        t = sdf
        if t.dim() != 3 or not t.is_cuda or t.dtype not in (torch.float32, torch.float64) or (t.shape[0] > 1 and t.stride(0) != 1):
            raise TypeError("torch field must be a 3-D float32/float64 CUDA tensor with x stride 1")
        return t, (t.stride(1) if t.shape[1] > 1 else max(t.shape[0], 1))
    a = np.asfortranarray(np.asarray(sdf))
    if a.ndim != 3 or a.dtype not in (np.float32, np.float64):
        raise TypeError("3-D Float32/Float64 field expected")
    t = torch.from_numpy(np.ascontiguousarray(a.transpose(2, 1, 0))).cuda().permute(2, 1, 0)  # memory order z, y, x
    return t, a.shape[0]


def vertex_normals(sdf, vertices, X=(-1, 1), Y=(-1, 1), Z=(-1, 1), x_offset=0, nx_global=0):
    """Unit normals (float32, (n, 3)) of `vertices` (numpy or CUDA torch, float32/float64 (n, 3)) from the gradient of
    `sdf` sampled on the ranges X, Y, Z (the same ones the extraction used): trilinear blend of central differences,
    pointing towards increasing field values.  Returns the same kind of array as `vertices`."""
    import torch
    t, ldx = _to_device_field(sdf)
    p = api.make_params(api.MarchingCubes(), X, Y, Z)
    p.field_is_f64 = int(t.dtype == torch.float64)
    p.x_offset, p.nx_global = x_offset, nx_global
    as_numpy = not api._is_torch(vertices)
    v = torch.from_numpy(np.ascontiguousarray(vertices)).cuda() if as_numpy else vertices.contiguous()
    if v.dim() != 2 or v.shape[1] != 3 or v.dtype not in (torch.float32, torch.float64):
        raise TypeError("vertices must be (n, 3) float32/float64")
    out = torch.empty((v.shape[0], 3), dtype=torch.float32, device=t.device)
    h = api.get_handle(t.device.index)
    nx, ny, nz = t.shape
    with torch.cuda.device(t.device):
        h.set_stream(torch.cuda.current_stream().cuda_stream)
        try:
            h.vertex_normals_async(p, t.data_ptr(), nx, ny, nz, ldx, v.data_ptr(), v.shape[0], v.dtype == torch.float64, out.data_ptr())
            torch.cuda.current_stream().synchronize()
        finally:
            h.use_own_stream()
    return out.cpu().numpy() if as_numpy else out


def isosurface_welded(sdf, method=None, X=(-1, 1), Y=(-1, 1), Z=(-1, 1)):
    """isosurface(sdf, MarchingCubes(...), X, Y, Z) as an INDEXED mesh: the vertex of every crossed grid edge once (the
    copy the reference's sweep creates first), faces renumbered.  Returns (vertices, faces) of the same kind as `sdf`
    (numpy in -> numpy out)."""
    import torch
    method = method or api.MarchingCubes()
    if not isinstance(method, api.MarchingCubes):
        raise TypeError("welding applies to Marching Cubes (Marching Tetrahedra meshes are indexed already)")
    t, ldx = _to_device_field(sdf)
    p = api.make_params(method, X, Y, Z)
    p.field_is_f64 = int(t.dtype == torch.float64)
    nx, ny, nz = t.shape
    h = api.get_handle(t.device.index)
    dev = t.device
    with torch.cuda.device(dev):
        h.set_stream(torch.cuda.current_stream().cuda_stream)
        try:
            nv, nf, f64 = h.count(p, t.data_ptr(), capi.DEVICE, nx, ny, nz, ldx)
            vt = torch.float64 if f64 else torch.float32
            verts = torch.empty((nv, 3), dtype=vt, device=dev)
            faces = torch.empty((nf, 3), dtype=torch.int64, device=dev)
            keys = torch.empty(max(nv, 1), dtype=torch.int64, device=dev)
            h.vertex_keys_async(keys.data_ptr(), nv)
            h.generate(verts.data_ptr(), faces.data_ptr(), capi.DEVICE, 0)
            wv = torch.empty((nv, 3), dtype=vt, device=dev)
            wf = torch.empty((nf, 3), dtype=torch.int64, device=dev)
            nw = h.weld(keys.data_ptr(), verts.data_ptr(), nv, f64, faces.data_ptr(), nf, 0, wv.data_ptr(), wf.data_ptr())
        finally:
            h.use_own_stream()
    wv = wv[:nw]
    if api._is_torch(sdf):
        return wv, wf
    return wv.cpu().numpy(), wf.cpu().numpy()


def write_ply(path, vertices, faces, normals=None):
    """Binary little-endian PLY of a mesh in host memory (faces 1-based as returned by isosurface; written 0-based)."""
    v = np.ascontiguousarray(vertices)
    f = np.ascontiguousarray(faces, dtype=np.int64)
    n = None if normals is None else np.ascontiguousarray(normals, dtype=np.float32)
    if v.dtype not in (np.float32, np.float64) or v.ndim != 2 or v.shape[1] != 3 or f.ndim != 2 or f.shape[1] != 3:
        raise TypeError("vertices (n, 3) float32/float64 and faces (m, 3) expected")
    if n is not None and n.shape != (len(v), 3):
        raise TypeError("normals must be (n, 3)")
    capi.write_ply(path, v.ctypes.data, len(v), v.dtype == np.float64, None if n is None else n.ctypes.data, f.ctypes.data, len(f))


def write_stl(path, vertices, faces):
    """Binary STL of a mesh in host memory (one facet per face, facet normal from the winding)."""
    v = np.ascontiguousarray(vertices)
    f = np.ascontiguousarray(faces, dtype=np.int64)
    if v.dtype not in (np.float32, np.float64) or v.ndim != 2 or v.shape[1] != 3 or f.ndim != 2 or f.shape[1] != 3:
        raise TypeError("vertices (n, 3) float32/float64 and faces (m, 3) expected")
    capi.write_stl(path, v.ctypes.data, len(v), v.dtype == np.float64, f.ctypes.data, len(f))
