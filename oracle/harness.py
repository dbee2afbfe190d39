"""ctypes harness for the CPU oracle (oracle/iso_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MC, MT = 0, 1
RANGE_INT, RANGE_F32, RANGE_F64 = 0, 1, 2


def build(force=False):
    """Compile liboracle.so with the committed Makefile (g++, no reference sources involved)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "iso_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        i64, dbl, vp, ci = ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_int
        L.oracle_isosurface.restype = vp
        L.oracle_isosurface.argtypes = [ci, vp, ci, i64, i64, i64, dbl, ci, dbl, ci, dbl, dbl, dbl, dbl, dbl, dbl, ci, ci, i64, i64]
        L.oracle_isosurface_slab.restype = vp
        L.oracle_isosurface_slab.argtypes = [ci, vp, ci, i64, i64, i64, dbl, ci, dbl, ci, dbl, dbl, dbl, dbl, dbl, dbl, ci, ci, i64, i64]
        L.oracle_nverts.restype = i64
        L.oracle_nverts.argtypes = [vp]
        L.oracle_nfaces.restype = i64
        L.oracle_nfaces.argtypes = [vp]
        L.oracle_vert_is_f64.restype = ci
        L.oracle_vert_is_f64.argtypes = [vp]
        L.oracle_copy.restype = None
        L.oracle_copy.argtypes = [vp, vp, vp]
        L.oracle_free.restype = None
        L.oracle_free.argtypes = [vp]
        L.oracle_case_indices.restype = None
        L.oracle_case_indices.argtypes = [ci, vp, ci, i64, i64, i64, dbl, ci, vp]
        L.oracle_get_cubeindex_f64.restype = ci
        L.oracle_get_cubeindex_f64.argtypes = [vp, dbl]
        L.oracle_vertex_interp_f64.restype = None
        L.oracle_vertex_interp_f64.argtypes = [dbl, vp, vp, dbl, dbl, vp]
        L.oracle_linrange_f64.restype = dbl
        L.oracle_linrange_f64.argtypes = [dbl, dbl, i64, i64]
        L.oracle_linrange_f32.restype = ctypes.c_float
        L.oracle_linrange_f32.argtypes = [dbl, dbl, i64, i64]
        _LIB = L
    return _LIB


def _field(sdf):
    """Column-major (Julia) view: sdf[x, y, z] with x contiguous.  Accepts an array of shape
    (nx, ny, nz); it is laid out Fortran-contiguous before the call."""
    a = np.asarray(sdf)
    assert a.ndim == 3 and a.dtype in (np.float32, np.float64)
    return np.asfortranarray(a)


def isosurface(sdf, algo=MC, iso=0.0, iso_is_f32=False, eps=1e-3, eps_is_f32=False,
               ranges=None, range_kind=RANGE_INT, nthreads=1, xrange=None, copy=True, slab=None):
    """Oracle isosurface.  `ranges` = ((x0,x1),(y0,y1),(z0,z1)) endpoints (default (-1,1)^3).
    `xrange` = (xlo, xhi) restricts the sweep to voxel x-planes [xlo, xhi) (bounded bench sample; face
    indices are then relative to the first vertex of the range).  `copy=False` returns only the counts.
    `slab` = (x_offset, nx_global), Marching Cubes only: `sdf` holds samples [x_offset, x_offset + nx) of a volume
    with nx_global samples along x; vertices get the coordinates of the whole volume (`ranges` are the whole
    volume's), face indices are relative to the slab's first vertex.
    Returns (vertices[nv,3] float32|float64, faces[nf,3] int64 1-based)."""
    a = sdf if (isinstance(sdf, np.ndarray) and sdf.flags.f_contiguous) else _field(sdf)
    xlo, xhi = xrange if xrange is not None else (-1, -1)
    nx, ny, nz = a.shape
    if ranges is None:
        ranges = ((-1.0, 1.0),) * 3
    (x0, x1), (y0, y1), (z0, z1) = ranges
    L = lib()
    if slab is not None:
        assert algo == MC and xrange is None
        h = L.oracle_isosurface_slab(algo, a.ctypes.data, int(a.dtype == np.float64), nx, ny, nz, float(iso), int(iso_is_f32),
                                     float(eps), int(eps_is_f32), x0, x1, y0, y1, z0, z1, range_kind, nthreads, int(slab[0]), int(slab[1]))
    else:
        h = L.oracle_isosurface(algo, a.ctypes.data, int(a.dtype == np.float64), nx, ny, nz, float(iso), int(iso_is_f32),
                                float(eps), int(eps_is_f32), x0, x1, y0, y1, z0, z1, range_kind, nthreads, xlo, xhi)
    try:
        nv, nf = L.oracle_nverts(h), L.oracle_nfaces(h)
        if not copy:
            return nv, nf
        vt = np.float64 if L.oracle_vert_is_f64(h) else np.float32
        verts = np.empty((nv, 3), dtype=vt)
        faces = np.empty((nf, 3), dtype=np.int64)
        L.oracle_copy(h, verts.ctypes.data, faces.ctypes.data)
    finally:
        L.oracle_free(h)
    return verts, faces


def case_indices(sdf, algo=MC, iso=0.0, iso_is_f32=False):
    a = _field(sdf)
    nx, ny, nz = a.shape
    out = np.empty(max(nx - 1, 0) * max(ny - 1, 0) * max(nz - 1, 0), dtype=np.uint8)
    lib().oracle_case_indices(algo, a.ctypes.data, int(a.dtype == np.float64), nx, ny, nz, float(iso), int(iso_is_f32),
                              out.ctypes.data)
    return out


def get_cubeindex(vals, iso):
    v = np.ascontiguousarray(vals, dtype=np.float64)
    return lib().oracle_get_cubeindex_f64(v.ctypes.data, float(iso))


def vertex_interp(iso, p1, p2, v1, v2):
    a = np.ascontiguousarray(p1, dtype=np.float64)
    b = np.ascontiguousarray(p2, dtype=np.float64)
    out = np.empty(3)
    lib().oracle_vertex_interp_f64(float(iso), a.ctypes.data, b.ctypes.data, float(v1), float(v2), out.ctypes.data)
    return out
