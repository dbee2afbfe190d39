"""Mesh consumers: normals, welded Marching Cubes, PLY / STL writers.
The writers are host code of the library and run without a GPU; the kernels are gpu-marked."""
import struct

import numpy as np
import pytest


def _read_ply(path):
    raw = open(path, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    lines = head.decode().split("\n")
    nv = int([l for l in lines if l.startswith("element vertex")][0].split()[-1])
    nf = int([l for l in lines if l.startswith("element face")][0].split()[-1])
    props = [l.split() for l in lines if l.startswith("property") and "list" not in l]
    dt = np.dtype([(p[2], "<f8" if p[1] == "double" else "<f4") for p in props])
    v = np.frombuffer(body, dtype=dt, count=nv)
    fdt = np.dtype([("n", "u1"), ("i", "<i4", 3)])
    f = np.frombuffer(body, dtype=fdt, count=nf, offset=nv * dt.itemsize)
    assert len(body) == nv * dt.itemsize + nf * fdt.itemsize
    return v, f


def test_ply_and_stl_writers_round_trip(pkg, tmp_path):
    rng = np.random.default_rng(0)
    for vt in (np.float32, np.float64):
        v = rng.standard_normal((50, 3)).astype(vt)
        f = rng.integers(1, 51, size=(80, 3)).astype(np.int64)
        n = rng.standard_normal((50, 3)).astype(np.float32)
        p = str(tmp_path / "m.ply")
        pkg.mesh.write_ply(p, v, f, normals=n)
        pv, pf = _read_ply(p)
        assert np.array_equal(np.stack([pv["x"], pv["y"], pv["z"]], 1), v) and np.array_equal(np.stack([pv["nx"], pv["ny"], pv["nz"]], 1), n)
        assert (pf["n"] == 3).all() and np.array_equal(pf["i"], (f - 1).astype(np.int32))
        pkg.mesh.write_ply(p, v, f)
        pv, pf = _read_ply(p)
        assert pv.dtype.names == ("x", "y", "z") and np.array_equal(pf["i"], (f - 1).astype(np.int32))
        s = str(tmp_path / "m.stl")
        pkg.mesh.write_stl(s, v, f)
        raw = open(s, "rb").read()
        assert len(raw) == 84 + 50 * len(f) and struct.unpack("<I", raw[80:84])[0] == len(f)
        rec = np.frombuffer(raw, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), offset=84)
        assert np.array_equal(rec["v"], v[f - 1].astype(np.float32))
        tri = v[f - 1].astype(np.float32)
        nn = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
        ln = np.linalg.norm(nn, axis=1, keepdims=True)
        assert np.allclose(rec["n"], np.where(ln > 0, nn / np.maximum(ln, 1e-30), 0), atol=1e-5)
    with pytest.raises(pkg.capi.B200IsoError):
        pkg.mesh.write_stl(str(tmp_path / "bad.stl"), v, np.array([[1, 2, 99]], np.int64))  # index out of range
    with pytest.raises(pkg.capi.B200IsoError):
        pkg.mesh.write_ply(str(tmp_path / "no_such_dir" / "m.ply"), v, f)


def _numpy_normals(s, v, lo=-1.0, hi=1.0):
    """the kernel's formula on the host: trilinear blend of central differences (one-sided at the borders)"""
    s = s.astype(np.float64)
    n = np.array(s.shape)
    h = (hi - lo) / (n - 1)
    g = np.stack(np.gradient(s, *h, edge_order=1), -1)  # central inside, one-sided at the borders
    q = (v.astype(np.float64) - lo) / h
    q = np.where(np.abs(q - np.rint(q)) < 1e-4, np.rint(q), q)  # (within 1e-4 of a grid plane: that plane's gradients)
    q = np.clip(q, 0, n - 1)
    c = np.minimum(q.astype(np.int64), n - 2)
    t = q - c
    out = np.zeros((len(v), 3))
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                w = (t[:, 0] if dx else 1 - t[:, 0]) * (t[:, 1] if dy else 1 - t[:, 1]) * (t[:, 2] if dz else 1 - t[:, 2])
                out += w[:, None] * g[np.minimum(c[:, 0] + dx, n[0] - 1), np.minimum(c[:, 1] + dy, n[1] - 1), np.minimum(c[:, 2] + dz, n[2] - 1)]
    return out / np.linalg.norm(out, axis=1, keepdims=True)


@pytest.mark.gpu
@pytest.mark.parametrize("algo", ["MC", "MT"])
def test_vertex_normals(pkg, algo):
    s = pkg.synth.sphere((48, 40, 56))
    m = pkg.MarchingCubes(iso=pkg.Float32(0)) if algo == "MC" else pkg.MarchingTetrahedra(iso=pkg.Float32(0), eps=pkg.Float32(1e-3))
    v, f = pkg.isosurface(s, m)
    nrm = pkg.mesh.vertex_normals(s, v)
    assert nrm.shape == v.shape and nrm.dtype == np.float32
    assert np.allclose(np.linalg.norm(nrm, axis=1), 1.0, atol=1e-5)
    # a sphere's normals are radial (the field is ||p|| - r: gradient = p / ||p|| up to the grid's discretisation error)
    radial = v / np.linalg.norm(v, axis=1, keepdims=True)
    assert (np.sum(nrm * radial, axis=1) > 0.999).all()
    assert np.allclose(nrm, _numpy_normals(s, v), atol=2e-5)
    # Float64 field and vertices, other ranges, device-resident vertices
    import torch
    s64 = np.asfortranarray(s.astype(np.float64))
    v64, _ = pkg.isosurface(s64, pkg.MarchingCubes(iso=0.0), (0.0, 2.0), (0.0, 2.0), (0.0, 2.0))
    n64 = pkg.mesh.vertex_normals(s64, torch.from_numpy(v64).cuda(), (0.0, 2.0), (0.0, 2.0), (0.0, 2.0))
    assert n64.is_cuda and np.allclose(n64.cpu().numpy(), _numpy_normals(s64, v64, 0.0, 2.0), atol=2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,shape", [("sphere", (40, 44, 36)), ("gyroid", (33, 20, 47)), ("noise", (12, 9, 40))])
def test_welded_marching_cubes(pkg, kind, shape):
    s = getattr(pkg.synth, kind)(shape)
    m = pkg.MarchingCubes(iso=pkg.Float32(0))
    v, f = pkg.isosurface(s, m)
    wv, wf = pkg.mesh.isosurface_welded(s, m)
    assert wf.shape == f.shape and wf.min() == 1 and wf.max() == len(wv)
    # one vertex per crossed grid edge (sign change between two neighbouring samples, NaN-free field)
    b = s < 0
    crossed = int((b[1:] != b[:-1]).sum() + (b[:, 1:] != b[:, :-1]).sum() + (b[:, :, 1:] != b[:, :, :-1]).sum())
    assert len(wv) == crossed
    # the faces are the same triangles (copies of a shared vertex may differ in the last bit: opposite interpolation directions)
    assert np.allclose(wv[wf - 1], v[f - 1], rtol=0, atol=1e-6)
    # kept vertices are first occurrences in output order: the first index of every welded vertex increases with its number
    first = np.full(len(wv), len(v), np.int64)
    np.minimum.at(first, (wf - 1).ravel(), (f - 1).ravel())
    assert (np.diff(first) > 0).all() and np.array_equal(wv, v[first])
    if kind == "sphere":  # closed surface: V - E + F = 2
        e = np.sort(np.concatenate([wf[:, [0, 1]], wf[:, [1, 2]], wf[:, [2, 0]]]), axis=1)
        ne = len(np.unique(e, axis=0))
        assert len(wv) - ne + len(wf) == 2 and 2 * ne == 3 * len(wf)
    # device-resident field in, device-resident mesh out
    import torch
    t = torch.from_numpy(np.ascontiguousarray(s.transpose(2, 1, 0))).cuda().permute(2, 1, 0)
    dv, df = pkg.mesh.isosurface_welded(t, m)
    assert np.array_equal(dv.cpu().numpy(), wv) and np.array_equal(df.cpu().numpy(), wf)
