"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.
Bar (BASELINE.json north_star): faces, vertex count/order and per-voxel case indices bit-exact; Float32
vertex coordinates within 1 ulp -- these tests demand bit-exact coordinates as well (tolerance 0 ulp)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ALGOS = {"MC": 0, "MT": 1}


def _method(pkg, algo, iso, f32, eps=1e-3):
    conv = pkg.Float32 if f32 else float
    if algo == "MC":
        return pkg.MarchingCubes(iso=conv(iso))
    return pkg.MarchingTetrahedra(iso=conv(iso), eps=conv(eps))


def _bits_equal(a, b):
    """Bit-exact equality of float arrays, NaNs compared by NaN-ness."""
    if a.dtype != b.dtype or a.shape != b.shape:
        return False
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return False
    u = np.uint32 if a.dtype == np.float32 else np.uint64
    return np.array_equal(a.view(u)[~na], b.view(u)[~nb])


def _check(pkg, oracle, s, algo, iso=0.0, f32=True, ranges=None, rk=None, eps=1e-3):
    rk = oracle.RANGE_INT if rk is None else rk
    conv = {oracle.RANGE_INT: int, oracle.RANGE_F32: np.float32, oracle.RANGE_F64: float}[rk]
    rr = ranges or ((-1, 1),) * 3
    args = [tuple(conv(e) for e in r) for r in rr]
    v, f = pkg.isosurface(s, _method(pkg, algo, iso, f32, eps), *args)
    vo, fo = oracle.isosurface(s, ALGOS[algo], iso=iso, iso_is_f32=f32, eps=eps, eps_is_f32=f32, ranges=rr, range_kind=rk)
    assert v.shape == vo.shape and f.shape == fo.shape, (v.shape, vo.shape, f.shape, fo.shape)
    assert v.dtype == vo.dtype
    assert np.array_equal(f, fo), "faces differ"
    assert _bits_equal(v, vo), "vertex coordinates differ (0 ulp demanded)"
    return v, f


@pytest.mark.parametrize("algo", ["MC", "MT"])
@pytest.mark.parametrize("shape", [(16, 16, 16), (33, 20, 47), (64, 64, 64), (5, 130, 37), (129, 7, 70), (12, 9, 260)])
def test_parity_shapes(pkg, oracle, algo, shape):
    _check(pkg, oracle, pkg.synth.gyroid(shape), algo)
    _check(pkg, oracle, pkg.synth.sphere(shape), algo)


@pytest.mark.parametrize("algo", ["MC", "MT"])
def test_parity_dense_noise(pkg, oracle, algo):
    """Worst case: nearly every voxel active, every one of the 254 cases occurs, multi-round blocks."""
    s = pkg.synth.noise((40, 37, 150), seed=3)
    _check(pkg, oracle, s, algo)
    c = pkg.api.case_indices(s, _method(pkg, algo, 0.0, True))
    assert np.array_equal(c, oracle.case_indices(s, ALGOS[algo], iso_is_f32=True))
    assert len(np.unique(c)) == 256


@pytest.mark.parametrize("algo", ["MC", "MT"])
@pytest.mark.parametrize("f32,rk", [(True, 0), (True, 1), (True, 2), (False, 0), (False, 1), (False, 2)])
def test_parity_type_combinations(pkg, oracle, algo, f32, rk):
    """typeof(iso) x eltype(X): the arithmetic modes of vertex_interp / vertPos."""
    s = pkg.synth.gyroid((30, 41, 36))
    _check(pkg, oracle, s, algo, iso=0.1, f32=f32, rk=rk, ranges=((-2, 3), (0, 1), (-7, -1)))
    s2 = pkg.synth.noise((20, 21, 40), seed=11)
    _check(pkg, oracle, s2, algo, iso=-0.3, f32=f32, rk=rk, ranges=((-2, 3), (0, 1), (-7, -1)))


@pytest.mark.parametrize("algo", ["MC", "MT"])
def test_case_indices_bit_exact(pkg, oracle, algo):
    s = pkg.synth.gyroid((70, 50, 90))
    for iso, f32 in ((0.0, True), (0.25, False), (1e-9, False)):
        c = pkg.api.case_indices(s, _method(pkg, algo, iso, f32))
        assert np.array_equal(c, oracle.case_indices(s, ALGOS[algo], iso=iso, iso_is_f32=f32))


def test_float64_iso_threshold_edge(pkg, oracle):
    """Float32 sample vs Float64 iso compares exactly (strict <): samples equal to / adjacent to iso."""
    base = np.float32(0.1)
    s = np.full((6, 6, 6), base, np.float32)
    s[::2, 1::2, ::3] = np.nextafter(base, np.float32(1))
    s[1::2, ::2, 1::3] = np.nextafter(base, np.float32(-1))
    for iso in (float(base), 0.1, float(np.nextafter(base, np.float32(1))), float(base) + 1e-12):
        for algo in ("MC", "MT"):
            c = pkg.api.case_indices(s, _method(pkg, algo, iso, False))
            assert np.array_equal(c, oracle.case_indices(s, ALGOS[algo], iso=iso, iso_is_f32=False)), iso


@pytest.mark.parametrize("algo", ["MC", "MT"])
def test_degenerate_and_tiny(pkg, oracle, algo):
    for shape in [(1, 5, 5), (5, 1, 5), (5, 5, 1), (1, 1, 1), (2, 2, 2), (2, 3, 2), (3, 2, 33)]:
        s = pkg.synth.noise(shape, seed=5)
        _check(pkg, oracle, s, algo)
    s = np.full((4, 4, 4), -1.0, np.float32)  # no surface
    v, f = _check(pkg, oracle, s, algo)
    assert len(v) == 0 and len(f) == 0


@pytest.mark.parametrize("algo", ["MC", "MT"])
def test_nan_and_inf_samples(pkg, oracle, algo):
    s = pkg.synth.gyroid((12, 12, 12)).copy(order="F")
    s[3, 4, 5] = np.nan
    s[7, 7, 7] = np.inf
    s[2, 9, 4] = -np.inf
    _check(pkg, oracle, s, algo)
    c = pkg.api.case_indices(s, _method(pkg, algo, 0.0, True))
    assert np.array_equal(c, oracle.case_indices(s, ALGOS[algo], iso_is_f32=True))


def test_reference_kat_through_gpu(pkg):
    """test/runtests.jl:15-33 through the CUDA path: cube index of a single voxel, and vertex_interp on a
    2x2x2 field whose only sign change reproduces (0, .5, 0)-style interpolation."""
    def voxel(vals):  # MC corner order -> field[x, y, z]
        s = np.empty((2, 2, 2), np.float32)
        for k, (x, y, z) in enumerate([(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]):
            s[x, y, z] = vals[k]
        return s
    f = np.float32
    assert pkg.api.case_indices(voxel([.1, .2, .3, .4, .5, .6, .7, .8]), pkg.MarchingCubes(iso=0.35))[0] == 0x07
    assert pkg.api.case_indices(voxel([.5, .6, .7, .8, .9, 1., 1.1, 1.2]), pkg.MarchingCubes(iso=0.75))[0] == 0x07
    assert pkg.api.case_indices(voxel([.9, .8, .7, .6, .5, .4, .3, .2]), pkg.MarchingCubes(iso=0.5))[0] == 0xE0
    # one corner inside: vertices on edges 1, 9, 4 at mu = 0.5 of a [0,1]^3 voxel
    s = voxel([-1, 1, 1, 1, 1, 1, 1, 1])
    v, fc = pkg.isosurface(s, pkg.MarchingCubes(iso=f(0)), (0, 1), (0, 1), (0, 1))
    assert fc.tolist() == [[3, 2, 1]]
    assert v.tolist() == [[0.5, 0, 0], [0, 0, 0.5], [0, 0.5, 0]]


def test_defaults_forwarder(pkg):
    """isosurface(A) == isosurface(A, MarchingCubes())  (test/runtests.jl:76-79)"""
    s = pkg.synth.noise((10, 10, 10), seed=1)
    v1, f1 = pkg.isosurface(s)
    v2, f2 = pkg.isosurface(s, pkg.MarchingCubes())
    assert v1.dtype == np.float64 and np.array_equal(v1, v2) and np.array_equal(f1, f2)


@pytest.mark.parametrize("algo", ["MC", "MT"])
def test_respect_origin_float32_field(pkg, oracle, algo):
    """test/runtests.jl:127-149 on a Float32 copy of norm_sdf (the GPU path is Float32-field only)."""
    g = np.arange(-100, 101) / 100.0
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    s = np.asfortranarray(np.sqrt(X * X + Y * Y + Z * Z).astype(np.float32))
    v, f = _check(pkg, oracle, s, algo, iso=0.5, f32=False)
    assert np.allclose(v.mean(0), 0, atol=0.015)
    assert np.allclose(v.max(0), 0.5, atol=1e-3) and np.allclose(v.min(0), -0.5, atol=1e-3)


@pytest.mark.parametrize("algo", ["MC", "MT"])
def test_noisy_spheres_float32(pkg, oracle, algo):
    """Input of test/runtests.jl:152-172 rounded to Float32 (its Float64 form pins the oracle in
    tests/test_oracle.py); MT result must be a closed manifold: V - F/2 == 2."""
    import os
    field = np.load(os.path.join(os.path.dirname(__file__), "golden", "noisy_spheres_input.npy"))
    s = np.asfortranarray(field.astype(np.float32))
    v, f = _check(pkg, oracle, s, algo, iso=8.0, f32=False)
    if algo == "MT":
        assert len(v) - len(f) // 2 == 2


def test_device_resident_torch_path(pkg, oracle):
    import torch
    s = pkg.synth.gyroid((65, 40, 50))
    for ldx in (65, 68):  # unaligned rows -> scalar-load classify; padded rows -> 128-bit loads
        store = torch.zeros((50, 40, ldx), dtype=torch.float32, device="cuda")
        store[:, :, :65] = torch.from_numpy(np.ascontiguousarray(s.transpose(2, 1, 0))).cuda()
        t = store.permute(2, 1, 0)[:65]
        for algo in ("MC", "MT"):
            v, f = pkg.isosurface(t, _method(pkg, algo, 0.0, True))
            vo, fo = oracle.isosurface(s, ALGOS[algo], iso_is_f32=True, eps_is_f32=True)
            assert np.array_equal(f.cpu().numpy(), fo) and _bits_equal(v.cpu().numpy(), vo)


def test_gyroid_synth_device_equals_host(pkg):
    import torch
    shape = (40, 33, 37)
    h = pkg.synth.gyroid(shape)
    d = pkg.synth.gyroid_torch(shape, "cuda", ldx=40)
    assert np.array_equal(d.cpu().numpy(), h)
    m = pkg.synth.multisphere_torus(shape)
    md = pkg.synth.multisphere_torus(shape, xp=torch, device="cuda")
    assert np.array_equal(md.cpu().numpy(), m)


def test_async_capacity_and_vertex_base(pkg, oracle):
    """Async device form: totals on the device, capacity guard, face index base (sharding hook)."""
    import torch
    s = pkg.synth.gyroid((48, 40, 56))
    vo, fo = oracle.isosurface(s, 0, iso_is_f32=True)
    t = torch.from_numpy(np.ascontiguousarray(s.transpose(2, 1, 0))).cuda().permute(2, 1, 0)
    h = pkg.capi.Handle(0)
    p = pkg.api.make_params(pkg.MarchingCubes(iso=pkg.Float32(0)))
    totals = torch.zeros(2, dtype=torch.int64, device="cuda")
    base = torch.tensor([1000], dtype=torch.int64, device="cuda")
    h.set_stream(torch.cuda.current_stream().cuda_stream)
    h.count_async(p, t.data_ptr(), 48, 40, 56, 48, totals.data_ptr())
    verts = torch.full((len(vo) + 7, 3), -7.0, dtype=torch.float32, device="cuda")
    faces = torch.full((len(fo) - 5, 3), -7, dtype=torch.int64, device="cuda")  # too small on purpose
    h.generate_async(verts.data_ptr(), verts.shape[0], faces.data_ptr(), faces.shape[0], base.data_ptr(), 234)
    torch.cuda.synchronize()
    assert totals.tolist() == [len(vo), len(fo)]
    assert h.totals()[:2] == (len(vo), len(fo))
    assert _bits_equal(verts[: len(vo)].cpu().numpy(), vo) and (verts[len(vo):] == -7).all()
    assert np.array_equal(faces.cpu().numpy(), fo[: len(fo) - 5] + 1234)
    h.close()


@pytest.mark.parametrize("algo,n", [("MC", 256), ("MT", 192)])
def test_parity_medium_gyroid(pkg, oracle, algo, n):
    _check(pkg, oracle, pkg.synth.gyroid(n), algo)


def test_parity_config1_sphere128(pkg, oracle):
    """BASELINE.json configs[0]: MC on the 128^3 sphere SDF."""
    v, f = _check(pkg, oracle, pkg.synth.sphere(128), "MC")
    assert v.shape == (76032, 3) and f.shape == (38012, 3)


@pytest.mark.parametrize("algo", ["MC", "MT"])
def test_parity_config_512_gyroid(pkg, oracle, algo):
    """BASELINE.json configs[1], configs[2]: 512^3 gyroid, full oracle comparison."""
    s = pkg.synth.gyroid(512)
    v, f = _check(pkg, oracle, s, algo)
    if algo == "MC":
        assert abs(len(v) - 10123197) < 2000 and abs(len(f) - 5061727) < 1000  # SURVEY.md Appendix C estimates


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_slabs_equal_unsharded(pkg, oracle, world):
    """N x-slabs (each with its halo plane, run through the GPU one after the other like N ranks) stitched
    with the exclusive prefix of their counts == the unsharded mesh, byte for byte."""
    shape = (70, 33, 45)
    s = pkg.synth.gyroid(shape)
    m = pkg.MarchingCubes(iso=pkg.Float32(0.05))
    X, Y, Z = (0, 3), (-1, 1), (2, 5)
    v1, f1 = pkg.isosurface(s, m, X, Y, Z)
    vo, fo = oracle.isosurface(s, 0, iso=0.05, iso_is_f32=True, ranges=(X, Y, Z))
    assert np.array_equal(f1, fo) and _bits_equal(v1, vo)
    counts, slabs = [], []
    for r in range(world):
        xa, xb = pkg.sharding.slab_bounds(shape[0], world, r)
        slabs.append((xa, xb))
        _, nv, nf, _ = pkg.api.slab_count(s[xa:xb], m, xa, shape[0], X, Y, Z)
        counts.append((nv, nf))
    parts = []
    for r, (xa, xb) in enumerate(slabs):
        vb, _ = pkg.sharding.exclusive_bases(counts, r)
        parts.append(pkg.api.isosurface_slab(s[xa:xb], m, xa, shape[0], vb, X, Y, Z))
    v, f = pkg.sharding.stitch(parts)
    assert np.array_equal(f, f1) and _bits_equal(v, v1)


@pytest.mark.parametrize("algo", ["MC", "MT"])
@pytest.mark.parametrize("shape", [(9, 40, 1100), (6, 70, 2200), (4, 5, 6000)])
def test_tall_grids(pkg, oracle, algo, shape):
    """nz > 1024: several quad-cells per thread; nz = 6000: bit-field too tall to stage in shared memory."""
    _check(pkg, oracle, pkg.synth.gyroid(shape), algo)
    _check(pkg, oracle, pkg.synth.noise(shape, seed=9), algo, iso=0.9)


@pytest.mark.parametrize("shape,kind", [((48, 40, 56), "gyroid"), ((33, 21, 300), "noise"), ((20, 130, 37), "sphere"), ((3, 3, 3), "noise"),
                                        ((300, 24, 70), "gyroid"), ((129, 9, 33), "noise"), ((515, 6, 40), "gyroid")])
def test_one_enqueue_extract(pkg, oracle, shape, kind):
    """b200iso_extract_async: classify, count + scan and generate in one enqueue into buffers with a capacity."""
    import torch
    s = getattr(pkg.synth, kind)(shape)
    vo, fo = oracle.isosurface(s, 0, iso_is_f32=True)
    nx, ny, nz = shape
    t = torch.from_numpy(np.ascontiguousarray(s.transpose(2, 1, 0))).cuda().permute(2, 1, 0)
    h = pkg.capi.Handle(0)
    p = pkg.api.make_params(pkg.MarchingCubes(iso=pkg.Float32(0)))
    totals = torch.zeros(2, dtype=torch.int64, device="cuda")
    verts = torch.full((len(vo) + 3, 3), -7.0, dtype=torch.float32, device="cuda")
    faces = torch.full((len(fo) + 3, 3), -7, dtype=torch.int64, device="cuda")
    h.set_stream(torch.cuda.current_stream().cuda_stream)
    for _ in range(2):  # twice: the scan state must be reset between calls
        h.extract_async(p, t.data_ptr(), nx, ny, nz, nx, verts.data_ptr(), verts.shape[0], faces.data_ptr(), faces.shape[0], 0, 0,
                        totals.data_ptr())
    torch.cuda.synchronize()
    assert totals.tolist() == [len(vo), len(fo)] and h.totals()[:2] == (len(vo), len(fo))
    assert _bits_equal(verts[: len(vo)].cpu().numpy(), vo) and (verts[len(vo):] == -7).all()
    assert np.array_equal(faces[: len(fo)].cpu().numpy(), fo) and (faces[len(fo):] == -7).all()
    # too-small buffers: totals are still exact, nothing is written past the capacity
    if len(fo) > 4:
        small = torch.full((len(fo) - 4, 3), -7, dtype=torch.int64, device="cuda")
        h.extract_async(p, t.data_ptr(), nx, ny, nz, nx, verts.data_ptr(), 2, small.data_ptr(), small.shape[0], 0, 0, totals.data_ptr())
        torch.cuda.synchronize()
        assert totals.tolist() == [len(vo), len(fo)]
        assert np.array_equal(small.cpu().numpy(), fo[: len(fo) - 4])
    # sharded fix-up: faces += base on the device
    base = torch.tensor([4321], dtype=torch.int64, device="cuda")
    h.extract_async(p, t.data_ptr(), nx, ny, nz, nx, verts.data_ptr(), verts.shape[0], faces.data_ptr(), faces.shape[0], 0, 0, totals.data_ptr())
    h.add_vertex_base_async(faces.data_ptr(), faces.shape[0], totals.data_ptr(), base.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(faces[: len(fo)].cpu().numpy(), fo + 4321) and (faces[len(fo):] == -7).all()
    h.close()


def test_one_enqueue_extract_mt(pkg, oracle):
    import torch
    s = pkg.synth.gyroid((30, 31, 64))
    vo, fo = oracle.isosurface(s, 1, iso_is_f32=True, eps_is_f32=True)
    t = torch.from_numpy(np.ascontiguousarray(s.transpose(2, 1, 0))).cuda().permute(2, 1, 0)
    h = pkg.capi.Handle(0)
    p = pkg.api.make_params(pkg.MarchingTetrahedra(iso=pkg.Float32(0), eps=pkg.Float32(1e-3)))
    verts = torch.empty((len(vo), 3), dtype=torch.float32, device="cuda")
    faces = torch.empty((len(fo), 3), dtype=torch.int64, device="cuda")
    h.set_stream(torch.cuda.current_stream().cuda_stream)
    h.extract_async(p, t.data_ptr(), 30, 31, 64, 30, verts.data_ptr(), len(vo), faces.data_ptr(), len(fo))
    torch.cuda.synchronize()
    assert h.totals()[:2] == (len(vo), len(fo))
    assert _bits_equal(verts.cpu().numpy(), vo) and np.array_equal(faces.cpu().numpy(), fo)
    h.close()


def test_full_size_1024_gyroid_properties(pkg, oracle):
    """BASELINE.json configs[3] at its full size (1024^3 gyroid, MC, Float32 vertices), checked through
    size-independent properties: (1) the totals equal the expectation of SURVEY.md Appendix C, (2) structural
    invariants of the face array, (3) the 8-slab sharded extraction equals the one-shot extraction byte for byte
    (a checksum of checksums over slabs), (4) the oracle agrees on sampled x-ranges of the volume."""
    import torch
    n = 1024
    t = pkg.synth.gyroid_torch(n, "cuda")
    m = pkg.MarchingCubes(iso=pkg.Float32(0))
    v, f = pkg.isosurface(t, m)
    assert pkg.api.get_handle(0).classify_path() == pkg.capi.CLASSIFY_TMA  # the kernel the benchmark times
    assert v.shape == (40573677, 3) and f.shape == (20286967, 3) and v.dtype == torch.float32
    assert int(f.min()) == 1 and int(f.max()) == v.shape[0]
    assert bool(torch.isfinite(v).all()) and float(v.min()) >= -1.0 and float(v.max()) <= 1.0
    # every voxel block starts with the face (fct+3, fct+2, fct+1): column 0 - column 2 == 2 exactly there
    assert int(((f[:, 0] - f[:, 2]) == 2).sum()) >= 10143355  # >= number of active voxels (Appendix C)
    # (3) sharded == unsharded
    counts, slabs = [], []
    for r in range(8):
        xa, xb = pkg.sharding.slab_bounds(n, 8, r)
        slabs.append((xa, xb))
        _, nv, nf, _ = pkg.api.slab_count(t[xa:xb], m, xa, n)
        counts.append((nv, nf))
    assert sum(c[0] for c in counts) == v.shape[0] and sum(c[1] for c in counts) == f.shape[0]
    vo = fo = 0
    for r, (xa, xb) in enumerate(slabs):
        vb, fb = pkg.sharding.exclusive_bases(counts, r)
        assert (vb, fb) == (vo, fo)
        sv, sf = pkg.api.isosurface_slab(t[xa:xb], m, xa, n, vb)
        assert torch.equal(sv, v[vo: vo + sv.shape[0]]) and torch.equal(sf, f[fo: fo + sf.shape[0]])
        vo += sv.shape[0]
        fo += sf.shape[0]
    # (4) the oracle's sweep of three x-ranges of the SAME host field, compared bit for bit with the corresponding
    # slices of the whole-volume result above -- i.e. with what the benchmarked kernels (TMA classify included) wrote
    host = t.cpu().numpy()
    assert host.flags.f_contiguous
    vh, fh = v.cpu().numpy(), f.cpu().numpy()
    for xlo in (0, 517, n - 7):
        xhi = xlo + 6
        ov, of = oracle.isosurface(host, 0, iso_is_f32=True, xrange=(xlo, xhi))
        v0 = f0 = 0
        if xlo > 0:  # vertices / faces of the voxel planes below the range
            _, v0, f0, _ = pkg.api.slab_count(t[0:xlo + 1], m, 0, n)
        assert len(ov) > 10000
        assert np.array_equal(fh[f0:f0 + len(of)], of + v0), xlo
        assert _bits_equal(vh[v0:v0 + len(ov)], ov), xlo
        if xhi == n - 1:
            assert v0 + len(ov) == len(vh) and f0 + len(of) == len(fh)


# ---- Float64 fields: the reference's own test inputs run through the CUDA path unchanged ----
def test_reference_noisy_spheres_golden_counts_on_gpu(pkg, oracle):
    """test/runtests.jl:152-172, the reference's only exact-count test, THROUGH THE CUDA PATH: Float64 field
    (Float32 distances + Julia's MersenneTwister(0) noise), MarchingTetrahedra(iso=8.0): 3466 vertices, 6928 faces."""
    import os
    field = np.load(os.path.join(os.path.dirname(__file__), "golden", "noisy_spheres_input.npy"))
    assert field.dtype == np.float64
    points, faces = pkg.isosurface(field, pkg.MarchingTetrahedra(iso=8.0))
    assert len(points) == 3466 and len(faces) == 6928 and points.dtype == np.float64
    vo, fo = oracle.isosurface(field, oracle.MT, iso=8.0)
    assert np.array_equal(faces, fo) and _bits_equal(points, vo)
    pm, fm = pkg.isosurface(field, pkg.MarchingCubes(iso=8.0))
    vo, fo = oracle.isosurface(field, oracle.MC, iso=8.0)
    assert np.array_equal(fm, fo) and _bits_equal(pm, vo)


@pytest.mark.parametrize("algo", ["MC", "MT"])
def test_reference_respect_origin_float64_on_gpu(pkg, oracle, algo):
    """test/runtests.jl:127-149 with its own Float64 norm_sdf and iso = 0.5, through the CUDA path."""
    g = np.arange(-100, 101) / 100.0
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    norm_sdf = np.asfortranarray(np.sqrt(X * X + Y * Y + Z * Z))
    v, f = _check(pkg, oracle, norm_sdf, algo, iso=0.5, f32=False)
    assert v.dtype == np.float64 and len(v) == (187800 if algo == "MC" else 140714)
    assert np.allclose(v.mean(0), 0, atol=0.015)
    assert np.allclose(v.max(0), 0.5, atol=1e-3) and np.allclose(v.min(0), -0.5, atol=1e-3)


@pytest.mark.parametrize("algo", ["MC", "MT"])
@pytest.mark.parametrize("f32,rk", [(True, 0), (False, 0), (True, 1), (False, 2)])
def test_float64_field_type_combinations(pkg, oracle, algo, f32, rk):
    rng = np.random.default_rng(5)
    s = np.asfortranarray(rng.standard_normal((23, 30, 41)))
    _check(pkg, oracle, s, algo, iso=0.2, f32=f32, rk=rk, ranges=((-2, 3), (0, 1), (-7, -1)))
    c = pkg.api.case_indices(s, _method(pkg, algo, 0.2, f32))
    assert np.array_equal(c, oracle.case_indices(s, ALGOS[algo], iso=0.2, iso_is_f32=f32))


def test_mixed_types_float64_vs_float32_on_gpu(pkg):
    """test/runtests.jl:81-97: same face topology for a Float64 field and its Float32 copy (MT, ranges 0:1)."""
    s = np.asfortranarray(np.random.default_rng(7).standard_normal((10, 10, 10)))
    p1, f1 = pkg.isosurface(s, pkg.MarchingTetrahedra(), (0, 1), (0, 1), (0, 1))
    p2, f2 = pkg.isosurface(s.astype(np.float32), pkg.MarchingTetrahedra(), (0, 1), (0, 1), (0, 1))
    assert len(p1) == len(p2) and np.array_equal(f1, f2)


@pytest.mark.parametrize("world", [2, 3, 5])
@pytest.mark.parametrize("kind", ["gyroid", "noise"])
def test_sharded_mt_slabs_equal_unsharded(pkg, oracle, world, kind):
    """Marching Tetrahedra x-slabs with a ghost voxel row: stitched == unsharded, byte for byte (shared vertices
    across the slab boundary keep their global ids)."""
    shape = (41, 19, 70)
    s = getattr(pkg.synth, kind)(shape)
    m = pkg.MarchingTetrahedra(iso=pkg.Float32(0.05), eps=pkg.Float32(1e-3))
    X, Y, Z = (0, 3), (-1, 1), (2, 5)
    v1, f1 = pkg.isosurface(s, m, X, Y, Z)
    vo, fo = oracle.isosurface(s, 1, iso=0.05, iso_is_f32=True, eps_is_f32=True, ranges=(X, Y, Z))
    assert np.array_equal(f1, fo) and _bits_equal(v1, vo)
    counts, slabs = [], []
    for r in range(world):
        xa, xb = pkg.sharding.slab_bounds(shape[0], world, r, ghost=True)
        slabs.append((xa, xb))
        _, nv, nf, _ = pkg.api.slab_count(s[xa:xb], m, xa, shape[0], X, Y, Z)
        counts.append((nv, nf))
    assert sum(c[0] for c in counts) == len(v1) and sum(c[1] for c in counts) == len(f1)
    parts = []
    for r, (xa, xb) in enumerate(slabs):
        vb, _ = pkg.sharding.exclusive_bases(counts, r)
        parts.append(pkg.api.isosurface_slab(s[xa:xb], m, xa, shape[0], vb, X, Y, Z))
    v, f = pkg.sharding.stitch(parts)
    assert np.array_equal(f, f1) and _bits_equal(v, v1)


# ---- one-shot host form: b200iso_extract_host (x-slab pipelined H2D || kernels || D2H) -------------------------
@pytest.mark.parametrize("algo", ["MC", "MT"])
@pytest.mark.parametrize("shape,kind,slabs", [((70, 33, 41), "gyroid", 1), ((70, 33, 41), "gyroid", 3), ((70, 33, 41), "noise", 7),
                                              ((131, 20, 37), "sphere", 16), ((9, 12, 300), "noise", 64), ((2, 9, 9), "noise", 5)])
def test_extract_host_equals_two_phase(pkg, oracle, monkeypatch, algo, shape, kind, slabs):
    """Every slab count gives the bytes of count + generate, which the tests above pin to the oracle."""
    s = getattr(pkg.synth, kind)(shape)
    m = _method(pkg, algo, 0.0, True)
    v0, f0 = pkg.api.isosurface_two_phase(s, m, (0, 2), (-1, 1), (3, 4))
    monkeypatch.setenv("B200ISO_HOST_SLABS", str(slabs))
    v1, f1 = pkg.isosurface(s, m, (0, 2), (-1, 1), (3, 4), capacity=(len(v0) + 5, len(f0) + 3))
    assert np.array_equal(f1, f0) and _bits_equal(v1, v0)
    vo, fo = oracle.isosurface(s, ALGOS[algo], iso=0.0, iso_is_f32=True, eps_is_f32=True, ranges=((0, 2), (-1, 1), (3, 4)))
    assert np.array_equal(f1, fo) and _bits_equal(v1, vo)


@pytest.mark.parametrize("algo", ["MC", "MT"])
def test_extract_host_natural_slabs_and_types(pkg, oracle, algo):
    """A field big enough (> 32 MB) for the default slab split; Float64 method => Float64 vertices; Float64 field."""
    s = pkg.synth.gyroid((530, 130, 140))
    for f32 in (True, False):
        m = _method(pkg, algo, 0.05, f32)
        v0, f0 = pkg.api.isosurface_two_phase(s, m)
        v1, f1 = pkg.isosurface(s, m, capacity=(len(v0), len(f0)))
        assert v1.dtype == (np.float32 if f32 else np.float64)
        assert np.array_equal(f1, f0) and _bits_equal(v1, v0)
    s64 = pkg.synth.gyroid((150, 190, 160)).astype(np.float64) * 1.000000001
    m = _method(pkg, algo, 0.0, False)
    v0, f0 = pkg.api.isosurface_two_phase(s64, m)
    v1, f1 = pkg.isosurface(s64, m, capacity=(len(v0), len(f0)))
    assert np.array_equal(f1, f0) and _bits_equal(v1, v0)
    v1, f1 = pkg.isosurface(s64, m)  # built-in guess (memo of the call above)
    assert np.array_equal(f1, f0) and _bits_equal(v1, v0)


def test_extract_host_capacity_protocol(pkg, monkeypatch):
    """Too small => B200ISO_ECAPACITY with the true totals; vcap = fcap = 0 is a pure count; isosurface(capacity=)
    re-runs exactly; isosurface_into writes into caller arrays."""
    s = pkg.synth.gyroid((90, 40, 40))
    m = pkg.MarchingCubes(iso=pkg.Float32(0))
    v0, f0 = pkg.isosurface(s, m)
    monkeypatch.setenv("B200ISO_HOST_SLABS", "4")
    h = pkg.api.get_handle(0)
    p = pkg.api.make_params(m)
    a = np.asfortranarray(s)
    nv, nf, f64, fits = h.extract_host(p, a.ctypes.data, *a.shape, a.shape[0], 0, 0, 0, 0)
    assert (nv, nf, f64, fits) == (len(v0), len(f0), False, False)
    vb, fb = np.empty((nv // 2, 3), np.float32), np.empty((nf, 3), np.int64)
    nv2, nf2, _, fits = h.extract_host(p, a.ctypes.data, *a.shape, a.shape[0], vb.ctypes.data, len(vb), fb.ctypes.data, len(fb))
    assert (nv2, nf2, fits) == (nv, nf, False)
    v1, f1 = pkg.isosurface(s, m, capacity=(10, 10))  # too small: exact arrays filled from the resident slabs
    assert np.array_equal(f1, f0) and _bits_equal(v1, v0)
    # the resident form by hand, and its state rules
    nv3, nf3, _, fits = h.extract_host(p, a.ctypes.data, *a.shape, a.shape[0], 0, 0, 0, 0)
    assert not fits
    vb, fb = np.full((nv3 + 2, 3), -1, np.float32), np.full((nf3 + 2, 3), -1, np.int64)
    assert h.extract_host_resident(vb.ctypes.data, len(vb), fb.ctypes.data, len(fb)) == (nv, nf, True)
    assert np.array_equal(fb[:nf], f0) and _bits_equal(vb[:nv], v0) and (fb[nf:] == -1).all()
    assert h.extract_host_resident(vb.ctypes.data, 5, fb.ctypes.data, len(fb)) == (nv, nf, False)  # still too small: still resident
    h.count(p, a.ctypes.data, pkg.capi.HOST, *a.shape, a.shape[0])  # another host call overwrites the staging
    with pytest.raises(pkg.capi.B200IsoError):
        h.extract_host_resident(vb.ctypes.data, len(vb), fb.ctypes.data, len(fb))
    vb, fb = np.full((nv + 7, 3), -1, np.float32), np.full((nf + 7, 3), -1, np.int64)
    assert pkg.api.isosurface_into(a, vb, fb, m) == (nv, nf)
    assert np.array_equal(fb[:nf], f0) and _bits_equal(vb[:nv], v0)
    assert (fb[nf:] == -1).all() and (vb[nv:] == -1).all()
    with pytest.raises(pkg.capi.B200IsoError):
        pkg.api.isosurface_into(a, vb[:10], fb, m)
    # the two-phase pair still works on the same handle afterwards
    v2, f2 = pkg.isosurface(s, m)
    assert np.array_equal(f2, f0) and _bits_equal(v2, v0)


@pytest.mark.parametrize("algo", ["MC", "MT"])
def test_extract_host_on_a_sharded_slab(pkg, algo):
    """A rank's slab (x_offset / nx_global / MT ghost row) through the one-shot call: slab-local indices."""
    s = pkg.synth.gyroid((120, 30, 34))
    m = _method(pkg, algo, 0.0, True)
    xa, xb = pkg.sharding.slab_bounds(120, 3, 1, ghost=(algo == "MT"))
    slab = np.asfortranarray(s[xa:xb])
    v0, f0 = pkg.api.isosurface_slab(slab, m, xa, 120, 0)
    p = pkg.api.make_params(m)
    p.x_offset, p.nx_global, p.x_ghost = xa, 120, int(algo == "MT")
    import os
    os.environ["B200ISO_HOST_SLABS"] = "3"
    try:
        vb, fb = np.empty((len(v0), 3), np.float32), np.empty((len(f0), 3), np.int64)
        nv, nf, _, fits = pkg.api.get_handle(0).extract_host(p, slab.ctypes.data, *slab.shape, slab.shape[0], vb.ctypes.data, len(vb), fb.ctypes.data, len(fb))
    finally:
        del os.environ["B200ISO_HOST_SLABS"]
    assert fits and (nv, nf) == (len(v0), len(f0))
    assert np.array_equal(fb, f0) and _bits_equal(vb, v0)


@pytest.mark.parametrize("slabs", [1, 2])
def test_extract_host_rows_longer_than_a_staging_chunk(pkg, oracle, monkeypatch, slabs):
    """A pageable field whose rows exceed one pinned staging chunk goes up in segments (the chunk is shrunk to 64 KB so
    that 80 KB rows do what 2.4 MB rows do by default)."""
    s = pkg.synth.noise((20011, 5, 4))
    m = _method(pkg, "MC", 0.0, True)
    vo, fo = oracle.isosurface(s, ALGOS["MC"], iso=0.0, iso_is_f32=True, eps_is_f32=True)
    monkeypatch.setenv("B200ISO_HOST_CHUNK_KB", "64")
    monkeypatch.setenv("B200ISO_HOST_SLABS", str(slabs))
    v1, f1 = pkg.isosurface(s, m, capacity=(len(vo) + 1, len(fo) + 1))
    assert np.array_equal(f1, fo) and _bits_equal(v1, vo)
    v2, f2 = pkg.api.isosurface_two_phase(s, m)
    assert np.array_equal(f2, fo) and _bits_equal(v2, vo)


@pytest.mark.parametrize("slabs", [1, 3])
def test_extract_host_never_reads_past_the_field(pkg, monkeypatch, slabs):
    """The staging copies round a row piece up to the staged pitch; the last row must still be read exactly.  The field
    ends on the last byte before an inaccessible page: one byte too far kills the process."""
    import ctypes, mmap
    shape = (70, 33, 41)  # 70 samples per row: the staged pitch (72) is wider than the row
    s = np.asfortranarray(pkg.synth.gyroid(shape))
    nbytes, page = s.nbytes, mmap.PAGESIZE
    span = (nbytes + page - 1) // page * page
    mm = mmap.mmap(-1, span + page)
    base = ctypes.addressof(ctypes.c_char.from_buffer(mm))
    libc = ctypes.CDLL(None, use_errno=True)
    libc.mprotect.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    assert libc.mprotect(base + span, page, 0) == 0  # PROT_NONE
    try:
        a = np.frombuffer(mm, dtype=np.float32, count=s.size, offset=span - nbytes).reshape(shape, order="F")
        a[...] = s
        assert a.ctypes.data + nbytes == base + span
        m = _method(pkg, "MC", 0.0, True)
        v0, f0 = pkg.api.isosurface_two_phase(s, m)
        monkeypatch.setenv("B200ISO_HOST_SLABS", str(slabs))
        p = pkg.api.make_params(m)
        vb, fb = np.empty((len(v0), 3), np.float32), np.empty((len(f0), 3), np.int64)
        nv, nf, _, fits = pkg.api.get_handle(0).extract_host(p, a.ctypes.data, *shape, shape[0], vb.ctypes.data, len(vb), fb.ctypes.data, len(fb))
        assert fits and np.array_equal(fb, f0) and _bits_equal(vb, v0)
        del a
    finally:
        libc.mprotect(base + span, page, 3)


@pytest.mark.parametrize("algo", ["MC", "MT"])
def test_host_paths_pinned_and_pageable_arrays(pkg, algo):
    """Pinned caller arrays take the direct-DMA lanes, pageable ones the threaded pinned staging (several 8 MB chunks
    per lane here); both entry points give the same bytes either way."""
    import torch
    n = (300, 260, 250)  # 78 MB field: multi-chunk, two slabs
    s = pkg.synth.gyroid(n)
    m = _method(pkg, algo, 0.0, True)
    v0, f0 = pkg.isosurface(s, m)  # pageable, two-phase
    pin = torch.empty(n[::-1], dtype=torch.float32).pin_memory()
    a = pin.numpy().transpose(2, 1, 0)  # Fortran-ordered view of the pinned block
    a[...] = s
    assert a.flags.f_contiguous
    vp = torch.empty((len(v0), 3), dtype=torch.float32).pin_memory().numpy()
    fp = torch.empty((len(f0), 3), dtype=torch.int64).pin_memory().numpy()
    h, p = pkg.api.get_handle(0), pkg.api.make_params(m)
    h.count(p, a.ctypes.data, pkg.capi.HOST, *n, n[0])
    h.generate(vp.ctypes.data, fp.ctypes.data, pkg.capi.HOST, 0)
    assert np.array_equal(fp, f0) and _bits_equal(vp, v0)
    vp[:], fp[:] = 0, 0
    assert pkg.api.isosurface_into(a, vp, fp, m) == (len(v0), len(f0))  # pinned in, pinned out
    assert np.array_equal(fp, f0) and _bits_equal(vp, v0)
    v1, f1 = np.zeros_like(v0), np.zeros_like(f0)
    assert pkg.api.isosurface_into(a, v1, f1, m) == (len(v0), len(f0))  # pinned in, pageable out
    assert np.array_equal(f1, f0) and _bits_equal(v1, v0)
    vp[:], fp[:] = 0, 0
    assert pkg.api.isosurface_into(np.asfortranarray(s), vp, fp, m) == (len(v0), len(f0))  # pageable in, pinned out
    assert np.array_equal(fp, f0) and _bits_equal(vp, v0)


@pytest.mark.parametrize("algo", ["MC", "MT"])
def test_peer_exchange_two_ranks_on_one_gpu(pkg, oracle, algo):
    """b200iso_exchange_async (the NVLink peer-memory form of the 16-byte all-gather): two handles play two ranks on
    one device, each publishing into both exchange buffers; several epochs exercise the parity double-buffering."""
    import torch
    s = pkg.synth.gyroid((61, 30, 33))
    m = _method(pkg, algo, 0.0, True)
    world = 2
    bufs = [torch.zeros(pkg.capi.PEER_BYTES // 8, dtype=torch.int64, device="cuda") for _ in range(world)]
    hs = [pkg.capi.Handle(0) for _ in range(world)]
    for r, h in enumerate(hs):
        h.set_peer_exchange(r, world, [b.data_ptr() for b in bufs])
    vo, fo = oracle.isosurface(s, ALGOS[algo], iso=0.0, iso_is_f32=True, eps_is_f32=True)
    slabs, prm = [], []
    for r in range(world):
        xa, xb = pkg.sharding.slab_bounds(61, world, r, ghost=(algo == "MT"))
        slabs.append(torch.from_numpy(np.asfortranarray(s[xa:xb]).transpose(2, 1, 0).copy()).cuda().permute(2, 1, 0))
        p = pkg.api.make_params(m)
        p.x_offset, p.nx_global, p.x_ghost = xa, 61, int(algo == "MT" and r > 0)
        prm.append(p)
    torch.cuda.synchronize()  # the handles run on their own streams
    bases = [torch.zeros(4, dtype=torch.int64, device="cuda") for _ in range(world)]
    allc = [torch.zeros((world, 2), dtype=torch.int64, device="cuda") for _ in range(world)]
    for epoch in range(3):
        for r, h in enumerate(hs):  # every "rank" counts and publishes before anyone has to wait
            t = slabs[r]
            h.count_async(prm[r], t.data_ptr(), t.shape[0], t.shape[1], t.shape[2], t.stride(1))
        for r, h in enumerate(hs):
            h.exchange_async(bases[r].data_ptr(), allc[r].data_ptr())
        parts = []
        for r, h in enumerate(hs):
            nv, nf, _ = h.totals()
            v = torch.empty((nv, 3), dtype=torch.float32, device="cuda")
            f = torch.empty((nf, 3), dtype=torch.int64, device="cuda")
            h.generate_async(v.data_ptr(), nv, f.data_ptr(), nf, bases[r].data_ptr(), 0)
            h.totals()  # (synchronises the handle's stream)
            torch.cuda.synchronize()
            parts.append((v.cpu().numpy(), f.cpu().numpy()))
        v, f = pkg.sharding.stitch(parts)
        assert np.array_equal(f, fo) and _bits_equal(v, vo)
        assert torch.equal(allc[0], allc[1]) and int(bases[1][0]) == len(parts[0][0]) and int(bases[0][2]) == len(vo)


def test_peer_exchange_timeout_reports_instead_of_hanging(pkg):
    """A rank that never publishes: the waiting rank gives up after about a second and the next totals() fails."""
    import torch
    s = pkg.synth.sphere((20, 20, 20))
    bufs = [torch.zeros(pkg.capi.PEER_BYTES // 8, dtype=torch.int64, device="cuda") for _ in range(2)]
    h = pkg.capi.Handle(0)
    h.set_peer_exchange(0, 2, [b.data_ptr() for b in bufs])
    t = torch.from_numpy(np.ascontiguousarray(s.transpose(2, 1, 0))).cuda().permute(2, 1, 0)
    p = pkg.api.make_params(pkg.MarchingCubes(iso=pkg.Float32(0)))
    torch.cuda.synchronize()
    h.count_async(p, t.data_ptr(), 20, 20, 20, t.stride(1))
    bases = torch.zeros(4, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    h.exchange_async(bases.data_ptr())
    with pytest.raises(pkg.capi.B200IsoError):
        h.totals()
    h.set_peer_exchange(0, 0, None)
    h.count_async(p, t.data_ptr(), 20, 20, 20, t.stride(1))
    assert h.totals()[0] > 0  # the handle is usable again


def test_c_example_runs_through_the_c_abi(pkg, c_example):
    """examples/extract.c: plain C against include/b200iso.h -- two-phase and one-shot host calls, same bytes."""
    import subprocess
    exe = c_example
    r = subprocess.run([exe, "96"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "identical bytes" in r.stdout


# ---- the TMA-staged classify kernel (what the 1024^3 benchmark times) against the oracle -------------------------
TMA_SHAPES = [(16, 16, 16), (33, 20, 47), (5, 130, 37), (129, 7, 70), (12, 9, 260), (131, 9, 40), (260, 5, 1030), (2, 2, 2),
              (127, 3, 33), (128, 4, 32), (256, 3, 513),
              (129, 20, 70), (257, 6, 40), (132, 40, 33), (130, 70, 600), (136, 5, 33)]  # 128 k + 1..8 samples: x-slabs with their halo plane (a nearly empty last x-segment)


@pytest.fixture
def tma_handle(pkg):
    """The cached handle of device 0 with the classify kernel forced to the TMA path."""
    h = pkg.api.get_handle(0)
    h.set_classify_mode(1)
    yield h
    h.set_classify_mode(-1)


@pytest.mark.parametrize("algo", ["MC", "MT"])
@pytest.mark.parametrize("shape", TMA_SHAPES)
def test_tma_classify_parity_shapes(pkg, oracle, tma_handle, algo, shape):
    """Ragged dimensions (nx not a multiple of 128, nz not a multiple of 8 / 32 / 512: NaN out-of-bounds fill of the
    tensor map), every shape through signpack_tma_kernel, bit-exact against the oracle incl. the case indices."""
    for s in (pkg.synth.gyroid(shape), pkg.synth.noise(shape, seed=17)):
        _check(pkg, oracle, s, algo)
        assert tma_handle.classify_path() == pkg.capi.CLASSIFY_TMA
        c = pkg.api.case_indices(s, _method(pkg, algo, 0.0, True))
        assert tma_handle.classify_path() == pkg.capi.CLASSIFY_TMA
        assert np.array_equal(c, oracle.case_indices(s, ALGOS[algo], iso_is_f32=True))


@pytest.mark.parametrize("algo", ["MC", "MT"])
def test_counting_warps_on_ragged_shapes(pkg, oracle, monkeypatch, algo):
    """The count riding in the TMA classify kernel (both algorithms) on ragged shapes small enough for the oracle: a fresh
    handle with the ride's minimum task count lowered to 1; the shapes have several waves of classify CTAs (more than
    1184 tasks), so that rows complete while the kernel runs and the counting warps do count."""
    monkeypatch.setenv("B200ISO_TMA", "1")
    monkeypatch.setenv("B200ISO_RIDE_MIN_TASKS", "1")
    h = pkg.capi.Handle(0)
    m = _method(pkg, algo, 0.0, True)
    p = pkg.api.make_params(m)
    for s in (pkg.synth.gyroid((129, 700, 1030)), pkg.synth.noise((129, 1300, 70), seed=5), pkg.synth.gyroid((260, 33, 97))):
        vo, fo = oracle.isosurface(s, ALGOS[algo], iso=0.0, iso_is_f32=True, eps_is_f32=True)
        big = s.shape[1] > 100
        for rep in range(2):  # (the second call reuses the queues the scan kernel reset)
            nv, nf, _ = h.count(p, s.ctypes.data, pkg.capi.HOST, *s.shape, s.shape[0])
            assert h.classify_path() == pkg.capi.CLASSIFY_TMA
            assert (nv, nf) == (len(vo), len(fo))
            v, f = np.empty((nv, 3), np.float32), np.empty((nf, 3), np.int64)
            h.generate(v.ctypes.data, f.ctypes.data, pkg.capi.HOST, 0)
            assert np.array_equal(f, fo) and _bits_equal(v, vo)
            assert h.ride_claimed() > 0 or not big  # the counting warps did count generate blocks
    del h


@pytest.mark.parametrize("algo", ["MC", "MT"])
def test_tma_classify_nan_inf_and_float64_iso(pkg, oracle, tma_handle, algo):
    s = pkg.synth.gyroid((70, 50, 90)).copy(order="F")
    s[3, 4, 5] = np.nan
    s[69, 49, 89] = np.nan
    s[7, 7, 7] = np.inf
    s[2, 9, 4] = -np.inf
    _check(pkg, oracle, s, algo)
    assert tma_handle.classify_path() == pkg.capi.CLASSIFY_TMA
    for iso, f32 in ((0.25, False), (1e-9, False), (-0.3, True)):
        c = pkg.api.case_indices(s, _method(pkg, algo, iso, f32))
        assert np.array_equal(c, oracle.case_indices(s, ALGOS[algo], iso=iso, iso_is_f32=f32))
    assert tma_handle.classify_path() == pkg.capi.CLASSIFY_TMA


def test_tma_classify_padded_device_rows(pkg, oracle, tma_handle):
    """Device-resident field whose leading dimension is larger than nx (an x-slab view of a bigger array)."""
    import torch
    s = pkg.synth.gyroid((150, 40, 50))
    store = torch.zeros((50, 40, 152), dtype=torch.float32, device="cuda")
    store[:, :, :150] = torch.from_numpy(np.ascontiguousarray(s.transpose(2, 1, 0))).cuda()
    for xa, xb in ((0, 150), (4, 137), (128, 150)):  # (16-byte aligned starts: the tensor map needs them)
        t = store.permute(2, 1, 0)[xa:xb]
        for algo in ("MC", "MT"):
            v, f = pkg.isosurface(t, _method(pkg, algo, 0.0, True))
            assert tma_handle.classify_path() == pkg.capi.CLASSIFY_TMA
            vo, fo = oracle.isosurface(s[xa:xb], ALGOS[algo], iso_is_f32=True, eps_is_f32=True)
            assert np.array_equal(f.cpu().numpy(), fo) and _bits_equal(v.cpu().numpy(), vo)
    t = store.permute(2, 1, 0)[3:150]  # unaligned start: no tensor map, the scalar-load kernel takes over
    v, f = pkg.isosurface(t, _method(pkg, "MC", 0.0, True))
    assert tma_handle.classify_path() == pkg.capi.CLASSIFY_SCALAR
    vo, fo = oracle.isosurface(s[3:150], 0, iso_is_f32=True)
    assert np.array_equal(f.cpu().numpy(), fo) and _bits_equal(v.cpu().numpy(), vo)


def test_tma_and_ldg_classify_give_the_same_bits_512(pkg, oracle):
    """configs[1] through both classify kernels: identical meshes (the LDG one is pinned to the oracle above)."""
    h = pkg.api.get_handle(0)
    s = pkg.synth.gyroid(512)
    m = pkg.MarchingCubes(iso=pkg.Float32(0))
    out = {}
    try:
        for mode in (0, 1):
            h.set_classify_mode(mode)
            out[mode] = pkg.isosurface(s, m)
            assert h.classify_path() == (pkg.capi.CLASSIFY_TMA if mode else pkg.capi.CLASSIFY_LDG128)
    finally:
        h.set_classify_mode(-1)
    assert np.array_equal(out[0][1], out[1][1]) and _bits_equal(out[0][0], out[1][0])
    vo, fo = oracle.isosurface(s, 0, iso_is_f32=True)
    assert np.array_equal(out[1][1], fo) and _bits_equal(out[1][0], vo)


# ---- BASELINE configs[4]: the multi-sphere/torus field ------------------------------------------------------------
@pytest.mark.parametrize("algo", ["MC", "MT"])
def test_parity_multisphere_torus_256(pkg, oracle, algo):
    s = pkg.synth.multisphere_torus(256)
    v, f = _check(pkg, oracle, s, algo)
    assert len(f) > 50000


@pytest.mark.parametrize("xa", [0, 700, 1021, 2042])
def test_parity_multisphere_torus_2048_slabs(pkg, oracle, xa):
    """configs[4] at its full 2048^3 size, slab-wise: Marching Cubes carries no state between voxels
    (src/marching_cubes.jl:40-62), so six sample planes of the 2048^3 field with the whole volume's coordinates are
    exactly that part of the whole sweep -- for the oracle (slab mode) and for the GPU (x_offset / nx_global)."""
    n, w = 2048, 6
    slab = pkg.synth.multisphere_torus((n, n, n), x_slice=(xa, xa + w))
    assert slab.shape == (w, n, n)
    m = pkg.MarchingCubes(iso=pkg.Float32(0))
    gv, gf = pkg.api.isosurface_slab(slab, m, xa, n, 0)
    assert pkg.api.get_handle(0).classify_path() == pkg.capi.CLASSIFY_TMA  # 8192 classify tasks: the automatic rule picks TMA
    ov, of = oracle.isosurface(slab, 0, iso_is_f32=True, slab=(xa, n))
    assert np.array_equal(gf, of) and _bits_equal(gv, ov)
    if xa == 1021:
        assert len(of) > 1000  # the torus crosses the middle of the volume


# ---- Marching Tetrahedra at 1024^3 -----------------------------------------------------------------
def test_full_size_1024_gyroid_mt_against_oracle_ranges(pkg, oracle):
    """MT on the 1024^3 gyroid: totals, manifold-style invariants, and the oracle's sweep of x-ranges compared with the
    corresponding slices of the whole-volume result.  MT shares vertices between voxels, so the oracle sweeps the range
    plus one voxel plane below it (which creates the shared vertices first); vertex ids then differ by a constant for
    vertices created inside the range, and every face must have bit-identical corner coordinates."""
    import torch
    n = 1024
    t = pkg.synth.gyroid_torch(n, "cuda")
    m = pkg.MarchingTetrahedra(iso=pkg.Float32(0), eps=pkg.Float32(1e-3))
    v, f = pkg.isosurface(t, m)
    assert v.dtype == torch.float32 and int(f.min()) == 1 and int(f.max()) == v.shape[0]
    assert bool(torch.isfinite(v).all())
    # MT vertex count identity: one vertex per sign-changing lattice edge of the 7 direction types
    b = t < 0
    nv_expect = 0
    for dx, dy, dz in ((1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 1, 1)):
        nv_expect += int((b[: n - dx, : n - dy, : n - dz] != b[dx:, dy:, dz:]).sum())
    assert v.shape[0] == nv_expect
    host = t.cpu().numpy()
    vh, fh = v.cpu().numpy(), f.cpu().numpy()
    del v, f, b
    for xlo in (1, 600, n - 4):
        xhi = xlo + 3
        # totals of the voxel planes [0, xlo) and [0, xlo - 1): where the range and its ghost plane start in the output
        _, v_lo, f_lo, _ = pkg.api.slab_count(t[0:xlo + 1], m, 0, n)
        ov, of = oracle.isosurface(host, 1, iso_is_f32=True, eps_is_f32=True, xrange=(xlo - 1, xhi))
        gv, gf = oracle.isosurface(host, 1, iso_is_f32=True, eps_is_f32=True, xrange=(xlo - 1, xlo), copy=False)  # ghost plane alone
        of_r = of[gf:]  # faces of the range proper
        fg = fh[f_lo:f_lo + len(of_r)]
        assert len(of_r) > 10000
        # bit-identical corner coordinates, face by face
        assert _bits_equal(vh[fg - 1], ov[of_r - 1]), xlo
        # vertices created inside the range: ids differ from the oracle's by one constant
        own = of_r > gv
        d = fg[own] - of_r[own]
        assert d.size > 0 and int(d.min()) == int(d.max()) == v_lo - gv, xlo
        # vertices of the ghost plane keep their relative order
        gh = ~own
        assert np.array_equal(np.argsort(fg[gh], kind="stable"), np.argsort(of_r[gh], kind="stable"))


# ---- randomised sweep (hypothesis, derandomised: the same examples every run) ---------------------------------------
def _random_field(shape, kind, seed, f64):
    rng = np.random.default_rng(seed)
    if kind == 0:  # smooth: a sum of a few random plane waves (a surface-like level set)
        x, y, z = np.meshgrid(*[np.linspace(0, 1, n) for n in shape], indexing="ij")
        s = np.zeros(shape)
        for _ in range(3):
            k = rng.uniform(-9, 9, 3)
            s += rng.uniform(0.3, 1.0) * np.sin(k[0] * x + k[1] * y + k[2] * z + rng.uniform(0, 6.28))
    elif kind == 1:  # dense noise
        s = rng.uniform(-1, 1, shape)
    else:  # few distinct values: many samples EQUAL to the level and to each other (degenerate interpolation)
        s = rng.integers(-2, 3, shape).astype(np.float64) * 0.25
    s = s.astype(np.float64 if f64 else np.float32)
    if seed % 5 == 0:  # a few NaN / Inf samples
        idx = tuple(rng.integers(0, n, 6) for n in shape)
        s[idx] = np.array([np.nan, np.inf, -np.inf, np.nan, np.inf, -np.inf], dtype=s.dtype)
    return np.asfortranarray(s)


try:
    from hypothesis import given, settings, strategies as st, HealthCheck

    @pytest.mark.gpu
    @settings(max_examples=400, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
    @given(nx=st.integers(2, 45), ny=st.integers(2, 30), nz=st.sampled_from([2, 3, 17, 32, 33, 64, 100, 129, 260]), kind=st.integers(0, 2),
           seed=st.integers(0, 10 ** 6), algo=st.sampled_from(["MC", "MT"]), f32=st.booleans(), f64_field=st.booleans(),
           rk=st.integers(0, 2), iso=st.sampled_from([0.0, 0.25, -0.25, 0.1, 0.5]), tma=st.booleans())
    def test_random_fields_against_the_oracle(pkg, oracle, nx, ny, nz, kind, seed, algo, f32, f64_field, rk, iso, tma):
        """Random shapes, field kinds (smooth / noise / few distinct values with samples equal to the level; NaN and Inf
        sprinkled in), level, method and range types, field precision, both classify kernels: bit-exact against the oracle."""
        s = _random_field((nx, ny, nz), kind, seed, f64_field)
        ranges = ((-1, 1), (0, 3), (-2, 5)) if rk == 0 else ((-1.0, 1.5), (0.25, 3.0), (-2.0, 5.5))
        h = pkg.api.get_handle(0)
        h.set_classify_mode(1 if tma else -1)
        try:
            _check(pkg, oracle, s, algo, iso=iso, f32=f32, ranges=ranges, rk=rk)
        finally:
            h.set_classify_mode(-1)
except ImportError:  # hypothesis missing: the parametrised tests above stand alone
    pass
