"""CPU tests that pin the oracle (oracle/iso_oracle.cpp) against everything the reference's own test-suite
holds for the hot path (test/runtests.jl), SURVEY.md §8c.  No GPU needed."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- test/runtests.jl:15-27 ------------------------------------------------------------------------------
def test_get_cubeindex_kat(oracle):
    assert oracle.get_cubeindex([0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8], 0.35) == 0x07
    assert oracle.get_cubeindex([0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2], 0.75) == 0x07
    assert oracle.get_cubeindex([0.9, 0.8, 0.7, 0.6, 0.5, 0.4, 0.3, 0.2], 0.5) == 0xE0  # strict <


# ---- test/runtests.jl:29-33 ------------------------------------------------------------------------------
def test_vertex_interp_kat(oracle):
    assert tuple(oracle.vertex_interp(0, (0, 0, 0), (0, 1, 0), -1, 1)) == (0, 0.5, 0)
    assert tuple(oracle.vertex_interp(-1, (0, 0, 0), (0, 1, 0), -1, 1)) == (0, 0, 0)
    assert tuple(oracle.vertex_interp(1, (0, 0, 0), (0, 1, 0), -1, 1)) == (0, 1, 0)


# ---- Julia's MersenneTwister restatement, pinned by Julia's own documented outputs ---------------------------
def test_julia_mersenne_twister_known_outputs():
    from oracle.julia_mt import MersenneTwister

    r = MersenneTwister(0)
    assert [r.rand() for _ in range(5)] == [0.8236475079774124, 0.9103565379264364, 0.16456579813368521,
                                            0.17732884646626457, 0.278880109331201]
    assert list(MersenneTwister(1234).rand(2)) == [0.5908446386657102, 0.7667970365022592]


# ---- test/runtests.jl:152-172 "noisy spheres": the reference's only exact-count test -------------------------
def test_noisy_spheres_golden_counts(oracle):
    from oracle.julia_mt import MersenneTwister

    N, sigma = 10, 1.0
    i = np.arange(-N, N + 1)
    I, J, K = np.meshgrid(i, i, i, indexing="ij")
    dist = np.sqrt((I * I + J * J + K * K).astype(np.float32)).astype(np.float32)
    field = dist.astype(np.float64) + sigma * MersenneTwister(0).rand(2 * N + 1, 2 * N + 1, 2 * N + 1)
    committed = np.load(os.path.join(GOLDEN, "noisy_spheres_input.npy"))
    assert np.array_equal(field, committed)
    points, faces = oracle.isosurface(field, oracle.MT, iso=N - 2 * sigma)
    assert len(points) == 3466
    assert len(faces) == 6928
    assert points.dtype == np.float64
    # watertight: every edge is used exactly twice, V - F/2 = 2 per closed surface component (one sphere)
    e = np.sort(np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]]), axis=1)
    _, cnt = np.unique(e, axis=0, return_counts=True)
    assert (cnt == 2).all()


# ---- test/runtests.jl:127-149 "respect origin" ------------------------------------------------------------
@pytest.mark.parametrize("algo,nv,nf", [(0, 187800, 93896), (1, 140714, 281424)])
def test_respect_origin(oracle, algo, nv, nf):
    g = np.arange(-100, 101) / 100.0  # -1:0.01:1
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    norm_sdf = np.sqrt(X * X + Y * Y + Z * Z)
    points, faces = oracle.isosurface(norm_sdf, algo, iso=0.5)
    assert points.shape == (nv, 3) and faces.shape == (nf, 3)  # SURVEY.md §4 [verified] counts
    assert np.allclose(points.mean(0), 0, atol=0.015)
    assert np.allclose(points.max(0), 0.5, atol=1e-3)
    assert np.allclose(points.min(0), -0.5, atol=1e-3)
    assert faces.min() == 1 and faces.max() == nv


# ---- test/runtests.jl:98-125 "forward diff" (value part) ---------------------------------------------------
@pytest.mark.parametrize("algo", [0, 1])
def test_surface_distance_from_origin(oracle, algo):
    g = np.arange(-100, 101) / 50.0  # -2:0.02:2 (coarser than the reference's 401^3, same property)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    sdf = np.sqrt(X * X + Y * Y + Z * Z)
    points, _ = oracle.isosurface(sdf, algo, iso=0.85, ranges=((-2, 2),) * 3, range_kind=oracle.RANGE_F64)
    assert abs(np.linalg.norm(points, axis=1).mean() - 0.85) < 1e-2


# ---- test/runtests.jl:81-97 "mixed types": face topology independent of the field precision ----------------
def test_mixed_types(oracle):
    rng = np.random.default_rng(7)
    s = rng.standard_normal((10, 10, 10))
    p1, f1 = oracle.isosurface(s, oracle.MT, ranges=((0, 1),) * 3)
    p2, f2 = oracle.isosurface(s.astype(np.float32), oracle.MT, ranges=((0, 1),) * 3)
    # (rounding to Float32 can flip a sample across iso only if |s| < 1e-8; not with this seed)
    assert len(p1) == len(p2) and np.array_equal(f1, f2)


# ---- type contracts, test/runtests.jl:35-75 and SURVEY.md Appendix A1 ------------------------------------------
def test_vertex_type_rule(oracle):
    s32 = np.random.default_rng(1).standard_normal((6, 7, 5)).astype(np.float32)
    for algo in (0, 1):
        assert oracle.isosurface(s32, algo)[0].dtype == np.float64  # default iso = 0.0::Float64
        assert oracle.isosurface(s32, algo, iso_is_f32=True, eps_is_f32=True)[0].dtype == np.float32
        assert oracle.isosurface(s32, algo, iso_is_f32=True, eps_is_f32=True, range_kind=oracle.RANGE_F32)[0].dtype == np.float32
        assert oracle.isosurface(s32, algo, iso_is_f32=True, eps_is_f32=True, range_kind=oracle.RANGE_F64)[0].dtype == np.float64
        assert oracle.isosurface(s32.astype(np.float64), algo, iso_is_f32=True, eps_is_f32=True)[0].dtype == np.float64
    assert oracle.isosurface(s32, 1, iso_is_f32=True, eps_is_f32=False)[0].dtype == np.float64  # eps promotes (MT only)
    assert oracle.isosurface(s32, 0, iso_is_f32=True, eps_is_f32=False)[0].dtype == np.float32


def test_linrange_endpoints_exact(oracle):
    L = oracle.lib()
    for n in (2, 3, 128, 1024):
        assert L.oracle_linrange_f64(-1.0, 1.0, n, 0) == -1.0
        assert L.oracle_linrange_f64(-1.0, 1.0, n, n - 1) == 1.0
    assert L.oracle_linrange_f64(-1.0, 1.0, 3, 1) == 0.0


# ---- structural properties --------------------------------------------------------------------------------
def test_mc_structure_and_order(oracle, pkg):
    """MC emits unshared per-voxel vertex blocks in x-outermost / z-innermost order."""
    s = pkg.synth.sphere((20, 17, 23))
    verts, faces = oracle.isosurface(s, oracle.MC, iso_is_f32=True)
    cases = oracle.case_indices(s, oracle.MC, iso_is_f32=True)
    assert cases.shape == (19 * 16 * 22,)
    active = np.flatnonzero((cases != 0) & (cases != 255))
    # vertex count per case = number of sign-changing cube edges
    edges = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
    c = cases[active].astype(np.int64)
    nv = sum((((c >> a) ^ (c >> b)) & 1) for a, b in edges)
    assert nv.sum() == len(verts)
    # first face of every voxel is (fct+3, fct+2, fct+1)
    fct = np.concatenate([[0], np.cumsum(nv)[:-1]])
    first = np.stack([fct + 3, fct + 2, fct + 1], axis=1)
    assert np.isin(first.view([("", first.dtype)] * 3), faces.view([("", faces.dtype)] * 3)).all()
    assert faces.min() == 1 and faces.max() == len(verts)
    # vertices lie inside the voxel that emitted them, voxels visited in scan-rank order
    z = active % 22
    y = (active // 22) % 16
    x = active // (22 * 16)
    owner = np.repeat(np.arange(len(active)), nv)
    lo = np.stack([-1 + 2 * x / 19, -1 + 2 * y / 16, -1 + 2 * z / 22], axis=1)[owner]
    hi = np.stack([-1 + 2 * (x + 1) / 19, -1 + 2 * (y + 1) / 16, -1 + 2 * (z + 1) / 22], axis=1)[owner]
    assert (verts >= lo - 1e-6).all() and (verts <= hi + 1e-6).all()


def test_mt_vertex_count_identity(oracle, pkg):
    """nverts_MT == number of lattice edges of the 7 direction types whose endpoints differ in (value < iso)
    (SURVEY.md Appendix C)."""
    s = pkg.synth.gyroid((24, 20, 28))
    verts, faces = oracle.isosurface(s, oracle.MT, iso_is_f32=True, eps_is_f32=True)
    b = s < 0
    n = 0
    for dx, dy, dz in [(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 1, 1)]:
        A = b[: b.shape[0] - dx, : b.shape[1] - dy, : b.shape[2] - dz]
        B = b[dx:, dy:, dz:]
        n += int((A != B).sum())
    assert n == len(verts)
    assert sorted(np.unique(faces)) == list(range(1, len(verts) + 1))


def test_threaded_mc_identical(oracle, pkg):
    s = pkg.synth.gyroid((40, 33, 37))
    v1, f1 = oracle.isosurface(s, oracle.MC, iso_is_f32=True)
    v4, f4 = oracle.isosurface(s, oracle.MC, iso_is_f32=True, nthreads=5)
    assert np.array_equal(v1, v4) and np.array_equal(f1, f4)


def test_degenerate_sizes(oracle):
    for shape in [(1, 5, 5), (5, 1, 5), (5, 5, 1), (1, 1, 1)]:
        v, f = oracle.isosurface(np.zeros(shape, np.float32) - 1, oracle.MC)
        assert len(v) == 0 and len(f) == 0
    s = np.full((2, 2, 2), 1.0, np.float32)
    s[0, 0, 0] = -1
    v, f = oracle.isosurface(s, oracle.MC, iso_is_f32=True)
    assert len(v) == 3 and f.tolist() == [[3, 2, 1]]
    assert np.array_equal(v, np.array([[0, -1, -1], [-1, -1, 0], [-1, 0, -1]], np.float32))  # edges 1, 9, 4


def test_nan_samples_are_outside(oracle):
    s = np.full((3, 3, 3), -1.0, np.float32)
    s[1, 1, 1] = np.nan
    c = oracle.case_indices(s, oracle.MC, iso_is_f32=True)
    assert ((c != 0xFF).sum()) == 8  # NaN < iso is false -> 8 voxels see one "outside" corner
    v, f = oracle.isosurface(s, oracle.MC, iso_is_f32=True)
    assert len(f) == 8 and np.isnan(v).any()


def test_mt_xrange_sample_is_the_same_sweep(oracle, pkg):
    """bench.py times the MT restatement on a bounded x-range: the full range must be the full sweep, and a partial
    range the sweep of that sub-volume (first-touch order restarts at the range's first plane)."""
    s = pkg.synth.gyroid((20, 14, 17))
    v, f = oracle.isosurface(s, oracle.MT, iso_is_f32=True, eps_is_f32=True)
    v1, f1 = oracle.isosurface(s, oracle.MT, iso_is_f32=True, eps_is_f32=True, xrange=(0, 19))
    assert np.array_equal(f, f1) and np.array_equal(v, v1)
    # planes [0, 7) of the volume == the whole sweep of the first 8 sample planes
    v2, f2 = oracle.isosurface(s, oracle.MT, iso_is_f32=True, eps_is_f32=True, xrange=(0, 7), ranges=((0, 19), (0, 1), (0, 1)))
    v3, f3 = oracle.isosurface(np.asfortranarray(s[:8]), oracle.MT, iso_is_f32=True, eps_is_f32=True, ranges=((0, 7), (0, 1), (0, 1)))
    assert np.array_equal(f2, f3) and np.allclose(v2, v3, atol=1e-5)
    # threaded throughput driver: same faces count, boundary vertices duplicated (never used for parity)
    v4, f4 = oracle.isosurface(s, oracle.MT, iso_is_f32=True, eps_is_f32=True, nthreads=3)
    assert len(f4) == len(f) and len(v4) >= len(v)


def test_mt_sphere_is_a_closed_manifold(oracle, pkg):
    """Marching Tetrahedra shares vertices through the reference's Dict: on a closed surface away from the volume
    boundary every undirected edge of the mesh belongs to exactly two faces, with opposite orientations."""
    s = pkg.synth.sphere((24, 25, 26))
    v, f = oracle.isosurface(s, oracle.MT, iso_is_f32=True, eps_is_f32=True)
    assert len(f) > 0 and f.min() == 1 and f.max() == len(v)
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    und, cnt = np.unique(np.sort(e, axis=1), axis=0, return_counts=True)
    assert (cnt == 2).all()
    directed = {(int(a), int(b)) for a, b in e}
    assert len(directed) == len(e) and all((b, a) in directed for a, b in directed)
    assert len(v) - len(und) + len(f) == 2  # Euler characteristic of a sphere


def test_slab_mode_equals_the_x_range_of_the_whole_sweep(oracle):
    """oracle.isosurface(slab=...) (used for the slab-wise parity of the 2048^3 config) is bit-identical to the
    corresponding x-range of the whole-volume sweep: Marching Cubes carries no state between voxels."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from __graft_entry__ import load_package
    synth = load_package().synth
    s = synth.gyroid((40, 21, 30))
    rr = ((0.0, 3.0), (-1.0, 1.0), (2.0, 5.0))
    for rk in (oracle.RANGE_INT, oracle.RANGE_F32, oracle.RANGE_F64):
        for xa, xb in ((0, 9), (10, 26), (33, 40)):
            vs, fs = oracle.isosurface(s[xa:xb], oracle.MC, iso=0.05, iso_is_f32=True, ranges=rr, range_kind=rk, slab=(xa, 40))
            vr, fr = oracle.isosurface(s, oracle.MC, iso=0.05, iso_is_f32=True, ranges=rr, range_kind=rk, xrange=(xa, xb - 1))
            assert len(vs) > 0 and np.array_equal(fs, fr)
            assert np.array_equal(vs.view(np.uint32 if vs.dtype == np.float32 else np.uint64),
                                  vr.view(np.uint32 if vr.dtype == np.float32 else np.uint64))
