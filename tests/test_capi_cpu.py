"""CPU-side checks of the product's boundary: the library loads, exports every declared symbol, and the
host mirror maps the reference's call forms onto struct b200iso_params (no compute without a GPU)."""
import numpy as np
import pytest


def test_library_exports_every_declared_symbol(pkg):
    L = pkg.capi.load()
    names = pkg.capi.declared_symbols()
    assert "b200iso_count" in names and "b200iso_generate" in names and len(names) >= 14
    for n in names:
        assert hasattr(L, n), n
    assert L.b200iso_version() >= 2000


def test_params_mirror_reference_call_forms(pkg):
    api, capi = pkg.api, pkg.capi
    p = api.make_params(api.MarchingCubes())  # MarchingCubes() => iso = 0.0::Float64, X = Y = Z = -1:1
    assert (p.algo, p.iso, p.iso_is_f32, p.range_kind) == (capi.MC, 0.0, 0, capi.RANGE_INT)
    assert (p.x0, p.x1, p.y0, p.y1, p.z0, p.z1) == (-1, 1, -1, 1, -1, 1)
    p = api.make_params(api.MarchingTetrahedra(iso=api.Float32(0.5), eps=api.Float32(1e-3)), (0, 1), (0, 1), (0, 1))
    assert (p.algo, p.iso_is_f32, p.eps_is_f32) == (capi.MT, 1, 1) and p.iso == 0.5 and p.eps == float(np.float32(1e-3))
    p = api.make_params(api.MarchingTetrahedra(iso=100), range(-2, 3), range(0, 2), range(5, 9))  # Int iso (examples/nrrd.jl:14)
    assert p.iso == 100.0 and p.iso_is_f32 == 1 and p.eps_is_f32 == 0 and (p.x0, p.x1, p.z0, p.z1) == (-2, 2, 5, 8)
    rng = np.arange(-2, 2.001, 0.01)
    p = api.make_params(api.MarchingCubes(iso=0.85), rng, rng, rng)
    assert p.range_kind == capi.RANGE_F64 and p.x0 == -2.0 and abs(p.x1 - 2.0) < 1e-9
    f = np.float32
    p = api.make_params(api.MarchingCubes(iso=f(0)), (f(-2), f(2)), (f(-2), f(2)), (f(-2), f(2)))
    assert p.range_kind == capi.RANGE_F32 and p.iso_is_f32 == 1
    # mixed endpoint types: the vertex type follows eltype(first(X)) only (src/marching_cubes.jl:31), the coordinates
    # follow LinRange(first(X), last(X), n), i.e. the promotion of BOTH endpoints (:36-38)
    for rng, kind in (((0, 1.0), capi.RANGE_INT), ((0.0, 1), capi.RANGE_F64), ((f(0), 1.0), capi.RANGE_INT),
                      ((0, f(1)), capi.RANGE_F32), ((f(0), 1), capi.RANGE_F32), ((0.0, f(1)), capi.RANGE_F64)):
        p = api.make_params(api.MarchingCubes(iso=f(0)), rng, rng, rng)
        assert p.range_kind == kind and (p.x0, p.x1) == (0.0, 1.0), rng


def test_argument_errors(pkg):
    api = pkg.api
    with pytest.raises(TypeError):
        api.make_params("MarchingCubes")
    with pytest.raises(TypeError):
        api.make_params(api.MarchingCubes(iso="0"))
    with pytest.raises(TypeError):
        api.make_params(api.MarchingCubes(), (0, 1), (0.0, 1.0), (0, 1))
    with pytest.raises(TypeError):  # Integer / Float16 fields: outside the accelerated path, no CPU fallback
        api.isosurface(np.zeros((4, 4, 4), np.int64))
    with pytest.raises(TypeError):
        api.isosurface(np.zeros((4, 4, 4), np.float16))
    with pytest.raises(TypeError):
        api.isosurface(np.zeros((4, 4), np.float32))


def test_no_gpu_fails_loudly(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.capi.B200IsoError):
        pkg.capi.Handle(0)


def test_product_never_imports_the_oracle():
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dp, _, files in os.walk(os.path.join(root, "meshing.jl_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl", "Makefile")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in src.lower(), os.path.join(dp, f)


def test_vertex_type_rule_of_the_host_mirror_matches_the_oracle(pkg, oracle):
    """api._vert_is_f64 sizes the caller's arrays for the one-shot host call before anything ran: it must be the
    reference's float(promote_type(...)) rule (src/marching_cubes.jl:31, src/marching_tetrahedra.jl:131)."""
    api = pkg.api
    s32 = pkg.synth.sphere((6, 6, 6))
    for field in (s32, s32.astype(np.float64)):
        for algo in ("MC", "MT"):
            for f32 in (True, False):
                for rk, conv in ((oracle.RANGE_INT, int), (oracle.RANGE_F32, np.float32), (oracle.RANGE_F64, float)):
                    cv = api.Float32 if f32 else float
                    m = api.MarchingCubes(iso=cv(0)) if algo == "MC" else api.MarchingTetrahedra(iso=cv(0), eps=cv(1e-3))
                    rng = tuple((conv(-1), conv(1)) for _ in range(3))
                    p = api.make_params(m, *rng)
                    p.field_is_f64 = int(field.dtype == np.float64)
                    v, _ = oracle.isosurface(field, 0 if algo == "MC" else 1, iso=0.0, iso_is_f32=f32, eps=1e-3, eps_is_f32=f32,
                                             ranges=((-1, 1),) * 3, range_kind=rk)
                    assert api._vert_is_f64(p) == (v.dtype == np.float64), (field.dtype, algo, f32, rk)


def test_header_is_plain_c_and_the_c_example_links(pkg, c_example):
    """include/b200iso.h must be usable from C (what ccall / cgo / JNI bind), not only from C++."""
    import subprocess
    exe = c_example
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe, "16"], capture_output=True, text=True)
        assert r.returncode != 0 and "b200iso_create" in r.stderr  # no GPU: fails loudly, no fallback
