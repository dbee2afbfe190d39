#!/usr/bin/env python3
"""Small workload for compute-sanitizer (memcheck / racecheck / initcheck): every kernel, both algorithms,
sparse and dense fields, Float32 and Float64, multi-window blocks, slabs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
pkg = load_package()
F = pkg.Float32
for s in (pkg.synth.gyroid((48, 40, 56)), pkg.synth.noise((33, 21, 300), seed=1), pkg.synth.noise((20, 20, 20), seed=2).astype(np.float64),
          pkg.synth.gyroid((140, 9, 70))):
    for m in (pkg.MarchingCubes(iso=F(0)), pkg.MarchingTetrahedra(iso=F(0), eps=F(1e-3)), pkg.MarchingCubes()):
        v, f = pkg.isosurface(s, m)
        c = pkg.api.case_indices(s, m)
        print(s.shape, s.dtype, type(m).__name__, len(v), len(f), int(c.sum()))
# the TMA-staged classify kernel on ragged shapes (NaN out-of-bounds fill)
_h = pkg.api.get_handle(0)
_h.set_classify_mode(1)
for shp in ((131, 9, 40), (33, 20, 47), (260, 5, 1030)):
    s = pkg.synth.gyroid(shp)
    for m in (pkg.MarchingCubes(iso=F(0)), pkg.MarchingTetrahedra(iso=F(0), eps=F(1e-3))):
        v, f = pkg.isosurface(s, m)
        assert _h.classify_path() == pkg.capi.CLASSIFY_TMA
        print("tma", shp, type(m).__name__, len(v), len(f))
_h.set_classify_mode(-1)
# the counting warps inside the TMA classify kernel, both algorithms (a fresh handle: the knobs are read at creation);
# 1400 rows = 2800 classify tasks, more than two waves of CTAs, so the warps do count
_os_env = {"B200ISO_TMA": "1", "B200ISO_RIDE_MIN_TASKS": "1"}
os.environ.update(_os_env)
_hr = pkg.capi.Handle(0)
for k in _os_env:
    del os.environ[k]
s = pkg.synth.gyroid((129, 1400, 70))
for m in (pkg.MarchingCubes(iso=F(0)), pkg.MarchingTetrahedra(iso=F(0), eps=F(1e-3))):
    p = pkg.api.make_params(m)
    nv, nf, _ = _hr.count(p, s.ctypes.data, pkg.capi.HOST, *s.shape, s.shape[0])
    v, f = np.empty((nv, 3), np.float32), np.empty((nf, 3), np.int64)
    _hr.generate(v.ctypes.data, f.ctypes.data, pkg.capi.HOST, 0)
    print("ride", type(m).__name__, nv, nf, "claimed", _hr.ride_claimed())
del _hr
s = pkg.synth.gyroid((41, 19, 70))
for m in (pkg.MarchingCubes(iso=F(0)), pkg.MarchingTetrahedra(iso=F(0), eps=F(1e-3))):
    gh = isinstance(m, pkg.MarchingTetrahedra)
    for r in range(3):
        xa, xb = pkg.sharding.slab_bounds(41, 3, r, ghost=gh)
        v, f = pkg.api.isosurface_slab(s[xa:xb], m, xa, 41, 100)
        print("slab", r, len(v), len(f))
# one-shot host path (slab pipeline, pinned staging threads) and the fused single pass
import os as _os
_os.environ["B200ISO_HOST_SLABS"] = "3"
s = pkg.synth.gyroid((70, 33, 41))
for m in (pkg.MarchingCubes(iso=F(0)), pkg.MarchingTetrahedra(iso=F(0), eps=F(1e-3))):
    v0, f0 = pkg.isosurface(s, m)
    v1, f1 = pkg.isosurface(s, m, capacity=(len(v0), len(f0)))
    assert np.array_equal(f0, f1)
    print("extract_host", type(m).__name__, len(v1), len(f1))
# peer exchange: two handles as two ranks on one device
import torch
bufs = [torch.zeros(pkg.capi.PEER_BYTES // 8, dtype=torch.int64, device="cuda") for _ in range(2)]
hs = [pkg.capi.Handle(0) for _ in range(2)]
t = torch.from_numpy(np.ascontiguousarray(pkg.synth.gyroid((30, 20, 40)).transpose(2, 1, 0))).cuda().permute(2, 1, 0)
p = pkg.api.make_params(pkg.MarchingCubes(iso=F(0)))
bases = [torch.zeros(4, dtype=torch.int64, device="cuda") for _ in range(2)]
torch.cuda.synchronize()
for r, h in enumerate(hs):
    h.set_peer_exchange(r, 2, [b.data_ptr() for b in bufs])
for ep in range(2):
    for h in hs:
        h.count_async(p, t.data_ptr(), 30, 20, 40, t.stride(1))
    for r, h in enumerate(hs):
        h.exchange_async(bases[r].data_ptr())
    print("peer", ep, [h.totals()[:2] for h in hs], [int(b[0]) for b in bases])
print("done")
