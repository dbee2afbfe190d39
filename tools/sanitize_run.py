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
s = pkg.synth.gyroid((41, 19, 70))
for m in (pkg.MarchingCubes(iso=F(0)), pkg.MarchingTetrahedra(iso=F(0), eps=F(1e-3))):
    gh = isinstance(m, pkg.MarchingTetrahedra)
    for r in range(3):
        xa, xb = pkg.sharding.slab_bounds(41, 3, r, ghost=gh)
        v, f = pkg.api.isosurface_slab(s[xa:xb], m, xa, 41, 100)
        print("slab", r, len(v), len(f))
print("done")
