#!/usr/bin/env python3
"""Minimal driver for ncu captures: a few device-resident steps of one workload (no timing, no checks)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
pkg = load_package()
algo = sys.argv[1] if len(sys.argv) > 1 else "MC"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
t = pkg.synth.gyroid_torch(n, "cuda")
m = pkg.MarchingCubes(iso=pkg.Float32(0)) if algo == "MC" else pkg.MarchingTetrahedra(iso=pkg.Float32(0), eps=pkg.Float32(1e-3))
p = pkg.api.make_params(m)
h = pkg.capi.Handle(0)
nv, nf, f64 = h.count(p, t.data_ptr(), pkg.capi.DEVICE, n, n, n, t.stride(1))
verts = torch.empty((nv, 3), dtype=torch.float32, device="cuda")
faces = torch.empty((nf, 3), dtype=torch.int64, device="cuda")
for _ in range(steps):
    h.count_async(p, t.data_ptr(), n, n, n, t.stride(1), 0)
    h.generate_async(verts.data_ptr(), nv, faces.data_ptr(), nf, 0, 0)
print(h.totals())
