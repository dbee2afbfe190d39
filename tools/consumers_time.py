#!/usr/bin/env python3
"""Device-resident timing of the mesh consumers on the gyroid (development aid): normals, edge keys, weld.
usage: consumers_time.py [n ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from __graft_entry__ import load_package
pkg = load_package()
capi = pkg.capi
sizes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [512, 1024]
h = pkg.api.get_handle(0)


def timed(fn, k=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


for n in sizes:
    t = pkg.synth.gyroid_torch(n, "cuda")
    m = pkg.MarchingCubes(iso=pkg.Float32(0))
    p = pkg.api.make_params(m)
    ldx = t.stride(1)
    h.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        nv, nf, f64 = h.count(p, t.data_ptr(), capi.DEVICE, n, n, n, ldx)
        verts = torch.empty((nv, 3), dtype=torch.float32, device="cuda")
        faces = torch.empty((nf, 3), dtype=torch.int64, device="cuda")
        keys = torch.empty(nv, dtype=torch.int64, device="cuda")
        normals = torch.empty((nv, 3), dtype=torch.float32, device="cuda")
        wv, wf = torch.empty_like(verts), torch.empty_like(faces)
        h.generate(verts.data_ptr(), faces.data_ptr(), capi.DEVICE, 0)
        t_keys = timed(lambda: h.vertex_keys_async(keys.data_ptr(), nv))
        t_norm = timed(lambda: h.vertex_normals_async(p, t.data_ptr(), n, n, n, ldx, verts.data_ptr(), nv, False, normals.data_ptr()))
        nw = [0]

        def weld():
            nw[0] = h.weld(keys.data_ptr(), verts.data_ptr(), nv, False, faces.data_ptr(), nf, 0, wv.data_ptr(), wf.data_ptr())
        t_weld = timed(weld, k=3)
    finally:
        h.use_own_stream()
    print(f"MC n={n}: {nv} vertices, {nf} faces -> welded {nw[0]} vertices ({nv / max(nw[0], 1):.2f}x fewer); "
          f"edge keys {t_keys:.3f} ms, normals {t_norm:.3f} ms ({nv * (12 + 12) / t_norm / 1e6:.0f} GB/s of vertex read + normal write), "
          f"weld {t_weld:.3f} ms (host-synchronous call)", flush=True)
    del t, verts, faces, keys, normals, wv, wf
    torch.cuda.empty_cache()
