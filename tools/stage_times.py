#!/usr/bin/env python3
"""Quick per-stage device timing of the pipeline on device-resident gyroids (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from __graft_entry__ import load_package
pkg = load_package()
algo = sys.argv[1] if len(sys.argv) > 1 else "MC"
f64 = "f64" in sys.argv  # Float64 field (and therefore Float64 vertices)
fused = "oneq" in sys.argv  # time b200iso_extract_async (one enqueue) end to end as well
mode = 0
sizes = [int(a) for a in sys.argv[2:] if a.isdigit()] or [256, 512, 1024]
h = pkg.capi.Handle(0)
h.enable_timing(True)
for n in sizes:
    t = pkg.synth.gyroid_torch(n, "cuda")
    if f64:
        t = (t.permute(2, 1, 0).contiguous().double() * 1.0000000001).permute(2, 1, 0)  # same layout, Float64 samples
    torch.cuda.synchronize()
    m = pkg.MarchingCubes(iso=pkg.Float32(0)) if algo == "MC" else pkg.MarchingTetrahedra(iso=pkg.Float32(0), eps=pkg.Float32(1e-3))
    p = pkg.api.make_params(m)
    p.field_is_f64 = int(f64)
    for it in range(8):
        if it == 3:
            h.enable_timing(True)  # resets the ring: average over the last 5 iterations only
        if fused and it > 0:
            h.extract_async(p, t.data_ptr(), n, n, n, t.stride(1), verts.data_ptr(), nv, faces.data_ptr(), nf)
        else:
            nv, nf, f64 = h.count(p, t.data_ptr(), pkg.capi.DEVICE, n, n, n, t.stride(1))
            verts = torch.empty((nv, 3), dtype=torch.float64 if f64 else torch.float32, device="cuda")
            faces = torch.empty((nf, 3), dtype=torch.int64, device="cuda")
            h.generate(verts.data_ptr(), faces.data_ptr(), pkg.capi.DEVICE, 0)
        tm = h.timings()
    tot = tm["classify_ms"] + tm["count_scan_ms"] + tm["generate_ms"]
    if fused:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h.set_stream(torch.cuda.current_stream().cuda_stream)
        e0.record()
        for _ in range(10):
            h.extract_async(p, t.data_ptr(), n, n, n, t.stride(1), verts.data_ptr(), nv, faces.data_ptr(), nf)
        e1.record(); torch.cuda.synchronize()
        tot = e0.elapsed_time(e1) / 10
        h.use_own_stream()
    esz = 8 if f64 else 4
    B = esz * n ** 3 + 3 * esz * nv + 24 * nf
    print(f"{algo}{" oneq" if fused else ""} n={n} nv={nv} nf={nf} classify={tm['classify_ms']:.3f} count={tm['count_scan_ms']:.3f} gen={tm['generate_ms']:.3f} "
          f"total={tot:.3f} ms  {(n-1)**3/tot/1e6:.1f} Gvox/s  {B/tot/1e6:.0f} GB/s  classify {esz*n**3/tm['classify_ms']/1e6:.0f} GB/s", flush=True)
    del t, verts, faces
