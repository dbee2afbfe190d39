#!/usr/bin/env python3
"""Counting warps inside the TMA classify kernel: stage times and claimed items against the number of warps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
pkg = load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
kind = sys.argv[2] if len(sys.argv) > 2 else "gyroid"
if kind == "gyroid":
    t = pkg.synth.gyroid_torch(n, "cuda")
elif kind == "gyroid129":  # one rank's slab of the 1024^3 volume split over 8 GPUs
    t = pkg.synth.gyroid_torch(n, "cuda", x_slice=(0, n // 8 + 1), ldx=n // 8 + 4)
elif kind == "gyroid257":  # one rank's slab of the 1024^3 volume split over 4 GPUs (also a slab of the host pipeline)
    t = pkg.synth.gyroid_torch(n, "cuda", x_slice=(0, n // 4 + 1), ldx=n // 4 + 4)
else:  # rank 3's slab of the 2048^3 multi-sphere/torus volume split over 8 GPUs
    t = pkg.synth.multisphere_torus((n, n, n), x_slice=(3 * n // 8, 4 * n // 8 + 1), xp=torch, device="cuda", ldx=n // 8 + 4)
nx = t.shape[0]
p = pkg.api.make_params(pkg.MarchingCubes(iso=pkg.Float32(0)))
h = pkg.capi.Handle(0)
nv, nf, _ = h.count(p, t.data_ptr(), pkg.capi.DEVICE, nx, n, n, t.stride(1))
verts = torch.empty((nv, 3), dtype=torch.float32, device="cuda")
faces = torch.empty((nf, 3), dtype=torch.int64, device="cuda")
stream = torch.cuda.current_stream()
h.set_stream(stream.cuda_stream)
for warps in [int(a) for a in sys.argv[3:]] or [0, 2, 4, 8]:
    h.set_ride_warps(warps)
    for _ in range(3):
        h.count_async(p, t.data_ptr(), nx, n, n, t.stride(1), 0)
        h.generate_async(verts.data_ptr(), nv, faces.data_ptr(), nf, 0, 0)
    torch.cuda.synchronize()
    h.enable_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(20):
        h.count_async(p, t.data_ptr(), nx, n, n, t.stride(1), 0)
        h.generate_async(verts.data_ptr(), nv, faces.data_ptr(), nf, 0, 0)
    e1.record(stream)
    torch.cuda.synchronize()
    tm = h.timings()
    h.enable_timing(False)
    print(f"ride warps {warps}: step {e0.elapsed_time(e1) / 20:.3f} ms  classify {tm['classify_ms']:.3f} count+scan {tm['count_scan_ms']:.3f} "
          f"generate {tm['generate_ms']:.3f}  counted inside classify: {h.ride_claimed()} generate blocks", flush=True)
