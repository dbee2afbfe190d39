#!/usr/bin/env python3
"""Development probe: how well do classify+count (HBM-bound) and generate (issue/latency-bound) co-run on one GPU?
Two handles on two streams: A classifies+counts the field while B generates from bits it counted before.
  python tools/overlap_probe.py [n]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
pkg = load_package()
capi = pkg.capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
m = pkg.MarchingCubes(iso=pkg.Float32(0))
p = pkg.api.make_params(m)
t = pkg.synth.gyroid_torch(n, "cuda")
hA, hB = capi.Handle(0), capi.Handle(0)
nv, nf, _ = hB.count(p, t.data_ptr(), capi.DEVICE, n, n, n, t.stride(1))
hA.count(p, t.data_ptr(), capi.DEVICE, n, n, n, t.stride(1))
verts = torch.empty((nv, 3), dtype=torch.float32, device="cuda")
faces = torch.empty((nf, 3), dtype=torch.int64, device="cuda")
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
hA.set_stream(sA.cuda_stream); hB.set_stream(sB.cuda_stream)
def run(mode, reps=10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if mode == "serial":
            hA.count_async(p, t.data_ptr(), n, n, n, t.stride(1))
            hA.generate_async(verts.data_ptr(), nv, faces.data_ptr(), nf)
        elif mode == "classify+count":
            hA.count_async(p, t.data_ptr(), n, n, n, t.stride(1))
        elif mode == "generate":
            hB.generate_async(verts.data_ptr(), nv, faces.data_ptr(), nf)
        else:  # overlapped: A's classify+count next to B's generate
            hA.count_async(p, t.data_ptr(), n, n, n, t.stride(1))
            hB.generate_async(verts.data_ptr(), nv, faces.data_ptr(), nf)
            # keep the pairs aligned: the next pair starts when both are done
            e1, e2 = torch.cuda.Event(), torch.cuda.Event()
            e1.record(sA); e2.record(sB); sA.wait_event(e2); sB.wait_event(e1)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
for mode in ("serial", "classify+count", "generate", "overlap"):
    run(mode, 3)
    print(f"TMA={os.environ.get('B200ISO_TMA','auto')} n={n} {mode:15s} {run(mode):7.3f} ms")
