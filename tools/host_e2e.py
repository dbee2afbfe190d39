#!/usr/bin/env python3
"""Development aid: wall time of the HOST entry points on pinned and pageable arrays (1024^3 gyroid by default).
  python tools/host_e2e.py [n] [MC|MT]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from __graft_entry__ import load_package
pkg = load_package()
capi = pkg.capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
algo = sys.argv[2] if len(sys.argv) > 2 else "MC"
m = pkg.MarchingCubes(iso=pkg.Float32(0)) if algo == "MC" else pkg.MarchingTetrahedra(iso=pkg.Float32(0), eps=pkg.Float32(1e-3))
p = pkg.api.make_params(m)
h = capi.Handle(0)
dev = pkg.synth.gyroid_torch(n, "cuda")                     # (n, n, n) with x stride 1
pin = torch.empty((n, n, n), dtype=torch.float32).pin_memory()
pin.copy_(dev.permute(2, 1, 0))                              # memory order == Fortran order of the field
pag = np.array(pin.numpy(), copy=True)
nv, nf, _ = h.count(p, pin.data_ptr(), capi.HOST, n, n, n, n)
print(f"{algo} {n}^3: {nv} verts {nf} faces; host threads: B200ISO_HOST_THREADS={os.environ.get('B200ISO_HOST_THREADS', 'default')}")
outs = {"pinned": (torch.empty((nv, 3), dtype=torch.float32).pin_memory().numpy(), torch.empty((nf, 3), dtype=torch.int64).pin_memory().numpy()),
        "pageable": (np.empty((nv, 3), np.float32), np.empty((nf, 3), np.int64))}
ins = {"pinned": pin.numpy(), "pageable": pag}
ref = None
for kind in ("pinned", "pageable"):
    a, (v, f) = ins[kind], outs[kind]
    def pair():
        h.count(p, a.ctypes.data, capi.HOST, n, n, n, n)
        h.generate(v.ctypes.data, f.ctypes.data, capi.HOST, 0)
    def oneshot():
        r = h.extract_host(p, a.ctypes.data, n, n, n, n, v.ctypes.data, nv, f.ctypes.data, nf)
        assert r[3] and r[:2] == (nv, nf)
    for name, fn in (("count+generate", pair), ("extract_host", oneshot)):
        f[:] = 0
        fn()
        ts = []
        for _ in range(4):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        if ref is None:
            ref = (v.copy(), f.copy())
        ok = np.array_equal(f, ref[1]) and np.array_equal(v.view(np.uint32), ref[0].view(np.uint32))
        t = min(ts)
        print(f"  {kind:9s} {name:15s} {t*1e3:8.1f} ms  {(n-1)**3/t/1e9:6.2f} Gvox/s  {(4*n**3 + 12*nv + 24*nf)/t/1e9:6.1f} GB/s over PCIe  same_bytes={ok}")
