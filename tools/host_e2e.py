#!/usr/bin/env python3
"""Host-array paths at 1024^3 (pageable arrays): where the milliseconds go.  usage: host_e2e.py [n]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from __graft_entry__ import load_package
pkg = load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
t = pkg.synth.gyroid_torch(n, "cuda")
hf = np.empty((n, n, n), dtype=np.float32)
torch.from_numpy(hf).copy_(t.permute(2, 1, 0))
hf = hf.transpose(2, 1, 0)
del t
torch.cuda.empty_cache()
m = pkg.MarchingCubes(iso=pkg.Float32(0))
p = pkg.api.make_params(m)
h = pkg.api.get_handle(0)

def timeit(name, fn, k=4):
    fn()
    t0 = time.perf_counter()
    for _ in range(k):
        fn()
    print(f"{name:70s} {(time.perf_counter() - t0) / k * 1e3:8.1f} ms  (threads {os.environ.get('B200ISO_HOST_THREADS', 'default')})", flush=True)

v, f = pkg.isosurface(hf, m)
nv, nf = len(v), len(f)
if len(sys.argv) > 2 and sys.argv[2] == "sweep":
    # usage: host_e2e.py 1024 sweep "A=1 B=2" "A=3" ...: the public call under each environment, in one process
    # (the library reads its B200ISO_HOST_* knobs per call)
    for cfg in sys.argv[3:]:
        kv = dict(x.split("=") for x in cfg.split())
        os.environ.update(kv)
        timeit(f"isosurface(pageable) [{cfg}]", lambda: pkg.isosurface(hf, m), k=5)
        timeit(f"  upload + count only     [{cfg}]", lambda: h.extract_host(p, hf.ctypes.data, n, n, n, n, 0, 0, 0, 0), k=3)
        for k_ in kv:
            del os.environ[k_]
    sys.exit(0)
timeit("public isosurface(pageable), fresh output arrays every call", lambda: pkg.isosurface(hf, m))
pv, pf = np.empty((nv + 1024, 3), np.float32), np.empty((nf + 1024, 3), np.int64)
pv[:] = 0; pf[:] = 0
timeit("b200iso_extract_host, pageable in, pre-touched pageable out", lambda: h.extract_host(p, hf.ctypes.data, n, n, n, n, pv.ctypes.data, len(pv), pf.ctypes.data, len(pf)))
timeit("b200iso_extract_host, pageable in, count only (no D2H)", lambda: h.extract_host(p, hf.ctypes.data, n, n, n, n, 0, 0, 0, 0))
timeit("two-phase count + generate, pre-touched pageable out", lambda: (h.count(p, hf.ctypes.data, pkg.capi.HOST, n, n, n, n), h.generate(pv.ctypes.data, pf.ctypes.data, pkg.capi.HOST, 0)))
def fresh():
    a = np.empty((nv + nv // 8, 3), np.float32); b = np.empty((nf + nf // 8, 3), np.int64)
    a[:nv] = 0; b[:nf] = 0
timeit("np.empty + first touch of the output arrays alone (1 thread)", fresh)
