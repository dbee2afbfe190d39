#!/usr/bin/env python3
"""How fast is cudaHostRegister on this box?  (would in-place pinning of a caller's pageable array beat staged copies?)"""
import time, threading, numpy as np, torch
rt = torch.cuda.cudart()
torch.cuda.init()
n = 1 << 30
a = np.ones(4 * n, dtype=np.uint8)  # 4 GiB, touched
for flags, name in ((0, "default"), (8, "read-only")):
    t0 = time.perf_counter(); rc = rt.cudaHostRegister(a.ctypes.data, a.nbytes, flags); t1 = time.perf_counter()
    rt.cudaHostUnregister(a.ctypes.data); t2 = time.perf_counter()
    print(f"cudaHostRegister 4 GiB ({name}) rc={int(rc)}: {(t1 - t0) * 1e3:.1f} ms = {a.nbytes / (t1 - t0) / 1e9:.1f} GB/s; unregister {(t2 - t1) * 1e3:.1f} ms", flush=True)
for nth in (4, 8, 16):
    chunk = a.nbytes // nth
    def work(i):
        rt.cudaHostRegister(a.ctypes.data + i * chunk, chunk, 0)
    ths = [threading.Thread(target=work, args=(i,)) for i in range(nth)]
    t0 = time.perf_counter(); [t.start() for t in ths]; [t.join() for t in ths]; t1 = time.perf_counter()
    for i in range(nth): rt.cudaHostUnregister(a.ctypes.data + i * chunk)
    print(f"{nth} threads x {chunk >> 20} MiB: {(t1 - t0) * 1e3:.1f} ms = {a.nbytes / (t1 - t0) / 1e9:.1f} GB/s", flush=True)
