#!/usr/bin/env python3
"""bench.py -- headline benchmark of the isosurface hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload ...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one full isosurface extraction (classify [+ count inside it] -> count tail + scan -> generate).
  value : whole-job Gvoxels/s with the field already resident in HBM and the mesh written to HBM
          (CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks)
  e2e   : the same metric through the call the drop-in makes -- `isosurface(::Array)` on ordinary PAGEABLE host
          arrays (N = 1: the public Python mirror `isosurface(field, method)`, output allocation included; N > 1: the
          same C call, b200iso_extract_host, on every rank's slab): H2D of the field and D2H of the mesh are inside
          the timed region
  roofline, cpu_baseline: see DESIGN.md "Measurement".
Workloads (config.workload):
  mc_gyroid   MarchingCubes(iso=0f0) on the Float32 gyroid (BASELINE configs[3]).  N GPUs: `value` is the weak-scaling
              form -- an (n*N) x n x n volume, one n-plane x-slab per GPU, n = 1024 -- and the line additionally
              carries `strong` (configs[3] read literally: ONE 1024^3 volume over N GPUs) and `configs4`
              (multi-sphere/torus, (256*N) x 2048 x 2048: BASELINE configs[4] exactly at N = 8)
  mc_gyroid_strong / mc_m2048 / mt_gyroid   the same measurements as stand-alone workloads
--impl reference times the CPU restatement of the reference loops (oracle/, Julia is not installed) on the
host cores, on bounded x-range samples spread over the same field.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


# ONE metric string for both arms (--impl b200 / reference), every workload and every N: the driver pairs lines by it.
# The workload, shape and method live in `config`.
METRIC = "isosurface extraction throughput (Gvoxels/s; Mtriangles/s alongside)"
UNIT = "Gvoxels/s"
DTYPE = "f32 field, f64 positions, f32 vertices, int64 faces"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="mc_gyroid", choices=["mc_gyroid", "mc_gyroid_strong", "mt_gyroid", "mc_m2048"])
    ap.add_argument("--n", type=int, default=0, help="override the base grid size (development)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="N > 1: skip the strong / configs4 measurements")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="sharded runs: how the slab totals travel")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target duration of the CPU baseline sample")
    return ap.parse_args()


def workload_spec(name, n_override, world):
    """-> dict(global shape, ranges for the synthetic field, algo, description)"""
    if name == "mc_gyroid":
        n = n_override or 1024
        return dict(name=name, algo="MC", shape=(n * world, n, n), kind="gyroid", n=n, scaling="weak",
                    desc=f"MarchingCubes(iso=0f0), Float32 gyroid cos x sin y + cos y sin z + cos z sin x, "
                         f"{n * world}x{n}x{n} samples on [0,4pi*{world}]x[0,4pi]^2, x-slabs of {n} voxel planes per GPU")
    if name == "mc_gyroid_strong":  # BASELINE configs[3] read literally: ONE 1024^3 volume sharded over the GPUs
        n = n_override or 1024
        return dict(name=name, algo="MC", shape=(n, n, n), kind="gyroid", n=n, scaling="strong",
                    desc=f"MarchingCubes(iso=0f0), Float32 gyroid, one {n}x{n}x{n} volume on [0,4pi]^3 split into "
                         f"{world} x-slabs of {n // world} voxel planes (strong scaling)")
    if name == "mt_gyroid":
        n = n_override or 512
        return dict(name=name, algo="MT", shape=(n * world, n, n), kind="gyroid", n=n, scaling="weak",
                    desc=f"MarchingTetrahedra(iso=0f0, eps=1f-3), Float32 gyroid {n * world}x{n}x{n} samples on "
                         f"[0,4pi*{world}]x[0,4pi]^2, x-slabs of {n} voxel planes per GPU (+1 ghost row)")
    n = n_override or 2048
    return dict(name=name, algo="MC", shape=(n // 8 * world, n, n), kind="mst", n=n, scaling="weak",
                desc=f"MarchingCubes(iso=0f0), multi-sphere/torus SDF (K=32, SplitMix64 seed 0x5EED2048), "
                     f"{n // 8 * world}x{n}x{n} samples, x-slabs of {n // 8} voxel planes per GPU"
                     + (" = BASELINE configs[4]" if world == 8 and n == 2048 else ""))


def slab_range(spec, rank, world):
    """sample x-range [xa, xb) of this rank's slab (one halo plane at high x except for the last rank)"""
    nxg = spec["shape"][0]
    per = nxg // world
    xa = rank * per
    xb = nxg if rank == world - 1 else (rank + 1) * per + 1
    if spec["algo"] == "MT" and rank > 0:
        xa -= 1  # Marching Tetrahedra: the previous slab's last voxel row rides along as a ghost row
    return xa, xb


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, str(e)

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                break
            time.sleep(0.002)

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, workload):
    """dram bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic.json)"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            tab = json.load(open(p)).get(workload, {})
            vals = [tab.get(k) for k in kernel.split("+")]  # (a stage of two kernels: the sum)
            return None if any(v is None for v in vals) else float(sum(vals))
        except Exception:
            return None
    return None


def build_field(pkg, spec, xa, xb, device, ldx=None):
    import torch
    if spec["kind"] == "gyroid":
        nxg = spec["shape"][0]
        world_x = max(nxg // spec["n"], 1)
        tabs = pkg.synth.gyroid_tables(spec["shape"], 0.0, 4.0 * np.pi)
        # the x axis spans [0, 4*pi*world]: recompute the x tables on that extent
        i = np.arange(nxg, dtype=np.float64)
        x = 0.0 + (4.0 * np.pi * world_x - 0.0) * i / max(nxg - 1, 1)
        tabs = ((np.cos(x), np.sin(x)), tabs[1], tabs[2])
        return pkg.synth.gyroid_torch(spec["shape"], device, x_slice=(xa, xb), tables=tabs, ldx=ldx)
    return pkg.synth.multisphere_torus(spec["shape"], x_slice=(xa, xb), xp=torch, device=device, ldx=ldx)


def host_planes(pkg, spec, groups):
    """Host field with the full leading dimensions (same strides as the full sweep), only the x-plane groups
    [a, b] (inclusive sample planes) filled -- the rest is never read by the x-range samples and never touched."""
    nx, ny, nz = spec["shape"]
    a = np.zeros((nx, ny, nz), dtype=np.float32, order="F")
    for lo, hi in groups:
        if spec["kind"] == "gyroid":
            world_x = max(nx // spec["n"], 1)
            tabs = pkg.synth.gyroid_tables(spec["shape"], 0.0, 4.0 * np.pi)
            i = np.arange(nx, dtype=np.float64)
            x = (4.0 * np.pi * world_x) * i / max(nx - 1, 1)
            tabs = ((np.cos(x), np.sin(x)), tabs[1], tabs[2])
            a[lo:hi + 1] = pkg.synth.gyroid(spec["shape"], 0.0, 4.0 * np.pi, x_slice=(lo, hi + 1), tables=tabs)
        else:
            a[lo:hi + 1] = pkg.synth.multisphere_torus(spec["shape"], x_slice=(lo, hi + 1))
    return a


def oracle_sweep(oracle, field, spec, ranges, threads):
    """restated reference sweep over the voxel x-plane ranges [(lo, hi), ...]; -> (seconds, voxels)"""
    nx, ny, nz = field.shape
    algo = 0 if spec["algo"] == "MC" else 1
    t0 = time.perf_counter()
    vox = 0
    for lo, hi in ranges:
        oracle.isosurface(field, algo, iso=0.0, iso_is_f32=True, eps=1e-3, eps_is_f32=True, nthreads=threads, xrange=(lo, hi), copy=False)
        vox += (hi - lo) * (ny - 1) * (nz - 1)
    return time.perf_counter() - t0, vox


def spread_ranges(nx, planes, groups=4):
    """`planes` voxel x-planes as `groups` ranges spread over the whole x extent (the surface density of the
    multi-sphere/torus field varies along x: a sample of the first planes only would not be representative)"""
    groups = max(1, min(groups, planes))
    per = max(1, planes // groups)
    out = []
    for g in range(groups):
        lo = int((nx - 1 - per) * g / max(groups - 1, 1)) if groups > 1 else 0
        out.append((lo, min(lo + per, nx - 1)))
    return out


def run_reference(args, rank):
    """Reference arm: the CPU restatement of the reference's loops on the host cores (Julia unavailable)."""
    if rank != 0:
        return
    from __graft_entry__ import load_package
    from oracle import harness as oracle
    oracle.build()
    pkg = load_package()
    spec = workload_spec(args.workload, args.n, max(args.gpus, 1))
    threads = os.cpu_count() or 1  # (MT: every x-slab thread keeps its own vertex dictionary -- a throughput measure)
    nx, ny, nz = spec["shape"]
    # size the per-step sample: probe a few planes per thread, then aim for ~2 s per step
    per_group = max(threads, 2)
    groups = spread_ranges(nx, 4 * per_group, 4)
    field = host_planes(pkg, spec, [(lo, hi) for lo, hi in groups])
    t, vox = oracle_sweep(oracle, field, spec, groups, threads)
    scale = min(8.0, max(1.0, 2.0 / max(t, 1e-6)))
    if scale >= 1.5:
        groups = spread_ranges(nx, int(4 * per_group * scale) // threads * threads or threads, 4)
        field = host_planes(pkg, spec, [(lo, hi) for lo, hi in groups])
    for _ in range(max(args.warmup, 1)):
        oracle_sweep(oracle, field, spec, groups, threads)
    t0 = time.perf_counter()
    vox = 0
    for _ in range(args.steps):
        _, v = oracle_sweep(oracle, field, spec, groups, threads)
        vox += v
    dt = (time.perf_counter() - t0) / args.steps
    vox //= args.steps
    value = vox / dt / 1e9
    # the reference itself is single-threaded (src/marching_cubes.jl:40): the same sweep on one thread, one group
    t1, v1 = oracle_sweep(oracle, field, spec, groups[:1], 1)
    sample = (f"voxel x-planes {groups} of the {nx}x{ny}x{nz} field ({vox} voxels per step, spread over the x extent), full-field strides")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": spec.get("scaling", "weak"), "vs_baseline": None, "dtype": DTYPE,
            "data": "synthetic", "config": {"workload": args.workload, "description": spec["desc"], "shape": [nx, ny, nz],
                                            "sample": sample, "host_threads": threads},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "single_thread": {"value": v1 / t1 / 1e9, "unit": UNIT, "cores": 1,
                                               "note": "the reference is single-threaded (src/marching_cubes.jl:40); all-core figure = x-slab threads"},
                             "note": "C++ restatement of Meshing.jl's loops (oracle/iso_oracle.cpp); Julia is not installed"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


class Ctx:
    pass


def device_run(ctx, spec, steps, warmup, want_clocks=False):
    """Device-resident timing of one workload on this rank's slab.  -> dict of measurements (all ranks), keeps the
    handle / field / outputs in the returned dict for the e2e leg."""
    import torch
    import torch.distributed as dist
    pkg, capi, device, rank, world = ctx.pkg, ctx.pkg.capi, ctx.device, ctx.rank, ctx.world
    nxg, ny, nz = spec["shape"]
    xa, xb = slab_range(spec, rank, world)
    nxl = xb - xa
    ldx = (nxl + 3) // 4 * 4
    field = build_field(pkg, spec, xa, xb, device, ldx=ldx)
    torch.cuda.synchronize()
    method = (pkg.MarchingCubes(iso=pkg.Float32(0)) if spec["algo"] == "MC"
              else pkg.MarchingTetrahedra(iso=pkg.Float32(0), eps=pkg.Float32(1e-3)))
    params = pkg.api.make_params(method)
    sharded = world > 1
    if sharded:
        params.x_offset, params.nx_global = xa, nxg  # slab vertices get the coordinates of the unsharded volume
        params.x_ghost = int(spec["algo"] == "MT" and rank > 0)
    h = ctx.handle
    stream = torch.cuda.current_stream()
    h.set_stream(stream.cuda_stream)
    # sizing pass (count -> allocate -> generate)
    nv, nf, f64 = h.count(params, field.data_ptr(), capi.DEVICE, nxl, ny, nz, ldx)
    verts = torch.empty((max(nv, 1), 3), dtype=torch.float64 if f64 else torch.float32, device=device)
    faces = torch.empty((max(nf, 1), 3), dtype=torch.int64, device=device)
    totals = torch.zeros(2, dtype=torch.int64, device=device)
    gathered = torch.zeros((world, 2), dtype=torch.int64, device=device)
    vbase = torch.zeros(1, dtype=torch.int64, device=device)
    px = ctx.px

    def step():
        # classify (+ count) -> count tail + scan -> [exchange of the slab totals] -> generate; all asynchronous
        h.count_async(params, field.data_ptr(), nxl, ny, nz, ldx, totals.data_ptr())
        if px is not None:
            base_ptr = px.exchange_async()
        elif sharded:
            dist.all_gather_into_tensor(gathered.view(-1), totals)
            torch.sum(gathered[:rank, 0], dim=0, keepdim=True, out=vbase)
            base_ptr = vbase.data_ptr()
        else:
            base_ptr = 0
        h.generate_async(verts.data_ptr(), verts.shape[0], faces.data_ptr(), faces.shape[0], base_ptr, 0)

    for _ in range(max(warmup, 3)):
        step()
    ctx.barrier()
    launches0 = h.launch_count()
    h.enable_timing(True)
    sampler = ClockSampler(physical_gpu_index(ctx.local_rank)) if want_clocks else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    ctx.barrier()
    clocks = sampler.result() if sampler else None
    ms_total = e0.elapsed_time(e1)
    stage = h.timings()
    claimed = h.ride_claimed() if spec["algo"] == "MC" else 0
    h.enable_timing(False)
    launches = h.launch_count() - launches0
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / steps
    own_rows = nxl - 1 - int(params.x_ghost)
    counts = torch.tensor([nv, nf, own_rows * (ny - 1) * (nz - 1)], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(counts)
    tot_nv, tot_nf, tot_vox = [int(v) for v in counts.tolist()]
    if px is not None:
        vbase.copy_(px.bases[:1])
        # the peer exchange must agree with an NCCL all-gather of the same totals
        dist.all_gather_into_tensor(gathered.view(-1), totals)
        assert torch.equal(px.all, gathered) and int(vbase.item()) == int(gathered[:rank, 0].sum().item())
    return dict(spec=spec, params=params, field=field, verts=verts, faces=faces, nv=nv, nf=nf, f64=f64, nxl=nxl, ny=ny, nz=nz, ldx=ldx,
                ms_step=ms_step, stage=stage, launches=int(launches), clocks=clocks, tot_nv=tot_nv, tot_nf=tot_nf, tot_vox=tot_vox,
                value=tot_vox / (ms_step * 1e-3) / 1e9, mtri=tot_nf / (ms_step * 1e-3) / 1e6, vbase=int(vbase.item()), sharded=sharded,
                claimed=int(claimed), nblocks_est=None)


def summarize(r, steps):
    """the JSON sub-object of an extra device-resident measurement (strong, configs4)"""
    st = r["stage"]
    return {"value": r["value"], "unit": UNIT, "mtriangles_per_s": r["mtri"], "ms_per_step": r["ms_step"], "steps": steps,
            "scaling": r["spec"]["scaling"], "workload": r["spec"]["name"], "description": r["spec"]["desc"],
            "shape": list(r["spec"]["shape"]), "per_gpu_shape": [r["nxl"], r["ny"], r["nz"]],
            "stage_ms": {"classify": st["classify_ms"], "count_scan": st["count_scan_ms"], "generate": st["generate_ms"]},
            "mesh": {"nverts": r["tot_nv"], "nfaces": r["tot_nf"]}}


def pcie_ceiling(ctx, nbytes=1 << 30):
    """Pinned H2D and D2H rates of this rank's GPU while EVERY rank copies at the same time (GB/s, this rank), and the
    job's aggregate: the ceiling of any host-array path on this box."""
    import torch
    import torch.distributed as dist
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dev = torch.empty(nbytes, dtype=torch.uint8, device=ctx.device)
    out = {}
    for name, dst, src in (("h2d", dev, host), ("d2h", host, dev)):
        dst.copy_(src, non_blocking=True)
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        ctx.barrier()
        rate = torch.tensor([nbytes / dt / 1e9], dtype=torch.float64, device=ctx.device)
        agg = rate.clone()
        if ctx.world > 1:
            dist.all_reduce(agg)
        out[name + "_gbs_this_gpu"] = float(rate.item())
        out[name + "_gbs_all_gpus"] = float(agg.item())
    del host, dev
    return out


def e2e_run(ctx, r, steps):
    """End to end through the drop-in's call on PAGEABLE host arrays."""
    import torch
    import torch.distributed as dist
    pkg, capi, h, world = ctx.pkg, ctx.pkg.capi, ctx.handle, ctx.world
    nxl, ny, nz, nv, nf, f64 = r["nxl"], r["ny"], r["nz"], r["nv"], r["nf"], r["f64"]
    spec = r["spec"]
    # an ordinary (pageable) Fortran-ordered host array with the slab's samples, like a Julia Array{Float32,3}
    hfield = np.empty((nz, ny, nxl), dtype=np.float32)
    torch.from_numpy(hfield).copy_(r["field"].permute(2, 1, 0)[:, :, :nxl])
    hfield = hfield.transpose(2, 1, 0)
    assert hfield.flags.f_contiguous
    method = (pkg.MarchingCubes(iso=pkg.Float32(0)) if spec["algo"] == "MC"
              else pkg.MarchingTetrahedra(iso=pkg.Float32(0), eps=pkg.Float32(1e-3)))
    h.use_own_stream()
    vsz = 8 if f64 else 4
    ksteps = max(3, min(steps, 6))
    res = {}
    if world == 1:
        def call():
            return pkg.isosurface(hfield, method)  # the public call: guess -> one-shot pipeline -> trim (api.py)
        api = "isosurface(field::pageable ndarray, method) -- the Python mirror of the drop-in: b200iso_extract_host into arrays sized by the previous call's totals (first call: surface-area estimate), output allocation included"
    else:
        params = r["params"]
        hv = np.empty((nv + nv // 8 + 1024, 3), dtype=np.float64 if f64 else np.float32)
        hf = np.empty((nf + nf // 8 + 1024, 3), dtype=np.int64)

        def call():
            a, b, _, fits = h.extract_host(params, hfield.ctypes.data, nxl, ny, nz, nxl, hv.ctypes.data, hv.shape[0], hf.ctypes.data, hf.shape[0])
            assert fits and (a, b) == (nv, nf)
            return hv[:a], hf[:b]
        api = "b200iso_extract_host on every rank's pageable slab (slab-local face indices; the vertex base of the 16-byte exchange is added by the caller)"
    v, f = call()  # warm-up (first call of the N = 1 form: estimate instead of memo)
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(ksteps):
        v, f = call()
    ctx.barrier()
    dt = torch.tensor([(time.perf_counter() - t0) / ksteps], dtype=torch.float64, device=ctx.device)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt.item())
    # sanity: the host result equals the device-resident result
    assert len(v) == nv and len(f) == nf
    assert torch.equal(torch.from_numpy(np.ascontiguousarray(f)), (r["faces"][:nf] - r["vbase"]).cpu() if r["sharded"] else r["faces"][:nf].cpu())
    assert torch.equal(torch.from_numpy(np.ascontiguousarray(v)), r["verts"][:nv].cpu())
    pageable = {"value": r["tot_vox"] / dt / 1e9, "ms_per_step": dt * 1e3, "api": api, "host_memory": "pageable"}
    # the same C call with pinned caller arrays (cudaHostAlloc): direct DMA, no staging threads
    pf = torch.empty((nz, ny, nxl), dtype=torch.float32).pin_memory()
    pf.copy_(torch.from_numpy(np.ascontiguousarray(hfield.transpose(2, 1, 0))))
    pv = torch.empty((nv, 3), dtype=torch.float64 if f64 else torch.float32).pin_memory()
    pff = torch.empty((nf, 3), dtype=torch.int64).pin_memory()
    params = r["params"]

    def pinned_call():
        a, b, _, fits = h.extract_host(params, pf.data_ptr(), nxl, ny, nz, nxl, pv.data_ptr(), nv, pff.data_ptr(), nf)
        assert fits
    pinned_call()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(ksteps):
        pinned_call()
    ctx.barrier()
    dtp = torch.tensor([(time.perf_counter() - t0) / ksteps], dtype=torch.float64, device=ctx.device)
    if world > 1:
        dist.all_reduce(dtp, op=dist.ReduceOp.MAX)
    dtp = float(dtp.item())
    assert torch.equal(pff, (r["faces"][:nf] - r["vbase"]).cpu() if r["sharded"] else r["faces"][:nf].cpu())
    pinned = {"value": r["tot_vox"] / dtp / 1e9, "ms_per_step": dtp * 1e3, "host_memory": "pinned",
              "api": "b200iso_extract_host, caller arrays pinned (cudaHostAlloc): direct DMA, no staging threads"}
    del pf, pv, pff
    # N = 1: the headline is the drop-in call as a Julia/numpy user makes it (pageable arrays).  N > 1: the reference has
    # no distributed API, the sharded host path is this repo's own (every rank owns its slab buffers): pinned slabs.
    head, other = (pageable, pinned) if world == 1 else (pinned, pageable)
    res = {"value": head["value"], "unit": UNIT, "ms_per_step": head["ms_per_step"], "steps": ksteps,
           "h2d_bytes_per_step": 4 * nxl * ny * nz, "d2h_bytes_per_step": 3 * vsz * nv + 24 * nf + 16,
           "host_memory": head["host_memory"], "api": head["api"], ("pinned" if world == 1 else "pageable"): other}
    dt = head["ms_per_step"] * 1e-3
    res["pcie_ceiling"] = pc = pcie_ceiling(ctx)
    h2d_rate = res["h2d_bytes_per_step"] / dt / 1e9
    res["h2d_gbs_this_gpu"] = h2d_rate
    res["frac_of_pinned_h2d_ceiling"] = h2d_rate / max(pc["h2d_gbs_this_gpu"], 1e-9)
    # what the box allows for this step's bytes at the rates measured above with every GPU copying at once:
    # both directions fully overlapped (duplex) / one after the other (sequential)
    t_up = res["h2d_bytes_per_step"] / max(pc["h2d_gbs_this_gpu"], 1e-9) / 1e6
    t_dn = res["d2h_bytes_per_step"] / max(pc["d2h_gbs_this_gpu"], 1e-9) / 1e6
    res["pcie_bound_ms"] = {"duplex": max(t_up, t_dn), "sequential": t_up + t_dn,
                            "note": "rank 0's bytes at rank 0's concurrent pinned H2D / D2H rates"}
    return res, hfield


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package

    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one process per GPU)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    ctx = Ctx()
    ctx.pkg = pkg = load_package()
    ctx.device, ctx.rank, ctx.world, ctx.local_rank = device, rank, world, local_rank
    ctx.handle = h = pkg.capi.Handle(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    ctx.barrier = barrier

    # the one exchange of the sharded path: every slab's (nverts, nfaces), 16 bytes per rank; the exclusive prefix is
    # the slab's global vertex base, added to its face indices inside generate.  Preferred form: peer-memory stores
    # over NVLink on the handle's own stream (b200iso_exchange_async); NCCL all-gather if the ranks cannot map each
    # other's memory (or --exchange nccl).
    ctx.px, exchange = None, "none"
    if world > 1:
        exchange = "nccl all-gather"
        if args.exchange == "peer":
            try:
                ctx.px = pkg.sharding.PeerExchange(h, device=device)
                exchange = "NVLink peer-memory stores (b200iso_exchange_async)"
                ok = torch.tensor([1], device=device)
            except Exception as e:  # pragma: no cover
                ok = torch.tensor([0], device=device)
                exchange = f"nccl all-gather (peer mapping failed: {type(e).__name__})"
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # all ranks take the same path
            if not bool(ok.item()) and ctx.px is not None:
                ctx.px.close()
                ctx.px, exchange = None, "nccl all-gather (peer mapping failed on another rank)"

    spec = workload_spec(args.workload, args.n, world)
    r = device_run(ctx, spec, args.steps, args.warmup, want_clocks=True)
    nxl, ny, nz, nv, nf, f64 = r["nxl"], r["ny"], r["nz"], r["nv"], r["nf"], r["f64"]

    e2e, hfield = None, None
    if not args.no_e2e:
        e2e, hfield = e2e_run(ctx, r, args.steps)

    extras = {}
    if world > 1 and args.workload == "mc_gyroid" and not args.no_extra:
        # free the primary workload's device memory first
        keep_cpu = hfield
        r_small = {k: v for k, v in r.items() if k not in ("field", "verts", "faces")}
        r = r_small
        torch.cuda.empty_cache()
        for key, wname in (("strong", "mc_gyroid_strong"), ("configs4", "mc_m2048")):
            sp = workload_spec(wname, args.n if wname != "mc_m2048" else (args.n * 2 if args.n else 0), world)
            rr = device_run(ctx, sp, args.steps, args.warmup)
            extras[key] = summarize(rr, args.steps)
            del rr
            torch.cuda.empty_cache()
        hfield = keep_cpu

    if rank == 0:
        peak, peak_src = measured_peak()
        vbytes = 24 if f64 else 12
        stage = r["stage"]
        mc = spec["algo"] == "MC"
        alg = {"classify": 4.0 * nxl * ny * nz, "count_scan": 0.0, "generate": float(vbytes * nv + 24 * nf)}
        w16 = ((nz + 31) // 32 + 15) // 16
        tma = ((nxl + 127) // 128) * ny * w16 >= 4096 and os.environ.get("B200ISO_TMA", "1") != "0"  # the library's rule
        kname = {"classify": "signpack_tma_kernel" if tma else "signpack_kernel",
                 "count_scan": ("mc_count_chunks_kernel+mc_scan_chunks_kernel" if mc else "mt_count_kernel+mt_scan_blocks_kernel"),
                 "generate": "mc_generate_kernel" if mc else "mt_generate_kernel"}
        stage_ms = {"classify": stage["classify_ms"], "count_scan": stage["count_scan_ms"], "generate": stage["generate_ms"]}
        dom = max(stage_ms, key=lambda k: stage_ms[k])
        achieved = alg[dom] / (stage_ms[dom] * 1e-3) / 1e9 if stage_ms[dom] > 0 else 0.0
        bytes_step = 4.0 * nxl * ny * nz + vbytes * nv + 24 * nf  # rank 0's slab
        pipe_gbs = bytes_step / (r["ms_step"] * 1e-3) / 1e9
        notes = {}
        if mc and tma:
            notes["signpack_tma_kernel"] = (f"carries the Marching Cubes count: its counting warps counted {r['claimed']} generate blocks "
                                            "while the field streamed; count_scan is the tail they left + the scan")
        line = {
            "metric": METRIC,
            "value": r["value"], "unit": UNIT, "mtriangles_per_s": r["mtri"],
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": r["ms_step"],
            "higher_is_better": True, "scaling": spec.get("scaling", "weak"), "vs_baseline": None,
            "dtype": DTYPE if not f64 else "f32 field, f64 positions and vertices, int64 faces",
            "data": "synthetic",
            "config": {"workload": args.workload, "description": spec["desc"], "shape": list(spec["shape"]), "per_gpu_shape": [nxl, ny, nz],
                       "algo": spec["algo"],
                       "sharding": ("x-slabs + one 16-byte exchange of counts" if world > 1 else "single GPU"),
                       "exchange": exchange,
                       "l2": "inputs larger than L2 (field %.2f GB per GPU, read once per step)" % (4.0 * nxl * ny * nz / 1e9),
                       "mesh": {"nverts": r["tot_nv"], "nfaces": r["tot_nf"]}},
            "roofline": {"bound": "hbm", "kernel": kname[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_src, "traffic": ncu_traffic(kname[dom], args.workload),
                         "algorithmic_bytes_per_launch": alg[dom], "kernel_ms": stage_ms[dom],
                         "stage_ms": stage_ms, "notes": notes,
                         "kernels": {kname[k]: {"ms": stage_ms[k], "algorithmic_bytes": alg[k],
                                                "achieved": (alg[k] / (stage_ms[k] * 1e-3) / 1e9 if stage_ms[k] > 0 else 0.0),
                                                "frac": (alg[k] / (stage_ms[k] * 1e-3) / 1e9 / peak if stage_ms[k] > 0 else 0.0),
                                                "traffic": ncu_traffic(kname[k], args.workload)} for k in stage_ms},
                         "pipeline": {"algorithmic_bytes_per_step": bytes_step, "achieved": pipe_gbs, "frac_of_measured": pipe_gbs / peak,
                                      "frac_of_nominal_8TBs": pipe_gbs / 8000.0}},
            "gpu_launches": r["launches"],
            "clocks": r["clocks"],
        }
        if e2e is not None:
            line["e2e"] = e2e
        line.update(extras)
        if world == 1 and not args.no_cpu_baseline:
            from oracle import harness as oracle
            oracle.build()
            host = hfield if hfield is not None else host_planes(pkg, spec, [(0, 65)])
            nxh = host.shape[0]
            # 1 thread (the reference is single-threaded): probe, then a sample of about --cpu-seconds
            t, vox = oracle_sweep(oracle, host, spec, [(0, 2)], 1)
            planes = int(max(2, min(nxh - 1, args.cpu_seconds / max(t / 2, 1e-6))))
            groups = spread_ranges(nxh, planes, 4) if hfield is not None else [(0, min(planes, 64))]
            t, vox = oracle_sweep(oracle, host, spec, groups, 1)
            line["cpu_baseline"] = {"value": vox / t / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"voxel x-planes {groups} of the same {nxl}x{ny}x{nz} field ({vox} voxels, {t:.1f} s), full-field strides",
                                    "note": "C++ restatement of Meshing.jl's single-threaded loops (oracle/iso_oracle.cpp); "
                                            "Julia is not installed in this image"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
