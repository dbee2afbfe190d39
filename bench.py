#!/usr/bin/env python3
"""bench.py -- headline benchmark of the isosurface hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload ...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one full isosurface extraction (classify -> count+scan -> generate) of the workload.
  value : whole-job Gvoxels/s with the field already resident in HBM and the mesh written to HBM
          (CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks)
  e2e   : the same metric through the C-ABI drop-in pair b200iso_count / b200iso_generate with HOST
          (pinned) buffers: H2D of the field and D2H of the mesh are inside the timed region
  roofline, cpu_baseline: see DESIGN.md "Measurement".
Workloads (config.workload):
  mc_gyroid   MarchingCubes(iso=0f0) on the Float32 gyroid; N GPUs hold an (n*N) x n x n volume split in
              x-slabs (x is the scan-outermost axis) -- weak scaling, n = 1024 (BASELINE configs[3])
  mt_gyroid   MarchingTetrahedra(iso=0f0, eps=1f-3) on the 512^3 gyroid (configs[2]); N GPUs: (512*N) x 512 x 512 in
              x-slabs with one ghost voxel row each
  mc_m2048    MarchingCubes on the multi-sphere/torus SDF, (256*N) x 2048 x 2048 (configs[4] at N = 8)
--impl reference times the CPU restatement of the reference loops (oracle/, Julia is not installed) on the
host cores, on a bounded x-range sample of the same field.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="mc_gyroid", choices=["mc_gyroid", "mc_gyroid_strong", "mt_gyroid", "mc_m2048"])
    ap.add_argument("--n", type=int, default=0, help="override the base grid size (development)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="sharded runs: how the slab totals travel")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target duration of the CPU baseline sample")
    return ap.parse_args()


def workload_spec(name, n_override, world):
    """-> dict(global shape, ranges for the synthetic field, algo, description)"""
    if name == "mc_gyroid":
        n = n_override or 1024
        return dict(algo="MC", shape=(n * world, n, n), kind="gyroid", n=n,
                    desc=f"MarchingCubes(iso=0f0), Float32 gyroid cos x sin y + cos y sin z + cos z sin x, "
                         f"{n * world}x{n}x{n} samples on [0,4pi*{world}]x[0,4pi]^2, x-slabs of {n} voxel planes per GPU")
    if name == "mc_gyroid_strong":  # BASELINE configs[3] read literally: ONE 1024^3 volume sharded over the GPUs
        n = n_override or 1024
        return dict(algo="MC", shape=(n, n, n), kind="gyroid", n=n, scaling="strong",
                    desc=f"MarchingCubes(iso=0f0), Float32 gyroid, one {n}x{n}x{n} volume on [0,4pi]^3 split into "
                         f"{world} x-slabs of {n // world} voxel planes (strong scaling)")
    if name == "mt_gyroid":
        n = n_override or 512
        return dict(algo="MT", shape=(n * world, n, n), kind="gyroid", n=n,
                    desc=f"MarchingTetrahedra(iso=0f0, eps=1f-3), Float32 gyroid {n * world}x{n}x{n} samples on "
                         f"[0,4pi*{world}]x[0,4pi]^2, x-slabs of {n} voxel planes per GPU (+1 ghost row)")
    n = n_override or 2048
    return dict(algo="MC", shape=(n // 8 * world, n, n), kind="mst", n=n,
                desc=f"MarchingCubes(iso=0f0), multi-sphere/torus SDF (K=32, SplitMix64 seed 0x5EED2048), "
                     f"{n // 8 * world}x{n}x{n} samples, x-slabs of {n // 8} voxel planes per GPU")


def slab_range(spec, rank, world):
    """sample x-range [xa, xb) of this rank's slab (one halo plane at high x except for the last rank)"""
    nxg = spec["shape"][0]
    if spec.get("replicas"):
        return 0, nxg
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
            return json.load(open(p)).get(workload, {}).get(kernel)
        except Exception:
            return None
    return None


def build_field(pkg, spec, xa, xb, device, ldx=None):
    import torch
    if spec["kind"] == "gyroid":
        nxg = spec["shape"][0]
        world_x = nxg // spec["n"]
        tabs = pkg.synth.gyroid_tables(spec["shape"], 0.0, 4.0 * np.pi)
        # the x axis spans [0, 4*pi*world]: recompute the x tables on that extent
        i = np.arange(nxg, dtype=np.float64)
        x = 0.0 + (4.0 * np.pi * world_x - 0.0) * i / max(nxg - 1, 1)
        tabs = ((np.cos(x), np.sin(x)), tabs[1], tabs[2])
        return pkg.synth.gyroid_torch(spec["shape"], device, x_slice=(xa, xb), tables=tabs, ldx=ldx)
    return pkg.synth.multisphere_torus(spec["shape"], x_slice=(xa, xb), xp=torch, device=device, ldx=ldx)


def cpu_sample(oracle, host_field, spec, seconds, nthreads):
    """Times the restated reference sweep (oracle) on voxel x-planes [0, S) of the host field."""
    nx, ny, nz = host_field.shape
    per_plane = (ny - 1) * (nz - 1)
    t0 = time.perf_counter()
    oracle.isosurface(host_field, 0 if spec["algo"] == "MC" else 1, iso=0.0, iso_is_f32=True, eps=1e-3, eps_is_f32=True,
                      nthreads=nthreads, xrange=(0, max(nthreads, 2)), copy=False)
    probe = (time.perf_counter() - t0) / max(nthreads, 2) * nthreads  # seconds per plane-batch
    planes = int(max(nthreads, min(nx - 1, seconds / max(probe, 1e-6) * 1)))
    planes = max(nthreads, planes // nthreads * nthreads)
    planes = min(planes, nx - 1)
    t0 = time.perf_counter()
    nv, nf = oracle.isosurface(host_field, 0 if spec["algo"] == "MC" else 1, iso=0.0, iso_is_f32=True, eps=1e-3, eps_is_f32=True,
                               nthreads=nthreads, xrange=(0, planes), copy=False)
    dt = time.perf_counter() - t0
    return planes * per_plane / dt / 1e9, planes, dt, nv, nf


def host_field_numpy(pkg, spec, planes):
    """Host field with the full leading dimensions (same strides as the full sweep), only the first
    `planes`+1 x-samples filled -- the rest is never read by the x-range sample."""
    nx, ny, nz = spec["shape"]
    a = np.zeros((nx, ny, nz), dtype=np.float32, order="F")
    if spec["kind"] == "gyroid":
        a[: planes + 1] = pkg.synth.gyroid(spec["shape"], 0.0, 4.0 * np.pi, x_slice=(0, planes + 1))
    else:
        a[: planes + 1] = pkg.synth.multisphere_torus(spec["shape"], x_slice=(0, planes + 1))
    return a


def run_reference(args, rank):
    """Reference arm: the CPU restatement of the reference's loops on the host cores (Julia unavailable)."""
    if rank != 0:
        return
    from __graft_entry__ import load_package
    from oracle import harness as oracle
    oracle.build()
    pkg = load_package()
    spec = workload_spec(args.workload, args.n, 1)
    threads = os.cpu_count() or 1  # (MT: every x-slab thread keeps its own vertex dictionary -- a throughput measure)
    nx, ny, nz = spec["shape"]
    per_plane = (ny - 1) * (nz - 1)
    # size the per-step sample: probe one batch of planes, then aim for ~2 s per step
    probe_planes = max(threads, 2)
    field = host_field_numpy(pkg, spec, min(nx - 1, 16 * probe_planes))
    algo = 0 if spec["algo"] == "MC" else 1

    def sweep(planes):
        t0 = time.perf_counter()
        oracle.isosurface(field, algo, iso=0.0, iso_is_f32=True, eps=1e-3, eps_is_f32=True, nthreads=threads,
                          xrange=(0, planes), copy=False)
        return time.perf_counter() - t0

    t = sweep(probe_planes)
    planes = int(min(16 * probe_planes, max(probe_planes, 2.0 / max(t, 1e-6) * probe_planes)))
    planes = max(threads, planes // threads * threads)
    for _ in range(args.warmup):
        sweep(planes)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sweep(planes)
    dt = (time.perf_counter() - t0) / args.steps
    value = planes * per_plane / dt / 1e9
    sample = f"voxel x-planes [0,{planes}) of the {nx}x{ny}x{nz} field ({planes * per_plane} voxels per step), full-field strides"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": spec.get("scaling", "weak"), "vs_baseline": None, "dtype": "f32 field, f64 positions, f32 vertices, int64 faces",
            "data": "synthetic", "config": {"workload": args.workload, "description": spec["desc"], "shape": [nx, ny, nz], "sample": sample, "host_threads": threads},
            "cpu_baseline": {"value": value, "unit": "Gvoxels/s", "cores": threads, "kind": "port", "sample": sample,
                             "note": "C++ restatement of Meshing.jl's loops (oracle/iso_oracle.cpp); Julia is not installed"},
            "e2e": {"value": value, "unit": "Gvoxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


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
    pkg = load_package()
    capi = pkg.capi
    spec = workload_spec(args.workload, args.n, world)
    nxg, ny, nz = spec["shape"]
    xa, xb = slab_range(spec, rank, world)
    nxl = xb - xa
    ldx = (nxl + 3) // 4 * 4
    field = build_field(pkg, spec, xa, xb, device, ldx=ldx)
    torch.cuda.synchronize()

    method = (pkg.MarchingCubes(iso=pkg.Float32(0)) if spec["algo"] == "MC"
              else pkg.MarchingTetrahedra(iso=pkg.Float32(0), eps=pkg.Float32(1e-3)))
    # physical extent of this slab (only used for vertex positions; the benchmark's ranges are the defaults)
    params = pkg.api.make_params(method)
    if world > 1 and not spec.get("replicas"):
        params.x_offset, params.nx_global = xa, nxg  # slab vertices get the coordinates of the unsharded volume
        params.x_ghost = int(spec["algo"] == "MT" and rank > 0)
    h = capi.Handle(local_rank)
    stream = torch.cuda.current_stream()
    h.set_stream(stream.cuda_stream)

    # sizing pass (what the Julia shim does: count -> allocate -> generate)
    nv, nf, f64 = h.count(params, field.data_ptr(), capi.DEVICE, nxl, ny, nz, ldx)
    verts = torch.empty((max(nv, 1), 3), dtype=torch.float64 if f64 else torch.float32, device=device)
    faces = torch.empty((max(nf, 1), 3), dtype=torch.int64, device=device)
    totals = torch.zeros(2, dtype=torch.int64, device=device)
    gathered = torch.zeros((world, 2), dtype=torch.int64, device=device)
    vbase = torch.zeros(1, dtype=torch.int64, device=device)
    sharded = world > 1 and not spec.get("replicas")

    # the one exchange of the sharded path: every slab's (nverts, nfaces), 16 bytes per rank; the exclusive prefix is
    # the slab's global vertex base, added to its face indices inside generate.  Preferred form: peer-memory stores
    # over NVLink on the handle's own stream (b200iso_exchange_async); NCCL all-gather if the ranks cannot map each
    # other's memory (or --exchange nccl).
    px, exchange = None, "none"
    if sharded:
        exchange = "nccl all-gather"
        if args.exchange == "peer":
            try:
                px = pkg.sharding.PeerExchange(h, device=device)
                exchange = "NVLink peer-memory stores (b200iso_exchange_async)"
            except Exception as e:  # pragma: no cover
                ok = torch.tensor([0], device=device)
                exchange = f"nccl all-gather (peer mapping failed: {type(e).__name__})"
            else:
                ok = torch.tensor([1], device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # all ranks take the same path
            if not bool(ok.item()) and px is not None:
                px.close()
                px, exchange = None, "nccl all-gather (peer mapping failed on another rank)"

    def step():
        # classify -> count + decoupled look-back scan -> [exchange of the slab totals] -> generate; all asynchronous
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = h.launch_count()
    h.enable_timing(True)
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    clocks = sampler.result()
    ms_total = e0.elapsed_time(e1)
    stage = h.timings()
    h.enable_timing(False)
    launches = h.launch_count() - launches0
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps

    # whole-job units
    own_rows = nxl - 1 - int(params.x_ghost)
    counts = torch.tensor([nv, nf, own_rows * (ny - 1) * (nz - 1)], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(counts)
    tot_nv, tot_nf, tot_vox = [int(v) for v in counts.tolist()]
    value = tot_vox / (ms_step * 1e-3) / 1e9
    mtri = tot_nf / (ms_step * 1e-3) / 1e6

    # ---- end-to-end through the drop-in C-ABI pair with pinned host buffers ----
    e2e = None
    if not args.no_e2e:
        hfield = torch.empty((nz, ny, nxl), dtype=torch.float32).pin_memory()
        hfield.copy_(field.permute(2, 1, 0))
        vsz = 8 if f64 else 4
        hverts = torch.empty((max(nv, 1), 3), dtype=torch.float64 if f64 else torch.float32).pin_memory()
        hfaces = torch.empty((max(nf, 1), 3), dtype=torch.int64).pin_memory()
        ksteps = max(3, min(args.steps, 10))

        def pair_step():
            a, b, _ = h.count(params, hfield.data_ptr(), capi.HOST, nxl, ny, nz, nxl)
            assert (a, b) == (nv, nf)
            h.generate(hverts.data_ptr(), hfaces.data_ptr(), capi.HOST, 0)

        def oneshot_step():
            a, b, _, fits = h.extract_host(params, hfield.data_ptr(), nxl, ny, nz, nxl, hverts.data_ptr(), hverts.shape[0],
                                           hfaces.data_ptr(), hfaces.shape[0])
            assert fits and (a, b) == (nv, nf)

        def time_host(step):
            step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(ksteps):
                step()
            barrier()
            dt = torch.tensor([(time.perf_counter() - t0) / ksteps], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return float(dt.item())

        oneshot = True
        dt_pair = time_host(pair_step)
        # the same pair on ordinary pageable arrays (what a Julia caller owns): threaded pinned staging inside the library
        pag_field = hfield.numpy().copy()
        pag_verts, pag_faces = np.empty_like(hverts.numpy()), np.empty_like(hfaces.numpy())

        def pageable_step():
            a, b, _ = h.count(params, pag_field.ctypes.data, capi.HOST, nxl, ny, nz, nxl)
            h.generate(pag_verts.ctypes.data, pag_faces.ctypes.data, capi.HOST, 0)

        ksteps_keep, ksteps = ksteps, 3
        dt_pag = time_host(pageable_step)
        ksteps = ksteps_keep
        assert np.array_equal(pag_faces[:nf], hfaces.numpy()[:nf])
        del pag_field, pag_verts, pag_faces
        hfaces.zero_()
        dt = time_host(oneshot_step) if oneshot else dt_pair
        e2e = {"value": tot_vox / dt / 1e9, "unit": "Gvoxels/s", "ms_per_step": dt * 1e3,
               "steps": ksteps, "h2d_bytes_per_step": 4 * nxl * ny * nz, "d2h_bytes_per_step": 3 * vsz * nv + 24 * nf + 16,
               "api": ("b200iso_extract_host (one-shot, x-slab pipelined H2D || kernels || D2H)" if oneshot else
                       "b200iso_count(HOST) + b200iso_generate(HOST)") + ", pinned host buffers, per GPU",
               "two_phase": {"value": tot_vox / dt_pair / 1e9, "ms_per_step": dt_pair * 1e3,
                             "api": "b200iso_count(HOST) + b200iso_generate(HOST)"},
               "pageable": {"value": tot_vox / dt_pag / 1e9, "ms_per_step": dt_pag * 1e3,
                            "api": "b200iso_count(HOST) + b200iso_generate(HOST) on pageable arrays (threaded pinned staging)"}}
        # sanity: host result of the last step equals the device-resident result
        if px is not None:
            vbase.copy_(px.bases[:1])
            # the peer exchange must agree with an NCCL all-gather of the same totals
            dist.all_gather_into_tensor(gathered.view(-1), totals)
            assert torch.equal(px.all, gathered) and int(vbase.item()) == int(gathered[:rank, 0].sum().item())
        assert torch.equal(hfaces[:nf], faces[:nf].cpu() - int(vbase.item()) if sharded else faces[:nf].cpu())

    if rank == 0:
        peak, peak_src = measured_peak()
        vbytes = 24 if f64 else 12
        alg = {"classify": 4.0 * nxl * ny * nz, "count_scan": 0.0, "generate": float(vbytes * nv + 24 * nf)}
        w16 = ((nz + 31) // 32 + 15) // 16
        tma = ((nxl + 127) // 128) * ny * w16 >= 4096 and os.environ.get("B200ISO_TMA", "1") != "0"  # the library's rule
        kname = {"classify": "signpack_tma_kernel" if tma else "signpack_kernel", "count_scan": "mc_count_chunks_kernel" if spec["algo"] == "MC" else "count_kernel", "generate": "mc_generate_kernel" if spec["algo"] == "MC" else "mt_generate_kernel"}
        stage_ms = {"classify": stage["classify_ms"], "count_scan": stage["count_scan_ms"], "generate": stage["generate_ms"]}
        dom = max(stage_ms, key=lambda k: stage_ms[k])
        achieved = alg[dom] / (stage_ms[dom] * 1e-3) / 1e9 if stage_ms[dom] > 0 else 0.0
        bytes_step = 4.0 * nxl * ny * nz + vbytes * nv + 24 * nf  # rank 0's slab
        pipe_gbs = bytes_step / (ms_step * 1e-3) / 1e9
        line = {
            "metric": METRIC,
            "value": value, "unit": UNIT, "mtriangles_per_s": mtri,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": spec.get("scaling", "weak"), "vs_baseline": None,
            "dtype": "f32 field, f64 positions, f32 vertices, int64 faces" if not f64 else "f32 field, f64 positions and vertices, int64 faces",
            "data": "synthetic",
            "config": {"workload": args.workload, "description": spec["desc"], "shape": [nxg, ny, nz], "per_gpu_shape": [nxl, ny, nz], "algo": spec["algo"],
                       "sharding": ("x-slabs + one 16-byte exchange of counts" if sharded else ("replicas" if world > 1 else "single GPU")),
                       "exchange": exchange,
                       "l2": "inputs larger than L2 (field %.2f GB per GPU, read once per step)" % (4.0 * nxl * ny * nz / 1e9),
                       "mesh": {"nverts": tot_nv, "nfaces": tot_nf}},
            "roofline": {"bound": "hbm", "kernel": kname[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_src, "traffic": ncu_traffic(kname[dom], args.workload),
                         "algorithmic_bytes_per_launch": alg[dom], "kernel_ms": stage_ms[dom],
                         "stage_ms": stage_ms,
                         "kernels": {kname[k]: {"ms": stage_ms[k], "algorithmic_bytes": alg[k],
                                                "achieved": (alg[k] / (stage_ms[k] * 1e-3) / 1e9 if stage_ms[k] > 0 else 0.0),
                                                "frac": (alg[k] / (stage_ms[k] * 1e-3) / 1e9 / peak if stage_ms[k] > 0 else 0.0),
                                                "traffic": ncu_traffic(kname[k], args.workload)} for k in stage_ms},
                         "pipeline": {"algorithmic_bytes_per_step": bytes_step, "achieved": pipe_gbs, "frac_of_measured": pipe_gbs / peak,
                                      "frac_of_nominal_8TBs": pipe_gbs / 8000.0}},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if e2e is not None:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            from oracle import harness as oracle
            oracle.build()
            if e2e is not None:
                host = hfield.numpy().transpose(2, 1, 0)  # (nx, ny, nz) view, x-contiguous
            else:
                host = host_field_numpy(pkg, spec, 64)
            v, planes, dt, _, _ = cpu_sample(oracle, host, spec, args.cpu_seconds, 1)
            line["cpu_baseline"] = {"value": v, "unit": "Gvoxels/s", "cores": 1, "kind": "port",
                                    "sample": f"voxel x-planes [0,{planes}) of the same {nxl}x{ny}x{nz} field "
                                              f"({planes * (ny - 1) * (nz - 1)} voxels, {dt:.1f} s), full-field strides",
                                    "note": "C++ restatement of Meshing.jl's single-threaded loops (oracle/iso_oracle.cpp); "
                                            "Julia is not installed in this image"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
