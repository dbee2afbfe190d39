"""ctypes binding of the C ABI in include/b200iso.h (libb200iso.so).  No torch types cross this boundary:
pointers are plain integers (numpy `.ctypes.data`, torch `.data_ptr()`).

The library is required: if it is missing or does not export a declared symbol, loading raises -- there
is no CPU fallback anywhere in the product path."""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("B200ISO_LIB") or os.path.join(_HERE, "lib", "libb200iso.so")  # env override: A/B builds
HEADER = os.path.join(ROOT, "include", "b200iso.h")

MC, MT = 0, 1
HOST, DEVICE = 0, 1
ECAPACITY = -5  # B200ISO_ECAPACITY
PEER_MAX = 16
PEER_BYTES = 2 * PEER_MAX * 4 * 8  # B200ISO_PEER_BYTES
RANGE_INT, RANGE_F32, RANGE_F64 = 0, 1, 2
CLASSIFY_LDG128, CLASSIFY_TMA, CLASSIFY_SCALAR, CLASSIFY_F64 = 0, 1, 2, 3  # B200ISO_CLASSIFY_*


class Params(ctypes.Structure):
    """struct b200iso_params"""
    _fields_ = [("algo", ctypes.c_int32), ("iso_is_f32", ctypes.c_int32), ("eps_is_f32", ctypes.c_int32),
                ("range_kind", ctypes.c_int32), ("iso", ctypes.c_double), ("eps", ctypes.c_double),
                ("x0", ctypes.c_double), ("x1", ctypes.c_double), ("y0", ctypes.c_double), ("y1", ctypes.c_double),
                ("z0", ctypes.c_double), ("z1", ctypes.c_double), ("x_offset", ctypes.c_int64), ("nx_global", ctypes.c_int64),
                ("field_is_f64", ctypes.c_int32), ("x_ghost", ctypes.c_int32)]


class B200IsoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b200iso error {code}: {msg}")
        self.code = code


_lib = None


def declared_symbols():
    """Every function name declared in include/b200iso.h."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200iso_[a-z_]+)\s*\(", src)))


def load():
    """dlopen libb200iso.so and type every entry point.  Raises if the library or a symbol is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    for name in declared_symbols():
        if not hasattr(L, name):
            raise ImportError(f"{LIB_PATH} does not export {name} declared in include/b200iso.h")
    vp, i64, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    pi64 = ctypes.POINTER(ctypes.c_int64)
    pci = ctypes.POINTER(ctypes.c_int)
    pp = ctypes.POINTER(Params)
    L.b200iso_create.argtypes = [ctypes.POINTER(vp), ci]
    L.b200iso_destroy.argtypes = [vp]
    L.b200iso_last_error.restype = ctypes.c_char_p
    L.b200iso_last_error.argtypes = []
    L.b200iso_version.argtypes = []
    L.b200iso_set_stream.argtypes = [vp, vp]
    L.b200iso_use_own_stream.argtypes = [vp]
    L.b200iso_count.argtypes = [vp, pp, vp, ci, i64, i64, i64, i64, pi64, pi64, pci]
    L.b200iso_generate.argtypes = [vp, vp, vp, ci, i64]
    L.b200iso_count_async.argtypes = [vp, pp, vp, i64, i64, i64, i64, vp]
    L.b200iso_generate_async.argtypes = [vp, vp, i64, vp, i64, vp, i64]
    L.b200iso_totals.argtypes = [vp, pi64, pi64, pci]
    L.b200iso_extract_async.argtypes = [vp, pp, vp, i64, i64, i64, i64, vp, i64, vp, i64, vp, i64, vp]
    L.b200iso_extract_host.argtypes = [vp, pp, vp, i64, i64, i64, i64, vp, i64, vp, i64, pi64, pi64, pci]
    L.b200iso_extract_host_resident.argtypes = [vp, vp, i64, vp, i64, pi64, pi64]
    L.b200iso_set_peer_exchange.argtypes = [vp, ci, ci, ctypes.POINTER(vp)]
    L.b200iso_exchange_async.argtypes = [vp, vp, vp]
    L.b200iso_add_vertex_base_async.argtypes = [vp, vp, i64, vp, vp]
    L.b200iso_set_classify_mode.argtypes = [vp, ci]
    L.b200iso_classify_path.argtypes = [vp]
    L.b200iso_set_ride_warps.argtypes = [vp, ci]
    L.b200iso_ride_claimed.argtypes = [vp]
    L.b200iso_ride_claimed.restype = i64
    L.b200iso_set_peer_timeout.argtypes = [vp, ctypes.c_double]
    L.b200iso_vertex_normals_async.argtypes = [vp, pp, vp, i64, i64, i64, i64, vp, i64, ci, vp]
    L.b200iso_vertex_keys_async.argtypes = [vp, vp, i64]
    L.b200iso_weld.argtypes = [vp, vp, vp, i64, ci, vp, i64, i64, vp, vp, pi64]
    L.b200iso_write_ply.argtypes = [ctypes.c_char_p, vp, i64, ci, vp, vp, i64]
    L.b200iso_write_stl.argtypes = [ctypes.c_char_p, vp, i64, ci, vp, i64]
    L.b200iso_case_indices.argtypes = [vp, vp, ci]
    L.b200iso_enable_timing.argtypes = [vp, ci]
    L.b200iso_timings.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ci]
    L.b200iso_launch_count.argtypes = [vp]
    L.b200iso_launch_count.restype = i64
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise B200IsoError(rc, load().b200iso_last_error().decode())


class Handle:
    """Owns one b200iso_handle (one CUDA device, one stream, all device scratch)."""

    def __init__(self, device=0):
        self.L = load()
        self.h = ctypes.c_void_p()
        self.device = device
        _check(self.L.b200iso_create(ctypes.byref(self.h), device))

    def close(self):
        if self.h:
            self.L.b200iso_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        """Run on the given cudaStream_t (0 = CUDA's legacy default stream, e.g. torch's default stream)."""
        _check(self.L.b200iso_set_stream(self.h, ctypes.c_void_p(cuda_stream_ptr or 0)))

    def use_own_stream(self):
        _check(self.L.b200iso_use_own_stream(self.h))

    def count(self, params, sdf_ptr, mem, nx, ny, nz, ldx):
        nv, nf, f64 = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int()
        _check(self.L.b200iso_count(self.h, ctypes.byref(params), ctypes.c_void_p(sdf_ptr), mem, nx, ny, nz, ldx,
                                    ctypes.byref(nv), ctypes.byref(nf), ctypes.byref(f64)))
        return nv.value, nf.value, bool(f64.value)

    def generate(self, verts_ptr, faces_ptr, mem, vertex_base=0):
        _check(self.L.b200iso_generate(self.h, ctypes.c_void_p(verts_ptr), ctypes.c_void_p(faces_ptr), mem, vertex_base))

    def count_async(self, params, sdf_dev_ptr, nx, ny, nz, ldx, totals_dev_ptr=0):
        _check(self.L.b200iso_count_async(self.h, ctypes.byref(params), ctypes.c_void_p(sdf_dev_ptr), nx, ny, nz, ldx,
                                          ctypes.c_void_p(totals_dev_ptr or 0)))

    def generate_async(self, verts_dev_ptr, vcap, faces_dev_ptr, fcap, vertex_base_dev_ptr=0, vertex_base=0):
        _check(self.L.b200iso_generate_async(self.h, ctypes.c_void_p(verts_dev_ptr), vcap, ctypes.c_void_p(faces_dev_ptr), fcap,
                                             ctypes.c_void_p(vertex_base_dev_ptr or 0), vertex_base))

    def extract_async(self, params, sdf_dev_ptr, nx, ny, nz, ldx, verts_dev_ptr, vcap, faces_dev_ptr, fcap,
                      vertex_base_dev_ptr=0, vertex_base=0, totals_dev_ptr=0):
        """classify + single-pass count/scan/generate, fully asynchronous (device buffers with capacity)."""
        _check(self.L.b200iso_extract_async(self.h, ctypes.byref(params), ctypes.c_void_p(sdf_dev_ptr), nx, ny, nz, ldx,
                                            ctypes.c_void_p(verts_dev_ptr), vcap, ctypes.c_void_p(faces_dev_ptr), fcap,
                                            ctypes.c_void_p(vertex_base_dev_ptr or 0), vertex_base,
                                            ctypes.c_void_p(totals_dev_ptr or 0)))

    def extract_host(self, params, sdf_ptr, nx, ny, nz, ldx, verts_ptr, vcap, faces_ptr, fcap):
        """One-shot slab-pipelined host call.  Returns (nverts, nfaces, vert_is_f64, fits); fits=False means the
        capacities were too small (B200ISO_ECAPACITY) and the totals say what to allocate."""
        nv, nf, f64 = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int()
        rc = self.L.b200iso_extract_host(self.h, ctypes.byref(params), ctypes.c_void_p(sdf_ptr), nx, ny, nz, ldx,
                                         ctypes.c_void_p(verts_ptr or 0), vcap, ctypes.c_void_p(faces_ptr or 0), fcap,
                                         ctypes.byref(nv), ctypes.byref(nf), ctypes.byref(f64))
        if rc != ECAPACITY:
            _check(rc)
        return nv.value, nf.value, bool(f64.value), rc == 0

    def extract_host_resident(self, verts_ptr, vcap, faces_ptr, fcap):
        """Finish an extract_host call whose capacities were short from the slabs still resident on the device.
        Returns (nverts, nfaces, fits)."""
        nv, nf = ctypes.c_int64(), ctypes.c_int64()
        rc = self.L.b200iso_extract_host_resident(self.h, ctypes.c_void_p(verts_ptr or 0), vcap, ctypes.c_void_p(faces_ptr or 0), fcap,
                                                  ctypes.byref(nv), ctypes.byref(nf))
        if rc != ECAPACITY:
            _check(rc)
        return nv.value, nf.value, rc == 0

    def set_peer_exchange(self, rank, world, slot_ptrs):
        """slot_ptrs[r] = device pointer (valid on this device) of rank r's PEER_BYTES exchange buffer; None switches it off."""
        if not slot_ptrs or world <= 1:
            _check(self.L.b200iso_set_peer_exchange(self.h, 0, 0, None))
            return
        arr = (ctypes.c_void_p * world)(*[ctypes.c_void_p(int(p)) for p in slot_ptrs])
        _check(self.L.b200iso_set_peer_exchange(self.h, rank, world, arr))

    def exchange_async(self, bases_dev_ptr, all_dev_ptr=0):
        _check(self.L.b200iso_exchange_async(self.h, ctypes.c_void_p(bases_dev_ptr), ctypes.c_void_p(all_dev_ptr or 0)))

    def set_classify_mode(self, mode):
        """-1 = automatic (default), 1 = TMA-staged classify whenever the field allows a tensor map, 0 = never TMA"""
        _check(self.L.b200iso_set_classify_mode(self.h, mode))

    def classify_path(self):
        """CLASSIFY_* of the last count (-1 before the first)"""
        return self.L.b200iso_classify_path(self.h)

    def set_ride_warps(self, warps):
        """counting warps per TMA classify CTA (0 = count kernel only)"""
        _check(self.L.b200iso_set_ride_warps(self.h, int(warps)))

    def ride_claimed(self):
        """generate blocks counted inside the classify kernel in the last count"""
        return self.L.b200iso_ride_claimed(self.h)

    def set_peer_timeout(self, seconds):
        """how long exchange_async waits for its peers; <= 0 waits for ever"""
        _check(self.L.b200iso_set_peer_timeout(self.h, float(seconds)))

    def add_vertex_base_async(self, faces_dev_ptr, fcap, totals_dev_ptr, vertex_base_dev_ptr):
        _check(self.L.b200iso_add_vertex_base_async(self.h, ctypes.c_void_p(faces_dev_ptr), fcap,
                                                    ctypes.c_void_p(totals_dev_ptr), ctypes.c_void_p(vertex_base_dev_ptr)))

    def vertex_normals_async(self, params, sdf_dev_ptr, nx, ny, nz, ldx, verts_dev_ptr, nverts, vert_is_f64, normals_dev_ptr):
        _check(self.L.b200iso_vertex_normals_async(self.h, ctypes.byref(params), ctypes.c_void_p(sdf_dev_ptr), nx, ny, nz, ldx,
                                                   ctypes.c_void_p(verts_dev_ptr), nverts, int(vert_is_f64), ctypes.c_void_p(normals_dev_ptr)))

    def vertex_keys_async(self, keys_dev_ptr, kcap):
        _check(self.L.b200iso_vertex_keys_async(self.h, ctypes.c_void_p(keys_dev_ptr), kcap))

    def weld(self, keys_dev_ptr, verts_dev_ptr, nverts, vert_is_f64, faces_dev_ptr, nfaces, vertex_base, verts_out_dev_ptr, faces_out_dev_ptr):
        """-> number of welded vertices"""
        n = ctypes.c_int64()
        _check(self.L.b200iso_weld(self.h, ctypes.c_void_p(keys_dev_ptr), ctypes.c_void_p(verts_dev_ptr), nverts, int(vert_is_f64),
                                   ctypes.c_void_p(faces_dev_ptr or 0), nfaces, vertex_base, ctypes.c_void_p(verts_out_dev_ptr),
                                   ctypes.c_void_p(faces_out_dev_ptr or 0), ctypes.byref(n)))
        return n.value

    def totals(self):
        nv, nf, f64 = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int()
        _check(self.L.b200iso_totals(self.h, ctypes.byref(nv), ctypes.byref(nf), ctypes.byref(f64)))
        return nv.value, nf.value, bool(f64.value)

    def case_indices(self, out_ptr, mem):
        _check(self.L.b200iso_case_indices(self.h, ctypes.c_void_p(out_ptr), mem))

    def enable_timing(self, on=True):
        _check(self.L.b200iso_enable_timing(self.h, int(on)))

    def timings(self):
        ms = (ctypes.c_float * 5)()
        _check(self.L.b200iso_timings(self.h, ms, 5))
        return dict(zip(("classify_ms", "count_scan_ms", "generate_ms", "h2d_ms", "d2h_ms"), list(ms)))

    def launch_count(self):
        return self.L.b200iso_launch_count(self.h)


def write_ply(path, verts_ptr, nverts, vert_is_f64, normals_ptr, faces_ptr, nfaces):
    _check(load().b200iso_write_ply(os.fsencode(path), ctypes.c_void_p(verts_ptr or 0), nverts, int(vert_is_f64), ctypes.c_void_p(normals_ptr or 0),
                                    ctypes.c_void_p(faces_ptr or 0), nfaces))


def write_stl(path, verts_ptr, nverts, vert_is_f64, faces_ptr, nfaces):
    _check(load().b200iso_write_stl(os.fsencode(path), ctypes.c_void_p(verts_ptr or 0), nverts, int(vert_is_f64), ctypes.c_void_p(faces_ptr or 0), nfaces))
