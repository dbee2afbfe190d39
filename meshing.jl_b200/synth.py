"""Deterministic synthetic scalar fields for parity tests and bench.py.

Every field is evaluated in Float64 at the nodes p = lo + (hi-lo)*i/(n-1) and rounded once to Float32.
The same operation order is used by the numpy (host) and torch (device) builders, all operations are
single IEEE operations (no FMA contraction in numpy or in torch's elementwise kernels), and the
transcendental parts (cos/sin of the 1-D node coordinates) are always computed on the host -- so the
host and device builders return bit-identical fields (tests/test_synth.py checks that on CPU torch).

Layout: arrays have shape (nx, ny, nz) and are Fortran-ordered / x-contiguous like a Julia Array{Float32,3}.
"""
import numpy as np


def _nodes(lo, hi, n):
    i = np.arange(n, dtype=np.float64)
    return lo + (hi - lo) * i / max(n - 1, 1)


def sphere(shape, radius=0.5, lo=-1.0, hi=1.0, dtype=np.float32):
    """||p|| - radius on [lo,hi]^3 (cf. norm_sdf / sphere_function, test/runtests.jl:10-13)."""
    nx, ny, nz = _shape3(shape)
    x, y, z = _nodes(lo, hi, nx), _nodes(lo, hi, ny), _nodes(lo, hi, nz)
    r2 = (x * x)[:, None, None] + (y * y)[None, :, None] + (z * z)[None, None, :]
    return np.asfortranarray((np.sqrt(r2) - radius).astype(dtype))


def gyroid_tables(shape, lo=0.0, hi=4.0 * np.pi):
    """1-D cos/sin tables of the gyroid cos x sin y + cos y sin z + cos z sin x (docs/src/examples.md:43-46)."""
    nx, ny, nz = _shape3(shape)
    x, y, z = _nodes(lo, hi, nx), _nodes(lo, hi, ny), _nodes(lo, hi, nz)
    return (np.cos(x), np.sin(x)), (np.cos(y), np.sin(y)), (np.cos(z), np.sin(z))


def gyroid(shape, lo=0.0, hi=4.0 * np.pi, x_slice=None, tables=None):
    """Float32 gyroid on [lo,hi]^3; `x_slice` = (xa, xb) returns only samples xa <= x < xb."""
    (cx, sx), (cy, sy), (cz, sz) = tables if tables is not None else gyroid_tables(shape, lo, hi)
    if x_slice is not None:
        cx, sx = cx[x_slice[0]:x_slice[1]], sx[x_slice[0]:x_slice[1]]
    t = cx[:, None, None] * sy[None, :, None]
    t = t + (cy[:, None] * sz[None, :])[None, :, :]
    t = t + cz[None, None, :] * sx[:, None, None]
    return np.asfortranarray(t.astype(np.float32))


def gyroid_torch(shape, device, lo=0.0, hi=4.0 * np.pi, x_slice=None, tables=None, chunk=32, ldx=None):
    """Same field built on `device` with torch (Float64 arithmetic, same operation order as gyroid()).
    Returns a torch.float32 tensor of logical shape (nx, ny, nz) that is x-contiguous (strides
    (1, ldx, ldx*ny)); `ldx` >= nx pads the leading dimension (padding is zero)."""
    import torch

    (cx, sx), (cy, sy), (cz, sz) = tables if tables is not None else gyroid_tables(shape, lo, hi)
    if x_slice is not None:
        cx, sx = cx[x_slice[0]:x_slice[1]], sx[x_slice[0]:x_slice[1]]
    nx, ny, nz = len(cx), len(cy), len(cz)
    ldx = nx if ldx is None else ldx
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    cx, sx, cy, sy, cz, sz = map(T, (cx, sx, cy, sy, cz, sz))
    store = torch.zeros((nz, ny, ldx), dtype=torch.float32, device=device)  # memory order z, y, x
    yz = (cy[None, :] * sz[:, None])  # (nz, ny)
    for z0 in range(0, nz, chunk):
        z1 = min(nz, z0 + chunk)
        t = cx[None, None, :] * sy[None, :, None]  # (1, ny, nx)
        t = t + yz[z0:z1, :, None]
        t = t + cz[z0:z1, None, None] * sx[None, None, :]
        store[z0:z1, :, :nx] = t.to(torch.float32)
    return store.permute(2, 1, 0)[:nx]  # logical (nx, ny, nz), x-contiguous


def splitmix64(seed):
    """SplitMix64 stream (Steele et al.); returns a function producing uniform doubles in [0,1)."""
    state = [seed & 0xFFFFFFFFFFFFFFFF]

    def nxt():
        state[0] = (state[0] + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = state[0]
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        z = z ^ (z >> 31)
        return (z >> 11) * (1.0 / 9007199254740992.0)

    return nxt


def multisphere_params(k=32, seed=0x5EED2048):
    u = splitmix64(seed)
    c = np.empty((k, 3))
    r = np.empty(k)
    for i in range(k):
        c[i] = [-0.8 + 1.6 * u(), -0.8 + 1.6 * u(), -0.8 + 1.6 * u()]
        r[i] = 0.05 + 0.10 * u()
    return c, r


def multisphere_torus(shape, lo=-1.0, hi=1.0, k=32, seed=0x5EED2048, x_slice=None, xp=np, device=None, chunk=16, ldx=None):
    """min( torus(2p), min_k(||p - c_k|| - r_k) ), torus(v) = (sqrt(v1^2+v2^2) - 0.5)^2 + v3^2 - 0.25
    (torus_function, test/runtests.jl:11).  xp = numpy (host) or torch (device), same operation order."""
    nx, ny, nz = _shape3(shape)
    x, y, z = _nodes(lo, hi, nx), _nodes(lo, hi, ny), _nodes(lo, hi, nz)
    if x_slice is not None:
        x = x[x_slice[0]:x_slice[1]]
    c, r = multisphere_params(k, seed)
    if xp is np:
        X, Y, Z = x[:, None, None], y[None, :, None], z[None, None, :]
        f = _mst_eval(np, X, Y, Z, c, r)
        return np.asfortranarray(f.astype(np.float32))
    import torch

    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    xt, yt, zt = T(x), T(y), T(z)
    nxl = len(x)
    ldx = nxl if ldx is None else ldx
    store = torch.zeros((nz, ny, ldx), dtype=torch.float32, device=device)
    for z0 in range(0, nz, chunk):
        z1 = min(nz, z0 + chunk)
        f = _mst_eval(torch, xt[None, None, :], yt[None, :, None], zt[z0:z1, None, None], c, r)
        store[z0:z1, :, :nxl] = f.to(torch.float32)
    return store.permute(2, 1, 0)[:nxl]


def _mst_eval(xp, X, Y, Z, c, r):
    vx, vy, vz = 2.0 * X, 2.0 * Y, 2.0 * Z
    q = xp.sqrt(vx * vx + vy * vy) - 0.5
    f = (q * q + vz * vz) - 0.25
    for k in range(len(r)):
        dx, dy, dz = X - float(c[k, 0]), Y - float(c[k, 1]), Z - float(c[k, 2])
        d = xp.sqrt((dx * dx + dy * dy) + dz * dz) - float(r[k])
        f = xp.minimum(f, d)
    return f


def noise(shape, seed=0, dtype=np.float32):
    """Uniform noise in [-1,1): the dense worst case (almost every voxel active)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return np.asfortranarray((rng.random(_shape3(shape)) * 2.0 - 1.0).astype(dtype))


def _shape3(shape):
    if isinstance(shape, int):
        return (shape, shape, shape)
    nx, ny, nz = shape
    return int(nx), int(ny), int(nz)
