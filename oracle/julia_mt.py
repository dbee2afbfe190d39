"""Restatement of Julia's `MersenneTwister(seed)` Float64 stream (Julia 1.9/1.10 stdlib Random).

TEST INFRASTRUCTURE ONLY.  The reference's single exact-count test ("noisy spheres",
test/runtests.jl:152-172) builds its input from `rand(Random.MersenneTwister(0), 21, 21, 21)`.
Julia is not installed here, so the generator is restated from its published algorithm:

  * Julia's MersenneTwister is dSFMT-19937 (Saito & Matsumoto, dSFMT v2.2, BSD licence; Julia vendors
    it as stdlib dependency `dSFMT_jll`), seeded with `dsfmt_init_by_array(state, UInt32[seed])`
    for a non-negative Int seed < 2^32 (stdlib/Random/src/RNGs.jl: `make_seed`, `seed!`; Julia 1.9/1.10
    -- Julia >= 1.11 hashes the seed first and produces a different stream).
  * `rand(rng, dims...)` for Float64 fills the first `(n-2) ÷ 2 * 2` values straight from the dSFMT
    word stream (`dsfmt_fill_array_close_open!`) and the remaining ones from the scalar path, which
    draws from a freshly generated 382-value cache (RNGs.jl `rand!(::MersenneTwister, ::Array{Float64})`).

Pinned by known outputs of Julia itself (from its documentation/REPL transcripts):
  rand(MersenneTwister(0))            == 0.8236475079774124
  rand(MersenneTwister(1234), 2)      == [0.5908446386657102, 0.7667970365022592]
(tests/test_oracle.py checks both) -- and by reproducing the reference's golden 3466 / 6928 counts.
"""
import struct

import numpy as np

MEXP = 19937
N = (MEXP - 128) // 104 + 1  # 191
N32 = N * 4
POS1 = 117
SL1 = 19
SR = 12
MSK1 = 0x000FFAFFFFFFFB3F
MSK2 = 0x000FFDFFFC90FFFD
FIX1 = 0x90014964B32F4329
FIX2 = 0x3B8D12AC548A7C7A
PCV1 = 0x3D84E1AC0DC82880
PCV2 = 0x0000000000000001
LOW_MASK = 0x000FFFFFFFFFFFFF
HIGH_CONST = 0x3FF0000000000000
M64 = (1 << 64) - 1
M32 = (1 << 32) - 1


class DSFMT:
    def __init__(self, key):
        """dsfmt_init_by_array(key) -- key: list of uint32."""
        size = (N + 1) * 4
        lag = 11 if size >= 623 else 7 if size >= 68 else 5 if size >= 39 else 3
        mid = (size - lag) // 2
        p = [0x8B8B8B8B] * size

        def f1(x):
            return ((x ^ (x >> 27)) * 1664525) & M32

        def f2(x):
            return ((x ^ (x >> 27)) * 1566083941) & M32

        klen = len(key)
        count = max(klen + 1, size)
        r = f1(p[0] ^ p[mid % size] ^ p[(size - 1) % size])
        p[mid % size] = (p[mid % size] + r) & M32
        r = (r + klen) & M32
        p[(mid + lag) % size] = (p[(mid + lag) % size] + r) & M32
        p[0] = r
        count -= 1
        i, j = 1, 0
        while j < count and j < klen:
            r = f1(p[i] ^ p[(i + mid) % size] ^ p[(i + size - 1) % size])
            p[(i + mid) % size] = (p[(i + mid) % size] + r) & M32
            r = (r + key[j] + i) & M32
            p[(i + mid + lag) % size] = (p[(i + mid + lag) % size] + r) & M32
            p[i] = r
            i = (i + 1) % size
            j += 1
        while j < count:
            r = f1(p[i] ^ p[(i + mid) % size] ^ p[(i + size - 1) % size])
            p[(i + mid) % size] = (p[(i + mid) % size] + r) & M32
            r = (r + i) & M32
            p[(i + mid + lag) % size] = (p[(i + mid + lag) % size] + r) & M32
            p[i] = r
            i = (i + 1) % size
            j += 1
        for j in range(size):
            r = f2((p[i] + p[(i + mid) % size] + p[(i + size - 1) % size]) & M32)
            p[(i + mid) % size] ^= r
            r = (r - i) & M32
            p[(i + mid + lag) % size] ^= r
            p[i] = r
            i = (i + 1) % size
        # little-endian: u64 k = p[2k] | p[2k+1] << 32
        u = [p[2 * k] | (p[2 * k + 1] << 32) for k in range(2 * (N + 1))]
        for k in range(2 * N):  # initial_mask
            u[k] = (u[k] & LOW_MASK) | HIGH_CONST
        # period_certification
        t0 = u[2 * N] ^ FIX1
        t1 = u[2 * N + 1] ^ FIX2
        inner = (t0 & PCV1) ^ (t1 & PCV2)
        s = 32
        while s > 0:
            inner ^= inner >> s
            s >>= 1
        if (inner & 1) != 1:
            u[2 * N + 1] ^= 1  # PCV2 & 1 == 1
        self.st = [[u[2 * k], u[2 * k + 1]] for k in range(N)]
        self.lung = [u[2 * N], u[2 * N + 1]]
        self.pos = 0  # index of the oldest word

    def next_words(self, nwords):
        """Next `nwords` 128-bit words of the stream as a flat list of 2*nwords uint64 ([1,2) doubles)."""
        out = []
        st, lung = self.st, self.lung
        i = self.pos
        for _ in range(nwords):
            a = st[i]
            b = st[(i + POS1) % N]
            t0, t1 = a
            L0, L1 = lung
            n0 = ((t0 << SL1) & M64) ^ (L1 >> 32) ^ ((L1 << 32) & M64) ^ b[0]
            n1 = ((t1 << SL1) & M64) ^ (L0 >> 32) ^ ((L0 << 32) & M64) ^ b[1]
            lung = [n0, n1]
            r0 = (n0 >> SR) ^ (n0 & MSK1) ^ t0
            r1 = (n1 >> SR) ^ (n1 & MSK2) ^ t1
            st[i] = [r0, r1]
            out.append(r0)
            out.append(r1)
            i = (i + 1) % N
        self.pos = i
        self.lung = lung
        return out


def _to_f64(u64s):
    return np.frombuffer(struct.pack("<%dQ" % len(u64s), *u64s), dtype=np.float64)


class MersenneTwister:
    """Julia 1.9/1.10 `Random.MersenneTwister(seed::Integer)`, Float64 `rand` only."""

    def __init__(self, seed):
        assert 0 <= seed < 2 ** 32
        self.g = DSFMT([seed])
        self.vals = np.empty(0)
        self.idx = 0

    def _scalar(self):
        if self.idx >= len(self.vals):  # gen_rand: refill the 382-value cache
            self.vals = _to_f64(self.g.next_words(N)) - 1.0
            self.idx = 0
        v = self.vals[self.idx]
        self.idx += 1
        return v

    def rand(self, *dims):
        """rand(rng) or rand(rng, dims...) -> Float64 in [0,1); array in column-major fill order."""
        if not dims:
            return float(self._scalar())
        n = int(np.prod(dims))
        n2 = (n - 2) // 2 * 2
        out = np.empty(n)
        if n2 < 382:  # _rand_max383!: everything through the cache
            for k in range(n):
                out[k] = self._scalar()
        else:
            out[:n2] = _to_f64(self.g.next_words(n2 // 2)) - 1.0
            for k in range(n2, n):
                out[k] = self._scalar()
        return out.reshape(dims, order="F")
