#!/usr/bin/env python3
"""Regenerates the committed golden fixtures (run from the repo root: python tests/golden/make_golden.py).

noisy_spheres_input.npy : the input of the reference's "noisy spheres" test (test/runtests.jl:152-172):
    Float32 distance field on -10:10 cubed + 1.0 * rand(MersenneTwister(0), 21, 21, 21) -> Float64 field.
    The reference asserts exactly 3466 vertices / 6928 faces for MarchingTetrahedra(iso=8.0).
    Julia's RNG stream is restated in oracle/julia_mt.py (Julia is not installed in this image).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.julia_mt import MersenneTwister  # noqa: E402

N, sigma = 10, 1.0
i = np.arange(-N, N + 1)
I, J, K = np.meshgrid(i, i, i, indexing="ij")
dist = np.sqrt((I * I + J * J + K * K).astype(np.float32)).astype(np.float32)
field = dist.astype(np.float64) + sigma * MersenneTwister(0).rand(2 * N + 1, 2 * N + 1, 2 * N + 1)
np.save(os.path.join(HERE, "noisy_spheres_input.npy"), np.asfortranarray(field))
print("wrote noisy_spheres_input.npy", field.shape, field.dtype)
