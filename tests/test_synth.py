"""The synthetic benchmark fields: host (numpy) and torch builders agree bit for bit (here on CPU torch)."""
import numpy as np


def test_gyroid_torch_equals_numpy(pkg):
    shape = (21, 18, 25)
    h = pkg.synth.gyroid(shape)
    t = pkg.synth.gyroid_torch(shape, "cpu", ldx=24)
    assert t.shape == shape and t.stride() == (1, 24, 24 * 18)
    assert np.array_equal(t.numpy(), h)
    tabs = pkg.synth.gyroid_tables(shape)
    assert np.array_equal(pkg.synth.gyroid(shape, x_slice=(5, 12), tables=tabs), h[5:12])
    assert np.array_equal(pkg.synth.gyroid_torch(shape, "cpu", x_slice=(5, 12), tables=tabs).numpy(), h[5:12])


def test_multisphere_torch_equals_numpy(pkg):
    import torch
    shape = (20, 22, 19)
    h = pkg.synth.multisphere_torus(shape)
    t = pkg.synth.multisphere_torus(shape, xp=torch, device="cpu")
    assert np.array_equal(t.numpy(), h)
    assert (h < 0).any() and (h > 0).any()
    c, r = pkg.synth.multisphere_params()
    assert c.shape == (32, 3) and (np.abs(c) <= 0.8).all() and (r >= 0.05).all() and (r <= 0.15).all()


def test_sphere_counts(pkg):
    s = pkg.synth.sphere(32)
    assert s.dtype == np.float32 and s.flags.f_contiguous and abs(float(s[0, 0, 0]) - (3 ** 0.5 - 0.5)) < 1e-6
