"""Planar Plummer generator of libgravity_host."""
import math

import numpy as np

import proj_entropy_b200 as g
from conftest import bits_equal

G, M, A = 6.673E-11, 1.989E30, 1.495978707E11
MASK = (1 << 64) - 1


class Xoshiro:
    """xoshiro256** seeded through splitmix64 -- independent restatement."""

    def __init__(self, seed):
        self.s = []
        for _ in range(4):
            seed = (seed + 0x9E3779B97F4A7C15) & MASK
            z = seed
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
            self.s.append(z ^ (z >> 31))

    @staticmethod
    def rotl(v, k):
        return ((v << k) | (v >> (64 - k))) & MASK

    def u(self):
        s = self.s
        res = (self.rotl((s[1] * 5) & MASK, 7) * 9) & MASK
        t = (s[1] << 17) & MASK
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t
        s[3] = self.rotl(s[3], 45)
        return (res >> 11) * 2.0 ** -53


def plummer_py(n, seed):
    r_ = Xoshiro(seed)
    x, y, vx, vy = [], [], [], []
    vs = math.sqrt(G * M / A)
    for _ in range(n):
        while True:
            u1 = r_.u()
            if u1 <= 0:
                continue
            r = A / math.sqrt(u1 ** (-2.0 / 3.0) - 1.0)
            if r <= 20 * A:
                break
        z = (1 - 2 * r_.u()) * r
        phi = 2 * math.pi * r_.u()
        rho = math.sqrt(r * r - z * z)
        x.append(rho * math.cos(phi)); y.append(rho * math.sin(phi))
        while True:
            u4 = r_.u(); u5 = r_.u()
            if 0.1 * u5 < u4 * u4 * (1 - u4 * u4) ** 3.5:
                break
        v = u4 * math.sqrt(2.0) * (1 + r * r / (A * A)) ** -0.25 * vs
        vz = (1 - 2 * r_.u()) * v
        phi = 2 * math.pi * r_.u()
        vr = math.sqrt(v * v - vz * vz)
        vx.append(vr * math.cos(phi)); vy.append(vr * math.sin(phi))
    return map(np.array, (x, y, vx, vy))


def test_matches_independent_restatement():
    n = 500
    m, x, y, vx, vy = g.plummer(n, seed=1)
    px, py, pvx, pvy = plummer_py(n, 1)
    for got, raw in ((x, px), (y, py), (vx, pvx), (vy, pvy)):
        want = raw - float(np.sum(raw.astype(np.longdouble)) / n)
        assert np.allclose(got, want, rtol=1e-13, atol=1e-13 * np.abs(raw).max())
    assert np.all(m == M / n)


def test_deterministic_and_seeded():
    a = g.plummer(4096, seed=1)
    b = g.plummer(4096, seed=1)
    c = g.plummer(4096, seed=2)
    assert all(bits_equal(p, q) for p, q in zip(a, b))
    assert not bits_equal(a[1], c[1])
    # strictly sequential draws: a shorter run is a prefix (before COM removal)
    d = g.plummer(1000, seed=1)
    shift = a[1][:1000] - d[1]
    assert np.allclose(shift, shift[0], rtol=0, atol=1e-3)


def test_centre_of_mass_and_shape():
    n = 65536
    m, x, y, vx, vy = g.plummer(n, seed=1)
    for a in (x, vx, y, vy):
        assert abs(a.mean()) < 1e-9 * np.abs(a).max()
    assert abs(m.sum() - M) / M < 1e-12
    r = np.hypot(x, y)
    # projected half-mass radius of a Plummer sphere is exactly its scale radius
    assert abs(np.median(r) / A - 1.0) < 0.03
    assert r.max() <= 20.5 * A
    v = np.hypot(vx, vy)
    assert v.max() < math.sqrt(2 * G * M / A) * 1.01     # below the central escape speed
