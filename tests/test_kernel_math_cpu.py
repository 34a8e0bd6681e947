"""The mathematical facts the optimised kernels lean on, checked on the CPU (numpy / the oracle), so that a
change to a kernel's algebra has a test that does not need a GPU:

* the fast power-8 map written as two complex 8th powers (csrc/march_kernel.cuh, de_mandelbulb8_fast) is the
  shader's triplex_pow8 (fragment.shd:74-99);
* the three-instruction division by a known divisor (csrc/julia.cu, div_known) is correctly rounded;
* the Cornell estimator's culling bounds hold for the oracle's distance field: the field is 1-Lipschitz (the
  caller's hint), the supporting plane is a lower bound and the nearest centroid an upper bound."""
import numpy as np


def _triplex_pow8(x, y, z):
    """fragment.shd:74-99 in float64."""
    x2, y2, z2 = x * x, y * y, z * z
    x4, y4, z4 = x2 * x2, y2 * y2, z2 * z2
    k3 = y2 + x2
    k2 = 1.0 / np.sqrt(k3 ** 7)
    k1 = y4 + z4 + x4 - 6.0 * z2 * x2 - 6.0 * y2 * z2 + 2.0 * x2 * y2
    k4 = y2 - z2 + x2
    return (-8.0 * z * k4 * (y4 * y4 - 28.0 * y4 * y2 * x2 + 70.0 * y4 * x4 - 28.0 * y2 * x2 * x4 + x4 * x4) * k1 * k2,
            64.0 * y * z * x * (y2 - x2) * k4 * (y4 - 6.0 * y2 * x2 + x4) * k1 * k2,
            -16.0 * z2 * k3 * k4 * k4 + k1 * k1)


def test_power8_map_is_two_complex_eighth_powers():
    """w' = (Im ζ⁸ · c⁸ / ρ⁸, Re ζ⁸) with ζ = z + iρ, c = x + iy, ρ = |c| — the form the fast kernel evaluates with
    three packed complex squarings — and r⁸ = |ζ⁸| (what the derivative's 8·r⁷ is built from)."""
    rng = np.random.default_rng(7)
    p = rng.uniform(-1.2, 1.2, (200000, 3))
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    rho = np.hypot(x, y)
    zeta8 = (z + 1j * rho) ** 8
    c8 = (x + 1j * y) ** 8
    xy = zeta8.imag * c8 / rho ** 8
    rx, ry, rz = _triplex_pow8(x, y, z)
    scale = (x * x + y * y + z * z) ** 4 + 1e-300
    assert np.max(np.abs(xy.real - rx) / scale) < 1e-11
    assert np.max(np.abs(xy.imag - ry) / scale) < 1e-11
    assert np.max(np.abs(zeta8.real - rz) / scale) < 1e-11
    assert np.allclose(np.abs(zeta8), (x * x + y * y + z * z) ** 4, rtol=1e-12)


def test_division_by_a_known_divisor_is_correctly_rounded():
    """div_known: y = RN(1/b), q = RN(a·y), r = fma(−b, q, a) (exact), RN(q + r·y) == RN(a/b) for the divisors the
    Julia kernel uses (log 2 from rg_log(2.0f), and 40) over the ranges it divides."""
    rng = np.random.default_rng(1)

    def fma(a, b, c):          # one rounding: the float64 product of two float32 is exact, the sums here are too
        return np.float32(np.float64(a) * np.float64(b) + np.float64(c))

    for b in (np.float32(40.0), np.float32(0.6931471824645996)):
        y = np.float32(1.0) / b
        a = np.concatenate([rng.uniform(0, 41, 1000000), 10.0 ** rng.uniform(-30, 3, 1000000), -rng.uniform(0, 5, 200000),
                            np.arange(0, 41, dtype=np.float64)]).astype(np.float32)
        q = (a * y).astype(np.float32)
        got = fma(fma(-b, q, a), y, q)
        assert np.array_equal(got, (a / b).astype(np.float32)), float(b)


def _cornell_planes_and_spheres(O):
    v = O.cornell_vertices().astype(np.float64).reshape(32, 3, 3)
    n = np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0])
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    w = -np.einsum("ij,ij->i", n, v[:, 0])
    c = v.mean(axis=1)
    return n, w, c


def test_cornell_culling_bounds_hold_for_the_oracle_field(O):
    """de_cornell_box (csrc/march_kernel.cuh) skips a triangle when its supporting plane or bounding sphere is farther
    than an upper bound of the result.  The bounds it uses, against the oracle's field d(p) (all 32 triangles):
    d(p + s·u) ≤ d(p) + s for a unit u (so after a march step of d(p) the distance is ≤ 2·d(p)), min_i |n_i·p + w_i| ≤
    d(p) (plane lower bound), d(p) ≤ min_i |p − centroid_i| (centroid upper bound) — each with the kernel's margin."""
    rng = np.random.default_rng(3)
    p = rng.uniform(-1.0, 1.0, (4000, 3))
    p = p[np.linalg.norm(p, axis=1) < 1.0][:2500]
    u = rng.normal(size=p.shape); u /= np.linalg.norm(u, axis=1, keepdims=True)
    d0 = O.distance(0, 0.0, p.astype(np.float32)).astype(np.float64)
    d1 = O.distance(0, 0.0, (p + u * d0[:, None]).astype(np.float32)).astype(np.float64)
    assert np.all(d1 <= d0 * 2.0002 + 1.0e-4)                              # the march hint
    for delta in (0.016, 0.081, 0.4):
        dn = O.distance(0, 0.0, (p + u * delta).astype(np.float32)).astype(np.float64)
        assert np.all(dn <= d0 + delta * 1.0001 + 1.0e-4)                  # the normal / AO hint
    n, w, c = _cornell_planes_and_spheres(O)
    plane = np.abs(p @ n.T + w)                                            # [points, triangles]
    assert np.all(plane.min(axis=1) <= d0 * 1.0001 + 1.0e-4)               # the nearest triangle's plane is never farther than it
    cent = np.linalg.norm(p[:, None, :] - c[None], axis=2).min(axis=1)
    assert np.all(d0 <= cent * 1.0001 + 1.0e-4)                            # phase A's bound
