"""Deterministic synthetic inputs for tests and benchmarks (no RNG, no files).

* ``test_case1`` -- the reference's analytic potential field (tests/integration_test/integration_test1.py:57-99).
* ``dipole``     -- sub-surface point dipole (BASELINE.json configs 0-2).
* ``charges``    -- two sub-surface magnetic charges, bipolar active-region-like magnetogram (config 3).
"""
import numpy as np


def mesh(nx, ny=None, nz=None):
    """x = linspace(0,1,nx); y, z continue with the same spacing (integration_test1.py:124-127)."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    x = np.linspace(0, 1, nx)
    dx = x[1] - x[0]
    return x, np.arange(ny) * dx, np.arange(nz) * dx


def test_case1(x, y, z):
    """Returns (A, b), both (3,nz,ny,nx): b = curl A, potential, k = pi, l = sqrt(2) k."""
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    wn = 1.0 * np.pi
    l = np.sqrt(2 * wn ** 2)
    b = np.zeros((3,) + X.shape)
    A = np.zeros((3,) + X.shape)
    e = np.exp(-l * Z)
    b[0] = +l * np.sin(wn * X) * np.cos(wn * Y) * e
    b[1] = +l * np.cos(wn * X) * np.sin(wn * Y) * e
    b[2] = +2 * wn * np.cos(wn * X) * np.cos(wn * Y) * e
    A[0] = -np.cos(wn * X) * np.sin(wn * Y) * e
    A[1] = +np.sin(wn * X) * np.cos(wn * Y) * e
    return A, b


def _faces_only(shape, fill):
    """Evaluate ``fill(k_slice, j_slice, i_slice)`` on the six boundary faces only (the solver never reads
    the interior of b; ndsm.py:82-83).  Keeps 513^3 inputs cheap to generate."""
    nz, ny, nx = shape
    b = np.zeros((3, nz, ny, nx))
    for sl in ((slice(None), slice(None), [0, nx - 1]), (slice(None), [0, ny - 1], slice(None)),
               ([0, nz - 1], slice(None), slice(None))):
        b[(slice(None),) + tuple(sl)] = fill(*sl)
    return b


def dipole(x, y, z, r0=(0.40, 0.55, -0.30), m=(0.3, -0.2, 1.0), faces_only=False):
    """B(r) = 3 (m.R) R / r^5 - m / r^3, R = r - r0.  Returns b (3,nz,ny,nx)."""
    m = np.asarray(m, dtype=np.float64)

    def field(ks, js, is_):
        Z, Y, X = np.meshgrid(z[ks], y[js], x[is_], indexing="ij")
        Rx, Ry, Rz = X - r0[0], Y - r0[1], Z - r0[2]
        r2 = Rx * Rx + Ry * Ry + Rz * Rz
        r = np.sqrt(r2)
        mR = m[0] * Rx + m[1] * Ry + m[2] * Rz
        r5 = r2 * r2 * r
        r3 = r2 * r
        return np.stack([3 * mR * Rx / r5 - m[0] / r3, 3 * mR * Ry / r5 - m[1] / r3, 3 * mR * Rz / r5 - m[2] / r3])

    if faces_only:
        return _faces_only((z.size, y.size, x.size), field)
    return field(slice(None), slice(None), slice(None))


def charges(x, y, z, q=((+1.0, (0.35, 0.45, -0.05)), (-0.7, (0.65, 0.55, -0.05))), faces_only=False):
    """B = sum_s q_s R_s / r_s^3 (two sub-surface magnetic charges)."""

    def field(ks, js, is_):
        Z, Y, X = np.meshgrid(z[ks], y[js], x[is_], indexing="ij")
        out = np.zeros((3,) + X.shape)
        for qs, (x0, y0, z0) in q:
            Rx, Ry, Rz = X - x0, Y - y0, Z - z0
            r3 = (Rx * Rx + Ry * Ry + Rz * Rz) ** 1.5
            out[0] += qs * Rx / r3
            out[1] += qs * Ry / r3
            out[2] += qs * Rz / r3
        return out

    if faces_only:
        return _faces_only((z.size, y.size, x.size), field)
    return field(slice(None), slice(None), slice(None))
