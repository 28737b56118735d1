"""Properties the reference's Fortran unit tests check (tests/unit_tests/*.f90), applied to the oracle,
plus internal consistency of the two transfer implementations."""
import numpy as np
import pytest

from conftest import aniso_mesh, rel_err


def rand(shape, seed):
    return np.random.default_rng(seed).standard_normal(shape[::-1])


def coarse_mesh(oracle, mesh, copt):
    o = oracle.OracleMG(mesh, copt, ngrids=2)
    return o.level_mesh(1)


@pytest.mark.parametrize("shape", [(22, 22, 22), (17, 20, 9), (33, 18), (9, 8)])
def test_literal_and_tabulated_transfers_are_bit_identical(oracle, shape):
    mesh = aniso_mesh(shape)
    mc = coarse_mesh(oracle, mesh, "N" * (2 * len(shape)))
    uf = rand(shape, 1)
    uc = rand(tuple(m.size for m in mc), 2)
    assert np.array_equal(oracle.mg_restrict(mesh, mc, uf, mode=0), oracle.mg_restrict(mesh, mc, uf, mode=1))
    assert np.array_equal(oracle.mg_interp(mesh, mc, uc, mode=0), oracle.mg_interp(mesh, mc, uc, mode=1))


def test_interp_exact_for_nlinear_function(oracle):
    """unit_test_interp.f90: N-linear interpolation reproduces an N-linear function to rounding error."""
    shape = (30, 22, 26)
    mesh = aniso_mesh(shape)
    mc = coarse_mesh(oracle, mesh, "NNNNNN")
    M, B = (0.7, -1.3, 0.4), (0.2, 0.5, -0.9)

    def f(ms):
        Z, Y, X = np.meshgrid(ms[2], ms[1], ms[0], indexing="ij")
        return (M[0] * X + B[0]) * (M[1] * Y + B[1]) * (M[2] * Z + B[2])

    got = oracle.mg_interp(mesh, mc, f(mc))
    assert np.abs(got - f(mesh)).max() <= 1e-14 * np.abs(f(mesh)).max()


@pytest.mark.parametrize("shape", [(32, 28, 36), (40, 31)])
def test_restriction_is_scaled_adjoint_of_interpolation(oracle, shape):
    """unit_test_galerkin.f90: <u_c, R u_f> dV_c == <P u_c, u_f> dV_f."""
    mesh = aniso_mesh(shape)
    mc = coarse_mesh(oracle, mesh, "N" * (2 * len(shape)))
    uf = rand(shape, 3)
    uc = rand(tuple(m.size for m in mc), 4)
    dV = lambda ms: np.prod([m[1] - m[0] for m in ms])
    lhs = (uc * oracle.mg_restrict(mesh, mc, uf)).sum() * dV(mc)
    rhs = (oracle.mg_interp(mesh, mc, uc) * uf).sum() * dV(mesh)
    assert abs(lhs - rhs) <= 1e-12 * max(abs(lhs), abs(rhs))


def test_2d_neumann_solve_is_second_order(oracle):
    """unit_test_2D_solve.f90: rhs = a1(2x-Lx)+b1(2y-Ly), exact cubic solution, error ~ h^2."""
    a1, b1, Lx, Ly = 1.0, 0.7, 1.0, 1.5
    errs, hs = [], []
    for s in (1, 2, 4):
        nx, ny = 27 * s, 36 * s
        x = np.linspace(0, Lx, nx)
        y = np.linspace(0, Ly, ny)
        Y, X = np.meshgrid(y, x, indexing="ij")
        rhs = a1 * (2 * X - Lx) + b1 * (2 * Y - Ly)
        uex = a1 * (X ** 3 / 3 - Lx * X ** 2 / 2) + b1 * (Y ** 3 / 3 - Ly * Y ** 2 / 2)
        uex -= uex.mean()
        o = oracle.OracleMG([x, y], "NNNN")
        ierr, u, du, nc = o.solve(np.zeros_like(rhs), rhs)
        assert ierr == 0
        errs.append(np.abs((u - u.mean()) - uex).max())
        hs.append(x[1] - x[0])
    gamma = np.polyfit(np.log10(hs), np.log10(errs), 1)[0]
    assert 1.7 < gamma < 2.3, (errs, gamma)


def test_relax3d_colour_order_depends_on_x_lower_bc(oracle):
    """Quirk Q6: with x-lower Neumann the first pass updates (i+j+k) even, with Dirichlet (i+j+k) odd
    (ndsm_optimized.f90:106).  One sweep from a delta function distinguishes the two orders."""
    shape = (8, 8, 8)
    mesh = aniso_mesh(shape, stretch=(1, 1, 1), origin=(0, 0, 0))
    u = np.zeros(shape[::-1])
    u[4, 4, 3] = 1.0  # (i,j,k) = (3,4,4): odd parity
    rhs = np.zeros_like(u)
    un = oracle.relax3d("NDDNDD", mesh, rhs, u)  # even first: the even neighbours see the delta in pass 1
    ud = oracle.relax3d("DNDDND", mesh, rhs, u)  # odd first: the delta itself is overwritten first
    assert un[4, 4, 4] != 0.0 and ud[4, 4, 4] == 0.0


def test_update_u_and_metrics(oracle):
    a, b = rand((9, 7, 5), 1), rand((9, 7, 5), 2)
    new, dmax, dmean = oracle.update_u(b, a)
    assert np.array_equal(new, b)
    assert dmax == np.abs(a - b).max()
    assert abs(dmean - np.abs(a - b).mean()) < 1e-15
