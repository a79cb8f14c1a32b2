"""Size-independent properties at the synthetic 3-D config's shape (extruded cylinder, lx1=8, lxd=12; here 3 periodic
z-layers = 5 988 elements, 3.07 M points -- the oracle is too slow at this size, so parity is checked through identities
the operators must satisfy): dssum idempotence, <D u, p> = <u, D^T p>, symmetry and positivity of the Helmholtz operator
and of E = D B^-1 D^T, and the projection property D u^{n+1} = 0 of one full time step."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big(nlk_lib):
    import bench
    from neklab_b200 import api
    case = bench.cylinder_inputs()
    coords, U, vertex, cbc = bench.extrude(case, 8, 3)
    mesh = api.Mesh(coords, vertex, cbc, 12)
    ctx = api.Context(mesh, api.default_params(viscosity=1 / 50.0, torder=3, vtol=1e-11, ptol=1e-10, gmres_maxit=500, pr_proj=20))
    yield coords, U, mesh, ctx
    ctx.close()


def _fields(coords, k, seed):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(k):
        f = np.sin(0.3 * coords[:, 0] + 0.5 * i) * np.cos(0.4 * coords[:, 1] - 0.2 * i) * np.cos(2 * np.pi * coords[:, 2] / 1.5 + i)
        out.append(f + 0.05 * rng.standard_normal(f.shape))
    return out


def test_dssum_idempotent_and_numbering(big):
    coords, U, mesh, ctx = big
    vmult = mesh.field("vmult")
    u = np.random.default_rng(0).integers(-9, 9, size=coords[:, 0].shape).astype(np.float64)
    s = ctx.dssum(u)
    # averaging then summing reproduces the sum (vertex valences 3, 5, 6 occur on this mesh, so only to round-off)
    assert np.abs(ctx.dssum(s * vmult) - s).max() < 1e-12
    assert np.array_equal(s, np.round(s))                      # sums of integers stay integers (no stray contributions)
    g = mesh.glo_num()
    # every coincidence class carries one coordinate (x, y; z modulo the periodic length)
    order = np.argsort(g.ravel(), kind="stable")
    gs = g.ravel()[order]; xs = coords[:, 0].ravel()[order]; ys = coords[:, 1].ravel()[order]
    same = gs[1:] == gs[:-1]
    assert np.abs(xs[1:][same] - xs[:-1][same]).max() < 1e-9
    dy = np.abs(ys[1:][same] - ys[:-1][same]); Ly = ys.max() - ys.min()
    assert np.minimum(dy, np.abs(dy - Ly)).max() < 1e-9          # equal, or one period apart if the case is y-periodic


def test_div_gradt_transpose_and_E_spd(big):
    coords, U, mesh, ctx = big
    u = _fields(coords, 3, 1)
    p = np.random.default_rng(2).standard_normal(mesh.shape2)
    lhs = float((ctx.opdiv(u) * p).sum())
    w = ctx.opgradt(p)
    rhs = sum(float((u[c] * w[c]).sum()) for c in range(3))
    assert abs(lhs - rhs) < 1e-11 * max(abs(lhs), abs(rhs))
    q = np.random.default_rng(3).standard_normal(mesh.shape2)
    Ep, Eq = ctx.cdabdtp(p), ctx.cdabdtp(q)
    assert abs(float((Ep * q).sum()) - float((p * Eq).sum())) < 1e-10 * abs(float((Ep * q).sum()))
    assert float((Ep * p).sum()) > 0


def test_helmholtz_symmetric(big):
    coords, U, mesh, ctx = big
    u, v = _fields(coords, 2, 4)
    vm = mesh.field("vmult")
    u = ctx.dssum(u) * vm; v = ctx.dssum(v) * vm                # continuous fields
    Au, Av = ctx.axhelm(u, 0.02, 150.0), ctx.axhelm(v, 0.02, 150.0)
    a, b = float((Au * v).sum()), float((u * Av).sum())
    assert abs(a - b) < 1e-11 * abs(a) and float((Au * u).sum()) > 0


def test_time_step_projection_property(big):
    coords, U, mesh, ctx = big
    from neklab_b200 import api
    import ctypes as C
    bf = ctx.vec(); bf.upload([U[:, 0], U[:, 1], U[:, 2]])
    x = ctx.vec(); x.rand(ifnorm=True, seed=3)
    A = api.exptA_linop(ctx, 1.0, bf); s = A.init()
    api.lib().nlk_exptA_set_tau(A.h, C.c_double(3 * s["dt"]))
    y = A.matvec(x)
    v, pr, _ = y.download()
    bm2 = mesh.field("bm2")
    div = ctx.opdiv(v)
    assert np.sqrt(float((div * div / bm2).sum()) / float(bm2.sum())) < 1e-8      # D u = 0 to the pressure tolerance
    assert y.nrst == 2 and np.isfinite(y.norm())
