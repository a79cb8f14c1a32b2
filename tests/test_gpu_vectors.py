"""GPU tests of the pieces of SURVEY.md §8(a) that had no test of their own: `nek_drand` (N5) and the neklab forcing
registry with its `ipert` slots (N19), through the C-ABI."""
import numpy as np
import pytest

from oracle.cref import CPertStepper
from oracle.precond import SchwarzCoarse
from oracle.stepper import ExptA, NekVec, StepParams, nonlinear_map, seeded_field
from tests.util import box_case, nlk_mesh

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small(nlk_lib):
    from neklab_b200 import api
    cbt = np.array([["t  ", "I  ", "t  ", "I  "]] * 12)
    om, _, _ = box_case(ndim=2, nel=(4, 3), n=6, lxd=9, bc={"xlo": "v  ", "xhi": "O  "}, cbc_t=cbt)
    ctx = api.Context(nlk_mesh(om), api.default_params(viscosity=0.05, torder=3, vtol=1e-13, ptol=1e-13, ttol=1e-13, gmres_maxit=2000, ifheat=1, conductivity=0.07))
    yield om, ctx
    ctx.close()


def test_rand_is_continuous_masked_normalised_and_seeded(small):
    """nek_drand (src/vectors/real_vectors.f90:52-123): random field -> opdssum * vmult (C0), bcdirvc / bcdirsc (Dirichlet values),
    optional normalisation, nrst = 0.  `random_number` is compiler-specific, so the generator is a seeded replacement; what the
    reference guarantees are these properties."""
    om, ctx = small
    v = ctx.vec(); v.rand(ifnorm=True, seed=7)
    f, pr, th = v.download()
    for c in range(2):
        assert np.abs(f[c]).max() > 0
        assert np.abs(om.dssum(f[c]) * om.vmult - f[c]).max() < 1e-14 * np.abs(f[c]).max()      # identical on every copy of a shared node
        assert np.abs(f[c][om.vmask[c] == 0]).max() == 0.0                                        # homogeneous Dirichlet values
    assert np.abs(om.dssum(th) * om.vmult - th).max() < 1e-14 * np.abs(th).max() and np.abs(th[om.tmask == 0]).max() == 0.0
    assert np.abs(pr).max() == 0.0 and v.nrst == 0
    assert abs(v.norm() - 1.0) < 1e-13                                                             # ifnorm: bm1-weighted norm over velocity + temperature
    w = ctx.vec(); w.rand(ifnorm=True, seed=7)
    g, _, _ = w.download()
    assert np.array_equal(g[0], f[0])                                                              # same seed, same field
    w.rand(ifnorm=True, seed=8)
    g, _, _ = w.download()
    assert np.abs(g[0] - f[0]).max() > 1e-3
    u = ctx.vec(); u.rand(ifnorm=False, seed=7)
    assert abs(u.norm() - 1.0) > 1e-3


def test_forcing_registry_slots(nlk_lib):
    """set/get/zero_neklab_forcing with ipert (src/neklab_nek_forcing.f90:36-114): slot 1 feeds the perturbation step, slot 0 the
    nonlinear one, they do not leak into each other, invalid ipert is an error, zeroing restores the unforced result."""
    from neklab_b200 import api
    om, _, _ = box_case(ndim=2, nel=(4, 3), n=6, lxd=9, bc={"xlo": "v  ", "xhi": "O  "})
    tight = dict(vtol=1e-13, ptol=1e-13, gmres_maxit=2000)
    prm = StepParams(viscosity=0.05, torder=3, **tight)
    x = om.coords
    bf = NekVec(om, 3); bf.v = [1.0 + 0.3 * np.sin(0.5 * x[:, 1]), 0.2 * np.cos(0.4 * x[:, 0])]
    force = [0.3 * np.cos(0.8 * x[:, 1]) * np.sin(0.5 * x[:, 0]), -0.2 * np.sin(0.7 * x[:, 0])]
    x0 = seeded_field(om, 3)
    pre = SchwarzCoarse(om)
    ctx = api.Context(nlk_mesh(om), api.default_params(viscosity=0.05, torder=3, **tight))
    to_dev = lambda nv: (lambda d: (d.upload(nv.v, nv.pr), d)[1])(ctx.vec())
    A = api.exptA_linop(ctx, 0.1, to_dev(bf))

    def wdiff(dev, ref):
        v, _, _ = dev.download()
        return np.sqrt(sum(float(((v[c] - ref.v[c]) ** 2 * om.bm1).sum()) for c in range(2)) / sum(float((ref.v[c] ** 2 * om.bm1).sum()) for c in range(2)))

    st = CPertStepper(om, prm, precond=pre)
    y_free = ExptA(st, 0.1, bf).matvec(x0)
    st_f = CPertStepper(om, prm, precond=pre); st_f.forcing = force
    y_forced = ExptA(st_f, 0.1, bf).matvec(x0)
    assert wdiff(A.matvec(to_dev(x0)), y_free) < 1e-10
    ctx.set_forcing(force, ipert=0)                       # nonlinear slot: the linear step must not see it
    assert wdiff(A.matvec(to_dev(x0)), y_free) < 1e-10
    ctx.set_forcing(force, ipert=1)
    assert wdiff(A.matvec(to_dev(x0)), y_forced) < 1e-10
    got = ctx.get_forcing(1)
    assert np.array_equal(got[0], force[0]) and np.array_equal(got[1], force[1])
    # nonlinear map with the slot-0 forcing (nek_system%response, fixed_point.f90:4-40)
    r_dev = api.nonlinear_map(ctx, 0.1, to_dev(bf), cfl_limit=0.4)
    r_or = nonlinear_map(st_f, bf, 0.1, 0.4)
    v, _, _ = r_dev.download()
    err = np.sqrt(sum(float(((v[c] - r_or.v[c]) ** 2 * om.bm1).sum()) for c in range(2))) / r_or.norm()
    assert err < 1e-9
    ctx.zero_forcing(1)
    assert wdiff(A.matvec(to_dev(x0)), y_free) < 1e-10
    ctx.zero_forcing()
    assert np.abs(ctx.get_forcing(0)[0]).max() == 0.0
    for bad in (-1, 2):
        with pytest.raises(api.NlkError):
            ctx.set_forcing(force, ipert=bad)
    ctx.close()


def test_ext_dvector_semantics(small):
    """nek_ext_dvector (src/vectors/real_extended_vectors.f90): fields on the device + the period T as one more degree of freedom."""
    from neklab_b200 import api
    om, ctx = small
    a = api.nek_ext_dvector(ctx); b = api.nek_ext_dvector(ctx)
    a.rand(False, 3); b.rand(False, 4)
    assert 0.0 <= a.T < 1.0 and a.T != b.T
    d0 = a.vec.dot(b.vec)
    assert abs(a.dot(b) - (d0 + a.T * b.T)) < 1e-14 * abs(d0)                  # :243
    assert a.get_size() == a.vec.get_size() + 1
    Ta, Tb = a.T, b.T
    a.axpby(2.0, b, 3.0)                                                       # :193
    assert abs(a.T - (3.0 * Ta + 2.0 * Tb)) < 1e-15
    a.scal(0.5); assert abs(a.T - 0.5 * (3.0 * Ta + 2.0 * Tb)) < 1e-15
    a.save_rst(b, 1); assert a.get_rst(1).T == b.T
    a.rand(True, 9); assert abs(a.norm() - 1.0) < 1e-13
    a.zero(); assert a.T == 0.0 and a.norm() == 0.0
