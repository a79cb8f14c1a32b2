"""Periodic-orbit system and Jacobian (src/systems/periodic_orbit.f90) through the C-ABI against the oracle: base flow and
perturbation advanced together on the device (second state bank), direct and adjoint, with rst fields on the input."""
import numpy as np
import pytest

from oracle.stepper import NekExtVec, UPOJacobian, StepParams, nonlinear_map
from tests.test_oracle_upo import steppers, upo_case, wnorm
from tests.util import nlk_mesh

pytestmark = pytest.mark.gpu


def _ext(api, ctx, nv, T):
    e = api.nek_ext_dvector(ctx, T); e.vec.upload(nv.v, nv.pr); return e


def _err(om, dev_ext, ref_ext):
    v, _, _ = dev_ext.vec.download()
    return wnorm(om, [v[c] - ref_ext.vec.v[c] for c in range(2)]) / wnorm(om, ref_ext.vec.v)


def test_upo_jacobian_parity(nlk_lib):
    from neklab_b200 import api
    om, X, dx, kw = upo_case()
    lin, nl = steppers(om, kw)
    T = 0.2
    Jo = UPOJacobian(lin, nl, NekExtVec(X, T))
    ctx = api.Context(nlk_mesh(om), api.default_params(**kw))
    Jd = api.nek_upo_jacobian(ctx, _ext(api, ctx, X, T))
    vin_o = NekExtVec(dx.copy(), 0.3)
    vin_d = _ext(api, ctx, dx, 0.3)
    out_o = Jo.matvec(vin_o)
    out_d = Jd.matvec(vin_d)
    assert _err(om, out_d, out_o) < 1e-9
    assert abs(out_d.T - out_o.T) < 1e-9 * max(1.0, abs(out_o.T))
    assert out_d.vec.nrst == 2
    # the tolerances come back as vtol = ptol = atol (periodic_orbit.f90:106-107)
    assert ctx.params.vtol == kw["vtol"] and ctx.params.ptol == kw["vtol"]
    # second application, now with rst fields on the input (jac_get_rst), direct and adjoint
    for name in ("matvec", "rmatvec"):
        o2 = getattr(Jo, name)(out_o)
        d2 = getattr(Jd, name)(out_d)
        assert _err(om, d2, o2) < 1e-8, name
        assert abs(d2.T - o2.T) < 1e-8 * max(1.0, abs(o2.T)), name
    # nek_upo_system%response: F_T(X) - X with T from the vector; the period component of the residual is zero
    st = steppers(om, kw)[0]
    st.prm.vtol = st.prm.ptol = 1e-9 * 0.1; st.ref.set_params(st.prm, 0)
    r_o = nonlinear_map(st, X, T, 0.4)
    r_d = api.nek_upo_system(ctx).response(_ext(api, ctx, X, T), atol=1e-9)
    v, _, _ = r_d.vec.download()
    assert wnorm(om, [v[c] - r_o.v[c] for c in range(2)]) / wnorm(om, r_o.v) < 1e-7
    assert r_d.T == 0.0
    # the ordinary exptA still works after the coupled run (U pointers and state bank restored)
    bf = ctx.vec(); bf.upload(X.v, X.pr)
    A = api.exptA_linop(ctx, 0.1, bf)
    y = A.matvec(vin_d.vec)
    assert np.isfinite(y.norm()) and y.norm() > 0
    ctx.close()
