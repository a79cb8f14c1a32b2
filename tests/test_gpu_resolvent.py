"""`resolvent_linop` (src/linops/resolvent.f90) through the C-ABI against the oracle: the two forced integrations (evaluate_rhs,
evaluate_imaginary_part) at 1e-10, the whole matvec / rmatvec at the accuracy the GMRES tolerance allows."""
import numpy as np
import pytest

from oracle.cref import CPertStepper
from oracle.precond import SchwarzCoarse
from oracle.stepper import Resolvent, StepParams
from tests.test_oracle_resolvent import resolvent_case, wdiff
from tests.util import nlk_mesh

pytestmark = pytest.mark.gpu


def _dev(ctx, nv):
    d = ctx.vec(); d.upload(nv.v, nv.pr); return d


class _Host:
    def __init__(self, dev):
        self.v, self.pr, _ = dev.download()


def test_resolvent_parity(nlk_lib):
    from neklab_b200 import api
    om, bf, fre, fim, kw = resolvent_case(torder=3)
    st = CPertStepper(om, StepParams(**kw), precond=SchwarzCoarse(om))
    omega = 2.0; tau = 2 * np.pi / omega
    ctx = api.Context(nlk_mesh(om), api.default_params(**kw))
    f = api.nek_zvector(ctx, _dev(ctx, fre), _dev(ctx, fim))
    Rd = api.resolvent_linop(ctx, omega, _dev(ctx, bf), rtol=1e-10)
    Ro = Resolvent(st, omega, bf, rtol=1e-10)
    # evaluate_rhs (direct and adjoint) and evaluate_imaginary_part from a given state
    for adj in (False, True):
        b_o = Ro.integrate(tau, fre, fim, None, adj)
        b_d = Rd.integrate(tau, f, None, adj)
        assert wdiff(om, _Host(b_d), b_o) < 1e-10
    y_o = Ro.integrate(tau / 4, fre, fim, b_o, False)
    y_d = Rd.integrate(tau / 4, f, _dev(ctx, b_o), False)
    assert wdiff(om, _Host(y_d), y_o) < 1e-10
    # the forcing registry is left zeroed (zero_neklab_forcing, resolvent.f90:107,161)
    assert np.abs(ctx.get_forcing(1)[0]).max() == 0.0
    # whole operator, tight GMRES on both sides
    re_o, im_o = Ro.matvec(fre, fim)
    out = Rd.matvec(f)
    assert Rd.info == 0
    assert wdiff(om, _Host(out.re), re_o) < 1e-7 and wdiff(om, _Host(out.im), im_o) < 1e-7
    re_a, im_a = Ro.rmatvec(fre, fim)
    out_a = Rd.rmatvec(f)
    assert wdiff(om, _Host(out_a.re), re_a) < 1e-7 and wdiff(om, _Host(out_a.im), im_a) < 1e-7
    # the reference's tolerance (rtol 1e-6): same answer to that accuracy
    R6 = api.resolvent_linop(ctx, omega, _dev(ctx, bf))
    out6 = R6.matvec(f)
    assert wdiff(om, _Host(out6.re), re_o) < 1e-4
    # input == output is an error, not silent aliasing
    with pytest.raises(api.NlkError):
        api.lib().nlk_resolvent_matvec  # symbol exists
        Rd._apply(f, f, False)
    ctx.close()
