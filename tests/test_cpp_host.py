"""The compiled host layer: include/neklab.hpp (C++17 mirror of nek_dvector / exptA_linop / the analysis drivers over the C-ABI)
and examples/channel_eigs.cpp.  CPU: it compiles warning-free against include/nlk.h, links against libnlk.so, builds its mesh
through the C-ABI and then stops loudly because there is no GPU (no CPU fallback).  GPU: the same binary reproduces the
Orr-Sommerfeld eigenvalue."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def channel_binary(nlk_lib, tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cpp") / "channel_eigs")
    libdir = os.path.join(ROOT, "neklab_b200")
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "channel_eigs.cpp"), "-L" + libdir, "-lnlk", "-Wl,-rpath," + libdir, "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def test_header_is_plain_cxx17(nlk_lib):
    """neklab.hpp needs nothing but the C header: no CUDA, no torch (the boundary stays a C-ABI)."""
    src = open(os.path.join(ROOT, "include", "neklab.hpp")).read()
    assert not re.search(r"#include\s*<(cuda|torch|ATen)", src)
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-x", "c++", "-I" + os.path.join(ROOT, "include"), "-"],
                       input='#include "neklab.hpp"\nint main() { return 0; }\n', capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_round2_host_types_compile_and_link(nlk_lib, tmp_path):
    """nek_zvector / resolvent_linop / nek_ext_dvector / nek_upo_jacobian of the host mirror compile warning-free and every C entry
    point they call resolves against libnlk.so."""
    src = tmp_path / "r2.cpp"
    src.write_text("""#include "neklab.hpp"
void use(neklab::context& c, neklab::nek_dvector& bf) {
  neklab::nek_zvector a(c), b(c); neklab::resolvent_linop R(2.0, bf); R.matvec(a, b); R.rmatvec(a, b); a.axpby({1, 0}, b, {0, 1}); (void)a.dot(b);
  neklab::nek_ext_dvector X(c, 1.0), x(c), y(c); neklab::nek_upo_jacobian J(c, X); J.matvec(x, y); J.rmatvec(x, y); y.axpby(1.0, x, 2.0); (void)y.norm();
}
int main() { return 0; }
""")
    libdir = os.path.join(ROOT, "neklab_b200")
    r = subprocess.run(["g++", "-std=c++17", "-O0", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), str(src), "-L" + libdir, "-lnlk",
                        "-Wl,-rpath," + libdir, "-Wl,--no-as-needed", "-o", str(tmp_path / "r2")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_channel_example_builds_and_fails_loudly_without_gpu(channel_binary):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([channel_binary], capture_output=True, text=True, timeout=120)
    assert "mesh: 80 elements, lx1 = 8, 3976 unique nodes" in r.stdout            # host setup ran through the C-ABI
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.skipif(not os.environ.get("NLK_LONG_TESTS"), reason="long (128 matvecs of the Re = 7500 channel): opt-in, passes on a B200 (r02)")
def test_channel_example_reproduces_orr_sommerfeld(channel_binary, tmp_path):
    from tests.util import orr_sommerfeld_leading
    r = subprocess.run([channel_binary, "7500", "100", "2", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"eig 0: .* sigma = ([-+0-9.eE]+) ([-+0-9.eE]+) i", r.stdout)
    lam = orr_sommerfeld_leading(7500.0)
    assert abs(float(m.group(1)) - lam.real) < 2e-3 and abs(abs(float(m.group(2))) - abs(lam.imag)) < 2e-3     # literal rst arithmetic: O(dt) bias
    assert os.path.exists(tmp_path / "eigs_output.txt")
