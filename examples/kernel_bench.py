"""Per-kernel roofline on the synthetic 3-D extruded-cylinder mesh (lx1=8, lxd=12): CUDA-event timing on the library
stream, algorithmic bytes from DESIGN.md §3, peak from MEASURED_PEAKS.json.

    python examples/kernel_bench.py [--layers 16] [--which 0,1,2,3,4,5] [--nrep 20]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from neklab_b200 import api, build  # noqa: E402

NAMES = {0: "axhelm (K1)", 1: "dssum (K2)", 2: "cdabdtp = opgradt+gs+opdiv (K6)", 3: "convect x d fields (K3)", 4: "precond Schwarz (K10)", 5: "vec dot (K12)", 6: "coarse solve, sparse PCG (K11)", 7: "Schwarz branch alone (K10)", 8: "fused Helmholtz apply of the PCG (K1/K8)", 9: "PCG update+reduce (K8)",
         10: "opgradt (K5)", 11: "opdiv, fused load scaling (K5)", 12: "dssum x 3 fields (K2)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=16)
    ap.add_argument("--which", default="0,1,2,3,4,5")
    ap.add_argument("--nrep", type=int, default=20)
    ap.add_argument("--lx1", type=int, default=8)
    ap.add_argument("--precond", type=int, default=2, help="2 Schwarz only (default), 4 Schwarz + sparse coarse (needed for --which 6)")
    a = ap.parse_args()
    build.build()
    peak, src = bench.load_peaks()
    case = bench.cylinder_inputs()
    coords, U, vertex, cbc = bench.extrude(case, a.lx1, a.layers)
    mesh = api.Mesh(coords, vertex, cbc, a.lx1 * 3 // 2)
    ctx = api.Context(mesh, api.default_params(viscosity=0.02, precond=a.precond))
    npts = coords.shape[0] * a.lx1 ** 3
    out = {"elements": int(coords.shape[0]), "points": int(npts), "peak_GBps": peak, "peak_source": src, "kernels": {}}
    for w in [int(x) for x in a.which.split(",")]:
        ms, by = ctx.bench_kernel(w, a.nrep)
        out["kernels"][NAMES[w]] = {"us": ms * 1e3, "algorithmic_MB": by / 1e6, "GBps": by / ms / 1e6, "frac_of_peak": by / ms / 1e6 / peak}
        print(f"{NAMES[w]:36s} {ms*1e3:10.1f} us  {by/1e6:9.1f} MB  {by/ms/1e6:8.1f} GB/s  {by/ms/1e6/peak*100:5.1f}% of {peak:.0f}", flush=True)
    print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
