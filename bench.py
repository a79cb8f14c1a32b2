#!/usr/bin/env python
"""bench.py -- exptA hot path on B200 (BASELINE.json metric: exptA matvec/s & GDOF.steps/s; Arnoldi time-to-leading-eigs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload synth3d|cylinder]
                    [--scaling strong|weak] [--layers-total 50] [--layers 6]

Workloads
  synth3d  : BASELINE.json configs[4]: the 2-D cylinder mesh (1 996 elements) extruded periodically in z, lx1 = 8, lxd = 12.
             Default = the headline size, 50 z-layers = 99 800 elements (51.1 M points per field), on ONE GPU at N = 1 and
             split into z-slabs of whole layers over the ranks at N > 1 (STRONG scaling: total work fixed).  `--scaling weak`
             keeps `--layers` z-layers per GPU instead.  One bench "step" = one perturbation time step (the body of the
             exptA loop, exponential_propagator.f90:39-46), timed after --spinup steady-state steps.  Metric GDOF*steps/s.
             At N = 1 the JSON line also carries the cylinder Re=50 block (exptA matvec/s and time-to-leading-eigs).
  cylinder : examples/cylinder/stability/direct of the reference (Re=50, 1996 el, lx1=6, lxd=9, bdf3, tau=1, 100+2 time steps
             per matvec).  One bench "step" = one exptA matvec.  Metric exptA matvec/s.  (N = 1 only)

Prints ONE JSON line (rank 0).  `value` is device-timed (CUDA events on the library's stream, max over ranks) with inputs
resident in HBM; `e2e` goes through the public C-ABI with host buffers (H2D of the input vector and D2H of the result inside
the timed region).  `--impl reference` times the reference's CPU path for the SAME workload on the host cores: the
Fortran/MPI reference cannot be built in this image (no Fortran compiler, no MPI), so it is the C++/OpenMP
restatement oracle/cpp/nekref.cpp (kind "port"), same discretisation and solver algorithm, on a bounded sample (a window
of the same extruded mesh), honouring --steps / --warmup.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LAYERS_HEADLINE = 50          # 50 x 1996 = 99 800 elements ("~100k elements")
# measured DRAM bytes per launch of the roofline kernel (k_axhelm8a<FUSE_CG>), keyed by points per GPU: ncu --set full,
# profiles/r02_ncu_axhelm8a_headline.md (4.0882 GB read + 0.8055 GB written; algorithmic 4.9054 GB)
NCU_TRAFFIC_AXHELM8A = {51097600: 4.8938e9}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, gpu=0):
        self.gpu = gpu; self.rows = []; self.p = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            pass
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- workloads
def cylinder_inputs():
    z = np.load(os.path.join(ROOT, "tests", "golden", "cylinder_case.npz"))
    return dict(coords=z["coords"], vertex=z["vertex"], cbc=z["cbc"], vel=z["vel"], pr=z["pr"], pid=z["pid"])


def zslab_partition(E2, L, world):
    """Strong-scaling partition of the extruded mesh: z-slabs of whole 2-D layers (uneven when world does not divide L).
    Returns (layer bounds per rank, element -> rank map or None for one rank)."""
    bounds = np.linspace(0, L, world + 1).round().astype(int)
    if world <= 1:
        return bounds, None
    gllnid = np.zeros(E2 * L, dtype=np.int32)
    for r in range(world):
        gllnid[E2 * bounds[r]:E2 * bounds[r + 1]] = r
    return bounds, gllnid


def extrude(case, n, layers_total, lz_per_layer=0.5, layer_range=None):
    """Extrude the 2-D cylinder mesh (bilinear re-interpolation to lx1=n) into `layers_total` periodic z-layers (>= 3: with two
    layers the vertical edges of both layers would carry the same vertex pair).  Connectivity (vertex ids, boundary codes) is
    always global; with layer_range = (l0, l1) the coordinates and the base flow are generated for those layers only (what one
    rank of a z-slab partition owns), in the same element order."""
    from neklab_b200.boxmesh import gll_points, lagrange_interp      # host-side input generation (no oracle on this path)
    c2 = case["coords"]; E2 = c2.shape[0]; n0 = c2.shape[-1]
    z0 = gll_points(n0); z1 = gll_points(n)
    I = lagrange_interp(z1, z0)
    up = lambda a: np.einsum("qj,pi,e...ji->e...qp", I, I, a)
    xy = up(c2[:, :, 0]); vel = up(case["vel"][:, :, 0])
    L = layers_total
    l0, l1 = (0, L) if layer_range is None else layer_range
    coords = np.zeros((E2 * (l1 - l0), 3, n, n, n)); U = np.zeros((E2 * (l1 - l0), 3, n, n, n))
    zz = 0.5 * (z1 + 1.0) * lz_per_layer
    nv2 = int(case["vertex"].max())
    vertex = np.zeros((E2 * L, 8), dtype=np.int64)
    cbc = np.full((E2 * L, 6), "E  ", dtype="U3")
    for l in range(L):
        sl = slice(l * E2, (l + 1) * E2)
        vertex[sl, :4] = case["vertex"] + l * nv2
        vertex[sl, 4:] = case["vertex"] + ((l + 1) % L) * nv2
        cbc[sl, :4] = case["cbc"]; cbc[sl, 4:] = "P  "
        if l0 <= l < l1:
            sl = slice((l - l0) * E2, (l - l0 + 1) * E2)
            coords[sl, 0] = xy[:, 0][:, None]; coords[sl, 1] = xy[:, 1][:, None]
            coords[sl, 2] = (l * lz_per_layer + zz)[None, :, None, None]
            U[sl, 0] = vel[:, 0][:, None]; U[sl, 1] = vel[:, 1][:, None]
    return coords, U, vertex, cbc


def window_case(case, half_width):
    """The 2-D elements of the cylinder mesh whose centroid lies within `half_width` of the cylinder, renumbered; cut faces
    become 'v  ' (inflow side and lateral) or 'O  ' (downstream side).  The bounded sample of the CPU arm and the
    stepper-parity test of tests/test_gpu_reference_configs.py extrude this window exactly like the full mesh."""
    c = case["coords"]; E = c.shape[0]
    cen = c[:, :, 0].reshape(E, 2, -1).mean(axis=2)
    sel = np.where((np.abs(cen[:, 0] - 0.5) < half_width) & (np.abs(cen[:, 1]) < half_width))[0]
    vert = case["vertex"][sel]
    _, inv = np.unique(vert, return_inverse=True)
    vert = (inv.reshape(vert.shape) + 1).astype(np.int64)
    cbc = case["cbc"][sel].copy()
    fc = {0: (0, 1), 1: (1, 3), 2: (2, 3), 3: (0, 2)}                      # preprocessor face -> lexicographic corner pair
    fidx = {0: (0, slice(None)), 1: (slice(None), -1), 2: (-1, slice(None)), 3: (slice(None), 0)}
    keys = {}
    for e in range(len(sel)):
        for f, (a, b) in fc.items():
            keys.setdefault(tuple(sorted((vert[e, a], vert[e, b]))), []).append((e, f))
    xs = c[sel][:, 0, 0]
    cut = [(e, f) for lst in keys.values() if len(lst) == 1 for e, f in lst if cbc[e, f] == "E  "]
    xcut = max(xs[e][fidx[f]].mean() for e, f in cut)
    for e, f in cut:
        cbc[e, f] = "O  " if xs[e][fidx[f]].mean() > xcut - 1e-6 else "v  "
    return dict(coords=c[sel], vel=case["vel"][sel], vertex=vert, cbc=cbc)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default=None, choices=[None, "cylinder", "synth3d"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--layers-total", type=int, default=LAYERS_HEADLINE, help="synth3d, strong scaling: z-layers of the whole mesh (1996 elements each)")
    ap.add_argument("--layers", type=int, default=6, help="synth3d, weak scaling: z-layers per GPU")
    ap.add_argument("--no-eigs", action="store_true", help="cylinder block: skip the Krylov-Schur run to the leading eigenpair (~25 s)")
    ap.add_argument("--spinup", type=int, default=30, help="synth3d: untimed time steps before the timed ones (>= warmup)")
    ap.add_argument("--no-cylinder", action="store_true", help="skip the extra cylinder Re=50 block at N=1")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--coarse-iters", type=int, default=8, help="synth3d: Jacobi-PCG iterations of the sparse coarse solve")
    ap.add_argument("--cpu-window", type=float, default=6.0, help="half width of the mesh window the CPU sample runs on")
    ap.add_argument("--e2e-tau", type=float, default=1.0, help="synth3d: integration horizon of the end-to-end matvec (reference configs: tau = 1)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    workload = a.workload or "synth3d"
    steps = a.steps if a.steps is not None else (5 if workload == "cylinder" else 10)
    if a.impl == "reference":
        if rank == 0:
            print(json.dumps(reference_arm(workload, steps, a.warmup, a)), flush=True)
        return
    out = native_arm(workload, steps, a.warmup, a, rank, world, local)
    if rank == 0:
        print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------------------------- CPU arm (C++/OpenMP oracle)
def cpu_sample(workload, nsteps, nwarm, half_width=6.0):
    """Time `nsteps` perturbation time steps (after `nwarm` untimed ones, >= 3 for the BDF3/EXT3 start-up) of the SAME
    workload with the C++/OpenMP restatement (oracle/cpp/nekref.cpp: same discretisation, Jacobi-PCG + Schwarz/coarse FGMRES(30)
    with the 20-vector residual projection, same tolerances) on every host core.
      synth3d : a window of the extruded mesh (|x - 0.5|, |y| < half_width around the cylinder, 3 periodic z-layers, lx1 = 8, lxd = 12)
      cylinder: the full 2-D config (1996 elements, lx1 = 6, lxd = 9).
    Returns dict(sec_per_step, points, steps_per_matvec, threads, sample)."""
    from oracle import ops
    from oracle.cref import CPertStepper
    from oracle.mesh import SEMesh
    from oracle.precond import SchwarzCoarse
    from oracle.stepper import NekVec, StepParams
    from oracle.cref import CRef
    CRef.use_all_cores()                                  # under torchrun the workers inherit OMP_NUM_THREADS=1
    case = cylinder_inputs()
    if workload == "cylinder":
        om = SEMesh(case["coords"], case["vertex"], case["cbc"], 9)
        U = [case["vel"][:, 0].copy(), case["vel"][:, 1].copy()]
        desc = "the full cylinder Re=50 config (1996 elements, lx1=6, lxd=9)"
    else:
        coords, U3, vertex, cbc = extrude(window_case(case, half_width), 8, 3)
        om = SEMesh(coords, vertex, cbc, 12)
        U = [U3[:, k].copy() for k in range(3)]
        desc = f"a window of the same extruded mesh ({om.E} elements = {om.E // 3} near-cylinder 2-D elements x 3 periodic z-layers, lx1=8, lxd=12, {om.bm1.size} points)"
    prm = StepParams(viscosity=1.0 / 50.0, torder=3, vtol=1e-9, ptol=1e-7, gmres_maxit=100)
    st = CPertStepper(om, prm, precond=SchwarzCoarse(om), pr_proj=20)
    st.U = U
    st.setup(1.0, 0.5, False)
    x = om.coords
    v = [om.vmask[c] * om.dssum(np.sin(0.7 * x[:, 0] + 0.3 * c) * np.cos(0.9 * x[:, 1] - 0.2 * c) * (np.cos(2 * np.pi * x[:, 2] / 1.5 + c) if om.ndim == 3 else 1.0)) * om.vmult
         for c in range(om.ndim)]
    st.set_state(v, np.zeros_like(om.bm2)); st.reset_history()
    nwarm = max(int(nwarm), 3)
    for i in range(1, nwarm + 1):
        st.advance(i)
    cg0, gm0 = st.ref.counters()
    t0 = time.perf_counter()
    for i in range(nwarm + 1, nwarm + 1 + nsteps):
        st.advance(i)
    sec = (time.perf_counter() - t0) / nsteps
    cg1, gm1 = st.ref.counters()
    return dict(sec_per_step=sec, points=int(om.bm1.size), steps_per_matvec=st.nsteps + 2, threads=st.ref.threads(),
                cg_iters_per_step=(cg1 - cg0) / nsteps, gmres_iters_per_step=(gm1 - gm0) / nsteps,
                sample=f"{nsteps} time steps (after {nwarm} untimed) of {desc}; C++/OpenMP restatement oracle/cpp/nekref.cpp, "
                       f"Jacobi-PCG + Schwarz/coarse FGMRES(30) + 20-vector pressure projection, tolerances 1e-9 / 1e-7")


def _cpu_value(workload, s):
    if workload == "cylinder":
        return 1.0 / (s["sec_per_step"] * s["steps_per_matvec"]), "matvec/s"
    return s["points"] * 1e-9 / s["sec_per_step"], "GDOF*steps/s"


def synth_config(a, world):
    L = a.layers_total if a.scaling == "strong" else a.layers * world
    return {"workload": "synth3d_extruded_cylinder_time_step", "elements": 1996 * L, "lx1": 8, "lxd": 12, "layers_total": L,
            "timestepper": "bdf3", "residualProj": 20, "tolerances": {"velocity": 1e-9, "pressure": 1e-7},
            "l2": "inputs larger than L2 (state + geometry + dealiasing metrics ~ %.1f GB in total)" % (1996 * L * 512 * 8 * 75 / 1e9)}


def reference_arm(workload, steps, warmup, a):
    s = cpu_sample(workload, steps, warmup, a.cpu_window)
    val, unit = _cpu_value(workload, s)
    cfg = synth_config(a, max(a.gpus, 1)) if workload == "synth3d" else {"workload": "cylinder_re50_exptA_matvec", "elements": 1996, "lx1": 6, "lxd": 9, "timestepper": "bdf3", "tau": 1.0, "residualProj": 20}
    ms = s["sec_per_step"] * 1e3 * (s["steps_per_matvec"] if workload == "cylinder" else 1)
    base = {"value": val, "unit": unit, "cores": s["threads"], "kind": "port", "sample": s["sample"]}
    return {"metric": "exptA matvec/s" if workload == "cylinder" else "GDOF*steps/s", "value": val, "unit": unit, "n_gpus": 0,
            "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference", "config": cfg, "cpu_baseline": base,
            "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "cg_iters_per_step": s["cg_iters_per_step"], "gmres_iters_per_step": s["gmres_iters_per_step"],
            "note": "CPU restatement (C++/OpenMP) of the Nek5000+LightKrylov path: the Fortran/MPI reference cannot be built in this image (no Fortran compiler, no MPI)"}


# ----------------------------------------------------------------------------------------------- native arm
def _cylinder_block(api, case, steps, warmup, device, do_eigs=True):
    """Cylinder Re=50 on ONE GPU (the reference's own config): device-timed and end-to-end exptA matvec/s, and the wall time
    of the Krylov-Schur run to the leading eigenpair (BASELINE metric 3)."""
    mesh = api.Mesh(case["coords"], case["vertex"], case["cbc"], 9)
    prm = api.default_params(viscosity=1.0 / 50.0, torder=3, vtol=1e-9, ptol=1e-7, pr_proj=20)   # 1cyl.par: residualProj = yes (mxprev 20)
    ctx = api.Context(mesh, prm, device=device)
    vel = case["vel"][:, :, 0]
    bf = ctx.vec(); bf.upload([vel[:, 0][:, None], vel[:, 1][:, None]])
    x = ctx.vec(); x.rand(ifnorm=True, seed=12345); y = ctx.vec()
    A = api.exptA_linop(ctx, 1.0, bf); A.init()
    for _ in range(warmup):
        A.matvec(x, y)
    ctx.sync()
    ms = 0.0; launches = 0; cg = gm = ts = 0
    for _ in range(steps):
        A.matvec(x, y)
        s = A.stats(); ms += s["ms_total"]; launches += s["launches"]; cg += s["cg_iters"]; gm += s["gmres_iters"]; ts += s["steps"]
    hv, hp, _ = x.download()
    t0 = time.perf_counter()
    for _ in range(steps):
        x.upload(hv, hp); A.matvec(x, y); ov, op, _ = y.download()
    ctx.sync()
    e2e_s = (time.perf_counter() - t0) / steps
    vec_bytes = sum(v.nbytes for v in hv) + hp.nbytes
    ms_ax, by_ax = ctx.bench_kernel(0, 200)
    out = {"matvec_per_s": 1e3 / (ms / steps), "ms_per_matvec": ms / steps, "e2e_matvec_per_s": 1.0 / e2e_s, "h2d_bytes": int(vec_bytes), "d2h_bytes": int(vec_bytes),
           "time_steps_per_matvec": ts // steps, "cg_iters_per_step": cg / max(ts, 1), "gmres_iters_per_step": gm / max(ts, 1),
           "launches_per_time_step": launches / max(ts, 1), "launches": int(launches),
           "gdof_steps_per_s": case["coords"].shape[0] * 36 * 1e-9 * (ts / steps) / (ms / steps * 1e-3),
           "axhelm_us": ms_ax * 1e3, "axhelm_GBps_L2_resident": by_ax / (ms_ax * 1e-3) / 1e9,
           "config": {"elements": int(case["coords"].shape[0]), "lx1": 6, "lxd": 9, "timestepper": "bdf3", "tau": 1.0, "residualProj": 20,
                      "note": "71 856 points (0.57 MB per field): L2-resident, launch/latency-bound by construction"}}
    out["time_to_leading_eigs_s"] = None
    if do_eigs:
        t0 = time.perf_counter()
        r = api.linear_stability_analysis_fixed_point(A, 128, 2)
        ctx.sync()
        out["time_to_leading_eigs_s"] = time.perf_counter() - t0
        out["time_to_leading_eigs_note"] = "LightKrylov eigs semantics (Krylov-Schur, kdim 128, nev 2, default tolerance), device-resident basis, restart arithmetic as in the reference"
        out["eigs"] = {"modulus": [float(abs(v)) for v in r["lam"]], "resid": [float(v) for v in r["resid"]], "niter": int(r["niter"]), "info": int(r["info"])}
    ctx.close()
    return out


def native_arm(workload, steps, warmup, a, rank, world, local):
    from neklab_b200 import api, build
    if rank == 0:
        build.build()
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_
        torch.cuda.set_device(local)
        dist_.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_
        dist.barrier()
    peak, peak_src = load_peaks()
    case = cylinder_inputs()
    sampler = ClockSampler(local)
    extra = {}
    other = {}
    if workload == "cylinder":
        if world > 1:
            raise SystemExit("the cylinder workload is a single-GPU config (71 856 points); use --workload synth3d for N > 1")
        sampler.start()
        cyl = _cylinder_block(api, case, steps, warmup, local, not a.no_eigs)
        clocks = sampler.stop()
        ms_step = cyl["ms_per_matvec"]; value = cyl["matvec_per_s"]; unit = "matvec/s"; metric = "exptA matvec/s"
        e2e = {"value": cyl["e2e_matvec_per_s"], "unit": unit, "h2d_bytes_per_step": cyl["h2d_bytes"], "d2h_bytes_per_step": cyl["d2h_bytes"]}
        launches = cyl["launches"]
        cfg = dict(cyl["config"]); cfg["workload"] = "cylinder_re50_exptA_matvec"
        extra = {k: cyl[k] for k in ("gdof_steps_per_s", "time_steps_per_matvec", "cg_iters_per_step", "gmres_iters_per_step", "launches_per_time_step", "time_to_leading_eigs_s")}
        if "eigs" in cyl:
            extra["eigs"] = cyl["eigs"]
        roof = {"kernel": "k_axhelm<6,2> (K1), timed alone", "bound": "hbm", "achieved": cyl["axhelm_GBps_L2_resident"], "peak": peak, "unit": "GB/s",
                "frac": cyl["axhelm_GBps_L2_resident"] / peak, "traffic": None, "peak_source": peak_src, "note": "L2-resident problem: not an HBM measurement"}
        scaling = "strong"
    else:
        cfg = synth_config(a, world)
        L = cfg["layers_total"]
        E2 = case["coords"].shape[0]
        bounds, gllnid = zslab_partition(E2, L, world)
        l0, l1 = int(bounds[rank]), int(bounds[rank + 1])
        coords, Uall, vertex, cbc = extrude(case, 8, L, layer_range=(l0, l1))
        t_setup = time.perf_counter()
        mesh = api.Mesh(coords, vertex, cbc, 12, gllnid=gllnid, rank=rank, nranks=world)
        prm = api.default_params(viscosity=1.0 / 50.0, torder=3, vtol=1e-9, ptol=1e-7, pr_proj=20, coarse_iters=a.coarse_iters)
        nccl_id = None
        if world > 1:
            buf = [api.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(buf, src=0)
            nccl_id = buf[0]
        ctx = api.Context(mesh, prm, device=local, nccl_id=nccl_id)
        t_setup = time.perf_counter() - t_setup
        bf = ctx.vec(); bf.upload([Uall[:, 0], Uall[:, 1], Uall[:, 2]])
        del coords, Uall
        x = ctx.vec(); x.rand(ifnorm=True, seed=12345); y = ctx.vec()
        npts = mesh.nel * 512
        npts_global = E2 * L * 512
        A = api.exptA_linop(ctx, 1.0, bf)
        s0 = A.init()
        spin = max(a.spinup, warmup)           # untimed steps: BDF start-up + projection space fill (W >= 3 always holds)
        ctx.sync()
        if dist is not None:
            dist.barrier()
        sampler.start()
        ms = A.time_steps(x, spin, steps)
        ctx.sync()
        if dist is not None:
            dist.barrier()
        clocks = sampler.stop()
        s = A.stats()
        ms_step = ms / steps
        state_norm = ctx.nek2vec(y).norm()      # global bm1 norm of the state after spin + steps time steps: identical for every N (strong scaling)
        # e2e: ONE full exptA matvec (tau = 1 as in the reference's configs: nsteps + 2 restart steps) through the public API --
        # host vector in (H2D), host vector out (D2H) inside the timed region; its per-step average includes the cold
        # BDF/projection start-up of the matvec.  --e2e-tau shortens it.
        import ctypes as C
        hv, hp, _ = x.download()
        api.lib().nlk_exptA_set_tau(A.h, C.c_double(a.e2e_tau))
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter(); x.upload(hv, hp); A.matvec(x, y); ov, op, _ = y.download(); ctx.sync()
        e2e_steps = A.stats()["steps"]
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        vec_bytes = sum(v.nbytes for v in hv) + hp.nbytes
        if dist is not None:
            import torch
            t = torch.tensor([ms_step, e2e_s], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_step, e2e_s = float(t[0]), float(t[1])
        value = npts_global * 1e-9 / (ms_step * 1e-3); unit = "GDOF*steps/s"; metric = "GDOF*steps/s"
        e2e = {"value": npts_global * 1e-9 / e2e_s, "unit": unit, "h2d_bytes_per_step": int(vec_bytes // e2e_steps), "d2h_bytes_per_step": int(vec_bytes // e2e_steps),
               "note": "one exptA matvec of %d time steps (tau = %g, incl. 2 restart steps) with host buffers; bytes are per time step and per rank" % (e2e_steps, a.e2e_tau)}
        launches = s["launches"]
        extra = {"time_steps_timed": int(s["steps"]), "spinup_steps": int(spin), "cg_iters_per_step": s["cg_iters"] / steps, "gmres_iters_per_step": s["gmres_iters"] / steps,
                 "launches_per_time_step": launches / steps, "dt": s0["dt"], "points_per_gpu": int(npts), "setup_s": t_setup,
                 "check": {"state_norm_after_timed_steps": state_norm, "note": "global bm1 norm of the perturbation after spinup+steps time steps from the same seeded start vector; equal across N up to solver tolerance"}}
        extra["partition"] = ("z-slabs of whole 2-D layers, layer bounds %s" % [int(b) for b in bounds]) if world > 1 else "single rank"
        # roofline of the dominant kernel of the step, at this problem size: the fused Helmholtz apply the Jacobi-PCG launches
        # (p = hd r + beta p on load, p.Ap on the way out); CUDA events on the library stream right after the timed region
        ms_ax, bytes_ax = ctx.bench_kernel(8, 50)
        share = s["cg_iters"] / steps * ms_ax / ms_step
        roof = {"kernel": "k_axhelm8a<FUSE_CG=true> (K1/K8: fused Helmholtz apply of the Jacobi-PCG, lx1 = 8, cp.async-staged factors), timed back-to-back on the step's own buffers",
                "bound": "hbm", "achieved": bytes_ax / (ms_ax * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": bytes_ax / (ms_ax * 1e-3) / 1e9 / peak,
                "traffic": NCU_TRAFFIC_AXHELM8A.get(int(npts)), "peak_source": peak_src, "us_per_launch": ms_ax * 1e3, "algorithmic_bytes_per_launch": bytes_ax,
                "launches_per_step": s["cg_iters"] / steps, "share_of_step": share,
                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel from one ncu --set full capture at this size "
                                "(profiles/r02_ncu_axhelm8a_headline.md); null at sizes that were not captured"}
        names = (("axhelm_plain", 0), ("cg_update_reduce", 9), ("dssum", 1), ("dssum3", 12), ("cdabdtp", 2), ("opgradt", 10), ("opdiv", 11), ("convect", 3), ("precond", 4), ("schwarz", 7), ("vec_dot", 5))
        for name, which in names:
            try:
                m_, b_ = ctx.bench_kernel(which, 10)
                other[name] = {"us": m_ * 1e3, "GBps": b_ / (m_ * 1e-3) / 1e9, "frac": b_ / (m_ * 1e-3) / 1e9 / peak}
            except Exception as e:          # e.g. no sparse coarse level at this size
                other[name] = {"error": str(e)}
        ctx.close()
        del ctx, mesh
        scaling = a.scaling
        if world == 1 and not a.no_cylinder:
            extra["cylinder_re50"] = _cylinder_block(api, case, 3, 3, local, not a.no_eigs)
            extra["time_to_leading_eigs_s"] = extra["cylinder_re50"]["time_to_leading_eigs_s"]
    out = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_step,
           "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "config": cfg, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
           "kernels": other, **extra}
    if world == 1 and rank == 0 and not a.no_cpu and os.environ.get("NLK_BENCH_NO_CPU") != "1":
        try:
            s = cpu_sample(workload, 5, 3, a.cpu_window)
            v, u = _cpu_value(workload, s)
            out["cpu_baseline"] = {"value": v, "unit": u, "cores": s["threads"], "kind": "port", "sample": s["sample"],
                                   "cg_iters_per_step": s["cg_iters_per_step"], "gmres_iters_per_step": s["gmres_iters_per_step"]}
        except Exception as e:  # the baseline is a reported number, never the product path
            out["cpu_baseline"] = {"value": None, "unit": unit, "cores": os.cpu_count() or 1, "kind": "port", "sample": f"failed: {e}"}
    if dist is not None:
        dist.destroy_process_group()
    return out


if __name__ == "__main__":
    main()
