#!/usr/bin/env python
"""bench.py -- exptA matvec throughput on B200 (BASELINE.json metric: exptA matvec/s & GDOF.steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload cylinder|synth3d]

Workloads
  synth3d  : the 2-D cylinder mesh extruded periodically in z (lx1=8, lxd=12), element-partitioned over the ranks
             (weak scaling: --layers z-layers PER GPU).  One bench "step" = one perturbation time step (the body of the
             exptA loop), timed after --spinup steady-state steps.  Metric GDOF*steps/s.   [default, every N; at N=1
             the JSON line also carries the cylinder Re=50 matvec/s block]
  cylinder : examples/cylinder/stability/direct of the reference (Re=50, 1996 el, lx1=6, lxd=9, bdf3, tau=1,
             100+2 time steps per matvec).  One bench "step" = one exptA matvec.  Metric exptA matvec/s.  (N=1 only)

Prints ONE JSON line (rank 0).  `value` is device-timed (CUDA events on the library's stream, max over ranks) with
inputs resident in HBM; `e2e` goes through the public C-ABI with host buffers (H2D of the input vector and D2H of the
result inside the timed region).  `--impl reference` times the CPU restatement oracle (the Fortran/MPI reference
cannot be built in this image: no Fortran compiler, no MPI) on a bounded sample.
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


def extrude(case, n, layers_total, lz_per_layer=0.5, layer_range=None):
    """Extrude the 2-D cylinder mesh (bilinear re-interpolation to lx1=n) into `layers_total` periodic z-layers.
    Connectivity (vertex ids, boundary codes) is always global; with layer_range = (l0, l1) the coordinates and the base flow
    are generated for those layers only (what one rank of a z-slab partition owns), in the same element order."""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default=None, choices=[None, "cylinder", "synth3d"])
    ap.add_argument("--eigs", action="store_true", help="cylinder block: also time the Krylov-Schur run to the leading eigenpair (kdim 128, nev 2; ~25 s)")
    ap.add_argument("--layers", type=int, default=6, help="synth3d: z-layers per GPU (1996 elements each)")
    ap.add_argument("--cpu-steps", type=int, default=6, help="time steps in the CPU-baseline sample")
    ap.add_argument("--spinup", type=int, default=30, help="synth3d: untimed time steps before the timed ones (>= warmup)")
    ap.add_argument("--no-cylinder", action="store_true", help="skip the extra cylinder Re=50 matvec block at N=1")
    ap.add_argument("--coarse-iters", type=int, default=8, help="synth3d: Jacobi-PCG iterations of the sparse coarse solve")
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


# ----------------------------------------------------------------------------------------------- CPU baseline (oracle)
def cpu_sample(nsteps_sample):
    """Time `nsteps_sample` perturbation time steps of the cylinder config with the numpy oracle (same algorithm:
    Jacobi-PCG + Schwarz/coarse FGMRES).  Returns (seconds per time step, description)."""
    from oracle.precond import SchwarzCoarse
    from oracle.stepper import PertStepper, seeded_field
    from tests.util import cylinder_case
    om, bf, prm, _ = cylinder_case()
    prm.gmres_maxit = 100
    pc = SchwarzCoarse(om)
    st = PertStepper(om, prm, precond=pc)
    st.U = [x.copy() for x in bf.v]
    st.setup(1.0, 0.5, False)
    x0 = seeded_field(om, 12345)
    st.set_state(x0.v, x0.pr, x0.theta); st.reset_history()
    st.advance(1)                                   # warm caches / BDF1 step not timed
    t0 = time.perf_counter()
    for i in range(2, 2 + nsteps_sample):
        st.advance(i)
    dt = (time.perf_counter() - t0) / nsteps_sample
    return dt, st.nsteps, om, st.stats


def reference_arm(workload, steps, warmup, a):
    cores = os.cpu_count() or 1
    sec_step, nsteps, om, stats = cpu_sample(max(2, a.cpu_steps))
    n_ts = nsteps + 2
    matvec_s = 1.0 / (sec_step * n_ts)
    unit = "matvec/s"
    val = matvec_s
    sample = f"{max(2, a.cpu_steps)} perturbation time steps of the cylinder Re=50 config (of {n_ts} per matvec), numpy oracle, extrapolated"
    if workload == "synth3d":
        unit = "GDOF*steps/s"; val = om.bm1.size * 1e-9 / sec_step
        sample += " (2-D cylinder slice; 3-D oracle not run)"
    return {"metric": "exptA matvec/s" if workload == "cylinder" else "GDOF*steps/s", "value": val, "unit": unit, "n_gpus": 0,
            "steps": steps, "warmup": warmup, "ms_per_step": 1e3 / val if workload == "cylinder" else sec_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "impl": "reference", "config": {"workload": "cylinder_re50_exptA_matvec" if workload == "cylinder" else "synth3d_extruded_cylinder_time_step",
                                            "cpu_sample_mesh": "2-D cylinder slice (1996 el, lx1=6): the 3-D mesh at numpy speed exceeds the bounded-sample budget"},
            "cpu_baseline": {"value": val, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU restatement oracle (numpy); the Fortran/MPI reference cannot be built in this image"}


# ----------------------------------------------------------------------------------------------- native arm
def _cylinder_block(api, case, steps, warmup, device, do_eigs=False):
    """Cylinder Re=50 exptA matvec on ONE GPU (the reference's own config): device-timed and end-to-end matvec/s."""
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
    # BASELINE metric 3, opt-in (--eigs, ~25 s): wall time of the Krylov-Schur run to nev = 2 converged eigenvalues, kdim = 128
    out["time_to_leading_eigs_s"] = None
    out["time_to_leading_eigs_note"] = "not run (pass --eigs); round-1 measurement 24.8 s: profiles/r01_cylinder_eigs_run/"
    if do_eigs:
        t0 = time.perf_counter()
        r = api.linear_stability_analysis_fixed_point(A, 128, 2)
        ctx.sync()
        out["time_to_leading_eigs_s"] = time.perf_counter() - t0
        out["time_to_leading_eigs_note"] = "LightKrylov eigs semantics (Krylov-Schur, kdim 128, nev 2, default tolerance), device-resident basis"
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
    if workload == "cylinder":
        if world > 1:
            raise SystemExit("the cylinder workload is a single-GPU config (71 856 points); use --workload synth3d for N > 1")
        sampler.start()
        cyl = _cylinder_block(api, case, steps, warmup, local, a.eigs)
        clocks = sampler.stop()
        ms_step = cyl["ms_per_matvec"]; value = cyl["matvec_per_s"]; unit = "matvec/s"; metric = "exptA matvec/s"
        e2e = {"value": cyl["e2e_matvec_per_s"], "unit": unit, "h2d_bytes_per_step": cyl["h2d_bytes"], "d2h_bytes_per_step": cyl["d2h_bytes"]}
        launches = cyl["launches"]
        cfg = dict(cyl["config"]); cfg["workload"] = "cylinder_re50_exptA_matvec"
        extra = {k: cyl[k] for k in ("gdof_steps_per_s", "time_steps_per_matvec", "cg_iters_per_step", "gmres_iters_per_step", "launches_per_time_step")}
        roof = {"kernel": "k_axhelm<6,2> (K1)", "bound": "hbm", "achieved": cyl["axhelm_GBps_L2_resident"], "peak": peak, "unit": "GB/s",
                "frac": cyl["axhelm_GBps_L2_resident"] / peak, "traffic": None, "peak_source": peak_src, "note": "L2-resident problem: not an HBM measurement"}
        other = {}
        scaling = "strong"
    else:
        # synthetic 3-D extruded cylinder, weak scaling: `layers` z-layers (1996 elements each) per GPU, z-slab partition
        L = a.layers * world
        # global connectivity, local coordinates / base flow: this rank owns layers [rank*layers, (rank+1)*layers)
        coords, Uall, vertex, cbc = extrude(case, 8, L, layer_range=(rank * a.layers, (rank + 1) * a.layers))
        E2 = case["coords"].shape[0]
        gllnid = (np.arange(E2 * L) // (E2 * a.layers)).astype(np.int32) if world > 1 else None
        sel = slice(None)
        mesh = api.Mesh(coords, vertex, cbc, 12, gllnid=gllnid, rank=rank, nranks=world)
        prm = api.default_params(viscosity=1.0 / 50.0, torder=3, vtol=1e-9, ptol=1e-7, pr_proj=20, coarse_iters=a.coarse_iters)
        nccl_id = None
        if world > 1:
            buf = [api.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(buf, src=0)
            nccl_id = buf[0]
        ctx = api.Context(mesh, prm, device=local, nccl_id=nccl_id)
        bf = ctx.vec(); bf.upload([Uall[sel, 0], Uall[sel, 1], Uall[sel, 2]])
        x = ctx.vec(); x.rand(ifnorm=True, seed=12345); y = ctx.vec()
        npts = mesh.nel * 512
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
        # e2e: ONE full exptA matvec (tau = 1: nsteps + 2 restart steps) through the public API -- host vector in (H2D),
        # host vector out (D2H) inside the timed region; its per-step average includes the cold BDF/projection start-up
        import ctypes as C
        hv, hp, _ = x.download()
        api.lib().nlk_exptA_set_tau(A.h, C.c_double(1.0))
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
        npts_global = npts * world
        value = npts_global * 1e-9 / (ms_step * 1e-3); unit = "GDOF*steps/s"; metric = "GDOF*steps/s"
        e2e = {"value": npts_global * 1e-9 / e2e_s, "unit": unit, "h2d_bytes_per_step": int(vec_bytes // e2e_steps), "d2h_bytes_per_step": int(vec_bytes // e2e_steps),
               "note": "one full tau=1 exptA matvec (%d time steps) with host buffers; bytes are per time step" % e2e_steps}
        launches = s["launches"]
        extra = {"time_steps_timed": int(s["steps"]), "spinup_steps": int(spin), "cg_iters_per_step": s["cg_iters"] / steps, "gmres_iters_per_step": s["gmres_iters"] / steps,
                 "launches_per_time_step": launches / steps, "dt": s0["dt"], "points_per_gpu": int(npts)}
        cfg = {"workload": "synth3d_extruded_cylinder_time_step", "elements": int(mesh.nel * world), "lx1": 8, "lxd": 12, "layers_per_gpu": a.layers,
               "partition": "z-slabs of whole 2-D layers (weak scaling)", "timestepper": "bdf3", "residualProj": 20,
               "l2": "inputs larger than L2 (state + geometry + dealiasing metrics ~ %.1f GB per GPU)" % (npts * 8 * 75 / 1e9)}
        # roofline of the dominant kernel at this problem size (CUDA events on the library stream, right after the timed region)
        ms_ax, bytes_ax = ctx.bench_kernel(0, 100)
        roof = {"kernel": "k_axhelm<8,3> (K1, Helmholtz apply inside the Jacobi-PCG)", "bound": "hbm", "achieved": bytes_ax / (ms_ax * 1e-3) / 1e9,
                "peak": peak, "unit": "GB/s", "frac": bytes_ax / (ms_ax * 1e-3) / 1e9 / peak, "traffic": bytes_ax * (1.1698 / 1.1773), "peak_source": peak_src,
                "us_per_launch": ms_ax * 1e3, "algorithmic_bytes_per_launch": bytes_ax,
                "traffic_note": "dram read+write / algorithmic = 0.994 from ncu --set full at 31 936 elements (profiles/r01_ncu_full_summary.md), scaled to this size"}
        other = {}
        for name, which in (("dssum", 1), ("cdabdtp", 2), ("convect", 3), ("precond", 4), ("vec_dot", 5)):
            m_, b_ = ctx.bench_kernel(which, 20)
            other[name] = {"us": m_ * 1e3, "GBps": b_ / (m_ * 1e-3) / 1e9, "frac": b_ / (m_ * 1e-3) / 1e9 / peak}
        ctx.close()
        scaling = "weak"
        if world == 1 and not a.no_cylinder:
            extra["cylinder_re50"] = _cylinder_block(api, case, 3, 3, local, a.eigs)
    out = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_step,
           "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "config": cfg, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
           "kernels": other, **extra}
    if world == 1 and rank == 0 and os.environ.get("NLK_BENCH_NO_CPU") != "1":
        try:
            sec_step, nsteps, om, stats = cpu_sample(max(2, a.cpu_steps))
            if workload == "cylinder":
                v = 1.0 / (sec_step * (nsteps + 2)); u = "matvec/s"
            else:
                v = om.bm1.size * 1e-9 / sec_step; u = "GDOF*steps/s"
            out["cpu_baseline"] = {"value": v, "unit": u, "cores": os.cpu_count() or 1, "kind": "port",
                                   "sample": f"{max(2, a.cpu_steps)} time steps of the 2-D cylinder Re=50 config with the numpy oracle (BLAS threads <= cores), "
                                             "same algorithm (Jacobi-PCG + Schwarz/coarse FGMRES), converted to the unit"}
        except Exception as e:  # the baseline is a reported number, never the product path
            out["cpu_baseline"] = {"value": None, "unit": unit, "cores": os.cpu_count() or 1, "kind": "port", "sample": f"failed: {e}"}
    if dist is not None:
        dist.destroy_process_group()
    return out


if __name__ == "__main__":
    main()
