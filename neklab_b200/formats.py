"""Readers/writers for Nek5000 on-disk formats used by the exptA hot path.

Formats (layouts verified against the reference's shipped fixtures under
``examples/``):

* ``.re2``  binary mesh  (``examples/cylinder/stability/direct/1cyl.re2``)
* ``.ma2``  genmap vertex map + partition ids (``.../1cyl.ma2``)
* ``*.f%05d`` / ``.fld`` field files (``.../BF_1cyl0.f00001``), the files the
  reference loads with ``load_fld`` (``examples/cylinder/stability/direct/1cyl.usr:15``)
  and writes with ``outpost`` (``src/neklab_utils.f90:305-361``).

Pure numpy; no device code here.  This module is host-side I/O only.
"""
from __future__ import annotations

import dataclasses
import numpy as np


@dataclasses.dataclass
class Re2Mesh:
    ndim: int
    nel: int
    xyz: np.ndarray          # (nel, ndim, 2**ndim) corner coords, preprocessor (ccw) order
    group: np.ndarray        # (nel,)
    curves: list             # [(iel0, iside0, params(5), ctype)]
    cbc: np.ndarray          # (nfld, nel, 2*ndim) dtype 'U3' boundary codes (velocity field = 0)
    bc: np.ndarray           # (nfld, nel, 2*ndim, 5) BC parameters


def read_re2(path: str) -> Re2Mesh:
    raw = open(path, "rb").read()
    hdr = raw[:80].decode("ascii", "replace")
    ver = hdr[:5]
    if ver not in ("#v001", "#v002", "#v003"):
        raise ValueError(f"{path}: not a re2 file (header {hdr[:5]!r})")
    tok = hdr[5:].split()
    nel, ndim, nelv = int(tok[0]), int(tok[1]), int(tok[2])
    tag = np.frombuffer(raw, dtype="<f4", count=1, offset=80)[0]
    en = "<" if abs(tag - 6.54321) < 1e-5 else ">"
    if en == ">" and abs(np.frombuffer(raw, dtype=">f4", count=1, offset=80)[0] - 6.54321) > 1e-5:
        raise ValueError(f"{path}: bad endian tag")
    if ver == "#v001":
        raise ValueError("re2 v001 (float32 records) not supported")
    off = 84
    nv = 2 ** ndim
    rec = 1 + ndim * nv
    a = np.frombuffer(raw, dtype=en + "f8", count=nel * rec, offset=off).reshape(nel, rec)
    off += nel * rec * 8
    group = a[:, 0].astype(np.int64)
    xyz = a[:, 1:].reshape(nel, ndim, nv).copy()
    ncurve = int(np.frombuffer(raw, dtype=en + "f8", count=1, offset=off)[0]); off += 8
    curves = []
    for _ in range(ncurve):
        r = np.frombuffer(raw, dtype=en + "f8", count=7, offset=off)
        ctype = raw[off + 56:off + 64].decode("ascii", "replace").strip()
        curves.append((int(r[0]) - 1, int(r[1]) - 1, r[2:7].copy(), ctype))
        off += 64
    nface = 2 * ndim
    cbcs, bcs = [], []
    while off + 8 <= len(raw):
        nbc = int(np.frombuffer(raw, dtype=en + "f8", count=1, offset=off)[0]); off += 8
        cbc = np.full((nel, nface), "E  ", dtype="U3")
        bc = np.zeros((nel, nface, 5))
        for _ in range(nbc):
            r = np.frombuffer(raw, dtype=en + "f8", count=7, offset=off)
            code = raw[off + 56:off + 59].decode("ascii", "replace")
            e, f = int(r[0]) - 1, int(r[1]) - 1
            cbc[e, f] = code
            bc[e, f] = r[2:7]
            off += 64
        cbcs.append(cbc); bcs.append(bc)
    if not cbcs:
        cbcs = [np.full((nel, nface), "E  ", dtype="U3")]; bcs = [np.zeros((nel, nface, 5))]
    return Re2Mesh(ndim, nel, xyz, group, curves, np.stack(cbcs), np.stack(bcs))


@dataclasses.dataclass
class Ma2Map:
    nel: int
    nactive: int
    depth: int
    d2: int
    npts: int
    nrank: int
    noutflow: int
    pid: np.ndarray          # (nel,) partition leaf id in [0, d2)
    vertex: np.ndarray       # (nel, 2**ndim) 1-based global vertex ids, lexicographic corner order


def read_ma2(path: str) -> Ma2Map:
    raw = open(path, "rb").read()
    hdr = raw[:132].decode("ascii", "replace")
    if hdr[:5] != "#v001":
        raise ValueError(f"{path}: not a ma2 file")
    nel, nactive, depth, d2, npts, nrank, noutflow = (int(t) for t in hdr[5:].split()[:7])
    tag = np.frombuffer(raw, dtype="<f4", count=1, offset=132)[0]
    en = "<" if abs(tag - 6.54321) < 1e-5 else ">"
    nv = npts // nel
    a = np.frombuffer(raw, dtype=en + "i4", count=nel * (nv + 1), offset=136).reshape(nel, nv + 1)
    return Ma2Map(nel, nactive, depth, d2, npts, nrank, noutflow,
                  a[:, 0].astype(np.int64).copy(), a[:, 1:].astype(np.int64).copy())


@dataclasses.dataclass
class FldFile:
    nx: int
    ny: int
    nz: int
    nel: int
    nelg: int
    time: float
    istep: int
    rdcode: str
    wdsize: int
    glel: np.ndarray                     # (nel,) 1-based global element ids in FILE order
    coords: np.ndarray | None            # (nelg, ndim, nz, ny, nx) sorted by global id
    vel: np.ndarray | None               # (nelg, ndim, nz, ny, nx)
    pr: np.ndarray | None                # (nelg, nz, ny, nx) on mesh 1
    temp: np.ndarray | None              # (nelg, nz, ny, nx)

    @property
    def ndim(self) -> int:
        return 3 if self.nz > 1 else 2


def read_fld(path: str) -> FldFile:
    raw = open(path, "rb").read()
    hdr = raw[:132].decode("ascii", "replace")
    if hdr[:4] != "#std":
        raise ValueError(f"{path}: not a nek field file")
    tok = hdr[4:].split()
    wd, nx, ny, nz, nel, nelg = (int(t) for t in tok[:6])
    time = float(tok[6]); istep = int(tok[7])
    rdcode = tok[10]
    tag = np.frombuffer(raw, dtype="<f4", count=1, offset=132)[0]
    en = "<" if abs(tag - 6.54321) < 1e-5 else ">"
    off = 136
    glel = np.frombuffer(raw, dtype=en + "i4", count=nel, offset=off).astype(np.int64); off += 4 * nel
    ndim = 3 if nz > 1 else 2
    npt = nx * ny * nz
    ft = en + ("f8" if wd == 8 else "f4")
    order = np.argsort(glel, kind="stable")
    out = {"X": None, "U": None, "P": None, "T": None}
    for c in rdcode:
        if c in "XU":
            a = np.frombuffer(raw, dtype=ft, count=nel * ndim * npt, offset=off).reshape(nel, ndim, nz, ny, nx)
            off += nel * ndim * npt * wd
            out[c] = a[order].astype(np.float64)
        elif c in "PT":
            a = np.frombuffer(raw, dtype=ft, count=nel * npt, offset=off).reshape(nel, nz, ny, nx)
            off += nel * npt * wd
            out[c] = a[order].astype(np.float64)
        elif c == "S":
            break
    return FldFile(nx, ny, nz, nel, nelg, time, istep, rdcode, wd, glel,
                   out["X"], out["U"], out["P"], out["T"])


def write_fld(path: str, *, coords=None, vel=None, pr=None, temp=None, time=0.0, istep=0,
              wdsize=8) -> None:
    """Write a single-file Nek field file (same layout read_fld parses).

    Arrays are (nel, [ndim,] nz, ny, nx), elements in ascending global-id order.
    """
    ref = next(a for a in (coords, vel, pr, temp) if a is not None)
    nel = ref.shape[0]
    nz, ny, nx = ref.shape[-3:]
    rd = ("X" if coords is not None else "") + ("U" if vel is not None else "") + \
         ("P" if pr is not None else "") + ("T" if temp is not None else "")
    hdr = "#std %1d %2d %2d %2d %10d %10d %20.13E %9d %6d %6d %-10s" % (
        wdsize, nx, ny, nz, nel, nel, time, istep, 0, 1, rd)
    hdr = hdr.ljust(132)[:132]
    ft = "<f8" if wdsize == 8 else "<f4"
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii"))
        f.write(np.array([6.54321], dtype="<f4").tobytes())
        f.write(np.arange(1, nel + 1, dtype="<i4").tobytes())
        for a in (coords, vel, pr, temp):
            if a is not None:
                f.write(np.ascontiguousarray(a, dtype=ft).tobytes())
