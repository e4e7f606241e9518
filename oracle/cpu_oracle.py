"""ctypes front-end of oracle/mkb_oracle.c -- the CPU restatement of the reference hot path.

TEST INFRASTRUCTURE ONLY (see the header of mkb_oracle.c).  Function names and argument
order mirror the reference's Cython modules so parity tests read like the reference's:

  moleculekit/occupancy_utils/occupancy_utils.pyx:34   calculate_occupancy
  moleculekit/distance_utils/distance_utils.pyx:59     contacts_trajectory
  moleculekit/distance_utils/distance_utils.pyx:98     get_collisions
  moleculekit/distance_utils/distance_utils.pyx:126    dist_trajectory
  moleculekit/distance_utils/distance_utils.pyx:211    dist_trajectory_reduction
  moleculekit/distance_utils/distance_utils.pyx:286    dist_trajectory_reduction_pairs
  moleculekit/distance_utils/distance_utils.pyx:355+   cdist / pdist / squareform

Parity status: PINNED (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "mkb_oracle.c")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libmkb_oracle.so")

_lib = None


def build(force: bool = False) -> str:
    """gcc -O3 -ffp-contract=off (no FMA contraction: the reference binary has none)."""
    if (not force) and os.path.isfile(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    cmd = ["gcc", "-O3", "-ffp-contract=off", "-fno-fast-math", "-fvisibility=hidden",
           "-shared", "-fPIC", SRC, "-o", LIB, "-lm"]
    subprocess.run(cmd, check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_contacts_trajectory.restype = C.c_int64
        _lib.oracle_get_collisions.restype = C.c_int64
        _lib.oracle_squareform_dim.restype = C.c_int64
        _lib.oracle_bond_grid_search.restype = C.c_int64
        _lib.oracle_wrap_compact.restype = C.c_int64
        _lib.oracle_hbonds.restype = C.c_int64
        _lib.oracle_ring_interactions.restype = C.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def calculate_occupancy(centers, coords, sigmas, results):
    centers = np.ascontiguousarray(centers, dtype=np.float64)
    coords = _f32(coords)
    sigmas = np.ascontiguousarray(sigmas, dtype=np.float64)
    assert results.dtype == np.float64 and results.flags["C_CONTIGUOUS"]
    M, N, Cn = centers.shape[0], coords.shape[0], sigmas.shape[1]
    assert results.shape == (M, Cn) and sigmas.shape[0] == N
    lib().oracle_calculate_occupancy(_p(centers), _p(coords), _p(sigmas), _p(results),
                                     C.c_int64(M), C.c_int64(N), C.c_int64(Cn))


def dist_trajectory(coords, box, sel1, sel2, digitized_chains, selfdist, pbc, results):
    coords, box = _f32(coords), _f32(box)
    sel1, sel2, ch = _u32(sel1), _u32(sel2), _u32(digitized_chains)
    assert results.dtype == np.float32 and results.flags["C_CONTIGUOUS"]
    F = coords.shape[2]
    lib().oracle_dist_trajectory(_p(coords), _p(box), _p(sel1), C.c_int64(len(sel1)),
                                 _p(sel2), C.c_int64(len(sel2)), _p(ch),
                                 C.c_int(bool(selfdist)), C.c_int(bool(pbc)), _p(results),
                                 C.c_int64(F), C.c_int64(results.shape[1]))


def contacts_trajectory(coords, box, sel1, sel2, digitized_chains, selfdist, pbc, dist_threshold=5):
    """Returns a list (per frame) of flat lists [a0, b0, a1, b1, ...] like the reference."""
    coords, box = _f32(coords), _f32(box)
    sel1, sel2, ch = _u32(sel1), _u32(sel2), _u32(digitized_chains)
    F = coords.shape[2]
    counts = np.zeros(F, dtype=np.int64)
    args = (_p(coords), _p(box), _p(sel1), C.c_int64(len(sel1)), _p(sel2), C.c_int64(len(sel2)),
            _p(ch), C.c_int(bool(selfdist)), C.c_int(bool(pbc)), C.c_float(dist_threshold),
            C.c_int64(F), _p(counts))
    total = lib().oracle_contacts_trajectory(*args, None)
    pairs = np.zeros((max(total, 1), 2), dtype=np.uint32)
    lib().oracle_contacts_trajectory(*args, _p(pairs))
    out, s = [], 0
    for f in range(F):
        out.append(pairs[s:s + counts[f]].reshape(-1).tolist())
        s += counts[f]
    return out


def get_collisions(coords1, coords2, dist_threshold):
    c1, c2 = _f32(coords1), _f32(coords2)
    n = lib().oracle_get_collisions(_p(c1), C.c_int64(len(c1)), _p(c2), C.c_int64(len(c2)),
                                    C.c_float(dist_threshold), None)
    pairs = np.zeros((max(n, 1), 2), dtype=np.uint32)
    lib().oracle_get_collisions(_p(c1), C.c_int64(len(c1)), _p(c2), C.c_int64(len(c2)),
                                C.c_float(dist_threshold), _p(pairs))
    return pairs[:n].reshape(-1).tolist()


def _csr(groups):
    off = np.zeros(len(groups) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(g) for g in groups])
    atoms = np.ascontiguousarray(np.concatenate([np.asarray(g, dtype=np.int32) for g in groups])
                                 if len(groups) else np.zeros(0, np.int32), dtype=np.int32)
    return off, atoms


def dist_trajectory_reduction(coords, box, groups1, groups2, digitized_chains1, digitized_chains2,
                              selfdist, pbc, masses, reduction1, reduction2, results):
    coords, box, masses = _f32(coords), _f32(box), _f32(masses)
    o1, a1 = _csr(groups1)
    o2, a2 = _csr(groups2)
    c1, c2 = _u32(digitized_chains1), _u32(digitized_chains2)
    F = coords.shape[2]
    lib().oracle_dist_trajectory_reduction(_p(coords), _p(box), _p(o1), _p(a1), C.c_int64(len(groups1)),
                                           _p(o2), _p(a2), C.c_int64(len(groups2)), _p(c1), _p(c2),
                                           C.c_int(bool(selfdist)), C.c_int(bool(pbc)), _p(masses),
                                           C.c_int(reduction1), C.c_int(reduction2), _p(results),
                                           C.c_int64(F), C.c_int64(results.shape[1]))
    return results


def dist_trajectory_reduction_pairs(coords, box, groups1, groups2, digitized_chains1, digitized_chains2,
                                    pbc, masses, reduction1, reduction2, results):
    coords, box, masses = _f32(coords), _f32(box), _f32(masses)
    o1, a1 = _csr(groups1)
    o2, a2 = _csr(groups2)
    c1, c2 = _u32(digitized_chains1), _u32(digitized_chains2)
    F = coords.shape[2]
    lib().oracle_dist_trajectory_reduction_pairs(_p(coords), _p(box), _p(o1), _p(a1), _p(o2), _p(a2),
                                                 C.c_int64(len(groups1)), _p(c1), _p(c2),
                                                 C.c_int(bool(pbc)), _p(masses),
                                                 C.c_int(reduction1), C.c_int(reduction2), _p(results),
                                                 C.c_int64(F), C.c_int64(results.shape[1]))
    return results


def cdist(coords1, coords2, results):
    c1, c2 = _f32(coords1), _f32(coords2)
    lib().oracle_cdist(_p(c1), C.c_int64(c1.shape[0]), _p(c2), C.c_int64(c2.shape[0]),
                       C.c_int64(c1.shape[1]), _p(results))


def pdist(coords, results):
    c = _f32(coords)
    lib().oracle_pdist(_p(c), C.c_int64(c.shape[0]), C.c_int64(c.shape[1]), _p(results))


def squareform(distances):
    d = _f32(distances)
    m = lib().oracle_squareform_dim(C.c_int64(len(d)))
    out = np.zeros((m, m), dtype=np.float32)
    lib().oracle_squareform(_p(d), C.c_int64(len(d)), _p(out))
    return out


def bond_grid_search(coords, grid_cutoff, is_hydrogen, radii, max_boxes: float = 4e6, cutoff_incr: float = 1.26):
    """moleculekit/bondguesser.py:259-392 (same arguments); returns (nbonds, 2) uint32 in the reference's order."""
    coords = _f32(coords)
    radii = _f32(radii)
    ish = _u32(is_hydrogen)
    n = coords.shape[0]
    if n == 0:
        return np.zeros((0, 2), np.uint32)
    rng = coords.max(axis=0) - coords.min(axis=0)

    def counts(pd):
        ax = (np.floor(rng / pd).astype(np.int64) + 1).tolist()
        return ax[0] * ax[1] * ax[2]

    pairdist = float(grid_cutoff)
    while counts(pairdist) > max_boxes:
        pairdist *= cutoff_incr
    args = (_p(coords), _p(radii), _p(ish), C.c_int64(n), C.c_double(pairdist))
    total = lib().oracle_bond_grid_search(*args, None)
    pairs = np.zeros((max(total, 1), 2), dtype=np.uint32)
    lib().oracle_bond_grid_search(*args, _p(pairs))
    return pairs[:total]


def metric_shell(coords, box, sel1, sel2, digitized_chains, selfdist, pbc, numshells=4, shellwidth=3, truncate=None):
    """moleculekit/projections/metricshell.py:183-202 (_shells) on top of dist_trajectory + the truncate post-op of
    projections/util.py:74-75: (F, ncenters * numshells) float64 densities, centre-major."""
    sel1, sel2 = _u32(sel1), _u32(sel2)
    F = coords.shape[2]
    n1, n2 = len(sel1), len(sel2)
    P = (n1 * (n2 - 1)) // 2 if selfdist else n1 * n2
    dist = np.zeros((F, P), dtype=np.float32)
    dist_trajectory(coords, box, sel1, sel2, digitized_chains, selfdist, pbc, dist)
    if truncate is not None:
        dist[dist > truncate] = truncate
    if selfdist:
        ii, jj = np.triu_indices(n1, k=1)
    else:
        ii, jj = np.repeat(np.arange(n1), n2), np.tile(np.arange(n2), n1)
    edges = np.arange(shellwidth * (numshells + 1), step=shellwidth)
    vol = 4 / 3 * np.pi * (edges[1:] ** 3 - edges[:-1] ** 3)
    out = np.ones((F, n1 * numshells)) * -1
    for c in range(n1):
        cols = ((ii == c) | (jj == c)) if selfdist else (ii == c)
        for e in range(numshells):
            inshell = (dist[:, cols] > edges[e]) & (dist[:, cols] <= edges[e + 1])
            out[:, c * numshells + e] = np.sum(inshell, axis=1) / vol[e]
    return out


def wrap_box(groups, coords, box, centersel, center):
    """moleculekit/wrapping/wrapping.pyx:91-144 (same arguments); coords (N, 3, F) float32 C-contiguous, in place."""
    assert coords.dtype == np.float32 and coords.flags["C_CONTIGUOUS"] and coords.ndim == 3 and coords.shape[1] == 3
    groups, centersel = _u32(groups), _u32(centersel)
    box, center = _f32(box), _f32(center)
    F = coords.shape[2]
    assert box.shape == (3, F)
    lib().oracle_wrap_box(_p(groups), C.c_int64(len(groups)), _p(coords), _p(box), C.c_int64(F),
                          _p(centersel), C.c_int64(len(centersel)), _p(center))


WRAP_MAX_ITER = 1 << 20  # bound of the reference's unbounded `while` loops (same bound in the CUDA kernels)


def _tric_args(groups, coords, boxvectors, centersel, center):
    assert coords.dtype == np.float32 and coords.flags["C_CONTIGUOUS"] and coords.ndim == 3 and coords.shape[1] == 3
    groups, centersel = _u32(groups), _u32(centersel)
    bv = np.ascontiguousarray(boxvectors, dtype=np.float64)
    center = _f32(center)
    F = coords.shape[2]
    assert bv.shape == (3, 3, F)
    return groups, bv, centersel, center, F


def wrap_triclinic_unitcell(groups, coords, boxvectors, centersel, center):
    """moleculekit/wrapping/wrapping.pyx:147-250 (same arguments); coords (N, 3, F) float32 C-contiguous, in place."""
    groups, bv, centersel, center, F = _tric_args(groups, coords, boxvectors, centersel, center)
    lib().oracle_wrap_triclinic(_p(groups), C.c_int64(len(groups)), _p(coords), C.c_int64(coords.shape[0]), _p(bv),
                                C.c_int64(F), _p(centersel), C.c_int64(len(centersel)), _p(center),
                                C.c_int64(WRAP_MAX_ITER))


def wrap_compact_unitcell(groups, coords, boxvectors, centersel, center, mode):
    """moleculekit/wrapping/wrapping.pyx:255-344 (same arguments); mode 0 = rectangular, 1 = compact; in place."""
    groups, bv, centersel, center, F = _tric_args(groups, coords, boxvectors, centersel, center)
    rc = lib().oracle_wrap_compact(_p(groups), C.c_int64(len(groups)), _p(coords), C.c_int64(coords.shape[0]), _p(bv),
                                   C.c_int64(F), _p(centersel), C.c_int64(len(centersel)), _p(center), C.c_int(mode),
                                   C.c_int64(WRAP_MAX_ITER))
    if rc < 0:
        raise ValueError("Too many triclinic vectors!!")


def hbonds_calculate(donors, acceptors, coords, box, sel1, sel2, dist_threshold=2.5, angle_threshold=120, intra=False,
                     ignore_hs=False):
    """moleculekit/interactions/hbonds/hbonds.pyx:25-134 `calculate` (same arguments): list over frames of flat int lists
    (heavy, hydrogen | -1, acceptor, ...)."""
    donors = np.ascontiguousarray(donors, dtype=np.uint32).reshape(len(donors), -1)
    if donors.shape[1] == 1:  # ignore_hs callers pass (n, 1)
        donors = np.ascontiguousarray(np.hstack([donors, donors]))
    acceptors, sel1, sel2 = _u32(acceptors), _u32(sel1), _u32(sel2)
    coords, box = _f32(coords), _f32(box)
    F = coords.shape[2]
    counts = np.zeros(F, dtype=np.int64)
    args = lambda out, cap: (_p(donors), C.c_int64(len(donors)), _p(acceptors), C.c_int64(len(acceptors)), _p(coords),
                             _p(box), C.c_int64(F), _p(sel1), _p(sel2), C.c_float(dist_threshold),
                             C.c_float(angle_threshold), C.c_int(int(intra)), C.c_int(int(ignore_hs)), _p(counts),
                             _p(out), C.c_int64(cap))
    dummy = np.zeros(3, dtype=np.int32)
    total = lib().oracle_hbonds(*args(dummy, 0))
    out = np.zeros(max(3 * total, 3), dtype=np.int32)
    lib().oracle_hbonds(*args(out, total))
    res, pos = [], 0
    for f in range(F):
        res.append(out[3 * pos:3 * (pos + counts[f])].tolist())
        pos += int(counts[f])
    return res


def ring_interactions(mode, rings_atoms, starts1, second, coords, box, p0, p1=0.0, p2=0.0, p3=0.0):
    """pipi.calculate (mode 0: pipi.pyx:86-185), cationpi.calculate (1: cationpi.pyx:91-173), sigmahole.calculate
    (2: sigmahole.pyx:91-174) with their positional arguments; returns (results, distangles) as the reference: per frame a flat
    int list and a flat float list."""
    rings_atoms, starts1 = _u32(rings_atoms), _u32(starts1)
    second = np.ascontiguousarray(second, dtype=np.uint32)
    coords, box = _f32(coords), _f32(box)
    F = coords.shape[2]
    n1 = len(starts1) - 1
    n2 = len(second) - 1 if mode == 0 else len(second)
    counts = np.zeros(F, dtype=np.int64)
    call = lambda pairs, da, cap: lib().oracle_ring_interactions(
        C.c_int(mode), _p(rings_atoms), _p(starts1), C.c_int64(n1), _p(second), C.c_int64(n2), _p(coords), _p(box),
        C.c_int64(F), C.c_float(p0), C.c_float(p1), C.c_float(p2), C.c_float(p3), _p(counts), _p(pairs), _p(da), C.c_int64(cap))
    total = call(np.zeros(2, np.int32), np.zeros(2, np.float32), 0)
    pairs, da = np.zeros(max(2 * total, 2), np.int32), np.zeros(max(2 * total, 2), np.float32)
    call(pairs, da, total)
    res, dist, pos = [], [], 0
    for f in range(F):
        res.append(pairs[2 * pos:2 * (pos + counts[f])].tolist())
        dist.append(da[2 * pos:2 * (pos + counts[f])].tolist())
        pos += int(counts[f])
    return res, dist


def within_distance(coords, cutoff, sel1, sel2, sel2_min_coords, sel2_max_coords, results):
    """moleculekit/atomselect_utils/atomselect_utils.pyx:612-620 (same arguments); ``results`` (bool, len(sel1)) is
    updated in place.  The min/max arguments only feed the reference's no-op pre-check."""
    coords = _f32(coords)
    sel1, sel2 = _u32(sel1), _u32(sel2)
    assert results.dtype == np.bool_ and results.flags["C_CONTIGUOUS"] and results.shape == (len(sel1),)
    lib().oracle_within_distance(_p(coords), C.c_float(cutoff), _p(sel1), C.c_int64(len(sel1)), _p(sel2),
                                 C.c_int64(len(sel2)), _p(results.view(np.uint8)))


def read_xtc(raw: bytes):
    """moleculekit/fileformats/xtc/xtc.pyx:34-55 (read_xtc) on a byte string: (coords (natoms,3,F) f32 [nm], box (3,3,F), time,
    step).  Own header walk (xdrfile_xtc.cpp:29-67) + oracle_xtc_decode_block per frame."""
    import struct

    lib().oracle_xtc_decode_block.restype = C.c_int
    pos, frames = 0, []
    while pos + 16 <= len(raw):
        magic, natoms, step = struct.unpack_from(">iii", raw, pos)
        assert magic == 1995
        (time,) = struct.unpack_from(">f", raw, pos + 12)
        box = np.array(struct.unpack_from(">9f", raw, pos + 16), dtype=np.float32).reshape(3, 3)
        (lsize,) = struct.unpack_from(">i", raw, pos + 52)
        pos += 56
        xyz = np.zeros((lsize, 3), dtype=np.float32)
        if lsize <= 9:
            xyz[:] = np.array(struct.unpack_from(f">{3 * lsize}f", raw, pos), dtype=np.float32).reshape(lsize, 3)
            pos += 12 * lsize
        else:
            (prec,) = struct.unpack_from(">f", raw, pos)
            ints = struct.unpack_from(">8i", raw, pos + 4)
            nbytes = ints[7]
            data = np.frombuffer(raw, dtype=np.uint8, count=nbytes, offset=pos + 36)
            mn, mx = np.array(ints[0:3], np.int32), np.array(ints[3:6], np.int32)
            rc = lib().oracle_xtc_decode_block(_p(data), C.c_int64(nbytes), C.c_int64(lsize), C.c_float(prec), _p(mn), _p(mx),
                                               C.c_int32(ints[6]), _p(xyz))
            assert rc == 0, rc
            pos += 36 + ((nbytes + 3) // 4) * 4
        frames.append((xyz, box, time, step))
    F = len(frames)
    n = frames[0][0].shape[0] if F else 0
    coords = np.zeros((n, 3, F), np.float32); bx = np.zeros((3, 3, F), np.float32)
    for f, (xyz, box, _, _) in enumerate(frames):
        coords[:, :, f] = xyz; bx[:, :, f] = box
    return coords, bx, np.array([t for _, _, t, _ in frames], np.float32), np.array([s for _, _, _, s in frames], np.int32)
