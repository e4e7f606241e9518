"""GPU parity of K13 (pi-pi, cation-pi, sigma-hole detectors; pipi.pyx / cationpi.pyx / sigmahole.pyx): identical pairs in
identical order and bit-identical distances against the reference's outputs (tests/golden/rings.npz) and the oracle; the
reported angles agree to 1e-5 degrees (they come from the device's acos, the reference's from glibc's acosf)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ANGLE_TOL = 2e-5  # degrees: a few float ulp of an angle <= 90


def _arrays(res):
    pairs, da = res
    return (np.array([len(x) // 2 for x in pairs]), np.array([v for x in pairs for v in x], dtype=np.int32).reshape(-1, 2),
            np.array([v for x in da for v in x], dtype=np.float32).reshape(-1, 2))


def _check(got, pairs, da, counts=None):
    gc, gp, gd = _arrays(got)
    assert np.array_equal(gp, pairs)
    if counts is not None:
        assert np.array_equal(gc, counts)
    assert np.array_equal(gd[:, 0].view(np.uint32), np.ascontiguousarray(da[:, 0]).view(np.uint32))   # sqrtf: bit-identical
    assert np.allclose(gd[:, 1], da[:, 1], rtol=0, atol=ANGLE_TOL)


def test_reference_outputs(g_rings):
    """every stored case: the protein of the reference's interaction tests and the seeded periodic systems, three detectors"""
    from moleculekit_b200 import ringpairs
    from test_oracle_golden import _ring_cases

    fns = {0: ringpairs.pipi_calculate, 1: ringpairs.cationpi_calculate, 2: ringpairs.sigmahole_calculate}
    total = 0
    for name, da_name, cnt_name, mode, args in _ring_cases(g_rings):
        _check(fns[mode](*args), g_rings[name], g_rings[da_name], g_rings[cnt_name] if cnt_name else None)
        total += len(g_rings[name])
    assert total > 300


class _Mol:
    def __init__(self, coords, box):
        self.coords, self.box = coords, box
        self.numAtoms, self.numFrames = coords.shape[0], coords.shape[2]


def test_calculate_mirrors(g_rings):
    """pipi_calculate / cationpi_calculate / sigmahole_calculate of interactions.py (lists of rings in, per-frame lists out;
    return_rings; argument checks) on the protein case"""
    from moleculekit_b200.interactions import cationpi_calculate, pipi_calculate, sigmahole_calculate

    g = g_rings
    mol = _Mol(g["p_coords"], g["p_box"])
    st = g["p_ring_starts"]
    rings = [g["p_ring_atoms"][st[i]:st[i + 1]] for i in range(len(st) - 1)]
    pp, da = pipi_calculate(mol, rings, rings, dist_threshold1=6.0, angle_threshold1_max=40, dist_threshold2=7.0,
                            angle_threshold2_min=50)
    for f in range(2):
        assert np.array_equal(np.array(pp[f], dtype=np.int32).reshape(-1, 2), g[f"p_pipi_{f}"])
        assert np.allclose(np.array(da[f], dtype=np.float32).reshape(-1, 2), g[f"p_pipi_da_{f}"], rtol=0, atol=ANGLE_TOL)
    cp, cda = cationpi_calculate(mol, rings, g["p_cations"], dist_threshold=7.0, angle_threshold_min=30)
    for f in range(2):
        assert np.array_equal(np.array(cp[f], dtype=np.int32).reshape(-1, 2), g[f"p_cat_{f}"])
    rr, _ = pipi_calculate(mol, rings, rings, dist_threshold1=6.0, angle_threshold1_max=40, dist_threshold2=7.0,
                           angle_threshold2_min=50, return_rings=True)
    a, b = g["p_pipi_0"][0]
    assert np.array_equal(rr[0][0][0], rings[a]) and np.array_equal(rr[0][0][1], rings[b])
    with pytest.raises(RuntimeError, match="Values for angles"):
        pipi_calculate(mol, rings, rings, angle_threshold1_max=100)
    with pytest.raises(RuntimeError, match="Values for angles"):
        sigmahole_calculate(mol, rings, [[1, 2]], angle_threshold_min=-1)
    assert cationpi_calculate(mol, rings, []) == ([[], []], [[], []])


def test_many_frames_vs_oracle(oracle):
    """40 rings x 300 cations / halogens / rings over 60 frames in a periodic box against the oracle, incl. thresholds at the
    ends of the angle range"""
    from moleculekit_b200 import ringpairs

    rng = np.random.default_rng(21)
    nr, N, F = 40, 240 + 300, 60
    L = (18.0 * (1 + 0.01 * rng.normal(size=(3, F)))).astype(np.float32)
    xyz = (rng.uniform(0, 18, size=(N, 3, F)) + rng.integers(-2, 3, size=(N, 3, 1)) * 18.0).astype(np.float32)
    for r in range(nr):
        ctr = rng.uniform(0, 18, size=(3, 1)) + np.cumsum(rng.normal(0, .2, size=(3, F)), axis=1)
        u = rng.normal(size=3); u /= np.linalg.norm(u); v = np.cross(u, rng.normal(size=3)); v /= np.linalg.norm(v)
        for j in range(6):
            xyz[6 * r + j] = (ctr + 1.39 * (np.cos(j * np.pi / 3) * u[:, None] + np.sin(j * np.pi / 3) * v[:, None])
                              + rng.normal(0, .05, size=(3, F))).astype(np.float32)
    ra = np.arange(6 * nr, dtype=np.uint32); st = np.arange(0, 6 * nr + 1, 6, dtype=np.uint32)
    cations = np.arange(240, N, dtype=np.uint32)
    hal = np.stack([cations, np.roll(cations, 1)], 1).astype(np.uint32)
    hits = 0
    for mode, fn, args in ((0, ringpairs.pipi_calculate, (ra, st, st, xyz, L, 5.5, 30.0, 7.5, 60.0)),
                           (0, ringpairs.pipi_calculate, (ra, st, st, xyz, L, 7.5, 90.0, 7.5, 0.0)),
                           (1, ringpairs.cationpi_calculate, (ra, st, cations, xyz, L, 6.0, 45.0)),
                           (1, ringpairs.cationpi_calculate, (ra, st, cations, xyz, L, 5.0, 0.0)),
                           (2, ringpairs.sigmahole_calculate, (ra, st, hal, xyz, L, 6.0, 30.0)),
                           (2, ringpairs.sigmahole_calculate, (ra, st, hal, xyz, L, 4.0, 90.0))):
        wc, wp, wd = _arrays(oracle.ring_interactions(mode, *args))
        _check(fn(*args), wp, wd, wc)
        hits += len(wp)
    assert hits > 20000


def test_edge_cases(oracle):
    """empty partner lists, zero frames, rings of 3 atoms, a ring set against a single ring, NaN coordinates: same lists as
    the oracle, no crash"""
    from moleculekit_b200 import ringpairs

    rng = np.random.default_rng(3)
    xyz = (rng.uniform(0, 9, size=(40, 3, 3))).astype(np.float32)
    box = np.full((3, 3), 9.0, np.float32)
    ra = np.arange(12, dtype=np.uint32); st = np.array([0, 3, 6, 12], np.uint32)       # two 3-rings and a 6-ring
    cat = np.arange(20, 30, dtype=np.uint32)
    assert ringpairs.cationpi_calculate(ra, st, cat[:0], xyz, box, 5.0, 10.0) == ([[], [], []], [[], [], []])
    assert ringpairs.pipi_calculate(ra, st, st, xyz[:, :, :0].copy(), box[:, :0].copy(), 9.0, 90.0, 9.0, 0.0) == ([], [])
    for args, mode, fn in (((ra, st, st, xyz, box, 9.0, 90.0, 9.0, 0.0), 0, ringpairs.pipi_calculate),
                           ((ra, st, cat, xyz, box, 9.0, 0.0), 1, ringpairs.cationpi_calculate),
                           ((ra, st, np.stack([cat, cat[::-1]], 1).astype(np.uint32), xyz, box, 9.0, 0.0), 2,
                            ringpairs.sigmahole_calculate)):
        wc, wp, wd = _arrays(oracle.ring_interactions(mode, *args))
        _check(fn(*args), wp, wd, wc)
        assert len(wp) > 0
    bad = xyz.copy(); bad[4, 1, 1] = np.nan
    wc, wp, wd = _arrays(oracle.ring_interactions(1, ra, st, cat, bad, box, 9.0, 0.0))
    _check(ringpairs.cationpi_calculate(ra, st, cat, bad, box, 9.0, 0.0), wp, wd, wc)
