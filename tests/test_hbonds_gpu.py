"""GPU parity of K12 (hydrogen bonds, hbonds.pyx:25-134 / interactions.py:365-467): identical triples in identical order
against the reference's own test expectations (tests/golden/hbonds.npz), seeded outputs of the compiled reference kernel and
the oracle on larger periodic systems."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class _Mol:
    def __init__(self, coords, box):
        self.coords, self.box = coords, box
        self.numAtoms, self.numFrames = coords.shape[0], coords.shape[2]


def test_reference_test_case(g_hbonds):
    """tests/test_interactions.py:7-55 through the hbonds_calculate mirror: the four protein-ligand bonds in both frames,
    178 bonds for 'all'; ignore_hs and wider thresholds equal the reference's live outputs."""
    from moleculekit_b200.interactions import hbonds_calculate

    g = g_hbonds
    mol = _Mol(g["coords"], g["box"])
    hb = hbonds_calculate(mol, g["donors"], g["acceptors"], g["protein"], g["ben"])
    ref = np.array([[3414, 3421, 2471], [3414, 3422, 2789], [3415, 3423, 2472], [3415, 3424, 2482]])
    assert len(hb) == 2 and hb[0].dtype == np.int64
    assert np.array_equal(hb[0], ref) and np.array_equal(hb[1], ref)
    everything = np.ones(mol.numAtoms, bool)
    hb = hbonds_calculate(mol, g["donors"], g["acceptors"], everything)
    assert len(hb) == 2 and hb[0].shape == (178, 3)
    assert np.array_equal(hb[0], g["hb_all"]) and np.array_equal(hb[1], g["hb_all"])
    assert np.array_equal(hbonds_calculate(mol, g["donors"], g["acceptors"], everything, ignore_hs=True)[0], g["hb_all_nohs"])
    assert np.array_equal(hbonds_calculate(mol, g["donors"], g["acceptors"], everything, dist_threshold=3.2,
                                           angle_threshold=100)[1], g["hb_all_wide"])
    assert np.array_equal(hbonds_calculate(mol, g["donors"], g["acceptors"], g["protein"], g["ben"], ignore_hs=True,
                                           dist_threshold=3.5)[0], g["hb_prot_ben_nohs"])
    # integer index selections and the empty cases (interactions.py:428-429)
    hb = hbonds_calculate(mol, g["donors"], g["acceptors"], np.where(g["protein"])[0], np.where(g["ben"])[0])
    assert np.array_equal(hb[0], ref)
    assert [x.shape for x in hbonds_calculate(mol, g["donors"][:0], g["acceptors"], everything)] == [(0, 3), (0, 3)]
    with pytest.raises(RuntimeError, match="same number of frames"):
        hbonds_calculate(_Mol(g["coords"], g["box"][:, :1]), g["donors"], g["acceptors"], everything)


def test_seeded_reference_cases(g_hbonds):
    """the drop-in `calculate` against stored outputs of the compiled reference: periodic wraps, a zero box component, an
    all-zero box, an overlapping donor pair (dist2_b == 0), a NaN coordinate (emitted with ignore_hs, as the reference)."""
    from moleculekit_b200 import hbonds

    g = g_hbonds
    for c in range(int(g["ncase"])):
        dth, ath = (float(x) for x in g[f"r{c}_thr"])
        for intra in (0, 1):
            for ign in (0, 1):
                dn = g[f"r{c}_donors"] if not ign else np.unique(g[f"r{c}_donors"][:, 0])[:, None].astype(np.uint32)
                res = hbonds.calculate(dn, g[f"r{c}_acceptors"], g[f"r{c}_coords"], g[f"r{c}_box"], g[f"r{c}_sel1"],
                                       g[f"r{c}_sel2"], dist_threshold=dth, angle_threshold=ath, intra=bool(intra),
                                       ignore_hs=bool(ign))
                counts = np.array([len(x) // 3 for x in res])
                tri = np.array([v for x in res for v in x], dtype=np.int32).reshape(-1, 3)
                assert np.array_equal(counts, g[f"r{c}_out_{intra}{ign}_counts"]), (c, intra, ign)
                assert np.array_equal(tri, g[f"r{c}_out_{intra}{ign}"]), (c, intra, ign)


def _water_box(rng, n_mol, F, L):
    """n_mol three-site waters on a jittered lattice in a periodic box of edge L, random orientations, moving frame to frame"""
    side = int(np.ceil(n_mol ** (1 / 3)))
    grid = np.stack(np.meshgrid(*[np.arange(side)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n_mol] * (L / side)
    O = grid[:, :, None] + rng.normal(0, 0.35, size=(n_mol, 3, F)) + rng.integers(-2, 3, size=(n_mol, 3, 1)) * L
    def unit(v):
        return v / np.linalg.norm(v, axis=1, keepdims=True)
    h1 = O + 0.9572 * unit(rng.normal(size=(n_mol, 3, F)))
    h2 = O + 0.9572 * unit(rng.normal(size=(n_mol, 3, F)))
    coords = np.empty((3 * n_mol, 3, F), np.float32)
    coords[0::3], coords[1::3], coords[2::3] = O, h1, h2
    o_idx = np.arange(0, 3 * n_mol, 3)
    donors = np.concatenate([np.stack([o_idx, o_idx + 1], 1), np.stack([o_idx, o_idx + 2], 1)]).astype(np.uint32)
    box = np.full((3, F), L, np.float32) * (1 + 0.003 * rng.normal(size=(3, F))).astype(np.float32)
    return coords, box, donors, o_idx.astype(np.uint32)


def test_water_box_vs_oracle(oracle):
    """1500 waters, 24 frames, molecules displaced by whole boxes (every pair needs the minimum image): 3000 donors x 1500
    acceptors per frame against the oracle, intra and two-selection modes, with and without hydrogens."""
    from moleculekit_b200 import hbonds

    rng = np.random.default_rng(41)
    coords, box, donors, acc = _water_box(rng, 1500, 24, 35.6)
    N = coords.shape[0]
    s1 = np.ones(N, np.uint32)
    half = (np.arange(N) // 3 % 2 == 0).astype(np.uint32)
    total = 0
    for sa, sb, intra in ((s1, s1, True), (half, 1 - half, False)):
        for ign in (False, True):
            dn = donors if not ign else np.unique(donors[:, 0])[:, None].astype(np.uint32)
            want = oracle.hbonds_calculate(dn, acc, coords, box, sa, sb, 2.5, 120, intra, ign)
            got = hbonds.calculate(dn, acc, coords, box, sa, sb, dist_threshold=2.5, angle_threshold=120, intra=intra,
                                   ignore_hs=ign)
            assert got == want, (intra, ign)
            total += sum(len(x) for x in got) // 3
    assert total > 5000  # the case is not vacuous


def test_dense_boundary_case_vs_oracle(oracle):
    """5e6 pair tests with minimum-image integers up to 8 and a third of the pairs bonded (the case that separates the
    reference's float overloads of round / sqrt / acos from a double restatement, tests/test_oracle_golden.py::
    test_hbonds_float_overloads): identical triples."""
    from moleculekit_b200 import hbonds
    from test_oracle_golden import _hb_dense_case

    don, acc, xyz, L, ones = _hb_dense_case(8, seed=19)
    want = oracle.hbonds_calculate(don, acc, xyz, L, ones, ones, 5.5, 95.0, True, False)
    got = hbonds.calculate(don, acc, xyz, L, ones, ones, dist_threshold=5.5, angle_threshold=95.0, intra=True)
    assert sum(map(len, want)) // 3 > 1_000_000 and got == want


def test_angle_threshold_sweep_vs_oracle(oracle):
    """the acos comparison is a precomputed cosine bound: sweep thresholds incl. 0, 180, > 180 and negative ones"""
    from moleculekit_b200 import hbonds

    rng = np.random.default_rng(43)
    coords, box, donors, acc = _water_box(rng, 300, 3, 20.8)
    s1 = np.ones(coords.shape[0], np.uint32)
    for ath in (-5.0, 0.0, 1e-3, 37.5, 90.0, 119.99999, 120.0, 150.0, 179.9, 180.0, 180.0001, 400.0, float("nan")):
        want = oracle.hbonds_calculate(donors, acc, coords, box, s1, s1, 3.0, ath, True, False)
        got = hbonds.calculate(donors, acc, coords, box, s1, s1, dist_threshold=3.0, angle_threshold=ath, intra=True)
        assert got == want, ath


def _paths(g, name):
    out, pos = [], 0
    for n in g[f"{name}_len"]:
        out.append([int(i) for i in g[f"{name}_idx"][pos:pos + n]])
        pos += int(n)
    return out


def test_waterbridge_reference_test_case(g_waterbridge):
    """tests/test_interactions.py:255-326 through the waterbridge_calculate mirror (K12 shells + the reference's graph walk):
    the paths written in the reference test for order 1 / 2 and for the whole protein, plus a with-hydrogens variant."""
    from moleculekit_b200.interactions import waterbridge_calculate

    g = g_waterbridge
    mol = _Mol(g["coords"], g["box"])
    kw = dict(dist_threshold=3.8, ignore_hs=True, water=g["water"])
    wb = waterbridge_calculate(mol, g["donors"], g["acceptors"], g["gol"], g["asn155"], order=1, **kw)
    assert [list(map(int, p)) for p in wb[0]] == [[3140, 2899, 2024]] == _paths(g, "wb1")
    wb = waterbridge_calculate(mol, g["donors"], g["acceptors"], g["gol"], g["asn155"], order=2, **kw)
    assert [list(map(int, p)) for p in wb[0]] == [[3140, 2899, 2944, 2023], [3140, 2899, 2024]] == _paths(g, "wb2")
    wb = waterbridge_calculate(mol, g["donors"], g["acceptors"], g["gol"], g["protein"], order=1, **kw)
    assert [list(map(int, p)) for p in wb[0]] == _paths(g, "wb3") and len(wb[0]) == 5
    wb = waterbridge_calculate(mol, g["donors"], g["acceptors"], g["gol"], g["protein"], order=1, dist_threshold=2.6,
                               water=g["water"])
    assert [list(map(int, p)) for p in wb[0]] == _paths(g, "wb4")


def test_edge_cases(oracle):
    """no donors / no acceptors / zero frames / a donor that is its own acceptor / selections that exclude everything"""
    from moleculekit_b200 import hbonds

    rng = np.random.default_rng(5)
    xyz = (rng.uniform(0, 6, size=(30, 3, 2))).astype(np.float32)
    box = np.full((3, 2), 6.0, np.float32)
    don = np.stack([np.arange(0, 20, 2), np.arange(1, 20, 2)], 1).astype(np.uint32)
    acc = np.arange(0, 30, 3, dtype=np.uint32)          # includes donor heavy atoms (a_idx == d_idx_d is skipped)
    ones, zeros = np.ones(30, np.uint32), np.zeros(30, np.uint32)
    assert hbonds.calculate(don[:0], acc, xyz, box, ones, ones) == [[], []]
    assert hbonds.calculate(don, acc[:0], xyz, box, ones, ones) == [[], []]
    assert hbonds.calculate(don, acc, xyz[:, :, :0].copy(), box[:, :0].copy(), ones, ones) == []
    assert hbonds.calculate(don, acc, xyz, box, zeros, ones, intra=True) == [[], []]
    for intra, s1, s2 in ((True, ones, ones), (False, ones, zeros), (False, (np.arange(30) % 2).astype(np.uint32),
                                                                   ((np.arange(30) + 1) % 2).astype(np.uint32))):
        want = oracle.hbonds_calculate(don, acc, xyz, box, s1, s2, 4.0, 30.0, intra, False)
        assert hbonds.calculate(don, acc, xyz, box, s1, s2, dist_threshold=4.0, angle_threshold=30.0, intra=intra) == want
