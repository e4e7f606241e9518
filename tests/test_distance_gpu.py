"""GPU parity of the distance path (K3-K6) against the oracle, the reference's stored goldens and the frozen outputs of
the reference's compiled kernels.  float32 / index outputs must be BIT-identical.  Mirrors the reference's
tests/test_metricdistance.py, tests/test_distance.py and tests/test_interactions.py::test_metal_coordination."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bits(a):
    """Bit pattern with NaNs canonicalised: x86 SSE produces the 'default NaN' 0xFFC00000 (sign set), the GPU
    0x7FFFFFFF; both are NaN, and no consumer can tell them apart except by reinterpreting the bits."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = a.view(np.uint32).copy()
    b[np.isnan(a)] = 0x7FC00000
    return b


def _groups(off, atoms):
    return [atoms[off[i]:off[i + 1]].tolist() for i in range(len(off) - 1)]


@pytest.fixture(scope="module")
def du():
    import torch

    assert torch.cuda.is_available()
    from moleculekit_b200 import distance_utils

    return distance_utils


@pytest.fixture(scope="module")
def mol(g_traj):
    from moleculekit_b200.molecule_lite import MolLite

    g = g_traj
    return MolLite(g["coords"], g["box"], element=g["element"], name=g["name"], resname=g["resname"], resid=g["resid"],
                   chain=g["chain"], segid=g["segid"],
                   named_selections={s: m for s, m in zip(g["sel_strings"].tolist(), g["sel_masks"])})


# ------------------------------------------------------------------------------------------ raw kernels, bit exact
def test_raw_dist_trajectory(du, g_raw):
    g = g_raw
    c, bx, ch, s1, s2 = g["coords"], g["box"], g["chains"], g["sel1"], g["sel2"]
    for key, a, b, selfd, pbc in (("dist_pbc", s1, s2, False, True), ("dist_nopbc", s1, s2, False, False),
                                  ("dist_self_pbc", s1, s1, True, True)):
        r = np.zeros_like(g[key])
        du.dist_trajectory(c, bx, a, b, ch, selfd, pbc, r)
        assert np.array_equal(_bits(r), _bits(g[key])), key


def test_raw_contacts(du, g_raw):
    g = g_raw
    c, bx, ch, s1, s2 = g["coords"], g["box"], g["chains"], g["sel1"], g["sel2"]
    off, pairs = du.contacts_trajectory_arrays(c, bx, s1, s2, ch, False, True, 6.5)
    assert np.diff(off).tolist() == g["ct_cnt"].tolist() and np.array_equal(pairs, g["ct_pairs"])
    off, pairs = du.contacts_trajectory_arrays(c, bx, s1, s1, ch, True, True, 7.25)
    assert np.diff(off).tolist() == g["ct_self_cnt"].tolist() and np.array_equal(pairs, g["ct_self_pairs"])
    lst = du.contacts_trajectory(c, bx, s1, s2, ch, False, True, 6.5)
    assert isinstance(lst, list) and len(lst) == c.shape[2] and lst[0] == g["ct_pairs"][:g["ct_cnt"][0]].reshape(-1).tolist()


def test_raw_reductions(du, g_raw):
    g = g_raw
    c, bx, ch = g["coords"], g["box"], g["chains"]
    F = c.shape[2]
    g1, g2 = _groups(g["g1_off"], g["g1_atoms"]), _groups(g["g2_off"], g["g2_atoms"])
    gc1 = np.array([ch[x[0]] for x in g1], np.uint32); gc2 = np.array([ch[x[0]] for x in g2], np.uint32)
    for r1 in (0, 1):
        for r2 in (0, 1):
            r = np.zeros((F, len(g1) * len(g2)), np.float32)
            du.dist_trajectory_reduction(c, bx, g1, g2, gc1, gc2, False, True, g["masses"], r1, r2, r)
            assert np.array_equal(_bits(r), _bits(g[f"red_{r1}{r2}"])), (r1, r2)
    r = np.zeros_like(g["red_self"])
    du.dist_trajectory_reduction(c, bx, g1, g1, gc1, gc1, True, True, g["masses"], 0, 0, r)
    assert np.array_equal(_bits(r), _bits(g["red_self"]))
    r = np.zeros_like(g["red_pairs_01"])
    du.dist_trajectory_reduction_pairs(c, bx, g1[:6], g2, gc1[:6], gc2, True, g["masses"], 0, 1, r)
    assert np.array_equal(_bits(r), _bits(g["red_pairs_01"]))


def test_reductions_many_small_groups_thread_kernel(du, oracle, monkeypatch):
    """Many small groups (residue-level minimum distances) take the one-thread-per-group-pair kernel; it must give the bits
    of the oracle and of the warp-per-pair kernel (MKB_K5_WARP=1): closest / COM combinations, self distances, a NaN first
    atom (sticks, pyx:260-275), an unusable box component."""
    rng = np.random.default_rng(23)
    N, F, NG = 420, 5, 90
    c = (rng.normal(size=(N, 3, F)) * 11).astype(np.float32)
    bx = np.abs(rng.normal(size=(3, F)) * 2 + 19).astype(np.float32)
    bx[2, 3] = 0.0
    cuts = np.sort(rng.choice(np.arange(1, N), NG - 1, replace=False))
    groups = [list(range(a, b)) for a, b in zip(np.concatenate([[0], cuts]), np.concatenate([cuts, [N]]))]
    c[groups[7][0], 1, 2] = np.nan  # first atom of group 7 in frame 2
    gch = rng.integers(0, 3, NG).astype(np.uint32)
    masses = rng.uniform(1, 16, N).astype(np.float32)
    g2 = groups[:70]
    for selfd, ga, gb, ca, cb in ((False, groups, g2, gch, gch[:70]), (True, groups, groups, gch, gch)):
        P = du.n_columns(len(ga), len(gb), selfd)
        for r1, r2 in ((0, 0), (1, 0), (1, 1)):
            if selfd and (r1, r2) != (0, 0):
                continue
            want = np.zeros((F, P), np.float32)
            with np.errstate(all="ignore"):
                oracle.dist_trajectory_reduction(c, bx, ga, gb, ca, cb, selfd, True, masses, r1, r2, want)
            got = np.zeros((F, P), np.float32)
            du.dist_trajectory_reduction(c, bx, ga, gb, ca, cb, selfd, True, masses, r1, r2, got)
            with monkeypatch.context() as m:
                m.setenv("MKB_K5_WARP", "1")
                warp = np.zeros((F, P), np.float32)
                du.dist_trajectory_reduction(c, bx, ga, gb, ca, cb, selfd, True, masses, r1, r2, warp)
            assert np.isnan(want).any() and np.array_equal(_bits(got), _bits(want)), (selfd, r1, r2)
            assert np.array_equal(_bits(warp), _bits(want)), (selfd, r1, r2)


def test_raw_cdist_pdist_squareform_collisions(du, g_raw):
    g = g_raw
    for D in (1, 2, 3, 5):
        a, b = g[f"cd{D}_a"], g[f"cd{D}_b"]
        r = np.zeros((len(a), len(b)), np.float32); du.cdist(a, b, r)
        assert np.array_equal(_bits(r), _bits(g[f"cd{D}_out"]))
        p = np.zeros(len(a) * (len(a) - 1) // 2, np.float32); du.pdist(a, p)
        assert np.array_equal(_bits(p), _bits(g[f"pd{D}_out"]))
    assert np.array_equal(du.squareform(g["pd3_out"]), g["sq_out"])
    coll = np.array(du.get_collisions(g["cd3_a"], g["cd3_b"], 6.0), np.uint32).reshape(-1, 2)
    assert np.array_equal(coll, g["coll_out"])


def test_reference_test_distance_py():
    """tests/test_distance.py:1-28 of the reference, verbatim expectations."""
    from moleculekit_b200.distance import cdist, pdist, squareform

    x = np.array([0, 1, 2])[:, None]; y = np.array([3, 4, 5])[:, None]
    assert np.allclose(cdist(x, y), [[3.0, 4.0, 5.0], [2.0, 3.0, 4.0], [1.0, 2.0, 3.0]])
    d = cdist(np.array([[0, 1], [2, 3]]), np.array([[4, 5], [6, 7], [8, 9]]))
    assert np.allclose(d, [[5.656854, 8.485281, 11.313708], [2.828427, 5.656854, 8.485281]])
    p = pdist(np.array([[4, 5], [6, 7], [8, 9]]))
    assert np.allclose(p, [2.828427, 5.656854, 2.828427]) and p.dtype == np.float32
    sq = squareform(p)
    assert sq.shape == (3, 3) and np.allclose(sq, sq.T) and np.allclose(np.diag(sq), 0) and sq[0, 2] == p[1]


def test_random_vs_oracle_large_wraps_and_half_ties(du, oracle):
    """Unwrapped walk with |n| up to ~6, integer lattice coordinates with d/box exactly k + 0.5 (roundf half away),
    a zero box component (NaN distances) -- all must match the oracle bit for bit, including NaN patterns."""
    rng = np.random.default_rng(11)
    N, F = 160, 37
    c = (rng.normal(size=(N, 3, F)) * 40).astype(np.float32)
    c[:40] = np.round(c[:40])  # integers: with box = 2, 4, 6 many quotients are exact half-integers
    bx = np.tile(np.array([[2.0], [4.0], [6.0]], np.float32), (1, F))
    bx[:, F // 2:] = np.abs(rng.normal(size=(3, F - F // 2)) * 3 + 17).astype(np.float32)
    bx[1, -1] = 0.0
    ch = rng.integers(0, 3, N).astype(np.uint32)
    s1 = np.sort(rng.choice(N, 70, replace=False)).astype(np.uint32)
    s2 = np.sort(rng.choice(N, 90, replace=False)).astype(np.uint32)
    for selfd, a, b in ((False, s1, s2), (True, s1, s1)):
        P = du.n_columns(len(a), len(b), selfd)
        want = np.zeros((F, P), np.float32); oracle.dist_trajectory(c, bx, a, b, ch, selfd, True, want)
        got = np.zeros((F, P), np.float32); du.dist_trajectory(c, bx, a, b, ch, selfd, True, got)
        assert np.isnan(want).any() and np.array_equal(_bits(got), _bits(want))
        assert du.contacts_trajectory(c, bx, a, b, ch, selfd, True, 9.5) == oracle.contacts_trajectory(c, bx, a, b, ch, selfd, True, 9.5)


def _ulps(got, want):
    """distance in float32 units in the last place (finite, non-negative values)"""
    return np.abs(got.view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64))


def test_fast_distances_within_tolerance(du, oracle, mol, g_traj):
    """exact=False (MKB_DIST_DISTANCES_FAST): the reference's minimum-image roundings, a fused sum of squares and MUFU.SQRT.
    Tolerance written here: 2e-6 relative (BASELINE's north star allows 1e-5), >= 99.9 % of the values within 4 ulp, the
    same NaN and zero pattern; pairs whose quotient |d/box| reaches 2^21 are outside the mode's stated domain."""
    rng = np.random.default_rng(23)
    N, F = 200, 41
    c = np.cumsum(rng.normal(size=(N, 3, F)).astype(np.float32) * 3, axis=2) + (rng.normal(size=(N, 3, 1)) * 60).astype(np.float32)
    c[:30] = np.round(c[:30])
    c[5] = c[4]  # coincident atoms: distance exactly 0
    bx = np.abs(rng.normal(size=(3, F)) * 6 + 31).astype(np.float32)
    bx[:, :5] = np.array([[2.0], [4.0], [6.0]], np.float32)  # exact half-integer quotients with the integer coordinates
    bx[2, -1] = 0.0   # NaN frame
    bx[0, -2] = np.inf
    ch = rng.integers(0, 3, N).astype(np.uint32)
    s1 = np.sort(rng.choice(N, 80, replace=False)).astype(np.uint32); s1[:2] = (4, 5); s1.sort()
    s2 = np.sort(rng.choice(N, 130, replace=False)).astype(np.uint32)
    worst = 0
    for selfd, a, b, tr in ((False, s1, s2, None), (True, s1, s1, None), (False, s1, s2, 17.25), (False, s1[:3], s2[:1], None)):
        P = du.n_columns(len(a), len(b), selfd)
        want = np.zeros((F, P), np.float32)
        with np.errstate(all="ignore"):
            oracle.dist_trajectory(c, bx, a, b, ch, selfd, True, want)
        if tr is not None:
            want[want > tr] = tr
        got = np.full((F, P), -1.0, np.float32)
        du.dist_trajectory(c, bx, a, b, ch, selfd, True, got, truncate=tr, exact=False)
        assert np.array_equal(np.isnan(got), np.isnan(want)) and np.isnan(want).any()
        ok = ~np.isnan(want)
        assert np.array_equal(got[ok] == 0, want[ok] == 0) and (selfd or tr is not None or P < 10 or (want[ok] == 0).any())
        np.testing.assert_allclose(got[ok], want[ok], rtol=2e-6, atol=0)
        u = _ulps(got[ok], want[ok])
        assert (u <= 4).mean() >= 0.999, (u.max(), (u > 4).mean())
        worst = max(worst, int(u.max()))
        exact = np.zeros((F, P), np.float32); du.dist_trajectory(c, bx, a, b, ch, selfd, True, exact, truncate=tr)
        assert np.array_equal(_bits(exact), _bits(want))   # the default stays bit-identical on the same inputs
    assert worst <= 32

    # differences within +-3 ulp of 0.5, 1.5, 2.5 ... box lengths (the image integer may differ from the reference's there)
    boxes = np.array([10.0, 7.3, 33.333, 8.0, 3e4, -5.0, np.nan, np.inf, 0.0], np.float32)
    Fb = len(boxes)
    mults = [0.0, 0.25, 0.5, 0.75, 1.0, 1.4998, 1.4999, 1.49995, 1.5, 2.0, 2.5, 3.5, 100.5]
    cols = []
    for f in range(Fb):
        bb = boxes[f] if np.isfinite(boxes[f]) and boxes[f] != 0 else np.float32(6.0)
        vals = []
        for m in mults:
            v = np.float32(np.float32(m) * np.abs(bb))
            for k in range(-3, 4):
                w = v
                for _ in range(abs(k)):
                    w = np.nextafter(w, np.float32(np.inf if k > 0 else -np.inf), dtype=np.float32)
                vals += [w, -w]
        cols.append(np.array(vals, np.float32))
    n2 = len(cols[0])
    cb = np.zeros((1 + n2, 3, Fb), np.float32)
    for f in range(Fb):
        cb[1:, 0, f] = cols[f]; cb[1:, 1, f] = cols[f][::-1]; cb[1:, 2, f] = np.roll(cols[f], 5)
    bxb = np.repeat(boxes[None, :], 3, axis=0).copy()
    sa = np.array([0], np.uint32); sb = np.arange(1, 1 + n2, dtype=np.uint32)
    chb = np.zeros(1 + n2, np.uint32); chb[1:] = 1
    want = np.zeros((Fb, n2), np.float32)
    with np.errstate(all="ignore"):
        oracle.dist_trajectory(cb, bxb, sa, sb, chb, False, True, want)
    got = np.zeros((Fb, n2), np.float32); du.dist_trajectory(cb, bxb, sa, sb, chb, False, True, got, exact=False)
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.isnan(want).any() and np.isfinite(want).any()
    ok = ~np.isnan(want)
    np.testing.assert_allclose(got[ok], want[ok], rtol=2e-6, atol=0)

    # through the projection: MetricDistance.exact = False against the frozen output of the reference's compiled kernel
    from moleculekit_b200.projections.metricdistance import MetricDistance

    m = MetricDistance("protein and name CA", "resname MOL and noh", metric="distances", periodic="selections")
    m.exact = False
    data = m.project(mol)
    np.testing.assert_allclose(data, g_traj["ref_distances"], rtol=2e-6, atol=0)
    assert (_ulps(data, g_traj["ref_distances"]) <= 4).mean() >= 0.999


def test_minimum_image_boundaries(du, oracle):
    """The dense kernels take n = rint(d * (1/b)) and verify it with |d - b n| < b/2 - 1e-6 |d|: differences placed
    within a few ulps of b/2, 1.4999 b, 1.5 b, 2.5 b (both signs), boxes that are powers of two / arbitrary / tiny / huge /
    zero / negative / NaN / Inf, and NaN coordinates must all reproduce the reference bits."""
    boxes = np.array([10.0, 7.3, 33.333, 8.0, 1e-3, 3e4, 0.0, -5.0, np.nan, np.inf, 1e-35, 1e32], np.float32)
    F = len(boxes)
    mults = [0.0, 0.25, 0.5, 0.75, 1.0, 1.4998, 1.4999, 1.49995, 1.5, 2.0, 2.5, 3.5, 100.5]
    cols = []
    for f in range(F):
        b = boxes[f] if np.isfinite(boxes[f]) and boxes[f] != 0 else np.float32(6.0)
        vals = []
        for m in mults:
            v = np.float32(np.float32(m) * np.abs(b))
            for k in range(-3, 4):
                w = v
                for _ in range(abs(k)):
                    w = np.nextafter(w, np.float32(np.inf if k > 0 else -np.inf), dtype=np.float32)
                vals += [w, -w]
        cols.append(np.array(vals, np.float32))
    n2 = len(cols[0])
    c = np.zeros((1 + n2, 3, F), np.float32)
    for f in range(F):
        c[1:, 0, f] = cols[f]
        c[1:, 1, f] = cols[f][::-1]
        c[1:, 2, f] = np.roll(cols[f], 5)
    c[7, 2, 0] = np.nan
    bx = np.repeat(boxes[None, :], 3, axis=0).copy()
    bx[1] = np.roll(boxes, 1)
    s1 = np.array([0], np.uint32); s2 = np.arange(1, 1 + n2, dtype=np.uint32)
    ch = np.zeros(1 + n2, np.uint32); ch[1:] = 1
    want = np.zeros((F, n2), np.float32)
    with np.errstate(all="ignore"):
        oracle.dist_trajectory(c, bx, s1, s2, ch, False, True, want)
    got = np.zeros((F, n2), np.float32); du.dist_trajectory(c, bx, s1, s2, ch, False, True, got)
    assert np.isnan(want).any() and np.isfinite(want).any()
    bad = np.where(_bits(got) != _bits(want))
    assert bad[0].size == 0, (bad[0][:5], bad[1][:5], got[bad][:5], want[bad][:5])
    for thr in (3.0, 5.5):
        assert du.contacts_trajectory(c, bx, s1, s2, ch, False, True, thr) == oracle.contacts_trajectory(c, bx, s1, s2, ch, False, True, thr)
    # the dense contact map decides most pairs on a fused estimate of d2 and re-checks a band around the threshold (and every
    # frame whose box is too small for the shortcut) with the exact sequence: same booleans as `reference distance <= thr`
    # on all of the above, for thresholds below, at and above half a box and exactly on values that occur
    with np.errstate(all="ignore"):
        for thr in (0.5, 2.5, 3.65, 5.0, 16.6665, 40.0, float(want[0, 30]), float(want[2, 11])):
            gotc = du.dist_trajectory(c, bx, s1, s2, ch, False, True, None, metric="contacts", threshold=thr)
            assert np.array_equal(gotc, want <= np.float32(thr)), thr


def test_contacts_fill_from_cached_masks_equals_recomputation(du, oracle, monkeypatch):
    """mkb_contacts_fill normally emits the pairs from the hit masks its count call left behind; with MKB_K4_NO_BALLOTS the
    distances are evaluated a second time.  Same ordered pairs either way (self and non-self, ragged column counts)."""
    rng = np.random.default_rng(19)
    N, F = 333, 9
    c = (rng.normal(size=(N, 3, F)) * 9).astype(np.float32)
    bx = np.abs(rng.normal(size=(3, F)) * 2 + 21).astype(np.float32)
    ch = rng.integers(0, 2, N).astype(np.uint32)
    s1 = np.sort(rng.choice(N, 77, replace=False)).astype(np.uint32)
    s2 = np.sort(rng.choice(N, 131, replace=False)).astype(np.uint32)
    for selfd, a, b in ((False, s1, s2), (True, s2, s2)):
        want = oracle.contacts_trajectory(c, bx, a, b, ch, selfd, True, 7.5)
        cached = du.contacts_trajectory(c, bx, a, b, ch, selfd, True, 7.5)
        with monkeypatch.context() as m:
            m.setenv("MKB_K4_NO_BALLOTS", "1")
            recomputed = du.contacts_trajectory(c, bx, a, b, ch, selfd, True, 7.5)
        assert cached == want and recomputed == want and sum(len(x) for x in want) > 100


def test_contact_threshold_ties(du, oracle):
    """metric="contacts" is decided on d2 (no sqrt): must equal sqrtf(d2) <= threshold of the reference incl. exact
    ties (3-4-5 triangles) and distances one ulp either side of the threshold."""
    rng = np.random.default_rng(2)
    n, F = 96, 3
    base = rng.integers(-6, 7, size=(n, 3)).astype(np.float32)
    c = np.repeat(base[:, :, None], F, axis=2)
    c[:, :, 1] += (rng.integers(-2, 3, size=(n, 3)) * np.float32(2.0 ** -20)).astype(np.float32)  # +- a few ulp
    c[:, :, 2] += rng.normal(0, 1e-3, size=(n, 3)).astype(np.float32)
    bx = np.zeros((3, F), np.float32)
    s1 = np.arange(0, 40, dtype=np.uint32); s2 = np.arange(40, n, dtype=np.uint32)
    ch = np.zeros(n, np.uint32)
    want_d = np.zeros((F, 40 * 56), np.float32); oracle.dist_trajectory(c, bx, s1, s2, ch, False, False, want_d)
    for thr in (5.0, 3.0, 7.0710678, 1e-3):
        got = du.dist_trajectory(c, bx, s1, s2, ch, False, False, None, metric="contacts", threshold=thr)
        assert got.dtype == bool and np.array_equal(got, want_d <= np.float32(thr)), thr
        assert (want_d == np.float32(thr)).any() or thr not in (5.0, 3.0)   # the tie cases really occur
        gt = du.dist_trajectory(c, bx, s1, s2, ch, False, False, None, metric="contacts", threshold=thr, truncate=4.0)
        wt = want_d.copy(); wt[wt > np.float32(4.0)] = np.float32(4.0)
        assert np.array_equal(gt, wt <= np.float32(thr))


def test_frame_shard_view_equals_full(du, g_raw):
    """A frame slice of a resident device trajectory (what a GPU shard sees) gives the same rows."""
    import torch

    g = g_raw
    dev = torch.device("cuda", 0)
    c = torch.from_numpy(g["coords"]).to(dev); bx = torch.from_numpy(g["box"]).to(dev)
    s1 = torch.from_numpy(g["sel1"].astype(np.int32)).to(dev); s2 = torch.from_numpy(g["sel2"].astype(np.int32)).to(dev)
    ch = torch.from_numpy(g["chains"].view(np.int32)).to(dev)
    full = du.dist_trajectory_device(c, bx, s1, s2, ch, False, True)
    part = du.dist_trajectory_device(c[:, :, 2:6], bx[:, 2:6], s1, s2, ch, False, True)
    assert torch.equal(full[2:6], part)
    assert np.array_equal(_bits(full.cpu().numpy()), _bits(g["dist_pbc"]))


def test_full_size_c4a_properties(du, oracle):
    """BASELINE config 4a at full size (10 000 frames x 256 x 1024 periodic pairs, 10.5 GB of float32 output) through
    size-independent properties: swapping the selections gives the transposed matrix bit for bit (the minimum image is
    symmetric), the fused contact map equals distances <= threshold, a frame shard equals the same rows, and ten full
    frames equal the oracle."""
    import torch

    n1, n2, F, L = 256, 1024, 10000, 36.84
    rng = np.random.default_rng(7)
    start = rng.uniform(0, L, size=(n1 + n2, 3, 1)).astype(np.float32)
    coords = start + np.cumsum(rng.normal(0, 0.3, size=(n1 + n2, 3, F)).astype(np.float32), axis=2)
    box = np.repeat((L * (1 + 0.002 * rng.normal(size=F))).astype(np.float32)[None, :], 3, axis=0)
    dev = torch.device("cuda", 0)
    d_c = torch.from_numpy(np.ascontiguousarray(coords)).to(dev); d_b = torch.from_numpy(np.ascontiguousarray(box)).to(dev)
    s1 = torch.arange(0, n1, dtype=torch.int32, device=dev); s2 = torch.arange(n1, n1 + n2, dtype=torch.int32, device=dev)
    ch = torch.ones(n1 + n2, dtype=torch.int32, device=dev); ch[n1:] = 2
    dist = du.dist_trajectory_device(d_c, d_b, s1, s2, ch, False, True)
    assert dist.shape == (F, n1 * n2) and bool(torch.isfinite(dist).all()) and float(dist.max()) <= L * 0.5 * 3 ** 0.5 * 1.01
    for f0 in range(0, F, 2500):  # swapped selections, in frame shards to bound memory
        sw = du.dist_trajectory_device(d_c[:, :, f0:f0 + 2500], d_b[:, f0:f0 + 2500], s2, s1, ch, False, True)
        assert torch.equal(sw.view(-1, n2, n1).transpose(1, 2).reshape(-1, n1 * n2), dist[f0:f0 + 2500])
        del sw
    con = du.dist_trajectory_device(d_c, d_b, s1, s2, ch, False, True, metric="contacts", threshold=12.0)
    assert con.dtype == torch.bool and torch.equal(con, dist <= 12.0) and 0.05 < float(con.float().mean()) < 0.5
    del con
    frames = [0, 1, 777, 2499, 2500, 4999, 5000, 7501, 9998, 9999]
    ch_h = np.ones(n1 + n2, np.uint32); ch_h[n1:] = 2
    sub = np.ascontiguousarray(coords[:, :, frames]); subb = np.ascontiguousarray(box[:, frames])
    want = np.zeros((len(frames), n1 * n2), np.float32)
    oracle.dist_trajectory(sub, subb, np.arange(n1, dtype=np.uint32), np.arange(n1, n1 + n2, dtype=np.uint32), ch_h, False,
                           True, want)
    assert np.array_equal(dist[frames].cpu().numpy().view(np.uint32), want.view(np.uint32))


# ------------------------------------------------------------------------------------------ MetricDistance API
def test_distances_and_contacts(mol, g_traj):
    """tests/test_metricdistance.py:37-51,182-193"""
    from moleculekit_b200.projections.metricdistance import MetricDistance

    data = MetricDistance("protein and name CA", "resname MOL and noh", metric="distances", periodic="selections").project(mol)
    assert data.dtype == np.float32 and data.shape == (20, 2493)
    assert np.allclose(data, g_traj["gold_distances"], atol=1e-3), "Distance calculation is broken"
    assert np.array_equal(_bits(data), _bits(g_traj["ref_distances"]))
    cont = MetricDistance("protein and name CA", "resname MOL and noh", periodic="selections", metric="contacts",
                          threshold=8).project(mol)
    assert cont.dtype == bool and np.allclose(cont, g_traj["gold_distances"] < 8, atol=1e-3)
    assert np.array_equal(cont, g_traj["ref_distances"] <= np.float32(8))
    trunc = MetricDistance("protein and name CA", "resname MOL and noh", periodic="selections", truncate=12.5).project(mol)
    ref = g_traj["ref_distances"].copy(); ref[ref > 12.5] = 12.5
    assert np.array_equal(_bits(trunc), _bits(ref))


def test_mindistances(mol, g_traj):
    """tests/test_metricdistance.py:196-228"""
    from moleculekit_b200.projections.metricdistance import MetricDistance

    kw = dict(periodic="selections", groupsel1="residue", groupsel2="all")
    data = MetricDistance("protein and noh", "resname MOL and noh", **kw).project(mol)
    assert data.shape == (20, 277) and np.allclose(data, g_traj["gold_mindistances"], atol=1e-3)
    assert np.array_equal(_bits(data), _bits(g_traj["ref_mindistances"]))
    data = MetricDistance("protein and noh", "resname MOL and noh", truncate=3, **kw).project(mol)
    assert np.allclose(data, np.clip(g_traj["gold_mindistances"], 0, 3), atol=1e-3)


def test_selfmindistance(mol, g_traj):
    """tests/test_metricdistance.py:231-278 (manual == auto == golden[::10])"""
    from moleculekit_b200.projections.metricdistance import MetricDistance, MetricSelfDistance

    sel = "protein and resid 1 to 50 and noh"
    manual = MetricDistance(sel, sel, periodic=None, groupsel1="residue", groupsel2="residue").project(mol)
    auto = MetricSelfDistance(sel, groupsel="residue").project(mol)
    assert manual.shape == (20, 1225) and np.array_equal(manual, auto)
    assert np.allclose(auto, g_traj["gold_selfmindistance"], atol=1e-3)
    assert np.array_equal(_bits(auto), _bits(g_traj["ref_selfmindistance"]))


def test_periodicity_and_com(mol, g_traj):
    """tests/test_metricdistance.py:329-352 + COM reductions frozen from the reference"""
    from moleculekit_b200.projections.metricdistance import MetricDistance

    a = "protein and resid 1 to 20 and noh"
    d1 = MetricDistance(a, "resname MOL and noh", periodic="selections").project(mol)
    d2 = MetricDistance(a, "resname MOL and noh", periodic="chains").project(mol)
    assert np.allclose(d1, d2) and np.array_equal(_bits(d2), _bits(g_traj["ref_chains_distances"]))
    m2 = mol.copy(); m2.chain[:] = ""
    d3 = MetricDistance(a, "resname MOL and noh", periodic="chains").project(m2)
    assert not np.allclose(d1, d3)
    b = "protein and resid 1 to 50 and noh"
    kw = dict(groupsel1="residue", groupsel2="all")
    cc = MetricDistance(b, "resname MOL and noh", "selections", groupreduce1="com", groupreduce2="com", **kw).project(mol)
    assert np.array_equal(_bits(cc), _bits(g_traj["ref_com_com"]))
    cl = MetricDistance(b, "resname MOL and noh", "selections", groupreduce1="com", groupreduce2="closest", **kw).project(mol)
    assert np.array_equal(_bits(cl), _bits(g_traj["ref_com_closest"]))


def test_atomselect_forms_and_mapping(mol):
    """tests/test_metricdistance.py:54-96: string / index / bool / manual groups, mapping shape"""
    from moleculekit_b200.projections.metricdistance import MetricSelfDistance

    data = MetricSelfDistance("protein and name CA", metric="contacts", threshold=8).project(mol)
    assert data.shape == (20, 38226) and data.dtype == bool
    ca = mol.atomselect("protein and name CA", indexes=True)
    m = MetricSelfDistance(ca, metric="contacts", threshold=8)
    assert np.array_equal(m.project(mol), data) and m.getMapping(mol).shape == (38226, 3)
    assert np.array_equal(MetricSelfDistance(mol.atomselect("protein and name CA"), metric="contacts").project(mol), data)
    m = MetricSelfDistance([ca[0::3], ca[1::3], ca[2::3]])
    d3 = m.project(mol); mp = m.getMapping(mol)
    assert d3.shape == (20, 3) and mp.shape == (3, 3)
    masks = [np.isin(np.arange(mol.numAtoms), ca[k::3]) for k in range(3)]
    assert np.array_equal(MetricSelfDistance(masks).project(mol), d3)


def test_distances_trivial():
    """tests/test_metricdistance.py:99-179 (analytic, incl. PBC wrap in a 2 A box and column ordering)"""
    from moleculekit_b200.molecule_lite import MolLite
    from moleculekit_b200.projections.metricdistance import MetricDistance

    coords = np.zeros((3, 3, 2), dtype=np.float32)
    coords[1, :, 0] = [3, 3, 3]; coords[2, :, 0] = [5, 5, 5]; coords[1, :, 1] = [7, 7, 7]; coords[2, :, 1] = [6, 6, 6]
    mol = MolLite(coords, element=["C"] * 3, name=["C"] * 3, chain=list("012"))
    i0, i12 = np.array([0]), np.array([1, 2])
    real = np.linalg.norm(coords[[1, 2], :, :], axis=1).T
    assert np.allclose(MetricDistance(i0, i12, metric="distances", periodic=None).project(mol), real)
    wrapped = np.linalg.norm(np.mod(coords, 2)[[1, 2], :, :], axis=1).T
    mol.box = np.full((3, 2), 2, dtype=np.float32)
    assert np.allclose(MetricDistance(i0, i12, metric="distances", periodic="selections").project(mol), wrapped)
    data = MetricDistance(i0, i12, metric="distances", periodic=None, groupsel1="all", groupsel2="all").project(mol)
    assert np.allclose(data.flatten(), np.min(real, axis=1))
    coords = np.zeros((4, 3, 2), dtype=np.float32)
    coords[1, :, 0] = [1, 1, 1]; coords[2, :, 0] = [3, 3, 3]; coords[3, :, 0] = [5, 5, 5]
    coords[1, :, 1] = [1, 1, 1]; coords[2, :, 1] = [7, 7, 7]; coords[3, :, 1] = [6, 6, 6]
    mol = MolLite(coords, element=["C"] * 4, chain=list("0123"))
    real = np.hstack((np.linalg.norm(coords[[2, 3]] - coords[0], axis=1).T, np.linalg.norm(coords[[2, 3]] - coords[1], axis=1).T))
    assert np.allclose(MetricDistance(np.array([0, 1]), np.array([2, 3]), metric="distances", periodic=None).project(mol), real)


def test_com_and_pair_distances(g_3ptb):
    """tests/test_metricdistance.py:355-493 (analytic COM cases, 3ptb constants, pairs=True)"""
    from moleculekit_b200.molecule_lite import MolLite
    from moleculekit_b200.projections.metricdistance import MetricDistance

    xyz = np.array([[0, 0, 0], [-1, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float32)[:, :, None]
    mol = MolLite(xyz, element=["H"] * 4, resid=[0, 1, 1, 0])
    fix = dict(periodic=None, groupsel1="all", groupsel2="all", groupreduce1="com", groupreduce2="com")
    I = lambda *a: np.array(a)
    assert abs(MetricDistance(I(0), I(1), **fix).project(mol)[0][0] - 1) < 1e-5
    assert abs(MetricDistance(I(0, 1), I(2), **fix).project(mol)[0][0] - 1.5) < 1e-5
    assert abs(MetricDistance(I(0, 1, 2), I(3), **fix).project(mol)[0][0] - 1) < 1e-5
    fix["groupsel1"] = "residue"
    assert np.allclose(MetricDistance(I(0, 1, 2), I(3), **fix).project(mol), [[1, 1]])
    fix["groupsel1"], fix["groupreduce1"] = "all", "closest"
    assert np.allclose(MetricDistance(I(0, 1, 2), I(3), **fix).project(mol), [[1]])
    fix["groupsel1"] = "residue"
    assert np.allclose(MetricDistance(I(0, 1, 2), I(3), **fix).project(mol), [[1, 1.4142135]])

    g = g_3ptb
    mol = MolLite(g["coords"], g["box"], element=g["element"], resid=g["resid"], resname=g["resname"], name=g["name"],
                  chain=g["chain"], named_selections={"protein": g["sel_protein"], "resname BEN": g["sel_ben"]})
    kw = dict(groupsel1="all", groupsel2="all")
    for r1, r2, key, const in (("com", "com", "ref_com_com", None), ("com", "closest", "ref_com_closest", 8.978174),
                               ("closest", "com", "ref_closest_com", 3.8286476),
                               ("closest", "closest", "ref_closest_closest", 2.8153415)):
        d = MetricDistance("protein", "resname BEN", None, groupreduce1=r1, groupreduce2=r2, **kw).project(mol)
        assert np.array_equal(_bits(d), _bits(g[key])), key
        if const is not None:
            assert abs(d[0][0] - const) < 1e-5
    sel1 = np.array([0, 1, 2]).reshape(-1, 1); sel2 = np.array([1, 2, 3]).reshape(-1, 1)
    res = MetricDistance(sel1, sel2, None, pairs=True).project(mol)
    ref = np.linalg.norm(g["coords"][sel1.flatten(), :, 0] - g["coords"][sel2.flatten(), :, 0], axis=1)
    assert np.allclose(res, ref) and MetricDistance(sel1, sel2, None, pairs=True).getMapping(mol).shape == (3, 3)
    res = MetricDistance([g["sel_residue_1"], g["sel_residue_2"]], [g["sel_residue_3"], g["sel_residue_4"]], None,
                         pairs=True).project(mol)
    assert np.array_equal(_bits(res), _bits(g["ref_pairs_residue"]))
    with pytest.raises(RuntimeError, match="Pairs calculation not implemented without groups"):
        MetricDistance(np.array([0, 1]), np.array([2, 3]), None, pairs=True).project(mol)


def test_calculate_contacts_and_metal_coordination(mol, g_traj, g_3ptb, g_5vl5):
    """ordered index pairs: tests/test_interactions.py:329-350 + frozen calculate_contacts outputs"""
    from moleculekit_b200.distance import calculate_contacts
    from moleculekit_b200.molecule_lite import MolLite

    sel = {s: m for s, m in zip(g_traj["sel_strings"].tolist(), g_traj["sel_masks"])}
    ca, lig, noh = sel["protein and name CA"], sel["resname MOL and noh"], sel["protein and noh"]
    for key, a, b, per, thr in (("ct_ca_lig_sel8", ca, lig, "selections", 8), ("ct_ca_ca_none6", ca, ca, None, 6),
                                ("ct_noh_lig_chains5", noh, lig, "chains", 5)):
        res = calculate_contacts(mol, a, b, per, thr)
        assert [len(r) for r in res] == g_traj[key + "_cnt"].tolist(), key
        assert all(r.dtype == np.uint32 and r.shape[1:] == (2,) for r in res)
        assert np.array_equal(np.vstack(res), g_traj[key + "_pairs"]), key
    m3 = MolLite(g_3ptb["coords"], g_3ptb["box"], element=g_3ptb["element"])
    res = calculate_contacts(m3, g_3ptb["mc_sel1"], g_3ptb["mc_sel2"], None, 3.5)
    assert np.array_equal(res[0], g_3ptb["mc_expected"])
    m5 = MolLite(g_5vl5["coords"], g_5vl5["box"], element=g_5vl5["element"])
    per = None if np.all(g_5vl5["box"] == 0) else "selections"
    r1 = calculate_contacts(m5, g_5vl5["mc_a_sel1"], g_5vl5["mc_a_sel2"], per, 3.5)
    r2 = calculate_contacts(m5, g_5vl5["mc_b_sel1"], g_5vl5["mc_b_sel2"], per, 3.5)
    assert np.array_equal(np.vstack((r1[0], r2[0])), g_5vl5["mc_expected"])


def test_errors(mol):
    from moleculekit_b200.molecule_lite import MolLite
    from moleculekit_b200.projections.metricdistance import MetricDistance

    with pytest.raises(RuntimeError, match="can only be None, 'chains' or 'selections'"):
        MetricDistance("a", "b", periodic="everything")
    nobox = MolLite(np.zeros((3, 3, 2), np.float32))
    with pytest.raises(RuntimeError, match="No periodic box dimensions given"):
        MetricDistance(np.array([0]), np.array([1, 2]), periodic="selections").project(nobox)
    with pytest.raises(RuntimeError, match="Selection returned 0 atoms"):
        MetricDistance(np.zeros(3, bool), np.array([1]), None).project(nobox)
    with pytest.raises(RuntimeError, match="not supported"):
        MetricDistance(np.array([0]), np.array([1]), None, metric="bananas").project(nobox)
    bad = MolLite(np.zeros((3, 3, 2), np.float32), box=np.ones((3, 5), np.float32))
    with pytest.raises(RuntimeError, match="Different number of frames"):
        MetricDistance(np.array([0]), np.array([1]), "selections").project(bad)


def test_metricshell(mol, g_traj, oracle):
    """tests/test_metricshell.py:8-56 -- stored golden (frames ::10), today's reference output (exact), the analytic
    3-atom cases, a symmetric selection with truncate, and a random case against the oracle restatement."""
    from moleculekit_b200.molecule_lite import MolLite
    from moleculekit_b200.projections.metricshell import MetricShell

    m = MetricShell("protein and name CA", "resname MOL and noh", periodic="selections")
    data = m.project(mol)
    assert data.shape == (20, 277 * 4) and data.dtype == np.float64
    assert np.allclose(data, g_traj["gold_shell"]), "Shell density calculation is broken"
    assert np.array_equal(data, g_traj["ref_shell"])
    assert m.getMapping(mol).shape == (277 * 4, 3)
    self_ = MetricShell("resname MOL and noh", "resname MOL and noh", periodic=None, numshells=6, shellwidth=1.5,
                        truncate=7.0).project(mol)
    assert np.array_equal(self_, g_traj["ref_shell_self"])

    xyz = np.zeros((3, 3, 1), np.float32); xyz[1, :, 0] = [0.5, 0, 0]; xyz[2, :, 0] = [0, 1.5, 0]
    tiny = MolLite(xyz, name=["CL"] * 3, resname=["CL"] * 3, element=["Cl"] * 3, resid=np.arange(3),
                   named_selections={"name CL": np.ones(3, bool)})
    d = MetricShell("name CL", "name CL", periodic=None).project(tiny)
    assert np.allclose(d, [[0.01768388256576615, 0, 0, 0, 0.01768388256576615, 0, 0, 0, 0.01768388256576615, 0, 0, 0]])
    d = MetricShell("name CL", "name CL", numshells=2, shellwidth=1, periodic=None).project(tiny)
    assert np.allclose(d, [[0.23873241, 0.03410463, 0.23873241, 0.03410463, 0.0, 0.06820926]])

    # two DIFFERENT strings that select the same atoms: the reference counts only later partners and has no row for the
    # last atom (its shell loop is keyed on string equality, the matrix is condensed) -- pinned here
    tiny2 = MolLite(xyz, name=["CL"] * 3, resname=["CL"] * 3, element=["Cl"] * 3, resid=np.arange(3),
                    named_selections={"name CL": np.ones(3, bool), "resname CL": np.ones(3, bool)})
    q = MetricShell("name CL", "resname CL", numshells=2, shellwidth=1, periodic=None)
    d = q.project(tiny2)
    vol = 4 / 3 * np.pi * np.array([1.0, 7.0])
    # pairs (0,1) d=0.5, (0,2) d=1.5, (1,2) d=sqrt(2.5): centre 0 sees one atom per shell, centre 1 one in shell 2
    assert np.allclose(d, [[1 / vol[0], 1 / vol[1], 0.0, 1 / vol[1]]])
    assert q.getMapping(tiny2).shape == (2 * 2, 3)

    rng = np.random.default_rng(4)
    c = (rng.normal(size=(70, 3, 5)) * 6).astype(np.float32)
    c[:20] = np.round(c[:20])          # distances exactly on shell edges (3-4-5 ...) must fall in the lower shell
    bx = np.full((3, 5), 21.0, np.float32)
    rm = MolLite(c, bx, chain=["A"] * 30 + ["B"] * 40, name=["X"] * 70, resname=["R"] * 70)
    s1 = np.arange(0, 30); s2 = np.arange(30, 70)
    ch = np.ones(70, np.uint32); ch[s2] = 2
    for ns, sw, tr in ((4, 3, None), (5, 2.5, 9.0), (3, 1, None)):
        got = MetricShell(s1, s2, periodic="selections", numshells=ns, shellwidth=sw, truncate=tr).project(rm)
        want = oracle.metric_shell(c, bx, s1.astype(np.uint32), s2.astype(np.uint32), ch, False, True, ns, sw, tr)
        assert np.array_equal(got, want), (ns, sw, tr)
