"""GPU parity of K7 (bond perception) against the reference's csv goldens and the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _canonical(b):
    b = np.sort(np.asarray(b, dtype=np.uint32).reshape(-1, 2), axis=1)
    return np.unique(b, axis=0)


def test_reference_goldens(g_bonds):
    """tests/test_bondguesser.py:27-44 for all 15 structures: guess_bonds == csv (bit-exact as a set, canonical order)."""
    from moleculekit_b200.bondguesser import guess_bonds
    from moleculekit_b200.molecule_lite import MolLite

    g = g_bonds
    for pid in g["pdbids"].tolist():
        mol = MolLite(g[f"{pid}_coords"], element=g[f"{pid}_element"], name=g[f"{pid}_name"])
        bonds = guess_bonds(mol)
        assert bonds.dtype == np.uint32 and bonds.shape[1] == 2
        assert np.array_equal(bonds, g[f"{pid}_bonds"]), pid


def test_random_vs_oracle_and_box_enlargement(oracle):
    """Random clouds incl. a sparse unwrapped-like system that triggers the max_boxes enlargement loop, H-H pairs,
    coincident atoms (d2 < 0.001) -- compared with the oracle as canonical sets."""
    from moleculekit_b200.bondguesser import bond_grid_search

    rng = np.random.default_rng(8)
    for n, span, max_boxes in ((400, 12.0, 4e6), (3000, 40.0, 4e6), (500, 900.0, 2e4), (64, 3.0, 4e6)):
        c = (rng.random((n, 3)) * span - span / 3).astype(np.float32)
        c[5] = c[4]                                  # identical coordinates: never bonded
        c[7] = c[6] + np.float32(0.02)               # d2 = 0.0012 > 0.001: bonded if radii allow
        radii = rng.choice([1.0, 1.5, 1.7, 1.9, 2.0], n).astype(np.float32)
        ish = (radii == 1.0).astype(np.uint32)
        want = _canonical(oracle.bond_grid_search(c, np.max(radii) * 1.2, ish, radii, max_boxes=max_boxes))
        got = bond_grid_search(c, np.max(radii) * 1.2, ish, radii, max_boxes=max_boxes)
        assert np.array_equal(got, want), (n, span)
        assert not ((ish[got[:, 0]] == 1) & (ish[got[:, 1]] == 1)).any()


def test_edge_cases_and_errors():
    from moleculekit_b200.bondguesser import bond_grid_search, guess_bonds
    from moleculekit_b200.molecule_lite import MolLite

    assert guess_bonds(MolLite(np.zeros((0, 3, 1), np.float32))).shape == (0, 2)
    one = guess_bonds(MolLite(np.zeros((1, 3, 1), np.float32), element=["C"], name=["C1"]))
    assert one.shape == (0, 2) and one.dtype == np.uint32
    two = MolLite(np.array([[0, 0, 0], [1.5, 0, 0]], np.float32), element=["C", "C"], name=["C1", "C2"])
    assert guess_bonds(two).tolist() == [[0, 1]]
    two.frame = 3
    with pytest.raises(RuntimeError, match="out of range"):
        guess_bonds(two)
    c = np.zeros((3, 3), np.float32); c[1, 0] = np.nan
    with pytest.raises(ValueError, match="non-finite coordinates"):
        bond_grid_search(c, 2.0, np.zeros(3, np.uint32), np.ones(3, np.float32))
    with pytest.raises(ValueError, match="positive, finite grid_cutoff"):
        bond_grid_search(np.zeros((3, 3), np.float32), -1.0, np.zeros(3, np.uint32), np.ones(3, np.float32))
