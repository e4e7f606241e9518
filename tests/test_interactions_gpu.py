"""GPU tests of the contact-type interaction detectors (second half): the reference's inline expected pairs
(tests/test_interactions.py:117-161, 329-351) and live reference outputs, reproduced exactly."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mol(g, prefix, coords_key="coords", box_key="box", **extra):
    from moleculekit_b200.molecule_lite import MolLite

    kw = {k: g[f"{prefix}_{k}"] for k in ("resname", "name", "element") if f"{prefix}_{k}" in g}
    return MolLite(g[f"{prefix}_{coords_key}"], box=g[f"{prefix}_{box_key}"], **kw, **extra)


def test_salt_bridges(g_interactions):
    from moleculekit_b200.interactions import get_protein_charged, saltbridge_calculate

    g = g_interactions
    mol = _mol(g, "me6", named_selections={"protein": g["me6_protein"]})
    pos, neg = get_protein_charged(mol)
    assert pos.dtype == np.uint32 and np.array_equal(pos, g["me6_pos"]) and np.array_equal(neg, g["me6_neg"])
    br = saltbridge_calculate(mol, pos, neg, "protein", "protein")
    assert len(br) == 1 and np.array_equal(br[0], np.array([[694, 725], [2146, 2183], [2158, 2346]]))
    assert np.array_equal(br[0], g["me6_bridges"])
    assert saltbridge_calculate(mol, pos, np.zeros(0, np.uint32)) == [[]]
    # default sel2 = sel1, index-array selections
    br_idx = saltbridge_calculate(mol, pos, neg, np.where(g["me6_protein"])[0])
    assert np.array_equal(br_idx[0], g["me6_bridges"])
    # two frames, periodic box, partners on opposite sides of the box in the second frame
    mol2 = _mol(g, "me6", "coords2", "box2")
    half, prot = g["me6_half"], g["me6_protein"]
    br2 = saltbridge_calculate(mol2, pos, neg, half & prot, ~half & prot)
    assert len(br2) == 2 and np.array_equal(br2[0], g["me6_bridges2"]) and np.array_equal(br2[1], g["me6_bridges2"])


def test_hydrophobic_contacts(g_interactions):
    from moleculekit_b200.interactions import hydrophobic_calculate

    g = g_interactions
    mol = _mol(g, "me6", named_selections={"protein": g["me6_protein"]})
    hy = hydrophobic_calculate(mol, g["me6_hyd_sel1"], "protein", 4.0)
    assert len(hy) == 1 and hy[0].dtype == np.uint32 and len(g["me6_hyd"]) > 100
    assert np.array_equal(hy[0], g["me6_hyd"])
    el = g["me6_element"]
    assert (el[hy[0]] == "C").all()


def test_metal_coordination(g_interactions):
    from moleculekit_b200.interactions import metal_coordination_calculate

    g = g_interactions
    for pid, ref in (("5vl5", [[933, 922], [933, 932], [933, 934], [933, 935], [933, 937], [933, 944]]),
                     ("3ptb", [[1629, 383], [1629, 396], [1629, 420], [1629, 460]])):
        mol = _mol(g, pid)
        res = metal_coordination_calculate(mol, g[pid + "_sel1"], g[pid + "_sel2"])
        assert np.array_equal(res[0], np.array(ref, dtype=np.uint32)) and np.array_equal(res[0], g[pid + "_metal"])
