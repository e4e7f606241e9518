"""The drop-in wrappers fed with the REFERENCE's own ``Molecule``, CPU only.

Runs in the build container, where /root/reference exists: the reference's pure-Python package is imported from there with
its compiled kernels taken from oracle/_ref (the reference's .pyx compiled by oracle/build_ref.py), a real ``Molecule`` is
read from the reference's test files, and the host prologues of ``MetricDistance.project`` / ``MetricSelfDistance`` -- string
selections through ``mol.atomselect``, group building, ``digitize_chains``, the arguments that reach the kernel -- are
compared with the reference's up to the kernel call (both kernels are replaced by recorders, nothing runs on a GPU).
Skipped on the GPU box, where the reference tree does not exist.
"""
import os
import shutil
import sys

import numpy as np
import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "moleculekit")), reason="reference tree not present")


@pytest.fixture(scope="module")
def refmol(tmp_path_factory):
    sys.path.insert(0, ROOT)
    from oracle import build_ref

    if not build_ref.build(verbose=False):
        pytest.skip("oracle/_ref could not be built")
    names = list(build_ref.MODULES)
    mods = dict(zip(names, build_ref.load()))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import moleculekit

    for name, m in mods.items():
        sys.modules["moleculekit." + name] = m
        setattr(moleculekit, name, m)
    from moleculekit.molecule import Molecule

    # the reference's XTC reader writes index caches next to the file it opens: work on a scratch copy
    d = tmp_path_factory.mktemp("traj")
    for f in ("filtered.pdb", "traj.xtc"):
        shutil.copy(os.path.join(REF, "tests", "test_projections", "trajectory", f), d / f)
    mol = Molecule(str(d / "filtered.pdb"))
    mol.read(str(d / "traj.xtc"))
    mol.dropFrames(keep=np.arange(0, mol.numFrames, 20))
    return mol


class _Recorder:
    def __init__(self):
        self.calls = []

    def __call__(self, *args, **kw):
        self.calls.append(args)
        return args[-1] if len(args) and isinstance(args[-1], np.ndarray) else None


def _capture_reference(monkeypatch, fn_name, build_and_project):
    rec = _Recorder()
    # the reference imports the kernel inside its function (util.py:23,99): patch the compiled module it imports from
    monkeypatch.setattr(sys.modules["moleculekit.distance_utils"], fn_name, rec)
    build_and_project()
    assert len(rec.calls) == 1
    return rec.calls[0]


def _capture_ours(monkeypatch, fn_name, build_and_project):
    from moleculekit_b200 import distance_utils as du

    rec = _Recorder()

    def wrapper(*args, **kw):
        rec.calls.append(args)
        out = args[-1] if isinstance(args[-1], np.ndarray) else None
        return out

    monkeypatch.setattr(du, fn_name, wrapper)
    build_and_project()
    assert len(rec.calls) == 1
    return rec.calls[0]


@pytest.mark.parametrize("periodic", [None, "chains", "selections"])
def test_metricdistance_prologue_with_reference_molecule(refmol, monkeypatch, periodic):
    """MetricDistance(string selections) on the reference Molecule: sel masks, chain ids, selfdist / pbc flags and the
    trajectory arrays that reach dist_trajectory are those of the reference (metricdistance.py:132-179, util.py:12-85)."""
    from moleculekit.projections.metricdistance import MetricDistance as RefMD
    from moleculekit_b200.projections.metricdistance import MetricDistance as OurMD

    sel1, sel2 = "protein and name CA and resid 10 to 40", "resname MOL and noh"
    ref_args = _capture_reference(monkeypatch, "dist_trajectory",
                                  lambda: RefMD(sel1, sel2, periodic=periodic, metric="distances").project(refmol))
    our_args = _capture_ours(monkeypatch, "dist_trajectory",
                             lambda: OurMD(sel1, sel2, periodic=periodic, metric="distances").project(refmol))
    # (coords, box, sel1, sel2, digitized_chains, selfdist, pbc, results)
    for k, (a, b) in enumerate(zip(ref_args[:7], our_args[:7])):
        if isinstance(a, np.ndarray):
            assert a.dtype == b.dtype and np.array_equal(a, b), f"argument {k} differs"
        else:
            assert bool(a) == bool(b), f"argument {k} differs"
    assert ref_args[7].shape == our_args[7].shape and ref_args[7].dtype == our_args[7].dtype


def test_selfdistance_and_mapping_with_reference_molecule(refmol, monkeypatch):
    """MetricSelfDistance + groupsel="residue" + getMapping on the reference Molecule (metricdistance.py:244-364)."""
    from moleculekit.projections.metricdistance import MetricSelfDistance as RefSD
    from moleculekit_b200.projections.metricdistance import MetricSelfDistance as OurSD

    sel = "protein and name CA and resid 5 to 30"
    ref_args = _capture_reference(monkeypatch, "dist_trajectory", lambda: RefSD(sel, periodic=None).project(refmol))
    our_args = _capture_ours(monkeypatch, "dist_trajectory", lambda: OurSD(sel, periodic=None).project(refmol))
    for a, b in zip(ref_args[:7], our_args[:7]):
        if isinstance(a, np.ndarray):
            assert np.array_equal(a, b)
        else:
            assert bool(a) == bool(b)
    rm = RefSD(sel, periodic=None, groupsel="residue").getMapping(refmol)
    om = OurSD(sel, periodic=None, groupsel="residue").getMapping(refmol)
    assert list(rm.columns) == list(om.columns) and len(rm) == len(om)
    assert rm["description"].tolist() == om["description"].tolist()
    assert [list(np.atleast_1d(x)) for x in rm["atomIndexes"]] == [list(np.atleast_1d(x)) for x in om["atomIndexes"]]


def test_reduction_prologue_with_reference_molecule(refmol, monkeypatch):
    """Residue groups against a ligand (get_reduced_distances, util.py:88-223): groups, group chain ids, masses, flags."""
    from moleculekit.projections.metricdistance import MetricDistance as RefMD
    from moleculekit_b200.projections.metricdistance import MetricDistance as OurMD

    kw = dict(periodic="selections", groupsel1="residue", groupsel2="all", metric="contacts", threshold=6)
    sel1, sel2 = "protein and resid 10 to 25 and noh", "resname MOL and noh"
    ref_args = _capture_reference(monkeypatch, "dist_trajectory_reduction", lambda: RefMD(sel1, sel2, **kw).project(refmol))
    our_args = _capture_ours(monkeypatch, "dist_trajectory_reduction", lambda: OurMD(sel1, sel2, **kw).project(refmol))
    # (coords, box, groups1, groups2, chains1, chains2, selfdist, pbc, masses, red1, red2, results)
    assert np.array_equal(ref_args[0], our_args[0]) and np.array_equal(ref_args[1], our_args[1])
    assert [list(g) for g in ref_args[2]] == [list(g) for g in our_args[2]]
    assert [list(g) for g in ref_args[3]] == [list(g) for g in our_args[3]]
    for k in (4, 5, 8):
        assert np.array_equal(np.asarray(ref_args[k]), np.asarray(our_args[k])), k
    for k in (6, 7, 9, 10):
        assert int(ref_args[k]) == int(our_args[k]), k
