"""Contact-type interaction detectors on top of the GPU contact kernel (K4) -- SURVEY 8f row 2, second half.

Mirrors of the three detectors in the reference that are thin consumers of ``calculate_contacts``
(moleculekit/interactions/interactions.py):

  saltbridge_calculate            :724-788     charged atoms within `threshold`, one positive + one negative per pair
  hydrophobic_calculate           :949-992     carbon - carbon contacts
  metal_coordination_calculate    :995-1056    metal - (N, O, S, halogen) contacts, both directions
  get_protein_charged / get_metal_charged      :292-306  (pure table look-ups on resname / name / element)

Selections are boolean masks / index arrays, or strings resolved by ``mol.atomselect``; the pair search itself is
``mkb_contacts_count`` + ``mkb_contacts_fill`` (bit-exact index output, reference order).  The ring / angle based detectors
(pi-pi, cation-pi, sigma holes, hydrogen bonds) are a different algorithm family and stay with moleculekit.
"""
from __future__ import annotations

import numpy as np

from .distance import calculate_contacts

# the element set moleculekit.periodictable.METAL_ELEMENTS (periodictable.py:317-407)
METAL_ELEMENTS = frozenset(
    "Ac Ag Al Am Au Ba Be Bi Bk Ca Cd Ce Cf Cm Co Cr Cs Cu Db Dy Er Es Eu Fe Fm Fr Ga Gd Ge Hf Hg Ho In Ir K La Li Lr Lu "
    "Md Mg Mn Mo Na Nb Nd Ni No Np Os Pa Pb Pd Pm Po Pr Pt Pu Ra Rb Re Rf Rh Ru Sb Sc Sg Sm Sn Sr Ta Tb Tc Th Ti Tl Tm U V "
    "W Y Yb Zn Zr".split())
COORDINATING_ELEMENTS = ["N", "O", "Cl", "F", "Br", "I", "CL", "BR", "S"]  # interactions.py:1034


def _mask(mol, sel) -> np.ndarray:
    """str -> mol.atomselect; bool mask as is; integer index array -> mask (what Molecule.atomselect accepts)."""
    if isinstance(sel, str):
        return np.asarray(mol.atomselect(sel), dtype=bool)
    sel = np.asarray(sel)
    if sel.dtype == bool:
        return sel.copy()
    m = np.zeros(int(mol.numAtoms), dtype=bool)
    m[sel] = True
    return m


def _periodic(mol):
    return "selections" if not np.all(mol.box == 0) else None


def get_protein_charged(mol):
    """interactions.py:296-306: (positive, negative) uint32 index arrays from residue / atom names."""
    resname, name = np.asarray(mol.resname), np.asarray(mol.name)
    pos = ((resname == "LYS") & (name == "NZ")) | ((resname == "ARG") & (name == "CZ")) | \
          ((resname == "HIP") & (name == "CE1"))
    neg = ((resname == "ASP") & (name == "CG")) | ((resname == "GLU") & (name == "CD"))
    return np.where(pos)[0].astype(np.uint32), np.where(neg)[0].astype(np.uint32)


def saltbridge_calculate(mol, pos, neg, sel1="all", sel2=None, threshold: float = 4, device=None):
    """interactions.py:724-788: per frame the (n, 2) pairs of charged atoms within ``threshold`` holding exactly one
    positive atom."""
    if len(pos) == 0 or len(neg) == 0:
        return [[] for _ in range(mol.numFrames)]
    m1 = _mask(mol, sel1)
    m2 = m1.copy() if sel2 is None else _mask(mol, sel2)
    charged = np.zeros(m1.shape, dtype=bool)
    charged[pos] = True
    charged[neg] = True
    inter = calculate_contacts(mol, m1 & charged, m2 & charged, _periodic(mol), threshold, device=device)
    return [fr[np.sum(np.isin(fr, pos), axis=1) == 1] for fr in inter]


def hydrophobic_calculate(mol, sel1, sel2, dist_threshold: float = 4.0, device=None):
    """interactions.py:949-992: carbon - carbon contacts between the two selections."""
    carbons = np.asarray(mol.element) == "C"
    return calculate_contacts(mol, _mask(mol, sel1) & carbons, _mask(mol, sel2) & carbons, _periodic(mol), dist_threshold,
                              device=device)


def metal_coordination_calculate(mol, sel1, sel2, dist_threshold: float = 3.5, device=None):
    """interactions.py:995-1056: metals of sel1 against coordinating atoms of sel2, then the other way round, stacked."""
    metals = sorted(METAL_ELEMENTS)
    m1, m2 = _mask(mol, sel1), _mask(mol, sel2)
    element = np.asarray(mol.element)
    is_metal, is_coord = np.isin(element, metals), np.isin(element, COORDINATING_ELEMENTS)
    periodic = _periodic(mol)
    inter1 = calculate_contacts(mol, m1 & is_metal, m2 & is_coord, periodic, dist_threshold, device=device)
    inter2 = calculate_contacts(mol, m1 & is_coord, m2 & is_metal, periodic, dist_threshold, device=device)
    return [np.vstack((inter1[f], inter2[f])) for f in range(mol.numFrames)]
