"""Per-element atomic mass (u) and van der Waals radius (Angstrom).

Physical constants needed by the hot path only: the vdW radius turns boolean channel masks into
per-channel sigmas (reference behaviour: moleculekit/tools/voxeldescriptors.py:117-121,332-335)
and the mass weights centre-of-mass group reductions (moleculekit/projections/util.py:180).
Values are the ones the reference's ``periodictable`` carries (moleculekit/periodictable.py:25-27,
sourced from the `mendeleev` package); elements without a tabulated radius map to ``None``.
"""
from __future__ import annotations

import numpy as np

# symbol -> (mass, vdw_radius)
ELEMENT_DATA = {
    "H": (1.00794, 1.1), "He": (4.002602, 1.4), "Li": (6.941, 1.82), "Be": (9.012182, 1.53),
    "B": (10.811, 1.92), "C": (12.0107, 1.7), "N": (14.0067, 1.55), "O": (15.9994, 1.52),
    "F": (18.9984032, 1.47), "Ne": (20.1797, 1.54), "Na": (22.98977, 2.27), "Mg": (24.305, 1.73),
    "Al": (26.981538, 1.84), "Si": (28.0855, 2.1), "P": (30.973761, 1.8), "S": (32.065, 1.8),
    "Cl": (35.453, 1.75), "Ar": (39.948, 1.88), "K": (39.0983, 2.75), "Ca": (40.078, 2.31),
    "Sc": (44.95591, 2.15), "Ti": (47.867, 2.11), "V": (50.9415, 2.07), "Cr": (51.9961, 2.06),
    "Mn": (54.938049, 2.05), "Fe": (55.845, 2.04), "Co": (58.9332, 2.0), "Ni": (58.6934, 1.97),
    "Cu": (63.546, 1.96), "Zn": (65.409, 2.01), "Ga": (69.723, 1.87), "Ge": (72.64, 2.11),
    "As": (74.9216, 1.85), "Se": (78.96, 1.9), "Br": (79.904, 1.85), "Kr": (83.798, 2.02),
    "Rb": (85.4678, 3.03), "Sr": (87.62, 2.49), "Y": (88.90585, 2.32), "Zr": (91.224, 2.23),
    "Nb": (92.90638, 2.18), "Mo": (95.94, 2.17), "Tc": (98, 2.16), "Ru": (101.07, 2.13),
    "Rh": (102.9055, 2.1), "Pd": (106.42, 2.1), "Ag": (107.8682, 2.11), "Cd": (112.411, 2.18),
    "In": (114.818, 1.93), "Sn": (118.71, 2.17), "Sb": (121.76, 2.06), "Te": (127.6, 2.06),
    "I": (126.90447, 1.98), "Xe": (131.293, 2.16), "Cs": (132.90545, 3.43), "Ba": (137.327, 2.68),
    "La": (138.9055, 2.43), "Ce": (140.116, 2.42), "Pr": (140.90765, 2.4), "Nd": (144.24, 2.39),
    "Pm": (145, 2.38), "Sm": (150.36, 2.36), "Eu": (151.964, 2.35), "Gd": (157.25, 2.34),
    "Tb": (158.92534, 2.33), "Dy": (162.5, 2.31), "Ho": (164.93032, 2.3), "Er": (167.259, 2.29),
    "Tm": (168.93421, 2.27), "Yb": (173.04, 2.26), "Lu": (174.967, 2.24), "Hf": (178.49, 2.23),
    "Ta": (180.9479, 2.22), "W": (183.84, 2.18), "Re": (186.207, 2.16), "Os": (190.23, 2.16),
    "Ir": (192.217, 2.13), "Pt": (195.078, 2.13), "Au": (196.96655, 2.14), "Hg": (200.59, 2.23),
    "Tl": (204.3833, 1.96), "Pb": (207.2, 2.02), "Bi": (208.98038, 2.07), "Po": (209, 1.97),
    "At": (210, 2.02), "Rn": (222, 2.2), "Fr": (223, 3.48), "Ra": (226, 2.83),
    "Ac": (227, 2.47), "Th": (232.0381, 2.45), "Pa": (231.03588, 2.43), "U": (238.02891, 2.41),
    "Np": (237, 2.39), "Pu": (244, 2.43), "Am": (243, 2.44), "Cm": (247, 2.45),
    "Bk": (247, 2.44), "Cf": (251, 2.45), "Es": (252, 2.45), "Fm": (257, 2.45),
    "Md": (258, 2.46), "No": (259, 2.46), "Lr": (262, 2.46), "Rf": (261, None),
    "Db": (262, None), "Sg": (266, None), "Bh": (264, None), "Hs": (277, None),
    "Mt": (268, None), "Ds": (281, None), "Rg": (272, None), "Cn": (285, None),
    "Nh": (286, None), "Fl": (289, None), "Mc": (289, None), "Lv": (293, None),
    "Ts": (294, None), "Og": (294, None),
}


def masses_of(elements) -> np.ndarray:
    """float32 masses for a sequence of element symbols (KeyError on unknown symbols, like the reference)."""
    return np.array([ELEMENT_DATA[e][0] for e in elements], dtype=np.float32)


def vdw_radii_of(elements) -> np.ndarray:
    """float64 van der Waals radii for a sequence of element symbols."""
    return np.array([ELEMENT_DATA[e][1] for e in elements])
