"""Seeded synthetic workloads of SURVEY.md section 8(d) / BASELINE.json `configs`.

Used by bench.py (timed region), the parity tests (same generators at reduced batch) and the CPU baseline leg.
All generators use numpy.random.default_rng(seed); coordinates are float32, sigmas float64 = vdW radius * mask
(the reference's default boolean-channel path, moleculekit/tools/voxeldescriptors.py:332-335).
"""
from __future__ import annotations

import numpy as np

# marginals of the 8 pharmacophoric channels measured on the reference's 3PTB_channels_inp.npy;
# donor raised 0 -> .05 so every channel is exercised; "occupancies" is always on.
CHANNEL_P_PROTEIN = np.array([.62, .08, .34, .05, .57, .39, .001, 1.0])
CHANNEL_P_LIGAND = np.array([.45, .30, .15, .10, .03, .03, 0.0, .55])
RADII_PROTEIN = np.array([1.37, 1.52, 1.55, 1.7, 1.8])
RADII_LIGAND = np.array([1.1, 1.52, 1.55, 1.7, 1.8])


def _sigmas(rng, n, radii, p):
    rad = rng.choice(radii, size=n)
    mask = rng.random((n, 8)) < p
    return rad[:, None] * mask.astype(np.float64)


def _random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def ligand_poses(B: int = 1024, n_atoms: int = 50, seed: int = 0):
    """C2: B poses of one ~50-atom ligand (self-avoiding 1.5 A walk), random rotation + N(0, 2 A) shift;
    24^3 box at 1 A centred on the pose centroid."""
    rng = np.random.default_rng(seed)
    pts = [np.zeros(3)]
    while len(pts) < n_atoms:
        step = rng.normal(size=3)
        cand = pts[-1] + 1.5 * step / np.linalg.norm(step)
        if min(np.linalg.norm(cand - p) for p in pts) > 1.2:
            pts.append(cand)
    conf = np.array(pts)
    conf -= conf.mean(axis=0)
    sig = _sigmas(rng, n_atoms, RADII_LIGAND, CHANNEL_P_LIGAND)
    coords, centers = [], []
    for b in range(B):
        r = np.random.default_rng(seed + 1 + b)
        xyz = (conf @ _random_rotation(r).T + r.normal(0, 2.0, size=3) + 30.0).astype(np.float32)
        coords.append(xyz)
        centers.append(xyz.astype(np.float64).mean(axis=0))
    return dict(coords=coords, sigmas=[sig] * B, centers=np.array(centers), boxsize=[24.0] * 3, voxelsize=1.0,
                name=f"C2: {B} ligand poses x {n_atoms} atoms, 24^3 @1A, 8 ch")


def _sphere_lattice(rng, n_atoms, radius, center, spacing=2.2, jitter=0.3):
    g = np.arange(-radius, radius + spacing, spacing)
    pts = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    pts = pts + rng.normal(0, jitter, size=pts.shape)
    pts = pts[np.linalg.norm(pts, axis=1) <= radius]
    if len(pts) >= n_atoms:
        pts = pts[rng.choice(len(pts), n_atoms, replace=False)]
    else:  # top up with uniform points in the sphere
        extra = rng.normal(size=(n_atoms - len(pts), 3))
        extra *= (radius * rng.random(len(extra)) ** (1 / 3) / np.linalg.norm(extra, axis=1))[:, None]
        pts = np.concatenate([pts, extra])
    return (pts + center).astype(np.float32)


def protein_pockets(B: int = 256, n_atoms: int = 3000, box: float = 64.0, radius: float = 20.0,
                    voxelsize: float = 1.0, seed: int = 1000):
    """C3 (primary): B pockets of ~3000 atoms (jittered 2.2 A lattice clipped to a sphere R=20 A centred in the
    box), 64^3 grid at 1 A, protein channel marginals."""
    coords, sigmas, centers = [], [], []
    for b in range(B):
        rng = np.random.default_rng(seed + b)
        ctr = rng.uniform(-50.0, 50.0, size=3)
        coords.append(_sphere_lattice(rng, n_atoms, radius, ctr))
        sigmas.append(_sigmas(rng, n_atoms, RADII_PROTEIN, CHANNEL_P_PROTEIN))
        centers.append(ctr)
    n = int(np.ceil(box / voxelsize))
    return dict(coords=coords, sigmas=sigmas, centers=np.array(centers), boxsize=[box] * 3, voxelsize=voxelsize,
                name=f"C3: {B} protein pockets x {n_atoms} atoms, {n}^3 @{voxelsize:g}A, 8 ch")


def fine_grids(B: int = 64, n_atoms: int = 8000, seed: int = 2000):
    """C5: whole-protein fine grids, 100 A box at 0.5 A (200^3), sphere R=28 A."""
    w = protein_pockets(B, n_atoms, box=100.0, radius=28.0, voxelsize=0.5, seed=seed)
    w["name"] = f"C5: {B} fine grids x {n_atoms} atoms, 200^3 @0.5A, 8 ch"
    return w


def occupancy_algorithmic_bytes(n_voxels_total: int, n_atoms_total: int, C: int = 8) -> int:
    """bytes = sum_b [ M_b*C*4 (fp32 grid written once) + N_b*(12 + 8*C) (coords f32 + sigmas f64
    read once) ]."""
    return n_voxels_total * C * 4 + n_atoms_total * (12 + 8 * C)


def periodic_trajectory(n_atoms: int = 5000, n_frames: int = 10000, seed: int = 7, L: float = 36.84):
    """C4: solvated-box random walk, (N,3,F) float32 frame-minor + (3,F) box."""
    rng = np.random.default_rng(seed)
    box = (L * (1 + 0.002 * rng.normal(size=n_frames))).astype(np.float32)
    coords = np.empty((n_atoms, 3, n_frames), dtype=np.float32)
    cur = rng.uniform(0, L, size=(n_atoms, 3))
    coords[:, :, 0] = cur
    for f in range(1, n_frames):
        cur = cur + rng.normal(0, 0.3, size=(n_atoms, 3))
        coords[:, :, f] = cur
    return coords, np.repeat(box[None, :], 3, axis=0).copy()
