"""Host-side logic of the mirrors that needs no GPU: rotation matrices, channel bit masks, grid specs, charged-atom tables."""
import numpy as np
import pytest


def test_rotation_matrices_equal_reference(g_rotate):
    """rotation_matrices reproduces moleculekit.util.rotationMatrix bit for bit (same scalar operations)."""
    from moleculekit_b200.tools.voxeldescriptors import rotation_matrices, rotationMatrix

    g = g_rotate
    for c in range(int(g["ncase"])):
        assert np.array_equal(rotation_matrices(g[f"c{c}_rot"])[0], g[f"c{c}_mats"])
    m = rotationMatrix([0, 0, 1], 1.5708)  # the reference's doctest (util.py:90-94)
    assert np.allclose(m.round(4), [[0, -1, 0], [1, 0, 0], [0, 0, 1]])
    assert np.allclose(np.dot(rotationMatrix([4.0, 4.0, 1.0], 1.2), [3.0, 5.0, 0.0]).round(2), [2.75, 4.77, 1.92])


def test_voxelbatch_channel_masks_and_grid_specs():
    from moleculekit_b200.tools import voxeldescriptors as vd

    rng = np.random.default_rng(0)
    coords = [rng.normal(0, 5, (7, 3)).astype(np.float32), rng.normal(9, 3, (4, 3)).astype(np.float32)]
    chans = [rng.random((7, 5)) < 0.5, rng.random((4, 5)) < 0.5]
    els = [np.array(["C", "N", "O", "S", "H", "C", "N"]), np.array(["C", "C", "O", "N"])]
    b = vd.VoxelBatch(coords, chans, boxsize=[6, 8, 10], centers=[c.mean(0) for c in coords], voxelsize=2.0, elements=els)
    assert b.sigmas is None and b.C == 5 and b.chanmask.dtype == np.int32 and b.radii.dtype == np.float64
    cat = np.concatenate(chans)
    for a in range(11):
        assert [(int(b.chanmask[a]) >> h) & 1 for h in range(5)] == cat[a].astype(int).tolist()
    assert np.array_equal(b.radii, np.concatenate([vd.vdw_radii_of(e) for e in els]))
    assert b.dims.tolist() == [[3, 4, 5]] * 2 and b.out_offsets.tolist() == [0, 60, 120] and b.total_voxels == 120
    assert np.allclose(b.origins[1], coords[1].mean(0).astype(np.float64) - np.array([3, 4, 5.0]))
    # float channels keep the sigma-matrix form; bounding-box grids follow getCenters (+1 voxel, float32 bbox)
    sig = [c * 1.7 for c in chans]
    b2 = vd.VoxelBatch(coords, sig, buffer=1.0, voxelsize=1.0)
    assert b2.sigmas is not None and b2.sigmas.dtype == np.float64 and b2.radii is None
    for i in range(2):
        bb_min = coords[i].min(0) - np.float32(1.0)
        bb_max = coords[i].max(0) + np.float32(1.0)
        assert b2.dims[i].tolist() == (np.ceil((bb_max - bb_min) / 1.0).astype(int) + 1).tolist()
        assert np.array_equal(b2.origins[i], bb_min.astype(np.float64))
    assert vd.VoxelBatch(coords, chans, boxsize=[4, 4, 4], centers=np.zeros((2, 3)), radii=[np.ones(7), np.ones(4)]).sigmas is None


def test_charged_atoms_and_metal_tables(g_interactions):
    from moleculekit_b200 import interactions as it
    from moleculekit_b200.molecule_lite import MolLite

    g = g_interactions
    mol = MolLite(g["me6_coords"], resname=g["me6_resname"], name=g["me6_name"], element=g["me6_element"])
    pos, neg = it.get_protein_charged(mol)
    assert np.array_equal(pos, g["me6_pos"]) and np.array_equal(neg, g["me6_neg"])
    assert len(it.METAL_ELEMENTS) == 84 and {"Zn", "Fe", "Ca", "Na"} <= it.METAL_ELEMENTS and "C" not in it.METAL_ELEMENTS
    m = it._mask(mol, np.array([1, 5]))
    assert m.dtype == bool and m.sum() == 2 and m[1] and m[5]


def test_wrap_argument_errors_without_gpu():
    from moleculekit_b200 import wrapping as wr
    from moleculekit_b200.molecule_lite import MolLite

    mol = MolLite(np.zeros((4, 3, 2), np.float32), box=np.ones((3, 2), np.float32))
    with pytest.raises(ValueError, match="Invalid unit cell type"):
        wr.wrap(mol, unitcell="hexagonal")
    mol.box = np.zeros((3, 2), np.float32)
    assert wr.wrap(mol) is None  # zero box: logged no-op, never reaches the GPU
    with pytest.raises(ValueError, match="Buffer dtype mismatch"):
        wr.wrap_box(np.array([0, 4]), mol.coords, np.ones((3, 2), np.float32), np.zeros(0, np.uint32), np.zeros(3, np.float32))
