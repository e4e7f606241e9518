"""GPU parity of K11 (XTC decoding on the device): bit-identical to what the reference's read_xtc returned for the committed
files, frame subsets, the nm -> Angstrom scaling of XTCread, corrupt input."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_reference_goldens(g_xtc):
    from moleculekit_b200 import xtc as px

    g = g_xtc
    for name in g["names"].tolist():
        fn = os.path.join(g["_dir"], name + ".xtc")
        coords, box, time, step = px.read_xtc(fn.encode("UTF-8"))  # bytes file name like the reference's call
        assert coords.dtype == np.float32 and coords.shape == g[f"{name}_coords"].shape
        assert np.array_equal(coords.view(np.uint32), g[f"{name}_coords"].view(np.uint32)), name
        assert np.array_equal(box, g[f"{name}_box"]) and np.array_equal(time, g[f"{name}_time"])
        assert np.array_equal(step, g[f"{name}_step"]) and step.dtype == np.int32


def test_frame_subsets_scaling_and_device_layout(g_xtc):
    import torch
    from moleculekit_b200 import xtc as px

    g = g_xtc
    fn = os.path.join(g["_dir"], "water.xtc")
    want = g["water_coords"]
    c, box, time, step = px.read_xtc_frames(fn, np.array([4, 0, 4, 2], dtype=np.int32))
    assert np.array_equal(c, want[:, :, [4, 0, 4, 2]]) and np.array_equal(step, g["water_step"][[4, 0, 4, 2]])
    assert box.shape == (3, 3, 4) and np.array_equal(box, g["water_box"][:, :, [4, 0, 4, 2]])
    d, _, _, _ = px.read_xtc_device(fn, scale=10.0)
    assert d.is_cuda and d.is_contiguous() and tuple(d.shape) == want.shape
    ang = want.copy(); ang *= 10.0  # readers.py:1846: a second float32 multiplication
    assert np.array_equal(d.cpu().numpy().view(np.uint32), ang.view(np.uint32))
    # the decoded trajectory feeds the distance kernels without leaving the device
    from moleculekit_b200 import distance_utils as du

    bx = torch.zeros((3, d.shape[2]), dtype=torch.float32, device=d.device)
    s1 = torch.arange(0, 3, dtype=torch.int32, device=d.device); s2 = torch.arange(3, 9, dtype=torch.int32, device=d.device)
    ch = torch.zeros(d.shape[0], dtype=torch.int32, device=d.device)
    dist = du.dist_trajectory_device(d, bx, s1, s2, ch, False, False)
    ref = np.linalg.norm(ang[0:3, None, :, :].astype(np.float64) - ang[None, 3:9, :, :], axis=2).reshape(18, -1).T
    assert np.allclose(dist.cpu().numpy(), ref, rtol=1e-6)


def test_corrupt_blocks_are_reported(g_xtc):
    import torch
    from moleculekit_b200 import xtc as px

    fn = os.path.join(g_xtc["_dir"], "globule.xtc")
    raw = bytearray(open(fn, "rb").read())
    idx = px.index_xtc(bytes(raw))
    dev = torch.device("cuda:0")
    frames = idx["frames"].copy()
    frames["smallidx"][1] = 3  # not a usable table entry
    with pytest.raises(RuntimeError, match="frame 1 could not be decoded"):
        px.decode_xtc_device(torch.frombuffer(raw, dtype=torch.uint8).to(dev), frames, idx["natoms"])
    frames = idx["frames"].copy()
    frames["data_offset"][0] = len(raw) - 4  # block runs past the end of the file
    with pytest.raises(RuntimeError, match="frame 0"):
        px.decode_xtc_device(torch.frombuffer(raw, dtype=torch.uint8).to(dev), frames, idx["natoms"])
    with pytest.raises(RuntimeError, match="truncated"):
        px.index_xtc(bytes(raw[:200]))
