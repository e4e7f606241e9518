"""Host-side check of the ARITHMETIC of the v10 fill kernel's gate (csrc/occ_runs.cuh).

The kernel decides `d2 < cut2` by float32 overflow: differences are scaled by lambda = 2^64 / cut, so U = dx^2 + dy^2 + dz^2 rounds to
+inf exactly when the pair is outside the gate, and r = U * w with w = 1 / (sigma lambda)^2.  This test restates those float32
operations in numpy (fused multiply-adds through float64, exact for float32 operands) and checks, on pairs placed within 1e-4 of the
gate, that the overflow decision differs from the float64 decision only inside the band the kernel's float64 fix-up re-evaluates
(R_FIND_BAND = 2e-6 relative) and that r is good to 1e-6 -- the two facts the parity argument rests on.  No GPU, no oracle."""
import numpy as np
import pytest

f32 = np.float32


def _fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


@pytest.mark.parametrize("cutv", [5.0, 10.0, 5.0 / 0.7, 1.0, 50.0])
def test_overflow_gate_matches_float64_outside_the_band(cutv):
    rng = np.random.default_rng(int(cutv * 1000))
    N = 400_000
    cut2 = f32(cutv * cutv)  # cut2v as the kernel receives it
    lamf = f32(2.0 ** 64 / np.sqrt(np.float64(cut2)))  # the root value (float), its exact double image, and cut up to 6e-8
    lam = np.float64(lamf)
    cwf = f32(2.0 ** 64 / lam)
    # block frame: voxel x_k in {-1.5 .. 1.5}, lane offsets fy in {-1.5 .. 1.5}, fz in {-3.5 .. 3.5}
    xk = (rng.integers(0, 4, N) - 1.5).astype(f32)
    fy = (rng.integers(0, 4, N) - 1.5).astype(f32)
    fz = (rng.integers(0, 8, N) - 3.5).astype(f32)
    delta = rng.choice([-1.0, 1.0], N) * 10.0 ** rng.uniform(-7.5, -4.0, N)
    d = np.sqrt(np.float64(cut2) * (1.0 + delta))
    v = rng.normal(size=(N, 3))
    v /= np.linalg.norm(v, axis=1)[:, None]
    ex, ey, ez = xk + v[:, 0] * d, fy + v[:, 1] * d, fz + v[:, 2] * d  # atom position (float64) in the block frame
    d2 = (ex - xk) ** 2 + (ey - fy) ** 2 + (ez - fz) ** 2
    truth = d2 < np.float64(cut2)
    xs, ys, zs = (ex * lam).astype(f32), (ey * lam).astype(f32), (ez * lam).astype(f32)  # the record pass: one rounding each
    L = np.full(N, lamf)
    D = _fma(xk, L, -xs)
    dys, dzs = _fma(fy, L, -ys), _fma(fz, L, -zs)
    with np.errstate(over="ignore"):
        s2 = (dys * dys + dzs * dzs).astype(f32)
        U = _fma(D, D, s2)
    gate = np.isfinite(U)
    rel = np.abs(d2 / np.float64(cut2) - 1.0)
    bad = gate != truth
    assert not bad.any() or rel[bad].max() < 1.0e-6, "overflow gate disagrees with float64 outside half the fix-up band"
    # the value path: r = U * w, w = ((cut / sigma) 2^-64)^2 -- sigma between 1 and 2.5 A at this voxel size
    sig = rng.uniform(1.0, 2.5, N) / (5.0 / cutv)
    sw = f32(1.0) / sig.astype(f32)
    wh = (sw * cwf) * f32(2.0 ** -64)
    r = U * (wh * wh)
    ok = gate & truth
    err = np.abs(r[ok].astype(np.float64) * (sig[ok].astype(f32).astype(np.float64) ** 2) / d2[ok] - 1.0)
    assert err.max() < 1.5e-6


def test_overflow_never_yields_nan():
    """inf * w and inf + inf stay inf; no inf - inf or 0 * inf can occur (w > 0, squares only)."""
    lamf = f32(2.0 ** 64 / 5.0)
    far = f32(13.5) * lamf  # the farthest candidate coordinate of a block list at 0.5 A voxels
    with np.errstate(over="ignore"):
        s2 = far * far + far * far
        U = _fma(np.array([far]), np.array([far]), np.array([s2]))
        r = U * f32(1e-38)
    assert np.isinf(U[0]) and np.isinf(r[0]) and not np.isnan(r[0])
