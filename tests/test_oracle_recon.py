"""CPU: the back-projection oracle (oracle/recon_np.py) against the golden vectors produced by the reference's own
kernels (tests/golden/recon_golden.npz, generator tests/golden/make_recon_golden.py) and, where the host build of those
kernels is available (oracle/_ref/librec_ref.so, built from /root/reference in this container), against it directly on
further seeded inputs.  fp32 with one rounding per operation on both sides: bit-exact."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import build_ref as R, recon_np as O

sys.path.insert(0, GOLDEN)
from make_recon_golden import CASES, case_inputs  # noqa: E402


@pytest.fixture(scope="module")
def rg():
    return dict(np.load(os.path.join(GOLDEN, "recon_golden.npz")))


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_kernels_golden(rg, name):
    seed, ntheta, nz, n, center = CASES[name]
    data, theta = case_inputs(seed, ntheta, nz, n)
    np.testing.assert_array_equal(O.rec_all(data, theta, center), rg[name + "_rec_all"])
    np.testing.assert_array_equal(O.rec_mask(rg[name + "_mask"], data, theta, center), rg[name + "_rec_mask"])
    np.testing.assert_array_equal(O.rec_patch(data, theta, center, *rg[name + "_box"]), rg[name + "_rec_patch"])
    np.testing.assert_array_equal(O.rec_pts(data, theta, center, rg[name + "_pts"]), rg[name + "_rec_pts"])
    np.testing.assert_array_equal(O.rec_pts_xy(data, theta, center, rg[name + "_pts"][:, 1:]), rg[name + "_rec_pts_xy"])
    # rec_mask keeps only voxels whose mask value is > 0 (negative entries are dropped too)
    assert np.all(rg[name + "_rec_mask"][rg[name + "_mask"] <= 0] == 0)


@pytest.mark.skipif(not (R.built() or R.reference_present()), reason="reference kernels not built and /root/reference absent")
def test_oracle_matches_host_build_of_reference_kernels():
    g = np.random.default_rng(7)
    for ntheta, nz, n, center in ((19, 3, 33, 16.0), (50, 7, 48, 24.5), (11, 18, 24, 9.75)):
        data = g.standard_normal((ntheta, nz, n)).astype(np.float32)
        theta = g.uniform(-1.0, 4.0, ntheta).astype(np.float32)
        np.testing.assert_array_equal(O.rec_all(data, theta, center), R.rec_all(data, theta, center))
        pts = np.stack([g.integers(0, nz, 64), g.integers(0, n, 64), g.integers(0, n, 64)], 1).astype(np.int32)
        np.testing.assert_array_equal(O.rec_pts(data, theta, center, pts), R.rec_pts(data, theta, center, pts))
        np.testing.assert_array_equal(O.rec_patch(data, theta, center, 2, n - 5, 1, n - 3, 1, nz - 1),
                                      R.rec_patch(data, theta, center, 2, n - 5, 1, n - 3, 1, nz - 1))


def test_calc_padding_known_answers():
    """prep.py:21-31: 1.5 n rounded up to a multiple of 8, split left / right."""
    assert O.calc_padding((1, 1, 2048)) == (512, 512)
    assert O.calc_padding((1, 1, 1500)) == (378, 378)       # 2250 -> 2256
    assert O.calc_padding((1, 1, 72)) == (20, 20)             # 108 -> 112
    assert O.calc_padding((1, 1, 33)) == (11, 12)           # 49.5 -> 56


def test_fbp_filter_is_a_ramp():
    """|f| filter: removes the mean of a row, scales a pure harmonic of the padded period by its frequency."""
    n = 64
    pl, pr = O.calc_padding((1, 1, n))
    n_pad = n + pl + pr
    const = np.full((2, 3, n), 3.25, np.float32)
    np.testing.assert_allclose(O.fbp_filter(const.copy()), 0.0, atol=1e-6)
    g = np.random.default_rng(0)
    d = g.standard_normal((2, 3, n)).astype(np.float32)
    ref = d.astype(np.float64)
    padded = np.pad(ref, ((0, 0), (0, 0), (pl, pr)), mode="edge")
    want = np.fft.irfft(np.fft.rfftfreq(n_pad) * np.fft.rfft(padded, axis=2), n=n_pad, axis=2)[..., pl:-pr]
    np.testing.assert_allclose(O.fbp_filter(d.copy()), want, atol=2e-6)


def test_preprocess_and_masks():
    g = np.random.default_rng(1)
    data = g.uniform(100, 1000, (4, 3, 8)).astype(np.float32)
    dark = g.uniform(0, 50, (3, 8)).astype(np.float32)
    flat = g.uniform(900, 1100, (3, 8)).astype(np.float32)
    flat[0, 0] = dark[0, 0]            # flat == dark: denominator clamps to 1e-6
    data[0, 1, 1] = dark[1, 1] - 5     # negative transmission clamps to 1e-6 before the log
    out = O.preprocess(data.copy(), dark, flat)
    want = -np.log(np.maximum((data - dark) / np.maximum(flat - dark, np.float32(1e-6)), np.float32(1e-6)))
    np.testing.assert_array_equal(out, want)
    assert np.isfinite(out).all() and out[0, 1, 1] == -np.log(np.float32(1e-6))
    m = O.make_mask((4, 8, 8), [[0, 0, 0], [0, 4, 4]], 4)
    assert m.sum() == 2 * 64 and m[0, 0, 0] == 1 and m[0, 0, 4] == 0
