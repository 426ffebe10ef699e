"""Feature front-end tail (started): splice + subsample on the device against the REAL reference functions
(goldens from tests/golden/make_golden_feature.py).  Pure data movement: compared for exact equality."""
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "feature_golden.npz"))
CASES = {"feat_T1003_F23": (1003, 23, 7, 10, 0), "feat_T40_F5_c2_s3": (40, 5, 2, 3, 1), "feat_T7_F23": (7, 23, 7, 10, 2),
         "feat_T500_F23_s1": (500, 23, 7, 1, 3)}


def dropin():
    """The device helper lives in the fseend_b200 package (NOT under a `datasets` package: that name belongs to the
    reference's namespace directory and must keep resolving there, see tests/test_dropin_boundary.py)."""
    from fseend_b200 import feature
    return feature


def restate(y, ctx, sub):
    """numpy restatement (feature.py:103-133): zero-pad ctx frames on both sides, concatenate 2 ctx + 1 frames, stride."""
    T, F = y.shape
    pad = np.concatenate([np.zeros((ctx, F), y.dtype), y, np.zeros((ctx, F), y.dtype)])
    return np.stack([pad[t:t + 2 * ctx + 1].reshape(-1) for t in range(0, T, sub)])


@pytest.mark.parametrize("name", list(CASES))
def test_restatement_matches_reference(name):
    T, F, ctx, sub, seed = CASES[name]
    y = np.random.default_rng(seed).standard_normal((T, F)).astype(np.float32)
    assert np.array_equal(restate(y, ctx, sub), GOLD[name])


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_gpu_splice_subsample_bit_exact(name):
    T, F, ctx, sub, seed = CASES[name]
    y = np.random.default_rng(seed).standard_normal((T, F)).astype(np.float32)
    out = dropin().splice_subsample(torch.from_numpy(y), ctx, sub)
    assert out.dtype == torch.float32 and tuple(out.shape) == GOLD[name].shape
    assert np.array_equal(out.cpu().numpy(), GOLD[name])


@pytest.mark.gpu
def test_gpu_splice_subsample_one_hour():
    """One hour of 10-ms log-mel frames (360 000 x 23) -> 36 000 x 345 model inputs."""
    y = np.random.default_rng(9).standard_normal((360000, 23)).astype(np.float32)
    out = dropin().splice_subsample(torch.from_numpy(y), 7, 10)
    assert tuple(out.shape) == (36000, 345)
    assert np.array_equal(out.cpu().numpy(), restate(y, 7, 10))
