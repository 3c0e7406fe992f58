"""Host-side pieces of ppo_b200.normalization that need no GPU: the episodic-discounting normalisation constants
(rl/wrappers.py:52-72) against the oracle's restatement of the reference."""
import numpy as np
import pytest
import torch

import oracle
from ppo_b200 import normalization as nz


@pytest.mark.parametrize("kind", ["quadratic", "power", "harmonic", "none"])
@pytest.mark.parametrize("bias", [1.0, 2.5])
def test_episodic_normalization_constant_matches_the_oracle(kind, bias):
    times = np.arange(0, 500, 7)
    with np.errstate(divide="ignore"):
        ref = np.broadcast_to(oracle.episodic_normalization_constant(times, kind, discount_bias=bias), times.shape)
    got = nz.episodic_normalization_constant(torch.from_numpy(times), kind, discount_bias=bias).numpy()
    assert got.dtype == np.float64 and got.shape == times.shape
    finite = np.isfinite(ref)
    assert np.array_equal(np.isinf(got), ~finite)  # harmonic at k = 1: 1 / log(1)
    assert np.allclose(got[finite], ref[finite], rtol=1e-14, atol=0)


def test_episodic_normalization_constant_errors_like_the_reference():
    t = torch.arange(4)
    with pytest.raises(ZeroDivisionError):  # gamma = 1.0 as the wrapper calls it: 1 / (1 - gamma) in plain Python
        nz.episodic_normalization_constant(t, "finite")
    assert bool(torch.isinf(nz.episodic_normalization_constant(t, "geometric")).all())
    assert torch.equal(nz.episodic_normalization_constant(t, "finite", gamma=0.5),
                       torch.from_numpy(oracle.episodic_normalization_constant(np.arange(4), "finite", gamma=0.5)))
    with pytest.raises(ValueError, match="Invalid discount_type"):
        nz.episodic_normalization_constant(t, "linear")
