"""CPU oracle for the return / advantage hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
leg may import it, and only as the checker / reported baseline.  The product
(``ppo_b200``) never imports this package and fails loudly when its CUDA
library is missing.

Every function is a NumPy restatement of the reference algorithm and cites
the reference ``file:line`` it follows (paths relative to the reference
checkout, dremovd/PPO).  Parity is PINNED: ``tests/golden/make_golden.py``
imports the real reference in the build container and stores seeded
input/output vectors under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks this oracle against them bit-for-bit (float compares are exact unless
the test states a tolerance), together with the reference's own known-answer
tests (``rl/unit_tests.py:203-210``, ``tests/test_tvf.py:121-129``).
"""
from .returns_np import gae, td_lambda, calculate_bootstrapped_returns, normalize_advantages  # noqa: F401
from .tvf_np import (  # noqa: F401
    get_return_estimate,
    sampled_return_multi,
    sampled_return_slow,
    interp_scalar,
    draw_n_step_samples,
    horizon_interpolate,
    get_rediscounted_value_estimate,
    rediscount_horizons,
    get_value_head_horizons,
    trim_horizons,
)
from .vtrace_np import importance_sampling_v_trace, v_trace_trust_region  # noqa: F401
from .normalization_np import (RunningMeanStd, intrinsic_returns, vec_normalize_rewards, return_tracker_update,  # noqa: F401
                               episodic_normalization_constant)
from .train_np import tvf_value_loss  # noqa: F401
