"""ppo_b200 -- B200-native return / advantage hot path for the PPO/TVF trainer dremovd/PPO.

Drop-in modules (same function names and signatures as the reference's ``rl`` package):

* :mod:`ppo_b200.returns`            <- ``rl/returns.py``            (``gae``, ``td_lambda``, ``calculate_bootstrapped_returns``)
* :mod:`ppo_b200.returns_truncated`  <- ``rl/returns_truncated.py``  (``get_return_estimate``, ``_calculate_sampled_return_multi_fast``)
* :mod:`ppo_b200.tvf`                <- ``rl/tvf.py``                (``get_value_head_horizons``, ``TVFRunnerModule`` return-side methods)
* :mod:`ppo_b200.vtrace`             <- ``rl/vtrace.py``             (``importance_sampling_v_trace``, ``v_trace_trust_region``)
* :mod:`ppo_b200.rollout`            <- the return stage of ``rl/rollout.py`` (``calculate_returns``, ``normalize_advantages``)
* :mod:`ppo_b200.normalization`      <- ``Runner.calculate_intrinsic_returns`` (``rl/rollout.py:1140-1180``), ``VecNormalizeRewardWrapper``
  (``rl/wrappers.py:795-919``) and ``RunningMeanStd`` (``rl/utils.py:414-443``) over whole rollouts

All arithmetic on rollout data runs in hand-written sm_100a CUDA kernels behind the C ABI of
``include/ppo_b200.h`` (``ppo_b200/csrc``).  There is no CPU fallback: without the built library or
without a CUDA device every call raises.
"""
__version__ = "0.1.0"

from ._device import pin_numpy  # noqa: E402,F401  (page-lock persistent NumPy rollout buffers once)
