"""CPU: the C-ABI library builds, loads and exports every symbol include/ppo_b200.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ppo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ppo_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(built_library):
    lib = ctypes.CDLL(built_library)
    names = declared_symbols()
    assert len(names) >= 8
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/ppo_b200.h but not exported"


def test_binding_matches_header(built_library):
    from ppo_b200 import _lib
    assert sorted(_lib.symbols()) == declared_symbols()
    lib = _lib.load()
    assert lib.ppo_version() >= 100
    assert lib.ppo_gae_workspace_bytes(4096) > 0


def test_bad_arguments_are_rejected_without_a_gpu(built_library):
    from ppo_b200 import _lib
    lib = _lib.load()
    rc = lib.ppo_gae(0, 0, 0, 0, 0, 0, 0, 0, 7, 0, 4, 4, 0.9, 0.9, 0.9, 0, 0, 0, 0, 0, 0)
    assert rc == _lib.PPO_ERR_BAD_ARG
    assert b"done_kind" in lib.ppo_last_error()
    rc = lib.ppo_gae(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, -1, 4, 0.9, 0.9, 0.9, 0, 0, 0, 0, 0, 0)
    assert rc == _lib.PPO_ERR_BAD_ARG


def test_no_cpu_fallback():
    import pytest
    import torch
    from ppo_b200 import returns, synthetic
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    r, d, v = synthetic.rollout(4, 4)
    with pytest.raises(RuntimeError):
        returns.gae(r, v[:-1], v[-1], d, 0.99)
