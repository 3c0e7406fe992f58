"""Device plumbing shared by the drop-in modules: torch is used for memory, streams and
``torch.distributed`` only; all arithmetic happens in the C-ABI kernels."""
import numpy as np
import torch

from . import _lib

_workspaces = {}


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("ppo_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def is_numpy_like(x):
    return not isinstance(x, torch.Tensor)


def to_device(x, dtype=None, device=None):
    """Tensor on the compute device.  NumPy / scalars / CPU tensors are copied, CUDA tensors are used in place."""
    device = device or require_cuda()
    if isinstance(x, torch.Tensor):
        t = x if x.device == device else x.to(device, non_blocking=True)
    else:
        arr = np.asarray(x)
        if not arr.flags.writeable:
            arr = arr.copy()
        t = torch.from_numpy(np.ascontiguousarray(arr)).to(device, non_blocking=True)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t


def done_tensor(dones, device=None):
    """(tensor, done_kind): bool/uint8 stay one byte per element, everything else becomes a float32 mask."""
    if dones is None:
        return None, _lib.DONE_NONE
    t = to_device(dones, device=device)
    if t.dtype == torch.bool:
        return t.view(torch.uint8), _lib.DONE_U8
    if t.dtype == torch.uint8:
        return t, _lib.DONE_U8
    return t.to(torch.float32), _lib.DONE_F32


def env_contiguous(t):
    """Make the last axis unit-stride (rows may stay strided)."""
    return t if t.dim() == 0 or t.stride(-1) == 1 or t.shape[-1] == 1 else t.contiguous()


def ptr(t):
    return 0 if t is None else t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def workspace(nbytes, device):
    """Zero-initialised scratch for the statistics tickets/partials, cached per device and stream."""
    need = max(int(nbytes), 1 << 16)
    key = (device.index, stream_ptr())
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(need, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


_registered = {}


def pin_numpy(array):
    """Page-lock the memory of a NumPy rollout buffer in place (``cudaHostRegister``) and return it unchanged.

    The reference keeps its rollout in persistent NumPy arrays (``rl/rollout.py:646-680``, ``rl/tvf.py:27-30``);
    registering them once lets every later call move them with asynchronous, full-rate PCIe copies (and the
    env-blocked pipeline of ``returns_truncated``) instead of pageable staging.  The registration is dropped
    when the array's memory owner is garbage-collected.  Only C-contiguous arrays that own their memory (or
    views whose base does) make sense here; call it on the base buffers at start-up."""
    import weakref
    require_cuda()
    owner = array
    while isinstance(getattr(owner, "base", None), np.ndarray):
        owner = owner.base
    ptr, nbytes = owner.ctypes.data, owner.nbytes
    if nbytes == 0 or ptr in _registered:
        return array
    rt = torch.cuda.cudart()
    err = rt.cudaHostRegister(ptr, nbytes, 0)
    if int(err) != 0:
        raise RuntimeError(f"cudaHostRegister failed with error {int(err)} for {nbytes} bytes at {ptr:#x}")
    _registered[ptr] = nbytes

    def release(p=ptr):
        if _registered.pop(p, None) is not None:
            try:
                torch.cuda.cudart().cudaHostUnregister(p)
            except Exception:
                pass
    weakref.finalize(owner, release)
    return array


def pinned_tensor_of(x):
    """CPU tensor view of ``x`` when its memory is page-locked (a pinned tensor, or a NumPy array registered with
    :func:`pin_numpy`), else None."""
    if isinstance(x, torch.Tensor):
        return x if x.device.type == "cpu" and x.is_pinned() else None
    if isinstance(x, np.ndarray) and x.flags.writeable and torch.cuda.is_available():
        owner = x
        while isinstance(getattr(owner, "base", None), np.ndarray):
            owner = owner.base
        if owner.ctypes.data in _registered:
            t = torch.from_numpy(x)
            return t if t.is_pinned() else None
    return None


def like_input(result, wants_numpy):
    """Reference semantics: NumPy in -> NumPy out; CUDA tensor in -> CUDA tensor out."""
    if wants_numpy:
        return result.cpu().numpy()
    return result
