"""PCIe probe (development aid): pinned H2D / D2H rates alone and concurrently, whole-buffer and env-blocked 2D."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ppo_b200 import _lib

def main():
    lib = _lib.load()
    N, A, K = 257, 4096, 128
    h_in = torch.empty((N, A, K), dtype=torch.float32).pin_memory()
    h_out = torch.empty((N, A, K), dtype=torch.float32).pin_memory()
    d_a = torch.empty((N, A, K), dtype=torch.float32, device="cuda")
    d_b = torch.empty((N, A, K), dtype=torch.float32, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    nbytes = h_in.numel() * 4
    def t(fn, n=5):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n
    dt = t(lambda: d_a.copy_(h_in, non_blocking=True)); print(f"H2D alone      {nbytes / dt / 1e9:6.1f} GB/s  {dt * 1e3:.2f} ms")
    dt = t(lambda: h_out.copy_(d_b, non_blocking=True)); print(f"D2H alone      {nbytes / dt / 1e9:6.1f} GB/s  {dt * 1e3:.2f} ms")
    def both():
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)
    dt = t(both); print(f"H2D+D2H concurrently: {dt * 1e3:.2f} ms  ({nbytes / dt / 1e9:.1f} GB/s each way)")
    for blocks in (4, 8, 16, 32):
        ab = A // blocks
        def blocked():
            for c in range(blocks):
                lib.ppo_memcpy_2d_async(d_a.data_ptr(), ab * K * 4, h_in.data_ptr() + c * ab * K * 4, A * K * 4, ab * K * 4, N, s1.cuda_stream)
                lib.ppo_memcpy_2d_async(h_out.data_ptr() + c * ab * K * 4, A * K * 4, d_b.data_ptr(), ab * K * 4, ab * K * 4, N, s2.cuda_stream)
        dt = t(blocked); print(f"blocked 2D x{blocks:2d} both ways concurrently: {dt * 1e3:.2f} ms")

if __name__ == "__main__":
    main()
