"""Back-to-back ppo_vtrace launches on rotating buffers (dev aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ppo_b200 import _lib

def run(n, b, act=18, sets=4, iters=40):
    lib = _lib.load(); dev = torch.device("cuda")
    g = torch.Generator(device=dev); g.manual_seed(1)
    bufs = []
    for _ in range(sets):
        beh = torch.log_softmax(torch.randn((n, b, act), device=dev, generator=g), dim=-1)
        tgt = torch.log_softmax(torch.randn((n, b, act), device=dev, generator=g), dim=-1)
        a = torch.randint(0, act, (n, b), device=dev, generator=g)
        r = torch.randint(-1, 3, (n, b), device=dev, generator=g).float()
        d = (torch.rand((n, b), device=dev, generator=g) < 0.03).view(torch.uint8)
        v = torch.randn((n + 1, b), device=dev, generator=g)
        bufs.append((beh, tgt, a, r, d, v, [torch.empty((n, b), device=dev) for _ in range(3)]))
    s = torch.cuda.current_stream().cuda_stream
    def launch(i):
        beh, tgt, a, r, d, v, o = bufs[i % sets]
        _lib.check(lib.ppo_vtrace(beh.data_ptr(), tgt.data_ptr(), a.data_ptr(), 8, r.data_ptr(), d.data_ptr(), 1, v.data_ptr(),
                                  v[n].data_ptr(), n, b, act, 0.999, 1.0, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), None, s), "vt")
    for i in range(6): launch(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): launch(i)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    print(f"vtrace {n}x{b}x{act}: {us:7.1f} us/launch  {n * b * 37 / us / 1e3:6.0f} GB/s useful, {n * b * 93 / us / 1e3:6.0f} GB/s sectors")

if __name__ == "__main__":
    run(256, 4096); run(256, 512, sets=16); run(256, 32768, sets=1)
