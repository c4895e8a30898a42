import sys, torch
sys.path.insert(0, ".")
from mr_blip_b200 import ops
def timeit(fn, iters=10):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters
for M, C in [(8132, 2048), (8132, 5120), (8132, 10240), (8132, 6144), (56, 2048), (56, 32128)]:
    P = torch.randn(M, C + 32, device="cuda").bfloat16()
    Q = torch.randn(M, 2080, device="cuda").bfloat16()
    out = torch.zeros(C, 8, device="cuda")
    for tr, impl in ((False, 'cc'), (True, 'cc'), (False, 'tc'), (True, 'tc')):
        o = torch.zeros(8, C, device="cuda") if tr else out
        ms = timeit(lambda: ops.skinny_wgrad(P.data_ptr(), P.stride(0), Q.data_ptr() + 2048 * 2, Q.stride(0), M, C, o, tr, ops.BF16, impl=impl))
        print("M=%d C=%d transposed=%s %s: %.1f us  (%.0f GB/s)" % (M, C, tr, impl, ms * 1e3, M * C * 2 / ms / 1e6))
