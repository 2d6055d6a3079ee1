"""Times the row kernels at the FLUX image+text stream shape (M = 4608, D = 3072), L2 flushed between launches.
`B200_COL_REDUCE_NARROW=1 python tools/time_rows.py` = the round-1 4-byte col_reduce for the A/B."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_toolkit_b200 import ops  # noqa: E402

dev = "cuda:0"
M, D, rps = 4608, 3072, 4608
torch.manual_seed(0)
x = torch.randn(M, D, device=dev).bfloat16()
dy = torch.randn(M, D, device=dev).bfloat16()
dres = torch.randn(M, D, device=dev).bfloat16()
mod = (torch.randn(1, 6 * D, device=dev) * 0.3).bfloat16()
out, mean, rstd = ops.ln_modulate_fwd(x, mod[:, :D], mod[:, D:2 * D], rps)
dmod = torch.zeros(1, 6 * D, device=dev)
mul = torch.empty_like(x)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    t = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        t.append(a.elapsed_time(b) * 1e3)
    t.sort()
    return t[len(t) // 2]


def warm(fn, n=50):  # back-to-back (inputs L2-resident where they fit)
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / n


cases = {
    "col_reduce stats (2 x 28 MB read)": (lambda: ops.col_reduce(dy, rps, b=x, mean=mean, rstd=rstd, sum_a=dmod[:, :D], sum_ab=dmod[:, D:2 * D]), 2 * M * D * 2),
    "col_reduce gate  (2 x 28 MB read, 28 MB write)": (lambda: ops.col_reduce(dy, rps, b=x, g=mod[:, 2 * D:3 * D], mul_out=mul, sum_ab=dmod[:, 2 * D:3 * D]), 3 * M * D * 2),
    "ln_modulate_fwd  (28 MB read, 28 MB write)": (lambda: ops.ln_modulate_fwd(x, mod[:, :D], mod[:, D:2 * D], rps), 2 * M * D * 2),
    "ln_modulate_bwd  (3 x 28 MB read, 28 MB write)": (lambda: ops.ln_modulate_bwd(dy, x, mean, rstd, mod[:, D:2 * D], rps, dres=dres), 4 * M * D * 2),
}
print(f"narrow={os.environ.get('B200_COL_REDUCE_NARROW', '0')}")
for name, (fn, nbytes) in cases.items():
    c, w = timeit(fn), warm(fn)
    print(f"{name:52s} cold {c:7.2f} us ({nbytes / c / 1e6:6.2f} TB/s)   back-to-back {w:7.2f} us ({nbytes / w / 1e6:6.2f} TB/s)")
