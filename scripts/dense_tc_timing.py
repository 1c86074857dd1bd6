"""Dev tool: CUDA-event timing of the tcgen05 dense layers (dense_tc.cu) on the two shapes that matter — the encoder backward's
[~250k x 256 x 128] layers (many M tiles per CTA) and the state policy's [2048 x 512 x 512] (one tile per CTA, latency-bound)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from partmanip_b200 import ops

dev = "cuda:0"


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for M, N, K in ((250000, 256, 128), (250000, 512, 256), (2048, 512, 512), (2048, 512, 53), (16384, 512, 512)):
    x, W, b = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev) / K ** 0.5, torch.randn(N, device=dev)
    dy = torch.randn(M, N, device=dev)
    out, dW, db, dx = torch.empty(M, N, device=dev), torch.empty(N, K, device=dev), torch.empty(N, device=dev), torch.empty(M, K, device=dev)
    for prec in ("fp32", "bf16"):
        f = timeit(lambda: ops.linear_forward_tc(x, W, b, "tanh", prec, out=out))
        bw = timeit(lambda: ops.linear_backward_tc(x, W, dy, dW, db, dx, "tanh", prec))
        fl = 2.0 * M * N * K
        print(f"{M:7d} x {N} x {K} {prec}: forward {f:8.1f} us ({fl / f / 1e6:6.1f} TFLOP/s)   backward (dX+dW+db) {bw:8.1f} us ({2 * fl / bw / 1e6:6.1f} TFLOP/s)")
ops.check_tc_errors()
