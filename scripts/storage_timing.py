"""Dev tool: CUDA-event timing of the rollout-storage row kernels (K8) at the reference's shapes, against the measured HBM copy peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from partmanip_b200 import ops
from partmanip_b200._lib import lib

dev = "cuda:0"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peaks = json.load(open(os.path.join(root, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(root, "MEASURED_PEAKS.json")) else {}
hbm = float(peaks.get("hbm_gbs", 6650.0))


def timeit(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def p(t):
    return t.data_ptr()


st = torch.cuda.current_stream().cuda_stream
for name, rows_total, D, B in (("PPO random sampler, T*E=32768 x D=3072, minibatch 2048", 32768, 3072, 2048),
                               ("DAgger ring, 32 steps x 2048 envs x D=6144, minibatch 2048", 65536, 6144, 2048)):
    buf = torch.randn(rows_total, D, device=dev)
    out = torch.empty(B, D, device=dev)
    ms = []
    for _ in range(4):                       # a different random minibatch each time: rows come from HBM, not L2
        idx = torch.randperm(rows_total, device=dev)[:B].contiguous()
        ms.append(timeit(lambda: lib.pm_gather_rows(p(buf), D, p(idx), p(out), D, B, D, st), 1))
    m = min(ms)
    print(f"gather_rows: {name}: {m * 1e3:.1f} us  {2 * B * D * 4 / m / 1e6:.0f} GB/s (read+write) = {2 * B * D * 4 / m / 1e6 / hbm:.2f} of the HBM copy peak {hbm:.0f}")
    del buf
E, D = 4096, 3072
obs = torch.randn(16, E, D, device=dev)
store = torch.empty(8, E, D, device=dev)
k = [0]


def add():
    lib.pm_copy_rows(p(obs[k[0] % 16]), D, p(store[k[0] % 8]), D, E, D, st)
    k[0] += 1


m = timeit(add, 32)
print(f"copy_rows (add_transitions, 4096 envs x D=3072 = 50 MB): {m * 1e3:.1f} us  {2 * E * D * 4 / m / 1e6:.0f} GB/s = {2 * E * D * 4 / m / 1e6 / hbm:.2f} of the HBM copy peak")
