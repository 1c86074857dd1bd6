"""Dev tool: CUDA-event timing of the observation normaliser (K6) at E=4096 x D=3072 against the measured HBM copy peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from partmanip_b200 import ops

dev = "cuda:0"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
hbm = float(json.load(open(os.path.join(root, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6550.0)) if os.path.exists(os.path.join(root, "MEASURED_PEAKS.json")) else 6550.0
E, D = 4096, 3072
pool = torch.randn(16, E, D, device=dev)                     # 16 x 50 MB: every step's observations come from HBM, not L2
out = torch.empty(E, D, device=dev)
mean = torch.zeros(1, D, device=dev); S = torch.full((1, D), 1e-4, device=dev); std = S.sqrt()
k = [0]


def step(update):
    k[0] += 1
    ops.rms_forward(pool[k[0] % 16], out, mean, S, std, k[0], update)


def timeit(fn, reps=32):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


mb = E * D * 4 / 1e6
for update, alg, what in ((True, 4 * mb, "3 reads (mean, squared deviation, normalise) + 1 write"), (False, 2 * mb, "1 read + 1 write")):
    ms = timeit(lambda: step(update))
    dram = 2 * mb                                            # the 2nd / 3rd read of the 50 MB batch can come from the 126 MB L2
    print(f"rms_forward update={update}: {ms * 1e3:.1f} us per env step; algorithmic {alg:.0f} MB ({what}) -> {alg / ms:.0f} GB/s = "
          f"{alg / ms / hbm:.2f} of the HBM copy peak; minimum DRAM traffic {dram:.0f} MB -> {dram / ms:.0f} GB/s = {dram / ms / hbm:.2f}")
