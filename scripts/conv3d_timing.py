"""Dev tool: CUDA-event timing of the Conv3DNet student (forward, backward) at DAgger minibatch sizes, 50^3 TSDF volumes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from partmanip_b200.algorithms.algo_utils.network import Conv3DNet

dev = "cuda:0"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pk = json.load(open(os.path.join(root, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(root, "MEASURED_PEAKS.json")) else {}
hbm = float(pk.get("hbm_gbs", 6550.0))
FLOPS = 2 * (17 ** 3 * 16 * 125 + 6 ** 3 * 32 * 432 + 27 * 32 * 864 + 864 * 256 + 256 * 10)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


precs = sys.argv[1].split(",") if len(sys.argv) > 1 else ("fp32", "bf16")
sizes = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else (256, 2048)
for prec in precs:
    for B in sizes:
        torch.manual_seed(0)
        net = Conv3DNet(125000, 10, dict(name="Conv3DNet", activation="tanh", precision=prec), 0).to(dev)
        xs = [torch.rand(B, 125000, device=dev) * 2 - 1 for _ in range(3)]
        grads = [torch.empty_like(p) for p in net.parameters()]
        dout = torch.randn(B, 10, device=dev)
        k = [0]

        def fwd():
            k[0] += 1
            return net.runner.forward(xs[k[0] % 3])
        f = timeit(fwd)
        b = timeit(lambda: net.runner.backward(xs[k[0] % 3], dout, grads))
        cols = B * (17 ** 3 * 128 + 216 * 432 + 27 * 864) * 4 / 1e6
        print(f"{prec} B={B}: forward {f:.3f} ms ({B / f:.0f} samples/ms, {B * FLOPS / f / 1e9:.1f} TFLOP/s algorithmic, input {B * 0.5:.0f} MB = "
              f"{B * 0.5e-3 / (f * 1e-3) / hbm:.2f} of HBM copy peak on the input alone; patch matrices {cols:.0f} MB written + read)  backward {b:.3f} ms")
