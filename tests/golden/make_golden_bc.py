"""Golden run of the UNMODIFIED reference behaviour-cloning runner (algorithms/bc.py:33-179) on CPU over the synthetic offline
dataset of tests/helpers_bc.py: initial weights, per-iteration logs, final weights, the saved optimizer layout.

    python tests/golden/make_golden_bc.py          # build container only (needs /root/reference)

The reference model (cfg/algos/bc.yaml) has action_std 0.0 -> log_std = -inf; it is never touched by the update."""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_golden as MG  # noqa: E402
from tests.helpers_bc import FakeBCEnv, Logger, bc_cfg, write_dataset  # noqa: E402


def main():
    MG._import_reference()
    from algorithms.bc import bc
    torch.set_num_threads(8)
    tmp = tempfile.mkdtemp(prefix="pm_bc_golden_")
    n = write_dataset(os.path.join(tmp, "data"), seed=5, scenes=3, steps=5)      # 15 samples, 3 minibatches of 5
    rec = {}
    for tag, over in (("step", dict(lr_schedule="step_decay", max_iterations=4)), ("lin", dict(lr_schedule="linear_decay", n_minibatches=4))):
        torch.manual_seed(123)
        log = Logger(os.path.join(tmp, "ckpt_" + tag))
        r = bc(FakeBCEnv(), bc_cfg(os.path.join(tmp, "data"), "cpu", **over), log)
        rec.update({f"{tag}.w0.{k}": v.clone() for k, v in r.student.state_dict().items()})
        torch.manual_seed(77)                     # the DataLoader's shuffles draw from the global generator
        r.run()
        rec.update({f"{tag}.w1.{k}": v.clone() for k, v in r.student.state_dict().items()})
        rec[f"{tag}.loss"] = torch.tensor([row["Train/bc_loss"] for _, row in log.rows])
        rec[f"{tag}.lr"] = torch.tensor([row["Train/learning_rate"] for _, row in log.rows], dtype=torch.float64)
        r.save(99)
        ck = torch.load(os.path.join(log.save_ckpt_dir, "model_99.pth"), weights_only=False)
        rec[f"{tag}.opt_keys"] = torch.tensor(sorted(ck["optimizer_state_dict"]["state"].keys()))
        rec[f"{tag}.opt_group_params"] = torch.tensor(ck["optimizer_state_dict"]["param_groups"][0]["params"])
        rec[f"{tag}.opt_step"] = torch.tensor([float(v["step"]) for v in ck["optimizer_state_dict"]["state"].values()])
        print(tag, "n =", n, "loss", rec[f"{tag}.loss"], "lr", rec[f"{tag}.lr"], "adam state on", rec[f"{tag}.opt_keys"].tolist())
    # slim the fixture: both runs start from the same weights, the critic and log_std never change (asserted), so store the
    # initial model once and only the updated actors
    out = {}
    for k, v in rec.items():
        if k.startswith("lin.w0."):
            assert torch.equal(v, rec["step.w0." + k[7:]])
        elif ".w1.critic." in k or k.endswith(".w1.log_std"):
            assert torch.equal(v, rec["step.w0." + k.split(".w1.")[1]]) or k.endswith("log_std")
        else:
            out[k.replace("step.w0.", "w0.")] = v
    MG.save("bc_conv3d.npz", **out)


if __name__ == "__main__":
    main()
