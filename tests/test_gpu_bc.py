"""GPU parity of the behaviour-cloning runner (partmanip_b200/algorithms/bc.py) against a recording of the UNMODIFIED reference
runner (algorithms/bc.py:33-179, tests/golden/make_golden_bc.py) on the same synthetic offline dataset (tests/helpers_bc.py), same
initial weights, same torch seed for the DataLoader's shuffles: per-iteration losses, learning-rate schedule, updated actor, untouched
critic / log_std, and the optimizer state layout of the checkpoint."""
import os

import pytest
import torch

from tests.helpers import load_golden
from tests.helpers_bc import FakeBCEnv, Logger, bc_cfg, write_dataset

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    root = tmp_path_factory.mktemp("bc_data")
    n = write_dataset(str(root), seed=5, scenes=3, steps=5)
    assert n == 15
    return str(root)


@pytest.mark.parametrize("tag,over", [("step", dict(lr_schedule="step_decay", max_iterations=4)), ("lin", dict(lr_schedule="linear_decay", n_minibatches=4))])
def test_bc_run_vs_reference_recording(dataset, tmp_path, tag, over):
    from partmanip_b200.algorithms import bc
    g = load_golden("bc_conv3d.npz")
    log = Logger(str(tmp_path))
    r = bc(FakeBCEnv(), bc_cfg(dataset, DEV, num_workers=0, **over), log)
    w0 = {k[3:]: v for k, v in g.items() if k.startswith("w0.")}
    r.student.load_state_dict(w0)
    torch.manual_seed(77)
    r.run()
    want_loss, want_lr = g[f"{tag}.loss"], g[f"{tag}.lr"]
    got_loss = torch.tensor([row["Train/bc_loss"] for _, row in log.rows])
    got_lr = torch.tensor([row["Train/learning_rate"] for _, row in log.rows], dtype=torch.float64)
    assert len(log.rows) == len(want_loss) and [it for it, _ in log.rows] == list(range(1, len(want_loss) + 1))
    assert torch.allclose(got_lr, want_lr, rtol=1e-12, atol=1e-15), (got_lr, want_lr)
    # the first iteration's loss is computed from the recorded weights alone: fp32 gate; later ones ride on the Adam trajectory
    assert abs(float(got_loss[0] - want_loss[0])) <= 2e-4 * float(want_loss[0]), (got_loss, want_loss)
    assert float((got_loss - want_loss).abs().max()) <= 2e-3 * float(want_loss.max()), (got_loss, want_loss)
    sd = {k: v.detach().cpu() for k, v in r.student.state_dict().items()}
    steps = int(g[f"{tag}.opt_step"][0])
    worst = 0.0
    for k, v in sd.items():
        if k.startswith("actor."):
            want = g[f"{tag}.w1.{k}"]
            moved = float((want - w0[k]).abs().max())
            assert moved > 0, k
            worst = max(worst, float((v - want).abs().max()) / (5e-4 * steps))
        else:                                        # critic and log_std (= -inf with action_std 0.0) are never updated
            assert torch.equal(v, w0[k]), k
    assert worst <= 0.1, worst                       # fraction of the lr * steps Adam displacement
    r.save(99)
    ck = torch.load(os.path.join(str(tmp_path), "model_99.pth"), weights_only=False)
    assert set(ck) == {"iteration", "model_state_dict", "optimizer_state_dict", "obs_mode", "total_steps", "tricks", "teacher"}
    opt = ck["optimizer_state_dict"]
    assert sorted(opt["state"].keys()) == g[f"{tag}.opt_keys"].tolist() and opt["param_groups"][0]["params"] == g[f"{tag}.opt_group_params"].tolist()
    assert [float(v["step"]) for v in opt["state"].values()] == g[f"{tag}.opt_step"].tolist()
    # resume: a second runner picks up iteration, weights and optimizer state
    r2 = bc(FakeBCEnv(), bc_cfg(dataset, DEV, num_workers=0, resume=os.path.join(str(tmp_path), "model_99.pth"), **over), Logger(str(tmp_path)))
    assert r2.curr_iter == 99 and all(torch.equal(a, b) for a, b in zip(r2.student.state_dict().values(), r.student.state_dict().values()))


def test_bc_test_only_raises(dataset, tmp_path):
    from partmanip_b200.algorithms import bc
    r = bc(FakeBCEnv(), bc_cfg(dataset, DEV, num_workers=0, test_only=True), Logger(str(tmp_path)))
    with pytest.raises(NotImplementedError):
        r.run()
