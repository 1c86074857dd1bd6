"""Round-2 additions to the golden fixtures, generated from the UNMODIFIED reference (build container only):

    python tests/golden/make_golden_extra.py      # needs /root/reference (read-only)

  pointnet_critic_out1.npz   PointNet with output_dim = 1 (the critic head, actor_critic.py:19) forward + gradients
  mlp_state_512.npz          the shipped state policy shape MLP 53 -> 512^3 -> 10 (cfg/algos/ppo.yaml:44-47) at a 96-row batch
  mlp_state_512_critic.npz   its critic, 53 -> 512^3 -> 1
Same stubbing of `utils` as make_golden.py."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402


def main():
    au, ppo, PointNet, MLP = MG._import_reference()
    torch.set_num_threads(8)
    base = dict(name="PointNet", activation="tanh", max_mean=False, sub_mean=False)
    torch.manual_seed(31)
    net = PointNet(3072, 1, base, 0)
    g = torch.Generator().manual_seed(32)
    x = torch.rand(5, 3072, generator=g) * 2 - 1
    x.view(-1)[::17] = 0.0
    x_in = x.clone()
    y = net(x_in)
    y.square().sum().backward()
    MG.save("pointnet_critic_out1.npz", x=x, x_after=x_in, y=y, **{"w." + k: v for k, v in net.state_dict().items()},
            **{"g." + k: p.grad for k, p in net.named_parameters()})
    torch.manual_seed(33)
    for tag, out in (("mlp_state_512", 10), ("mlp_state_512_critic", 1)):
        net = MLP(53, out, dict(hid_dim=[512, 512, 512], activation="tanh"), 0)
        x = torch.randn(96, 53)
        y = net(x)
        y.square().sum().backward()
        MG.save(f"{tag}.npz", x=x, y=y, **{"w." + k: v for k, v in net.state_dict().items()},
                **{"g." + k: p.grad for k, p in net.named_parameters()})


if __name__ == "__main__":
    main()
