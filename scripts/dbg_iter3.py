"""Dev tool: shows that the 40-step Adam trajectory of the full-iteration test is chaotic in the max-pool arg-max — two of our own
kernel variants whose single-step gradients agree to 1e-7 drift apart as much as either does from the reference recording
(the reason the full-iteration weight gate in tests/test_gpu_pointnet_ppo.py is statistical)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import test_gpu_pointnet_ppo as T
from partmanip_b200 import ops
from partmanip_b200.algorithms.algo_utils import network as NW
cu = T.cu
runners = []
for mode in (True, False):
    NW._FUSED_HEAD = mode
    g, cfg, env, r = T._runner("ppo_iter_pointnet_e16.npz")
    curr = r._ingest(env.reset()["obs"], r.storage.obs_slot())
    last_obs, last_values = r.collect(curr, None, eps=cu(g["eps"]))
    st = r.storage
    st.compute_returns(last_values, cfg["gamma"], cfg["lam"])
    for k in ("observations", "actions", "values", "returns", "advantages", "actions_log_prob", "mu", "sigma"):
        getattr(st, k).copy_(cu(g["buf." + k]))
    runners.append(r)
step = 0
for epoch in range(5):
    for i in range(8):
        outs = []
        for mode, r in zip((True, False), runners):
            NW._FUSED_HEAD = mode
            ac = r.actor_critic
            batch = r.storage.mini_batch_generator(r.num_mini_batches)
            mb = r._minibatch(batch[i])
            B = mb['obs'].shape[0]
            dv = torch.empty(B, 1, device="cuda:0")
            v = ac.critic.runner.forward(mb['obs'])
            ops.value_loss(v, mb['ret'].reshape(-1), mb['val'].reshape(-1), None, 1.0 / B, r._stats_v, dv)
            ac.critic.runner.backward(mb['obs'], dv, r._critic_grads)
            gr = r.optimizer_critic.grad.clone()
            r.optimizer_critic.step(None)
            outs.append((gr, ac.critic_flat.clone(), ac.critic.runner._bufs[B]["argmax"].clone()))
        (ga, pa, aa), (gb, pb, ab) = outs
        print(f"step {step:2d} grad rel {float((ga-gb).norm()/gb.norm()):.2e} max|dgrad| {float((ga-gb).abs().max()):.2e}  param maxdiff/lr {float((pa-pb).abs().max())/5e-5:.4f}  argmax diff {int((aa!=ab).sum())}")
        step += 1
