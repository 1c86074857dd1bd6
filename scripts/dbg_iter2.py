import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.helpers import sub
from tests import test_gpu_pointnet_ppo as T
from partmanip_b200 import ops
from partmanip_b200.algorithms.algo_utils import network as NW
g, cfg, env, r = T._runner("ppo_iter_pointnet_e16.npz")
cu = T.cu
curr = r._ingest(env.reset()["obs"], r.storage.obs_slot())
last_obs, last_values = r.collect(curr, None, eps=cu(g["eps"]))
st = r.storage
st.compute_returns(last_values, cfg["gamma"], cfg["lam"])
for k in ("observations", "actions", "values", "returns", "advantages", "actions_log_prob", "mu", "sigma"):
    getattr(st, k).copy_(cu(g["buf." + k]))
ac = r.actor_critic
batch = st.mini_batch_generator(r.num_mini_batches)
mb = r._minibatch(batch[0])
B = mb['obs'].shape[0]
res = {}
for mode in (True, False):
    NW._FUSED_HEAD = mode
    dv = torch.empty(B, 1, device="cuda:0")
    v = ac.critic.runner.forward(mb['obs'])
    ops.value_loss(v, mb['ret'].reshape(-1), mb['val'].reshape(-1), None, 1.0 / B, r._stats_v, dv)
    r.optimizer_critic.grad.zero_()
    ac.critic.runner.backward(mb['obs'], dv, r._critic_grads)
    buf = ac.critic.runner._bufs[B]
    res[mode] = dict(v=v.clone(), dv=dv.clone(), grad=r.optimizer_critic.grad.clone(), dfeat=buf["dfeat"].clone(), h1=buf["h1"].clone(), h2=buf["h2"].clone(),
                     feat=buf["feat"].clone())
for k in res[True]:
    a, b = res[True][k], res[False][k]
    print(k, tuple(a.shape), "rel", float((a - b).norm() / (b.norm() + 1e-30)), "maxabs", float((a - b).abs().max()), "ref max", float(b.abs().max()))
names = [n for n, _ in ac.critic.named_parameters()]
for n, a, b in zip(names, ac.grad_views(res[True]["grad"], "critic"), ac.grad_views(res[False]["grad"], "critic")):
    print(f"{n:22s} rel {float((a - b).norm() / (b.norm() + 1e-30)):.2e}  ref norm {float(b.norm()):.3e}")
