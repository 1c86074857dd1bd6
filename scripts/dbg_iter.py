"""Dev tool: per-tensor deviation of one full PPO iteration from the reference recording (fraction of elements
beyond 2 % of the 40*lr Adam displacement), for A/B-ing kernel variants."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.helpers import sub
from tests import test_gpu_pointnet_ppo as T
name = sys.argv[1] if len(sys.argv) > 1 else "ppo_iter_pointnet_e16.npz"
g, cfg, env, r = T._runner(name)
cu = T.cu
curr = r._ingest(env.reset()["obs"], r.storage.obs_slot())
last_obs, last_values = r.collect(curr, None, eps=cu(g["eps"]))
st = r.storage
st.compute_returns(last_values, cfg["gamma"], cfg["lam"])
for k in ("observations", "actions", "values", "returns", "advantages", "actions_log_prob", "mu", "sigma"):
    getattr(st, k).copy_(cu(g["buf." + k]))
r.update(1)
fin = sub(g, "final")
sd = {k: v.cpu() for k, v in r.actor_critic.state_dict().items()}
disp = 40 * cfg["lr"]
for k, v in fin.items():
    d = (sd[k] - v).abs()
    print(f"{k:28s} max {float(d.max())/disp:.4f}  frac>2% {float((d > 0.02*disp).float().mean()):.4f}  rms {float(d.pow(2).mean().sqrt())/disp:.5f}")
print({k: float(v) for k, v in r.log_dict.items() if k.startswith("Train/")})
