# mirrors /root/reference/algorithms/__init__.py:1-3 — classes are named exactly like cfg['algo_name'] (train.py:70)
from .ppo import ppo  # noqa: F401
from .dagger import dagger  # noqa: F401
from .bc import bc  # noqa: F401
