# mirrors /root/reference/algorithms/algo_utils/__init__.py:1-3
from .storage import RolloutStorage  # noqa: F401
from .actor_critic import ActorCritic  # noqa: F401
from .RMS import Normalization, RunningMeanStd  # noqa: F401
