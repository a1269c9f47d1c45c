"""cee-us_b200: B200-native (sm_100a) iCEM planning hot path of martius-lab/cee-us.

Importable as ``ceeus_b200`` (the directory name carries the upstream project's hyphen; ``ceeus_b200.py`` at the repo
root aliases it).  Everything arithmetic lives in ``libceeus_b200.so`` (csrc/, C ABI in include/ceeus_b200.h)."""
from . import _lib  # noqa: F401
from .controllers import B200CuriosityMpcICem, B200MpcICem  # noqa: F401
from .engine import Engine  # noqa: F401
from .envs import TensorEnv, construction_env, describe_env, playground_env, robodesk_env  # noqa: F401
from .models import B200GNNEnsemble, B200MLPEnsemble  # noqa: F401
from .rnd import B200RndMpcICem  # noqa: F401
from .rollout_buffer import RolloutView  # noqa: F401
from .rollout_manager import B200RolloutManager, BatchedEnvView, Rollout  # noqa: F401
