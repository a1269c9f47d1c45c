"""Stub modules so the *unmodified* reference (``/root/reference/mbrl``) imports in this container.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may import it.

The reference needs gym / mujoco_py / forwardable / tensorboardX / h5py / imageio, none of which exist
here (no network).  The hot path never calls into them, so empty stand-ins are
enough.  ``import ref_shims`` must precede any ``import mbrl``.  This file is only used by the golden
generator and the oracle validation script, both of which run in the build container where
``/root/reference`` is mounted; it is never needed on the GPU box.
"""
import sys
import types

import numpy as np


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class Space:
    pass


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            low = np.asarray(low, dtype=dtype)
            high = np.asarray(high, dtype=dtype)
            shape = low.shape
        else:
            low = np.full(shape, low, dtype=dtype)
            high = np.full(shape, high, dtype=dtype)
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype


class Dict(Space):
    def __init__(self, spaces):
        self.spaces = spaces


class Discrete(Space):
    def __init__(self, n):
        self.n = n


class Env:
    def __init__(self, **kw):
        pass


class Wrapper(Env):
    pass


class EzPickle:
    def __init__(self, *a, **k):
        pass


class _MujocoEnv(Env):
    pass


class _FetchEnv(Env):
    pass


class _RobotEnv(Env):
    pass


class _AnyEnv(Env):
    pass


def install():
    if "gym" in sys.modules and getattr(sys.modules["gym"], "__ceeus_shim__", False):
        return
    spaces = _mod("gym.spaces", Box=Box, Dict=Dict, Discrete=Discrete, Space=Space)
    seeding = _mod("gym.utils.seeding", np_random=lambda seed=None: (None, seed))
    gutils = _mod("gym.utils", EzPickle=EzPickle, seeding=seeding)
    gerror = _mod("gym.error", DependencyNotInstalled=ImportError, Error=Exception)  # robodesk.py:9-20
    gwrappers = _mod("gym.wrappers")
    gmj = _mod("gym.envs.mujoco", MujocoEnv=_MujocoEnv, mujoco_env=types.SimpleNamespace(MujocoEnv=_MujocoEnv))
    fetch_env = _mod("gym.envs.robotics.fetch_env", FetchEnv=_FetchEnv)
    rotations = _mod("gym.envs.robotics.rotations", euler2quat=lambda e: e, mat2euler=lambda m: m)
    rutils = _mod("gym.envs.robotics.utils")
    robot_env = _mod("gym.envs.robotics.robot_env", RobotEnv=_RobotEnv)
    pp = _mod("gym.envs.robotics.fetch.pick_and_place", FetchPickAndPlaceEnv=_AnyEnv)
    rc = _mod("gym.envs.robotics.fetch.reach", FetchReachEnv=_AnyEnv)
    fetchpkg = _mod("gym.envs.robotics.fetch", pick_and_place=pp, reach=rc)
    grob = _mod(
        "gym.envs.robotics",
        fetch=fetchpkg,
        fetch_env=fetch_env,
        rotations=rotations,
        utils=rutils,
        robot_env=robot_env,
    )
    genvs = _mod("gym.envs", mujoco=gmj, robotics=grob)
    _mod("gym", spaces=spaces, Env=Env, Wrapper=Wrapper, utils=gutils, wrappers=gwrappers, envs=genvs, error=gerror,
         __ceeus_shim__=True)
    const = _mod("mujoco_py.generated.const")
    gen = _mod("mujoco_py.generated", const=const)
    builder = _mod("mujoco_py.builder", MujocoException=Exception)
    _mod("mujoco_py", generated=gen, builder=builder, MujocoException=Exception)

    def forwardable():
        def deco(cls):
            for attr, names in getattr(cls, "__delegators__", []):
                for n in names:
                    if n == "__class__":
                        continue
                    setattr(cls, n, property(lambda self, attr=attr, n=n: getattr(getattr(self, attr), n)))
            return cls

        return deco

    def def_delegators(attr, names):
        import inspect

        frame = inspect.currentframe().f_back
        names = [s.strip() for s in names.split(",")] if isinstance(names, str) else list(names)
        frame.f_locals.setdefault("__delegators__", []).append((attr, names))

    _mod("forwardable", forwardable=forwardable, def_delegators=def_delegators)

    class SummaryWriter:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, name):
            return lambda *a, **k: None

    _mod("tensorboardX", SummaryWriter=SummaryWriter)
    _mod("h5py")
    _mod("imageio")


install()
