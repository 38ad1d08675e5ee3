"""``create_env`` -- the reference's ``environment`` module surface (environment.py:66-74 upstream).

The emulator is outside the learner hot path.  When gym + ALE are importable the reference's
environment is reproduced (grayscale, frameskip 4, 84x84 INTER_AREA warp, no-op starts).  Like the
reference, ``create_env`` RAISES when the emulator cannot be created (missing package or ROM, wrong game
name ...): training on noise by accident is worse than not starting.  Only when the caller opts in
explicitly -- ``R2D2_SYNTHETIC_ENV=1`` in the environment or ``create_env(..., synthetic=True)`` -- and gym /
ALE are not importable, a synthetic stand-in with the same interface (``action_space.n``,
``reset() -> (1,84,84) u8``, ``step(a) -> (obs, reward, done, info)``) is returned, with a warning, so that
train.py-style pipelines and the benchmarks can run on a box without an emulator (this image has none).
"""
import os
import warnings

import numpy as np

from . import config


class _Discrete:
    def __init__(self, n, rng):
        self.n, self._rng = n, rng

    def sample(self):
        return int(self._rng.integers(0, self.n))


class SyntheticAtariEnv:
    """Random-frame episodic environment with MsPacman's action count (9)."""

    def __init__(self, action_dim: int = 9, obs_shape=config.obs_shape, mean_episode_len: int = 600, seed: int = 0):
        self._rng = np.random.default_rng(seed)
        self.action_space = _Discrete(action_dim, self._rng)
        self.obs_shape = tuple(obs_shape)
        self.mean_episode_len = mean_episode_len
        self._t = 0

    def _obs(self):
        return self._rng.integers(0, 256, size=self.obs_shape, dtype=np.uint8)

    def reset(self, **kwargs):
        self._t = 0
        self._len = int(self._rng.integers(self.mean_episode_len // 2, self.mean_episode_len * 3 // 2))
        return self._obs()

    def step(self, action):
        self._t += 1
        reward = float(self._rng.random() < 0.1)
        return self._obs(), reward, self._t >= self._len, {}


def _make_ale(env_name, noop_start):
    import cv2
    import gym

    class WarpFrame(gym.ObservationWrapper):
        def __init__(self, env):
            super().__init__(env)
            self.observation_space = gym.spaces.Box(low=0, high=255, shape=(1, 84, 84), dtype=np.uint8)

        def observation(self, obs):
            return np.expand_dims(cv2.resize(obs, (84, 84), interpolation=cv2.INTER_AREA), 0)

    class NoopReset(gym.Wrapper):
        def reset(self, **kwargs):
            self.env.reset(**kwargs)
            obs = None
            for _ in range(np.random.randint(1, 31)):
                obs, _, done, _ = self.env.step(0)
                if done:
                    obs = self.env.reset(**kwargs)
            return obs

    env = gym.make(f'ALE/{env_name}-v5', obs_type='grayscale', frameskip=4, repeat_action_probability=0, full_action_space=False)
    env = WarpFrame(env)
    if noop_start:
        assert env.unwrapped.get_action_meanings()[0] == 'NOOP'          # environment.py:24 upstream
        env = NoopReset(env)
    return env


def create_env(env_name=config.game_name, noop_start=True, synthetic=None):
    if synthetic is None:
        synthetic = os.environ.get("R2D2_SYNTHETIC_ENV") == "1"
    try:
        return _make_ale(env_name, noop_start)
    except ImportError as e:                                             # gym / ale_py / cv2 not installed
        if not synthetic:
            raise ImportError(f"{e}; no Atari emulator in this environment -- set R2D2_SYNTHETIC_ENV=1 (or pass synthetic=True) "
                              f"to run on a synthetic random-frame environment instead") from e
        warnings.warn(f"create_env({env_name!r}): gym/ALE not importable ({e}); using SyntheticAtariEnv (random frames, 9 actions)",
                      RuntimeWarning, stacklevel=2)
        return SyntheticAtariEnv(seed=np.random.randint(0, 2 ** 31 - 1))
