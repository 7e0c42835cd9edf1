"""In-repo CartPole with the OLD gym API the reference expects (actor.py:73,88; utils.py:95-107).

`reset() -> obs`, `step(a) -> (obs, reward, done, info)`, `observation_space.shape`,
`action_space.n`, and an action-space class literally named `Discrete`.  Standard cart-pole
dynamics (Barto, Sutton & Anderson 1983), Euler integration, 500-step limit.  Test helper only.
"""
import math

import numpy as np


class Discrete:
    def __init__(self, n):
        self.n = n


class Box:
    def __init__(self, shape):
        self.shape = shape


class CartPoleEnv:
    GRAVITY, M_CART, M_POLE, HALF_LEN, FORCE, DT = 9.8, 1.0, 0.1, 0.5, 10.0, 0.02
    X_LIMIT, THETA_LIMIT, MAX_STEPS = 2.4, 12 * 2 * math.pi / 360, 500

    def __init__(self, seed=0):
        self.rng = np.random.default_rng(seed)
        self.observation_space = Box((4,))
        self.action_space = Discrete(2)
        self.state = None
        self.steps = 0

    def reset(self):
        self.state = self.rng.uniform(-0.05, 0.05, size=4)
        self.steps = 0
        return self.state.copy()

    def step(self, action):
        x, x_dot, th, th_dot = self.state
        f = self.FORCE if int(action) == 1 else -self.FORCE
        total_m = self.M_CART + self.M_POLE
        pm_l = self.M_POLE * self.HALF_LEN
        tmp = (f + pm_l * th_dot ** 2 * math.sin(th)) / total_m
        th_acc = (self.GRAVITY * math.sin(th) - math.cos(th) * tmp) / (
            self.HALF_LEN * (4.0 / 3.0 - self.M_POLE * math.cos(th) ** 2 / total_m))
        x_acc = tmp - pm_l * th_acc * math.cos(th) / total_m
        self.state = np.array([x + self.DT * x_dot, x_dot + self.DT * x_acc,
                               th + self.DT * th_dot, th_dot + self.DT * th_acc])
        self.steps += 1
        done = bool(abs(self.state[0]) > self.X_LIMIT or abs(self.state[2]) > self.THETA_LIMIT
                    or self.steps >= self.MAX_STEPS)
        return self.state.copy(), 1.0, done, {}

    def render(self):
        pass

    def close(self):
        pass


def make(env_name, **kwargs):
    if "CartPole" not in env_name:
        raise ValueError(f"test stub only provides CartPole, not {env_name}")
    return CartPoleEnv(seed=kwargs.get("seed", 0))
