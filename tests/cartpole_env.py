"""The old-gym-API CartPole lives in oracle/ (test infrastructure shared with oracle/_ref's gym stub)."""
from oracle.cartpole_env import Box, CartPoleEnv, Discrete, make  # noqa: F401
