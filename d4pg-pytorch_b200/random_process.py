"""Exploration-noise processes constructed by DDPG.__init__ (ddpg.py:75).  Actor-side helpers,
outside the learner hot path; kept so `ddpg.noise.sample()` / `.reset()` callers keep working
(reference: random_process.py:4-45)."""
import numpy as np


class _DecayingNoise(object):
    min_epsilon = 0.01

    def _decayed(self, rate, it):
        return self.min_epsilon + (1.0 - self.min_epsilon) * np.exp(-rate * it)


class GaussianNoise(_DecayingNoise):
    def __init__(self, dimension, num_epochs, mu=0.0, var=1):
        self.mu, self.var, self.dimension = mu, var, dimension
        self.epochs, self.num_epochs = 0, num_epochs
        self.epsilon = 0.3
        self.decay_rate = 5.0 / num_epochs
        self.iter = 0

    def sample(self):
        return self.epsilon * np.random.normal(self.mu, self.var, size=self.dimension)

    def reset(self):
        self.epsilon = self._decayed(self.decay_rate, self.iter)


class OrnsteinUhlenbeckProcess(_DecayingNoise):
    def __init__(self, dimension, num_steps, theta=0.25, mu=0.0, sigma=0.05, dt=0.01):
        self.theta, self.mu, self.sigma, self.dt = theta, mu, sigma, dt
        self.dimension, self.num_steps = dimension, num_steps
        self.x = np.zeros((dimension,))
        self.iter = 0
        self.epsilon = 1.0
        self.decay_rate = 5.0 / num_steps

    def sample(self):
        drift = self.theta * (self.mu - self.x) * self.dt
        diffusion = self.sigma * np.sqrt(self.dt) * np.random.normal(size=self.dimension)
        self.x = self.x + drift + diffusion
        return self.epsilon * self.x

    def reset(self):
        self.x = np.zeros_like(self.x)
        self.iter += 1
        self.epsilon = self._decayed(self.decay_rate, self.iter)
