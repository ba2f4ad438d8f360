import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    binding.lib()
    return binding


@pytest.fixture(scope="session")
def sv_series():
    """Synthetic SV series: x_1 ~ N(0, s^2/(1-phi^2)), x_t = phi x_{t-1} + s e_t, y_t = beta e^{x_t/2} n_t."""
    def make(T, seed=20260101, beta=1.0, phi=0.95, sigma=0.25):
        rng = np.random.default_rng(seed)
        x = np.empty(T)
        x[0] = rng.standard_normal() * sigma / np.sqrt(1 - phi * phi)
        for t in range(1, T):
            x[t] = phi * x[t - 1] + sigma * rng.standard_normal()
        return beta * np.exp(0.5 * x) * rng.standard_normal(T)
    return make


@pytest.fixture(scope="session")
def gpu_backend_factory():
    import ssme_b200 as sb
    made = []

    def make(**kw):
        be = sb.ParticleFilterBackend(sb.FilterConfig(**kw))
        made.append(be)
        return be
    yield make
    for be in made:
        be.close()
