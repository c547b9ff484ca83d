import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


class AlgoFixture:
    """tests/golden/algo_*.npz.  The wide (E = 128) fixtures store the shared encoder once (oracle/gen_golden_algo.py drops the
    identical critic.shared_net.* copies of weights and gradients); this view gives them back under both names."""

    def __init__(self, path):
        self._fx = np.load(path)
        have = set(self._fx.files)
        self._alias = {}
        for k in self._fx.files:
            if ".actor.shared_net." in k:
                twin = k.replace(".actor.", ".critic.")
                if twin not in have:
                    self._alias[twin] = k
        order = []
        for k in self._fx.files:                      # keep the state_dict order: critic.shared_net.* first among the critic keys
            if k.startswith(("w.critic.", "grad.critic.")) and not any(o.startswith(k.split(".")[0] + ".critic.") for o in order):
                order += [a for a in self._alias if a.startswith(k.split(".")[0] + ".critic.")]
            order.append(k)
        self.files = order

    def __getitem__(self, k):
        return self._fx[self._alias.get(k, k)]

    def __contains__(self, k):
        return k in self._alias or k in self._fx.files


def env_fixture_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "env_*.npz")))


def unpack_bits(packed, n):
    return np.unpackbits(packed, axis=-1, bitorder="little")[..., :n]


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc
