"""TEST INFRASTRUCTURE - golden vectors for RewardScaling / RunningMeanStd on non-integer samples, produced by EXECUTING THE
UNMODIFIED REFERENCE `DHGN/normalization.py` (numpy only) in the build container.

Two episodes of 40 steps for 5 agents: per-step rewards in, scaled rewards out (x / (std of the running discounted return + 1e-8)),
`reset()` between the episodes (the running estimate is kept), plus the final n / mean / S / std of the estimate."""
import importlib.util
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_bootstrap import REFERENCE_ROOT  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    spec = importlib.util.spec_from_file_location("ref_normalization", os.path.join(REFERENCE_ROOT, "DHGN", "normalization.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    g = np.random.default_rng(41)
    N, T, gamma = 5, 40, 0.99
    rs = ref.RewardScaling(shape=N, gamma=gamma)
    x = np.concatenate([g.integers(-12, 3, (T, N)).astype(np.float64), g.normal(0.0, 3.0, (T, N))])      # integer rewards, then real-valued ones
    out = []
    for t in range(2 * T):
        if t == T:
            rs.reset()
        out.append(rs(x[t]))
    ms = rs.running_ms
    np.savez_compressed(os.path.join(GOLDEN_DIR, "reward_scaling.npz"), x=x, out=np.stack(out), gamma=np.float64(gamma),
                        n=np.int64(ms.n), mean=ms.mean, S=ms.S, std=ms.std, R=rs.R)
    print("reward_scaling: n =", ms.n, "std =", ms.std)


if __name__ == "__main__":
    main()
