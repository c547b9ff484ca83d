"""Running mean/std normalisation — mirror of the reference `DHGN/normalization.py` (RunningMeanStd, Normalization,
RewardScaling) with the statistics updated by the Welford kernel (csrc/stats_kernels.cu, marl_welford_update).

Semantics kept (DHGN/normalization.py:11-22): on the first sample `mean = x` and `std = x` (so the first output is 0),
afterwards the population std `sqrt(S/n)`; the estimate is never reset across episodes (one per worker)."""
import numpy as np
import torch

from . import _lib


class RunningMeanStd:
    def __init__(self, shape, device=None):
        self.shape = int(shape)
        # default: the CURRENT device (one process per GPU: rank k must not allocate, or launch, on cuda:0)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._n = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._mean = torch.zeros(1, self.shape, dtype=torch.float64, device=self.device)
        self._S = torch.zeros_like(self._mean)
        self._std = torch.zeros_like(self._mean)
        self._out = torch.zeros(1, self.shape, dtype=torch.float32, device=self.device)

    n = property(lambda s: int(s._n.item()))
    mean = property(lambda s: s._mean[0].cpu().numpy())
    S = property(lambda s: s._S[0].cpu().numpy())
    std = property(lambda s: s._std[0].cpu().numpy())

    def update(self, x):
        self._run(x, True)

    def update_f64(self, x):
        """RunningMeanStd.update for samples that are not integers (the discounted return of RewardScaling)."""
        x = torch.as_tensor(np.asarray(x, dtype=np.float64).reshape(1, self.shape), device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().marl_welford_update_f64(1, self.shape, _lib.ptr(x), _lib.ptr(self._n), _lib.ptr(self._mean),
                                                          _lib.ptr(self._S), _lib.ptr(self._std), _lib.ptr(self._out), 1,
                                                          _lib.stream_ptr()), "marl_welford_update_f64")

    def _run(self, x, update):
        a = np.asarray(x, dtype=np.float64)
        if not np.array_equal(a, np.rint(a)):             # checked BEFORE the cast: a non-integer reward must not be truncated silently
            raise _lib.MarlError("Normalization: rewards of this env are integers")
        x = torch.as_tensor(a.astype(np.int32).reshape(1, self.shape), device=self.device)
        with torch.cuda.device(self.device):
            return self._launch(x, update)

    def _launch(self, x, update):
        _lib.check(_lib.lib().marl_welford_update(1, self.shape, _lib.ptr(x), _lib.ptr(self._n), _lib.ptr(self._mean),
                                                  _lib.ptr(self._S), _lib.ptr(self._std), _lib.ptr(self._out),
                                                  1 if update else 0, _lib.stream_ptr()), "marl_welford_update")
        return x


class Normalization:
    def __init__(self, shape, device=None):
        self.running_ms = RunningMeanStd(shape=shape, device=device)

    def __call__(self, x, update=True):
        """-> np.ndarray float64 [(x - mean) / (std + 1e-8)], like the reference."""
        ms = self.running_ms
        ms._run(x, update)
        xs = np.asarray(x, dtype=np.float64)
        return (xs - ms.mean) / (ms.std + 1e-8)

    # the batched engine keeps one estimate per env; these move the single-env estimate in and out of a B=1 engine
    def to_engine(self, eng):
        ms = self.running_ms
        eng.wf_n.copy_(ms._n)
        eng.wf_mean.copy_(ms._mean)
        eng.wf_S.copy_(ms._S)
        eng.wf_std.copy_(ms._std)

    def from_engine(self, eng):
        ms = self.running_ms
        ms._n.copy_(eng.wf_n)
        ms._mean.copy_(eng.wf_mean)
        ms._S.copy_(eng.wf_S)
        ms._std.copy_(eng.wf_std)


class RewardScaling:
    """DHGN/normalization.py:38-52: x / (std of the running discounted return + 1e-8).  The reference's training loop never
    builds one (it normalises with `Normalization`); same class, same arithmetic, the estimate updated by the f64 Welford entry."""

    def __init__(self, shape, gamma, device=None):
        self.shape, self.gamma = shape, gamma
        self.running_ms = RunningMeanStd(shape=shape, device=device)
        self.R = np.zeros(self.shape)

    def __call__(self, x):
        x = np.asarray(x, dtype=np.float64)
        self.R = self.gamma * self.R + x
        self.running_ms.update_f64(self.R)
        return x / (self.running_ms.std + 1e-8)

    def reset(self):
        self.R = np.zeros(self.shape)
