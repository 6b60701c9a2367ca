"""
Callbacks and replica plumbing (pysages/methods/utils.py:22-216).

`HistogramLogger` keeps the reference's interface but logs on the device: every `period` steps it
copies `state.xi` into a pre-allocated device ring with one async device-to-device copy on the
step's stream (the reference pulls xi to the host and `np.vstack`s, an O(n^2) host loop with a
device sync per sample, methods/utils.py:82-92).  Data reaches the host only when asked for.
"""

from concurrent.futures import Executor, Future

import numpy as np
import torch


class SerialExecutor(Executor):
    """methods/utils.py:22-34"""

    def submit(self, fn, *args, **kwargs):  # pylint: disable=arguments-differ
        future = Future()
        future.set_result(fn(*args, **kwargs))
        return future


class ReplicasConfiguration:
    """methods/utils.py:37-58"""

    def __init__(self, copies: int = 1, executor=SerialExecutor()):
        self.copies = copies
        self.executor = executor


class HistogramLogger:
    """methods/utils.py:61-127"""

    def __init__(self, period: int, offset: int = 0, capacity: int = 4096):
        self.period = period
        self.counter = 0
        self.offset = offset
        self._buf = None
        self._n = 0
        self._capacity = capacity
        self._host = None

    def __call__(self, snapshot, state, timestep):
        self.counter += 1
        if self.counter > self.offset and self.counter % self.period == 0:
            xi = state.xi[0]  # (k,) row, methods/utils.py:88
            if self._buf is None:
                self._buf = torch.empty((self._capacity, xi.shape[0]), dtype=xi.dtype, device=xi.device)
            elif self._n == self._buf.shape[0]:
                grown = torch.empty((2 * self._buf.shape[0], xi.shape[0]), dtype=xi.dtype, device=xi.device)
                grown[: self._n].copy_(self._buf)
                self._buf = grown
            self._buf[self._n].copy_(xi, non_blocking=True)
            self._n += 1
            self._host = None

    @property
    def data(self):
        if self._buf is None:
            return None
        if self._host is None:
            self._host = self._buf[: self._n].cpu().numpy()
        return self._host[0] if self._n == 1 else self._host

    def get_histograms(self, **kwargs):
        data = np.asarray(self.data)
        if "density" not in kwargs:
            kwargs["density"] = True
        return np.histogramdd(data, **kwargs)

    def get_means(self):
        return np.mean(np.asarray(self.data), axis=0)

    def get_cov(self):
        return np.cov(np.asarray(self.data).T)

    def reset(self):
        self.counter = 0
        self._buf, self._n, self._host = None, 0, None

    def numpyfy(self):
        _ = self.data

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_host"] = None if self._buf is None else self._buf[: self._n].cpu().numpy()
        d["_buf"] = None
        return d

    def __setstate__(self, d):
        host = d.pop("_host")
        self.__dict__.update(d)
        self._host = host
        if host is not None:
            self._buf = torch.as_tensor(host)
            self._n = host.shape[0]


class MetaDLogger:
    """methods/utils.py:130-169: appends the last deposited hill to a text file."""

    def __init__(self, hills_file, log_period):
        self.hills_file = hills_file
        self.log_period = log_period
        self.counter = 0

    def save_hills(self, xi, sigma, height):
        with open(self.hills_file, "a+", encoding="utf8") as f:
            f.write(str(self.counter) + "\t")
            f.write("\t".join(map(str, np.asarray(xi).flatten())) + "\t")
            f.write("\t".join(map(str, np.asarray(sigma).flatten())) + "\t")
            f.write(str(float(height)) + "\n")

    def __call__(self, snapshot, state, timestep):
        if self.counter >= self.log_period and self.counter % self.log_period == 0:
            idx = int(state.idx)
            idx = idx - 1 if idx > 0 else 0
            self.save_hills(state.centers[idx].cpu().numpy(), state.sigmas.cpu().numpy(),
                            state.heights[idx].item())
        self.counter += 1


def listify(arg, replicas, name, dtype):
    """methods/utils.py:172-185"""
    if isinstance(arg, list):
        if len(arg) != replicas:
            raise RuntimeError(f"Invalid length for argument {name} (got {len(arg)}, expected {replicas})")
        return arg
    return [dtype(arg) for i in range(replicas)]
