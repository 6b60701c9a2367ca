"""
Multi-GPU plumbing of the bias path: one process per GPU (`torchrun`), `torch.distributed` over
NCCL (gloo for the CPU tests).  The path shards only across independent simulations:

* umbrella / string windows: `partition_windows` assigns windows to ranks round-robin -- replicas
  only, no data-path collective; `gather_window_means` collects the k doubles per window that
  `analyze` needs (umbrella_integration.py:235-248) at the very end.
* multiple-walker metadynamics sharing hills (extension, absent from the reference): on deposit
  steps every walker contributes one record (centre[k], height, valid); `RecordExchange` all-gathers
  them in rank order on the engine's stream and every rank appends all W records, so all walkers
  hold identical hill lists / grids.  One exchange of (k + 2) doubles per rank every `stride` steps.

`walker_step` is written against a small engine interface (`next_step_deposits`, `walker_begin`,
`walker_finish`, `step`) implemented by `NativeWalkerEngine` (CUDA, through the C ABI) and by the
oracle (tests), so the exchange logic is testable on CPU with gloo.
"""

import torch


def partition_windows(nwindows, world_size, rank):
    """Window indices owned by `rank`: round-robin, so neighbouring (similar-cost) windows spread out."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"invalid rank {rank} for world size {world_size}")
    return list(range(rank, int(nwindows), world_size))


def _dist():
    import torch.distributed as dist

    return dist if (dist.is_available() and dist.is_initialized()) else None


class RecordExchange:
    """All-gather of one hill record per rank, returned as a (W, k + 2) tensor in rank order."""

    def __init__(self, group=None):
        self.group = group
        dist = _dist()
        self.world_size = dist.get_world_size(group) if dist else 1
        self.rank = dist.get_rank(group) if dist else 0
        self._out = None

    def __call__(self, record):
        record = record.reshape(-1)
        if self.world_size == 1:
            return record.reshape(1, -1)
        n = record.numel()
        if self._out is None or self._out.numel() != self.world_size * n or self._out.device != record.device:
            self._out = torch.empty(self.world_size * n, dtype=record.dtype, device=record.device)
        # NCCL enqueues on the current CUDA stream: ordered after walker_begin, before walker_finish
        _dist().all_gather_into_tensor(self._out, record.contiguous(), group=self.group)
        return self._out.view(self.world_size, n)


class NativeWalkerEngine:
    """Two-phase deposit steps of one walker on the GPU (psb_walker_begin / psb_walker_finish)."""

    def __init__(self, native):
        if not native.method.shared_hills:
            raise ValueError("build the method with Metadynamics(..., shared_hills=True)")
        self.native = native
        self.record = torch.zeros(native.record_doubles, dtype=torch.float64, device=native.device)

    def next_step_deposits(self):
        return self.native.next_step_deposits()

    def walker_begin(self, snapshot, timestep=0):
        self.native.walker_begin(snapshot, self.record, timestep)
        return self.record

    def walker_finish(self, snapshot, records):
        records = records.contiguous()
        self.native.walker_finish(snapshot, records, records.shape[0])

    def step(self, snapshot, timestep=0):
        self.native.step(snapshot, timestep)


def walker_step(engine, snapshot, exchange, timestep=0):
    """One bias step of one walker; all walkers must call it for the same step index."""
    if engine.next_step_deposits():
        record = engine.walker_begin(snapshot)
        engine.walker_finish(snapshot, exchange(record))
    else:
        engine.step(snapshot)


def gather_window_means(local_means, nwindows, group=None):
    """{window index: (k,) mean CV} of this rank -> (nwindows, k) tensor on every rank."""
    dist = _dist()
    items = sorted(local_means.items())
    k = len(next(iter(local_means.values()))) if local_means else 0
    if dist is None:
        out = torch.zeros((nwindows, k), dtype=torch.float64)
        for w, m in items:
            out[w] = torch.as_tensor(m, dtype=torch.float64)
        return out
    gathered = [None] * dist.get_world_size(group)
    dist.all_gather_object(gathered, [(int(w), [float(x) for x in m]) for w, m in items], group=group)
    k = max((len(m) for part in gathered for _, m in part), default=0)
    out = torch.zeros((nwindows, k), dtype=torch.float64)
    for part in gathered:
        for w, m in part:
            out[w] = torch.as_tensor(m, dtype=torch.float64)
    return out
