"""
Multi-GPU plumbing of the bias path: one process per GPU (`torchrun`), `torch.distributed` over
NCCL (gloo for the CPU tests).  The path shards only across independent simulations:

* umbrella / string windows: `partition_windows` assigns windows to ranks round-robin -- replicas
  only, no data-path collective; `gather_window_means` collects the k doubles per window that
  `analyze` needs (umbrella_integration.py:235-248) at the very end.  The windows of ONE rank step in
  lockstep through `WindowBatch`: one cooperative launch per MD step for all of them (psb_batch_step)
  instead of one launch, stream and per-step latency chain per window.
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


class WindowBatch:
    """The windows (independent simulations) of one process stepped by ONE launch per MD step (psb_batch_*,
    include/pysages_b200.h).  `natives` are the `update.native` objects of the windows' built methods; every
    window keeps its own handle and state, results are bit-identical to stepping it alone."""

    def __init__(self, natives):
        import ctypes

        from . import _native as N

        natives = list(natives)
        if not natives:
            raise ValueError("a batch needs at least one simulation")
        self.natives = natives
        self.lib = natives[0].lib
        self.device = natives[0].device
        handles = (ctypes.c_void_p * len(natives))(*[n.handle for n in natives])
        out = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            N.check(self.lib.psb_batch_create(handles, len(natives), ctypes.byref(out)))
        self.handle = out
        self._descs = (N.SnapshotDesc * len(natives))()

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                self.lib.psb_batch_destroy(h)
            except Exception:  # interpreter shutdown
                pass

    def _stream(self):
        import ctypes

        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def describe(self, snapshots, out=None, offset=0):
        from .backends.snapshot import describe_snapshot

        out = self._descs if out is None else out
        for i, (nat, snap) in enumerate(zip(self.natives, snapshots)):
            describe_snapshot(snap, nat.layout, out[offset + i])
        return out

    def step(self, snapshots, timestep=0):
        """One bias step of every window: snapshots[i] belongs to natives[i]."""
        from . import _native as N

        N.check(self.lib.psb_batch_step(self.handle, self.describe(snapshots), int(timestep), self._stream()))
        for nat in self.natives:
            nat.ncalls += 1

    def step_described(self, timestep=0):
        """psb_batch_step on descriptors the caller keeps current in `self._descs` (adapters: raw pointers)."""
        from . import _native as N

        N.check(self.lib.psb_batch_step(self.handle, self._descs, int(timestep), self._stream()))
        for nat in self.natives:
            nat.ncalls += 1

    def ring(self, snapshot_rows):
        """(nsnaps, n) descriptor array for run_cycle: snapshot_rows[r][i] is window i at ring position r."""
        from . import _native as N

        n = len(self.natives)
        arr = (N.SnapshotDesc * (n * len(snapshot_rows)))()
        for r, row in enumerate(snapshot_rows):
            self.describe(row, arr, r * n)
        return arr

    def run_cycle(self, ring, nsnaps, nsteps):
        from . import _native as N

        N.check(self.lib.psb_batch_run_cycle(self.handle, ring, int(nsnaps), int(nsteps), self._stream()))
        for nat in self.natives:
            nat.ncalls += int(nsteps)

    @property
    def launches(self):
        return int(self.lib.psb_batch_launch_count(self.handle))


def run_lockstep(sampling_contexts, timesteps):
    """Drive several bound simulations one MD step at a time, all bias steps of a step in one launch.  Needs
    engines that can be stepped from outside their own run loop (`MockEngine`: `_advance()`); engines that call
    the hook from inside their run loop (HOOMD, OpenMM, LAMMPS) would need one host thread each and a rendezvous
    per step, which is not provided: run those per window."""
    samplers = [c.sampler for c in sampling_contexts]
    engines = [c.context for c in sampling_contexts]
    for e in engines:
        if not hasattr(e, "_advance"):
            raise NotImplementedError("lockstep windows need engines that can be advanced one step from outside")
    for s in samplers:
        if not s._fused:
            raise NotImplementedError("lockstep windows need methods whose update is the fused native step only")
    batch = WindowBatch([s._native for s in samplers])
    for _ in range(int(timesteps)):
        for e in engines:
            e._advance()
        views = [s._views() for s in samplers]
        batch.step(views, engines[0].timestep)
        for s, e, v in zip(samplers, engines, views):
            s._state_stale = True
            s.after_step()
            if s.callback:
                s.callback(v, s.state, e.timestep)
            e.timestep += 1
    return batch
