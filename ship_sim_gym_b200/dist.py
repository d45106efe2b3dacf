"""One process per GPU.  Environments are independent units (each ShipEnv owns its own ShipGame,
ship_env.py:32), so the batch shards into contiguous blocks of global env ids with NO data-path collective; the
only exchange is one all-reduce(SUM) of the 16-double episode-statistics vector per rollout (NCCL over
NVLink on GPUs, gloo in the CPU tests)."""
import os

import torch
import torch.distributed as dist


def env_info():
    """(rank, world_size, local_rank) from the torchrun environment (defaults: single process)."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def init(backend=None):
    """Join the process group described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT, if any."""
    rank, world, local = env_info()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            # The statistics all-reduce is a tiny kernel racing a step kernel that fills every SM: on a high-priority
            # stream its one CTA is placed first instead of queueing behind the whole next rollout.
            os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, **kw)
    return rank, world, local


def shard(total_envs, rank, world):
    """Contiguous block [offset, offset+count) of global env ids owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(int(total_envs), int(world))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def all_reduce_stats(stats):
    """In-place SUM over ranks of the float64[16] statistics vector; no-op in a single process."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


class StatsReducer(object):
    """The per-rollout statistics all-reduce, off the critical path: the reduction of rollout i runs on the
    communication stream while the step kernel of rollout i+1 is already executing (the kernel does not depend on it).
    `depth` buffers rotate; `next_buffer` only makes the CURRENT STREAM wait for the all-reduce that last used the
    buffer about to be overwritten -- the host never blocks, and ranks may drift up to depth - 1 rollouts apart, so a
    rank that happens to draw a slow rollout (or starts a few milliseconds late) does not stall the others.  The
    default depth of 32 rollouts is 4 KB of device memory: with 4 buffers a rank could lead by only 3 rollouts (1.4 ms
    at the bench shape), and every start skew beyond that was charged to the early rank (round-1 driver run: 1 -> 8
    GPU efficiency 0.72 with --steps 20)."""

    def __init__(self, device, n=16, depth=32):
        self.bufs = [torch.zeros(n, dtype=torch.float64, device=device) for _ in range(depth)]
        self.work = [None] * depth
        self.i = 0
        self.on = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def next_buffer(self):
        """The buffer the next rollout's local statistics go into (e.g. BatchedShipEnv.stats_tensor(out=...))."""
        k = self.i % len(self.bufs)
        if self.work[k] is not None:
            self.work[k].wait()
            self.work[k] = None
        return self.bufs[k]

    def reduce(self):
        """Start the all-reduce of the buffer handed out by the last next_buffer(); returns that buffer."""
        k = self.i % len(self.bufs)
        self.i += 1
        if self.on:
            self.work[k] = dist.all_reduce(self.bufs[k], op=dist.ReduceOp.SUM, async_op=True)
        return self.bufs[k]

    def submit(self, stats_vec):
        """Copy this rank's float64[16] vector into the next buffer and start its all-reduce."""
        self.next_buffer().copy_(stats_vec)
        return self.reduce()

    def latest(self):
        """Global sums of the most recently submitted rollout (the current stream waits for its all-reduce)."""
        self.wait_all()
        return self.bufs[(self.i - 1) % len(self.bufs)]

    def wait_all(self):
        for k in range(len(self.bufs)):
            if self.work[k] is not None:
                self.work[k].wait()
                self.work[k] = None


def verify_all_reduce(local_vec, reduced_vec, rtol=1e-12):
    """Check one all-reduce(SUM) result against the all-gathered per-rank inputs (every rank calls this; collective).
    Returns (ok, max relative error, per-rank matrix [world, n]).  Counters are integers in float64, so their sums
    are exact; the two real-valued sums may differ in the last bits with the reduction order."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return True, 0.0, local_vec[None].clone()
    parts = [torch.empty_like(local_vec) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, local_vec.contiguous())
    mat = torch.stack(parts)
    want = mat.sum(0)
    err = ((reduced_vec - want).abs() / want.abs().clamp_min(1.0)).max()
    return bool(err <= rtol), float(err), mat


def summarize(stats_vec, names):
    """Global means from the reduced vector (what Curriculum.progress is fed)."""
    v = [float(x) for x in stats_vec.tolist()]
    d = dict(zip(names, v))
    ep = max(d.get("episodes", 0.0), 1.0)
    d["mean_return"] = d.get("return_sum", 0.0) / ep
    d["mean_length"] = d.get("length_sum", 0.0) / ep
    return d
