"""CPU, world_size 2, gloo: the host-side multi-process logic (ship_sim_gym_b200/dist.py) -- contiguous sharding by
global env id, and the one collective the path has (all-reduce of the 16-double episode-statistics vector).  The
per-rank "engine" here is the float64 oracle (a GPU is needed for the product kernel); what is under test is that
sharded runs + all-reduce reproduce the single-process run, i.e. that env ids / RNG keys are global."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_ENVS, K, N_SCEN, SEED = 96, 40, 8, 5


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _run_shard(offset, count):
    import oracle
    from ship_sim_gym_b200 import ScenarioBank
    bank = ScenarioBank.generate(N_SCEN, (600, 600), seed=SEED).as_dict()
    env = oracle.OracleEnv(count, bank, auto_reset=True, seed=SEED, env_id_offset=offset)
    env.reset()
    out = env.step(None, K=K, want=("obs", "reward", "done"))
    stats = np.zeros(16)
    stats[:len(env.stats)] = env.stats
    return out, stats


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from ship_sim_gym_b200 import dist as sdist
    r, w, _ = sdist.init("gloo")
    assert (r, w) == (rank, world) and dist.is_initialized()
    off, cnt = sdist.shard(N_ENVS, r, w)
    out, stats = _run_shard(off, cnt)
    t = torch.from_numpy(stats.copy())
    sdist.all_reduce_stats(t)
    # the overlapped reducer (what bench.py uses): three submissions through two alternating buffers
    red = sdist.StatsReducer("cpu")
    for i in range(3):
        red.submit(torch.from_numpy(stats * (i + 1)))
    assert torch.allclose(red.latest(), t * 3)
    q.put((rank, off, cnt, out["reward"].sum(), out["done"].sum(), out["obs"][-1].copy(), t.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_covers_every_env_once():
    from ship_sim_gym_b200 import dist as sdist
    for total in (1, 7, 96, 4096, 1048576):
        for world in (1, 2, 3, 8):
            spans = [sdist.shard(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (o0, c0), (o1, _) in zip(spans, spans[1:]):
                assert o0 + c0 == o1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_all_reduce_is_a_noop_in_a_single_process():
    from ship_sim_gym_b200 import dist as sdist
    t = torch.arange(16, dtype=torch.float64)
    assert torch.equal(sdist.all_reduce_stats(t.clone()), t)
    d = sdist.summarize(torch.tensor([4.0, -2.0, 100.0] + [0.0] * 13), ("episodes", "return_sum", "length_sum"))
    assert d["mean_return"] == -0.5 and d["mean_length"] == 25.0


@pytest.mark.timeout(300)
def test_two_ranks_reproduce_the_single_process_run():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full, full_stats = _run_shard(0, N_ENVS)
    # every rank holds the same reduced vector, equal to the single-process statistics
    for _, _, _, _, _, _, red in res:
        np.testing.assert_allclose(red, full_stats, rtol=0, atol=1e-9)
    assert full_stats[0] > 0                                         # episodes did finish
    # and the shards are exactly the corresponding slices of the full run (RNG keyed by GLOBAL env id)
    for _, off, cnt, rsum, dsum, last_obs, _ in res:
        assert rsum == pytest.approx(full["reward"][:, off:off + cnt].sum(), abs=1e-9)
        assert dsum == full["done"][:, off:off + cnt].sum()
        np.testing.assert_array_equal(last_obs, full["obs"][-1, off:off + cnt])
