"""World-size-2 (and 3) gloo tests of the multi-GPU host logic: cell-balanced sharding of the pair
list, variable-length all_gather of score slices, scatter into the score matrix."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_score(pairs):
    p = np.asarray(pairs, dtype=np.int64)
    return ((p[:, 0] * 131 + p[:, 1] * 7) % 1000).astype(np.float32) / 8.0


def _worker(rank, world, port, tmp, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    os.chdir(tmp)
    import torch.distributed as dist
    from acoss_b200.distributed import all_pairwise_distributed
    from acoss_b200.serra09 import Serra09
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    feats = [dict(hpcp=rng.random((int(n), 12)).astype(np.float32), label=str(i // 3))
             for i, n in enumerate(rng.integers(20, 200, size=17))]
    alg = Serra09(None, None, features=feats, downsample_fac=1, shortname="r%d" % rank, cachedir="cache%d" % rank)
    bounds = all_pairwise_distributed(alg, symmetric=True, score_fn=_fake_score)
    q.put((rank, np.array(alg.Ds["main"]), bounds))
    dist.destroy_process_group()


def _worker_shared(rank, world, port, tmp, q):
    """Every rank constructs the plugin with the SAME cache prefix: all ranks map the same memmap file."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    os.chdir(tmp)
    import torch.distributed as dist
    from acoss_b200.distributed import all_pairwise_distributed
    from acoss_b200.serra09 import Serra09
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    feats = [dict(hpcp=rng.random((int(n), 12)).astype(np.float32), label=str(i // 3))
             for i, n in enumerate(rng.integers(20, 200, size=17))]
    alg = Serra09(None, None, features=feats, downsample_fac=1, shortname="shared", cachedir="cache_shared")
    dist.barrier()                                          # every constructor has truncated the file by now
    all_pairwise_distributed(alg, symmetric=True, score_fn=_fake_score)
    dist.barrier()                                          # every rank has written
    q.put((rank, np.array(alg.Ds["main"])))
    dist.destroy_process_group()


def test_distributed_shared_cache_prefix_is_not_doubled(tmp_path):
    """ranks that share one memmap file must not symmetrise it twice (plain idempotent assignment)."""
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_shared, args=(r, world, port, str(tmp_path), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = 17
    i, j = np.triu_indices(n, k=1)
    want = np.zeros((n, n), np.float32)
    want[i, j] = _fake_score(np.stack([i, j], 1))
    want = want + want.T
    for rank, D in res:
        assert np.array_equal(D, want)


_CALLS = []


def _worker_resume(rank, world, port, tmp, q):
    """Tile checkpoints: a second run of the same job scores nothing (all tiles resumed), a different job ignores them."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    os.chdir(tmp)
    import torch.distributed as dist
    from acoss_b200.distributed import all_pairwise_distributed
    from acoss_b200.serra09 import Serra09
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    feats = [dict(hpcp=rng.random((int(n), 12)).astype(np.float32), label=str(i // 3))
             for i, n in enumerate(rng.integers(20, 200, size=17))]
    calls = []

    def counting(p):
        calls.append(len(p))
        return _fake_score(p)
    out = []
    for run in range(2):
        alg = Serra09(None, None, features=feats, downsample_fac=1, shortname="res%d" % rank, cachedir="cacheres%d" % rank,
                      tile_pairs=16)
        tm = {}
        all_pairwise_distributed(alg, symmetric=True, score_fn=counting, checkpoint_dir="ckpt", timings=tm)
        out.append((np.array(alg.Ds["main"]), tm["resumed_tiles"], tm["tiles"], sum(calls)))
        calls.clear()
    q.put((rank, out))
    dist.destroy_process_group()


def test_distributed_tile_resume(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_resume, args=(r, world, port, str(tmp_path), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = 17
    i, j = np.triu_indices(n, k=1)
    want = np.zeros((n, n), np.float32)
    want[i, j] = _fake_score(np.stack([i, j], 1))
    want = want + want.T
    for rank, out in res:
        (D1, res1, tiles1, scored1), (D2, res2, tiles2, scored2) = out
        assert np.array_equal(D1, want) and np.array_equal(D2, want)
        assert res1 == 0 and scored1 > 0 and tiles1 >= 4        # first run scores every tile
        assert res2 == tiles2 == tiles1 and scored2 == 0         # second run resumes all of them


def _fake_score4(pairs):
    p = np.asarray(pairs, dtype=np.int64)
    base = ((p[:, 0] * 37 + p[:, 1] * 11) % 500).astype(np.float32)
    return np.stack([base + 0.1 * k for k in range(4)]).astype(np.float32)


def _worker_ef(rank, world, port, tmp, q):
    """Multi-key plugin (EarlyFusion: four score rows per pair) through the same sharding + single gather."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    os.chdir(tmp)
    import torch.distributed as dist
    from acoss_b200.distributed import all_pairwise_distributed
    from acoss_b200.earlyfusion import EarlyFusion
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(4)
    feats = []
    for i, n in enumerate(rng.integers(12, 60, size=11)):
        feats.append(dict(mfccs=np.zeros((int(n), 4), np.float32), ssms=np.zeros((int(n), 3), np.float32),
                          chromas=np.zeros((int(n), 12), np.float32), chroma_med=np.zeros(12), label=str(i // 2)))
    alg = EarlyFusion(None, None, features=feats, shortname="ef%d" % rank, cachedir="cacheef%d" % rank)
    bounds = all_pairwise_distributed(alg, symmetric=True, score_fn=_fake_score4)
    q.put((rank, {k: np.array(v) for k, v in alg.Ds.items()}, bounds, alg.pair_weights(alg._pair_array(True))))
    dist.destroy_process_group()


def test_distributed_multikey_gloo(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_ef, args=(r, world, port, str(tmp_path), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = 11
    i, j = np.triu_indices(n, k=1)
    sc = _fake_score4(np.stack([i, j], 1))
    for rank, Ds, bounds, w in res:
        assert list(Ds) == ["mfccs", "ssms", "chromas", "early"]
        for k, key in enumerate(Ds):
            want = np.zeros((n, n), np.float32)
            want[i, j] = sc[k]
            assert np.array_equal(Ds[key], want + want.T), key
        loads = [w[bounds[r]:bounds[r + 1]].sum() for r in range(world)]
        assert abs(loads[0] - loads[1]) <= 2 * w.max()      # balanced by cross-similarity cells


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_all_pairwise_gloo(tmp_path, world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, str(tmp_path), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = 17
    i, j = np.triu_indices(n, k=1)
    want = np.zeros((n, n), np.float32)
    want[i, j] = _fake_score(np.stack([i, j], 1))
    want = want + want.T
    for rank, D, bounds in res:
        assert np.array_equal(D, want)                      # byte-identical gathered matrix on every rank
        assert bounds[0] == 0 and bounds[-1] == n * (n - 1) // 2 and (np.diff(bounds) > 0).all()


def _worker_fail(rank, world, port, tmp, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    os.chdir(tmp)
    import torch.distributed as dist
    from acoss_b200.distributed import all_pairwise_distributed
    from acoss_b200.serra09 import Serra09
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    feats = [dict(hpcp=rng.random((int(n), 12)).astype(np.float32), label=str(i // 3))
             for i, n in enumerate(rng.integers(20, 200, size=17))]
    alg = Serra09(None, None, features=feats, downsample_fac=1, shortname="f%d" % rank, cachedir="cachef%d" % rank)

    def score(pairs):
        if rank == 1:
            raise RuntimeError("device lost")              # what an ACOSS_E_CUDA return code turns into
        return _fake_score(pairs)

    try:
        all_pairwise_distributed(alg, symmetric=True, score_fn=score)
        q.put((rank, "ok"))
    except RuntimeError as e:
        q.put((rank, str(e)))
        dist.destroy_process_group()
        sys.exit(3)
    dist.destroy_process_group()


def test_distributed_rank_failure_surfaces_everywhere(tmp_path):
    """A scoring failure on one rank raises on EVERY rank before the gather ('rank-failure surfaces as
    non-zero rc'): no rank hangs in the collective, every process exits non-zero."""
    import multiprocessing as mp
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_fail, args=(r, world, port, str(tmp_path), q)) for r in range(world)]
    for p in procs:
        p.start()
    msgs = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 3
    assert "failed on rank 1" in msgs[1] and "device lost" in msgs[1]
    assert "another rank" in msgs[0]


def test_triangle_shard_equals_list_split():
    """The list-free split of the upper triangle gives the bounds and the slices of shard_bounds over the full pair list."""
    from acoss_b200.distributed import shard_bounds, triangle_shard
    rng = np.random.default_rng(0)
    for n in (2, 3, 5, 17, 160, 700):
        a = rng.integers(1, 2500, size=n).astype(np.int64)
        i, j = np.triu_indices(n, 1)
        pairs = np.stack([i, j], 1)
        w = (a[i] * a[j]).astype(np.float64)
        for world in (1, 2, 3, 8):
            b0 = shard_bounds(w, world)
            for r in range(world):
                b, mine = triangle_shard(a, world, r)
                assert np.array_equal(b, b0)
                assert np.array_equal(mine, pairs[b0[r]:b0[r + 1]])


def test_shard_bounds_balance():
    from acoss_b200.distributed import shard_bounds
    rng = np.random.default_rng(0)
    w = rng.integers(1, 100, size=1000).astype(np.float64)
    for world in (1, 2, 4, 8):
        b = shard_bounds(w, world)
        assert len(b) == world + 1 and b[0] == 0 and b[-1] == 1000 and (np.diff(b) >= 0).all()
        loads = [w[b[r]:b[r + 1]].sum() for r in range(world)]
        assert max(loads) - min(loads) <= 2 * w.max()
    assert list(shard_bounds([], 4)) == [0, 0, 0, 0, 0]
    assert list(shard_bounds([5.0], 2)) in ([0, 0, 1], [0, 1, 1])
