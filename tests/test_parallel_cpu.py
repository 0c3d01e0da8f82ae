"""N > 1 host logic on CPU: world_size-2 gloo run of the sharding / gather / fold plumbing in
fusion_cryptography_b200.parallel, with the CPU oracle standing in for the kernels (no GPU here)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fusion_cryptography_b200 import parallel

Q = 2147465729
H = (Q - 1) // 2


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 9, 2818, 10**6):
        for world in (1, 2, 3, 4, 8):
            spans = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    counts = [3, 1, 4, 1, 5, 9, 2, 6]
    seen = []
    for r in range(3):
        j0, j1, p0, p1 = parallel.shard_counts(counts, r, 3)
        assert p1 - p0 == sum(counts[j0:j1])
        seen += list(range(j0, j1))
    assert seen == list(range(len(counts)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_fold(parts: torch.Tensor, n_parts: int, n_aggs: int) -> torch.Tensor:
    s = parts.to(torch.int64).sum(dim=1)
    return (((s + H) % Q) - H).to(torch.int32)


def _worker(rank: int, world: int, port: int, secpar: int, n: int, out_dir: str):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc

    d, ell = orc.param(secpar, "d"), orc.param(secpar, "ell")
    rng = np.random.default_rng(1234)  # same data on every rank; each rank uses its shard
    sigs = rng.integers(-H, H + 1, size=(n, ell, d)).astype(np.int32)
    alpha = rng.integers(-H, H + 1, size=(n, d)).astype(np.int32)
    lo, hi = parallel.shard_bounds(n, rank, world)
    # this rank's partial sum over its signers (oracle arithmetic)
    acc = np.zeros((ell, d), dtype=np.int64)
    for i in range(lo, hi):
        acc += orc.pointwise_mul(sigs[i], np.broadcast_to(alpha[i], (ell, d))).astype(np.int64)
    partial = torch.from_numpy((((acc + H) % Q) - H).astype(np.int32)).reshape(1, ell, d)
    total = parallel.reduce_partial_aggregates(partial, _oracle_fold)
    # ragged all-gather of per-signer rows
    rows = torch.from_numpy(alpha[lo:hi].copy())
    sizes = [parallel.shard_bounds(n, r, world)[1] - parallel.shard_bounds(n, r, world)[0] for r in range(world)]
    gathered = parallel.all_gather_rows(rows, sizes)
    assert np.array_equal(gathered.numpy(), alpha)
    want = np.zeros((ell, d), dtype=np.int64)
    for i in range(n):
        want += orc.pointwise_mul(sigs[i], np.broadcast_to(alpha[i], (ell, d))).astype(np.int64)
    want = (((want + H) % Q) - H).astype(np.int32)
    assert np.array_equal(total[0].numpy(), want)
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([1]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("secpar,n", [(128, 5), (256, 7)])
def test_partial_aggregate_reduce_world2(tmp_path, secpar, n):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, secpar, n, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0.npy") and os.path.exists(tmp_path / "ok1.npy")
