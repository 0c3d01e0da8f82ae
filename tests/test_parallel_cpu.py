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


class _OracleCtx:
    """Host stand-in for native.Context in the gloo tests: every kernel call of the two sharded single-aggregate paths is answered by the CPU
    oracle (numpy), so that the N > 1 HOST logic (slicing, gathers, folds) is what these tests exercise.  "Digests" are indices into the
    message list (the oracle hashes messages itself)."""

    def __init__(self, secpar, A, msgs):
        from oracle import oracle as orc

        self.orc, self.secpar, self.A, self.msgs = orc, secpar, A, msgs
        self.d, self.ell, self.capacity, self.device = orc.param(secpar, "d"), orc.param(secpar, "ell"), orc.param(secpar, "capacity"), torch.device("cpu")

    def hash_msg(self, msgs):
        return torch.arange(len(msgs), dtype=torch.int64).reshape(-1, 1)

    def sort_keys(self, vk_l, vk_r, counts):
        return torch.from_numpy(self.orc.sort_by_repr(self.secpar, vk_l.numpy(), vk_r.numpy()).astype(np.int64))

    def gather_sorted(self, vk_l, vk_r, dg, perm):
        return vk_l[perm].contiguous(), vk_r[perm].contiguous(), dg[perm].contiguous()

    def hash_ch(self, vl, vr, dg):
        out = [self.orc.hash_ch(self.secpar, vl[i].numpy(), vr[i].numpy(), self.msgs[int(dg[i, 0])]) for i in range(vl.shape[0])]
        return torch.from_numpy(np.stack(out)) if out else torch.zeros((0, self.d), dtype=torch.int32)

    def hash_ag_fast(self, vl_s, vr_s, dg_s, c_s, counts):
        alpha, ch = self.orc.hash_ag(self.secpar, vl_s.numpy(), vr_s.numpy(), [self.msgs[int(i)] for i in dg_s[:, 0]], want_ch=True)
        assert np.array_equal(ch, c_s.numpy()), "gathered challenges differ from the oracle's"
        return torch.from_numpy(alpha)

    @staticmethod
    def _cent(x):
        return (((x + H) % Q) - H).astype(np.int32)

    def _mul(self, a, b):
        return self.orc.pointwise_mul(np.ascontiguousarray(a), np.ascontiguousarray(b)).astype(np.int64)

    def aggregate_sum(self, sigs, rows, alpha, counts):
        acc = np.zeros((self.ell, self.d), dtype=np.int64)
        for p, r in enumerate(rows.tolist()):
            acc += self._mul(sigs[r].numpy(), np.broadcast_to(alpha[p].numpy(), (self.ell, self.d)))
        return torch.from_numpy(self._cent(acc)).reshape(1, self.ell, self.d)

    def fold_partials(self, parts, n_parts, n_aggs):
        return _oracle_fold(parts, n_parts, n_aggs)

    def verify_target(self, vl, vr, c, alpha, win_off, n_aggs):
        acc = np.zeros(self.d, dtype=np.int64)
        for i in range(vl.shape[0]):
            t = self._cent(self._mul(vl[i].numpy(), c[i].numpy()) + vr[i].numpy())
            acc += self._mul(t, alpha[i].numpy())
        return torch.from_numpy(self._cent(acc)).reshape(1, self.d)

    def pointwise_add(self, a, b):
        return torch.from_numpy(self._cent(a.numpy().astype(np.int64) + b.numpy()))

    def verify_finish(self, A, target, agg):
        obs = self.orc.matvec(self.secpar, A, agg.numpy().reshape(1, self.ell, self.d))[0]
        if not np.array_equal(obs, target.numpy().reshape(-1)):
            return torch.tensor([3], dtype=torch.uint8)
        coef = self.orc.ntt_inv(self.secpar, agg.numpy().reshape(self.ell, self.d))
        return torch.tensor([4 if int(np.abs(coef.astype(np.int64)).max()) > self.orc.param(self.secpar, "beta_vf") else 0], dtype=torch.uint8)


class _Eng:
    def __init__(self, ctx):
        self.ctx, self.A = ctx, ctx.A


def _one_aggregate_worker(rank: int, world: int, port: int, secpar: int, n: int, out_dir: str):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc

    A = orc.setup_A(secpar, 42)
    seeds = (3 * 10**6 + 2 * np.arange(n)).astype(np.uint64)
    msgs = [f"shard-{i}".encode() for i in range(n)]
    sk_l, sk_r, vl, vr = orc.keygen_batch(secpar, A, seeds)
    sigs = orc.sign_batch(secpar, sk_l, sk_r, vl, vr, msgs)
    want = orc.aggregate(secpar, vl, vr, msgs, sigs)
    eng = _Eng(_OracleCtx(secpar, A, msgs))
    lo, hi = parallel.shard_bounds(n, rank, world)
    t_vl, t_vr = torch.from_numpy(vl), torch.from_numpy(vr)
    agg = parallel.aggregate_one_sharded(eng, t_vl, t_vr, msgs, torch.from_numpy(sigs[lo:hi].copy()))
    assert np.array_equal(agg[0].numpy(), want), "sharded aggregate differs from the oracle's aggregate"
    code = parallel.verify_one_sharded(eng, t_vl, t_vr, msgs, agg)
    assert code.tolist() == [0] and orc.verify(secpar, A, vl, vr, msgs, want) == 0
    bad = agg.clone()
    bad[0, 1, 2] += 1
    assert parallel.verify_one_sharded(eng, t_vl, t_vr, msgs, bad).tolist() == [3]
    np.save(os.path.join(out_dir, f"one{rank}.npy"), np.array([1]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("secpar,n", [(128, 5), (256, 3)])
def test_one_aggregate_sharded_world2(tmp_path, secpar, n):
    """aggregate and verify of ONE aggregate whose signers are split over two ranks."""
    port = _free_port()
    mp.spawn(_one_aggregate_worker, args=(2, port, secpar, n, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "one0.npy") and os.path.exists(tmp_path / "one1.npy")


def _one_aggregate_worker_emulated(rank: int, world: int, port: int, secpar: int, n: int, out_dir: str):
    """The same two sharded paths with NOTHING stood in: the real native.Context / batch.Engine over the library compiled for the CUDA-on-CPU
    emulator (tests/emu_device.py), gloo between the two ranks."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import emu_device

    emu_device.install(emu_device.Permanent)
    from fusion_cryptography_b200.batch import Engine
    from oracle import oracle as orc

    eng = Engine(secpar, seed=42, device=0)
    A = eng.A.numpy()
    assert np.array_equal(A, orc.setup_A(secpar, 42))
    seeds = (3 * 10**6 + 2 * np.arange(n)).astype(np.uint64)
    msgs = [f"shard-{i}".encode() for i in range(n)]
    sk_l, sk_r, vl, vr = orc.keygen_batch(secpar, A, seeds)
    sigs = orc.sign_batch(secpar, sk_l, sk_r, vl, vr, msgs)
    want = orc.aggregate(secpar, vl, vr, msgs, sigs)
    lo, hi = parallel.shard_bounds(n, rank, world)
    t_vl, t_vr = torch.from_numpy(vl), torch.from_numpy(vr)
    agg = parallel.aggregate_one_sharded(eng, t_vl, t_vr, msgs, torch.from_numpy(sigs[lo:hi].copy()))
    assert np.array_equal(agg[0].numpy(), want), "sharded aggregate differs from the oracle's aggregate"
    assert parallel.verify_one_sharded(eng, t_vl, t_vr, msgs, agg).tolist() == [0]
    bad = agg.clone()
    bad[0, 1, 2] += 1
    assert parallel.verify_one_sharded(eng, t_vl, t_vr, msgs, bad).tolist() == [3]
    np.save(os.path.join(out_dir, f"emu{rank}.npy"), np.array([1]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("secpar,n", [(128, 5), (256, 3)])
def test_one_aggregate_sharded_world2_on_the_emulated_device(tmp_path, secpar, n):
    """SURVEY 8(e) rows 3 and 4 through the product's own kernels and host code on two gloo ranks (no stand-in)."""
    import hostlib

    os.environ["FZ_EMU_LIBRARY"] = hostlib.load_emulated_library()._name     # built once here; the two ranks load it
    port = _free_port()
    mp.spawn(_one_aggregate_worker_emulated, args=(2, port, secpar, n, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "emu0.npy") and os.path.exists(tmp_path / "emu1.npy")
