"""Multi-GPU sharding of the hot path: one process per GPU, ``torch.distributed`` (NCCL over NVLink on
the B200 box; gloo in the CPU tests) for the plumbing.

SURVEY 8(e): NTT batches, keygen, sign and verify-of-many-aggregates are embarrassingly parallel -- they
shard by polynomial / signer / aggregate index with NO data-path collective.  The one exchange step is
the aggregation of a single large aggregate whose signatures are spread over the ranks:

    1. every rank hashes its slice of the (sorted) signers' challenges            (no communication)
    2. all-gather of the challenges c_hat (1 KiB per signer)                      (NCCL all_gather)
    3. the aggregation hash is one serial sponge: every rank runs it redundantly  (no communication)
    4. every rank forms the partial sum over ITS signers                          (83 KiB read per signer)
    5. all-gather of the world_size partial aggregates (ell*d*4 B each) and a modular fold

Centred partials are below 2^30 in magnitude, so an int32 ncclSum over 8 ranks could overflow; the
partials are gathered and folded by fz_fold_partials instead (exact, and cheaper than a 64-bit reduce).

verify_one_sharded has the same shape: steps 1-3, then every rank forms the partial target
sum_i (vk_l_i * c_i + vk_r_i) * alpha_i over ITS slice of the sorted signers (fusion.py:706-714; d values), the
world_size partial targets are gathered and added modulo q, and fusion.py:715-728 runs on the complete target.

The serial sponge of step 3 (279 K dependent Keccak-f for a capacity aggregate at secpar 256) is the latency floor of
both: sharding one aggregate spreads its DATA (83 KiB per signature) and its per-signer work, not its hash.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition of n items: ranks [0, n % world) get one extra item."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_counts(counts: Sequence[int], rank: int, world: int) -> Tuple[int, int, int, int]:
    """Round-robin-free block partition of whole aggregates: returns (j0, j1, p0, p1) = aggregate and position ranges."""
    j0, j1 = shard_bounds(len(counts), rank, world)
    p0 = int(sum(counts[:j0]))
    return j0, j1, p0, p0 + int(sum(counts[j0:j1]))


def all_gather_rows(t: torch.Tensor, sizes: Sequence[int], group=None) -> torch.Tensor:
    """all-gather of row blocks of different length (sizes[r] rows from rank r): pads to the largest block."""
    world = dist.get_world_size(group)
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[: sizes[r]] for r, b in enumerate(bufs)], dim=0)


def reduce_partial_aggregates(partial: torch.Tensor, fold: Callable[[torch.Tensor, int, int], torch.Tensor], group=None) -> torch.Tensor:
    """partial: [n_aggs, ell, d] this rank's partial sums.  Gathers every rank's partials and folds them modulo q.
    fold(parts [n_aggs, world, ell, d], n_parts, n_aggs) -> [n_aggs, ell, d]  (Context.fold_partials on the GPU)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return partial
    bufs = [torch.empty_like(partial) for _ in range(world)]
    dist.all_gather(bufs, partial.contiguous(), group=group)
    parts = torch.stack(bufs, dim=1).contiguous()  # [n_aggs, world, ell, d]
    return fold(parts, world, partial.shape[0])


def aggregate_one_sharded(engine, vk_l: torch.Tensor, vk_r: torch.Tensor, msgs: Sequence[bytes], local_sigs: torch.Tensor,
                          group=None) -> torch.Tensor:
    """BASELINE config[3]: ONE aggregate of N signers whose signatures are sharded by signer index over the ranks
    (rank r holds local_sigs = sigs[lo_r:hi_r] in the ORIGINAL signer order); all ranks hold all keys and messages.
    Returns the full aggregate [1, ell, d] on every rank."""
    ctx = engine.ctx
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = vk_l.shape[0]
    # (1) challenges for this rank's slice of the sorted positions, (2) all-gather, (3) serial aggregation hash on every rank
    perm, _, _, _, alpha, _ = _sorted_challenges(ctx, vk_l, vk_r, msgs, rank, world, group)
    # (4) partial sum over the signers whose signatures live here (original index in [lo, hi))
    lo, hi = shard_bounds(n, rank, world)
    mine = (perm >= lo) & (perm < hi)
    pos = torch.nonzero(mine).flatten()
    if pos.numel():
        rows = (perm[pos] - lo).to(torch.int32).contiguous()
        partial = ctx.aggregate_sum(local_sigs, rows, alpha[pos].contiguous(), [int(pos.numel())])
    else:
        partial = torch.zeros((1, ctx.ell, ctx.d), dtype=torch.int32, device=ctx.device)
    # (5) gather + fold
    return reduce_partial_aggregates(partial, ctx.fold_partials, group)


def _sorted_challenges(ctx, vk_l, vk_r, msgs, rank: int, world: int, group=None):
    """Steps 1-3 shared by the two sharded paths: sort, this rank's slice of the challenges, all-gather, serial hash."""
    n = vk_l.shape[0]
    dg = ctx.hash_msg(msgs)
    perm = ctx.sort_keys(vk_l, vk_r, [n])                              # sorted position -> original signer
    vl_s, vr_s, dg_s = ctx.gather_sorted(vk_l, vk_r, dg, perm)
    s0, s1 = shard_bounds(n, rank, world)
    c_loc = ctx.hash_ch(vl_s[s0:s1], vr_s[s0:s1], dg_s[s0:s1])
    if world > 1:
        sizes = [shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0] for r in range(world)]
        c_s = all_gather_rows(c_loc, sizes, group)
    else:
        c_s = c_loc
    alpha = ctx.hash_ag_fast(vl_s, vr_s, dg_s, c_s, [n])
    return perm, vl_s, vr_s, c_s, alpha, (s0, s1)


def verify_one_sharded(engine, vk_l: torch.Tensor, vk_r: torch.Tensor, msgs: Sequence[bytes], agg: torch.Tensor, group=None) -> torch.Tensor:
    """SURVEY 8(e) row 4 / fusion.py:680-728 for ONE aggregate of N signers whose target terms are sharded over the ranks:
    rank r owns the sorted signer positions shard_bounds(N, r, world).  All ranks hold the keys, the messages and the
    aggregate signature agg [1, ell, d].  Returns the uint8 code [1] (batch.VERIFY_REASONS) on every rank.
    The argument checks of fusion.py:686-689 are the caller's (as in Engine.verify)."""
    ctx = engine.ctx
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = vk_l.shape[0]
    if n > ctx.capacity:
        raise ValueError("verify_one_sharded: Too many keys.")
    if len(msgs) != n:
        raise ValueError("verify_one_sharded: Number of keys and messages must be equal.")
    perm, vl_s, vr_s, c_s, alpha, (s0, s1) = _sorted_challenges(ctx, vk_l, vk_r, msgs, rank, world, group)
    win_off = torch.tensor([0, s1 - s0], dtype=torch.int64, device=vl_s.device)
    if s1 > s0:
        part = ctx.verify_target(vl_s[s0:s1].contiguous(), vr_s[s0:s1].contiguous(), c_s[s0:s1].contiguous(), alpha[s0:s1].contiguous(), win_off, 1)
    else:
        part = torch.zeros((1, ctx.d), dtype=torch.int32, device=vl_s.device)
    if world > 1:
        bufs = [torch.empty_like(part) for _ in range(world)]
        dist.all_gather(bufs, part.contiguous(), group=group)
        target = bufs[0]
        for b in bufs[1:]:
            target = ctx.pointwise_add(target, b)                     # exact: every partial is a centred residue
    else:
        target = part
    return ctx.verify_finish(engine.A, target, agg)
