"""Batched device pipelines of the Fusion hot path (the throughput API next to the drop-in
``fusion.fusion`` single-item API).  Everything stays in HBM as int32 ``[batch, rows, d]`` tensors;
every step is a kernel of libfusion_b200 called through the C ABI.

    eng = Engine(256, seed=42)                       # fusion_setup(256, 42)
    keys = eng.keygen(seeds)                         # fusion.py:338-373, n keys at once
    sigs = eng.sign(keys, msgs)                      # fusion.py:534-557
    aggs = eng.aggregate(keys.vk_l, keys.vk_r, msgs, sigs, counts, sel)    # fusion.py:655-677, many aggregates
    codes = eng.verify(keys.vk_l, keys.vk_r, msgs, aggs, counts, sel)      # fusion.py:680-728
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import native

VERIFY_REASONS = {
    0: "",
    1: "Too many keys.",
    2: "Number of keys and messages must be equal.",
    3: "Target doesn't match image of aggregate signature.",
    4: "Norm of aggregate signature too large.",
    5: "Weight of aggregate signature too large.",
}


@dataclass
class KeyBatch:
    sk_hat: torch.Tensor  # [n, 2, d]: the one distinct row of sk_l / sk_r (all ell rows of a seeded key are equal)
    vk_l: torch.Tensor    # [n, d]
    vk_r: torch.Tensor    # [n, d]
    sk_l: Optional[torch.Tensor] = None  # [n, ell, d] when materialised
    sk_r: Optional[torch.Tensor] = None


class Engine:
    def __init__(self, secpar: int, seed: Optional[int] = 42, device: Optional[int] = None, A: Optional[torch.Tensor] = None):
        self.ctx = native.get_context(secpar, device)
        self.secpar, self.d, self.ell = secpar, self.ctx.d, self.ctx.ell
        if A is None:
            # fusion.py:273-282: the 1 x ell public matrix, every column sampled with the same seed
            A = self.ctx.sample_ntt([seed])[0].reshape(1, self.d).expand(self.ell, self.d).contiguous()
        self.A = self.ctx._dev(A).reshape(self.ell, self.d)

    # ------------------------------------------------------------------ keygen / sign
    def keygen(self, seeds, want_sk: bool = False) -> KeyBatch:
        sk_l, sk_r, sk_hat, vk_l, vk_r = self.ctx.keygen(self.A, seeds, want_sk=want_sk)
        return KeyBatch(sk_hat=sk_hat, vk_l=vk_l, vk_r=vk_r, sk_l=sk_l, sk_r=sk_r)

    def challenges(self, vk_l: torch.Tensor, vk_r: torch.Tensor, msgs: Sequence[bytes]) -> torch.Tensor:
        return self.ctx.hash_ch(vk_l, vk_r, self.ctx.hash_msg(msgs))

    def sign(self, keys: KeyBatch, msgs: Sequence[bytes], out: Optional[torch.Tensor] = None) -> torch.Tensor:
        c_hat = self.challenges(keys.vk_l, keys.vk_r, msgs)
        if keys.sk_l is not None and keys.sk_r is not None:
            return self.ctx.sign_mac(keys.sk_l, keys.sk_r, c_hat, out=out)   # general: ell distinct rows
        return self.ctx.sign_mac_rows(keys.sk_hat, c_hat, out=out)

    def keygen_sign(self, seeds, msgs: Sequence[bytes], chunk: int = 16384, keep_sigs: bool = True):
        """BASELINE config[2]: keygen + sign of many one-time keys; secret keys never reach HBM as ell rows.
        Returns (vk_l, vk_r, sigs or None); with keep_sigs=False signatures of each chunk overwrite one buffer."""
        import numpy as np

        seeds = np.asarray(seeds, dtype=np.uint64)
        n = len(seeds)
        vk_l = torch.empty((n, self.d), dtype=torch.int32, device=self.ctx.device)
        vk_r = torch.empty((n, self.d), dtype=torch.int32, device=self.ctx.device)
        sigs = torch.empty((n if keep_sigs else min(chunk, n), self.ell, self.d), dtype=torch.int32, device=self.ctx.device)
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            kb = self.keygen(seeds[lo:hi])
            vk_l[lo:hi], vk_r[lo:hi] = kb.vk_l, kb.vk_r
            self.sign(kb, msgs[lo:hi], out=sigs[lo:hi] if keep_sigs else sigs[: hi - lo])
        return vk_l, vk_r, (sigs if keep_sigs else None)

    # ------------------------------------------------------------------ aggregate / verify
    def _sorted_hash(self, vk_l, vk_r, digests_pos, counts, sel):
        """sort -> gather -> hash_ch -> hash_ag for a chunk of aggregates; position p uses key row sel[p] (identity if None)."""
        perm = self.ctx.sort_keys(vk_l, vk_r, counts, sel=sel)
        vl_s, vr_s, dg_s = self.ctx.gather_sorted(vk_l, vk_r, digests_pos, perm, sel=sel)
        c_s = self.ctx.hash_ch(vl_s, vr_s, dg_s)
        alpha = self.ctx.hash_ag(vl_s, vr_s, dg_s, c_s, counts)
        return perm, vl_s, vr_s, c_s, alpha

    @staticmethod
    def _chunks(counts: Sequence[int], max_positions: int):
        lo, acc, start = 0, 0, 0
        for j, c in enumerate(counts):
            if j > lo and acc + c > max_positions:
                yield lo, j, start, start + acc
                lo, start, acc = j, start + acc, 0
            acc += c
        if len(counts) > lo:
            yield lo, len(counts), start, start + acc

    def aggregate(self, vk_l, vk_r, msgs: Optional[Sequence[bytes]], sigs, counts: Sequence[int], sel: Optional[torch.Tensor] = None,
                  sig_sel: Optional[torch.Tensor] = None, max_positions: int = 1 << 22, digests: Optional[torch.Tensor] = None) -> torch.Tensor:
        """counts[j] signers per aggregate; position p (0 <= p < sum(counts)) signs msgs[p] with key row sel[p] and
        signature row sig_sel[p] (both default to p).  `digests` ([positions, 32] prehashes from hash_msg) may replace msgs.
        Returns [n_aggs, ell, d]."""
        total = int(sum(counts))
        if any(c <= 0 for c in counts):
            raise IndexError("list index out of range")  # the reference indexes sorted_signatures[0] (fusion.py:670)
        if (digests.shape[0] if digests is not None else len(msgs)) != total:
            raise ValueError("one message per position is required")
        dev = self.ctx.device
        out = torch.empty((len(counts), self.ell, self.d), dtype=torch.int32, device=dev)
        dg_all = digests if digests is not None else self.ctx.hash_msg(msgs)
        for j0, j1, p0, p1 in self._chunks(counts, max_positions):
            cs = list(counts[j0:j1])
            s = None if sel is None else sel[p0:p1].to(dev)
            # without sel, position p of this chunk is key row p0 + p
            s_eff = s if s is not None else torch.arange(p0, p1, dtype=torch.int32, device=dev)
            perm, _, _, _, alpha = self._sorted_hash(vk_l, vk_r, dg_all[p0:p1], cs, s_eff)
            zsel = (sig_sel[p0:p1].to(dev) if sig_sel is not None else torch.arange(p0, p1, dtype=torch.int32, device=dev))
            # perm is relative to the chunk: map sorted position -> signature row
            sig_rows = zsel[perm.long()].to(torch.int32).contiguous()
            out[j0:j1] = self.ctx.aggregate_sum(sigs, sig_rows, alpha, cs)
        return out

    def verify(self, vk_l, vk_r, msgs: Optional[Sequence[bytes]], aggs, counts: Sequence[int], sel: Optional[torch.Tensor] = None,
               n_msgs: Optional[Sequence[int]] = None, max_positions: int = 1 << 22, digests: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Returns uint8 codes per aggregate (0 = (True, ""), else VERIFY_REASONS[code])."""
        dev = self.ctx.device
        codes = torch.zeros(len(counts), dtype=torch.uint8, device=dev)
        cap = self.ctx.capacity
        host_codes = [1 if c > cap else (2 if (n_msgs is not None and n_msgs[j] != c) else 0) for j, c in enumerate(counts)]
        if any(host_codes):  # reasons 1 / 2 are argument checks (fusion.py:686-689); those aggregates are not hashed
            raise ValueError("verify: aggregates with too many keys or a key/message count mismatch must be filtered by the caller: "
                             + ", ".join(f"{j}:{VERIFY_REASONS[c]}" for j, c in enumerate(host_codes) if c))
        total = int(sum(counts))
        if (digests.shape[0] if digests is not None else len(msgs)) != total:
            raise ValueError("one message per position is required")
        dg_all = digests if digests is not None else self.ctx.hash_msg(msgs)
        for j0, j1, p0, p1 in self._chunks(counts, max_positions):
            cs = list(counts[j0:j1])
            s = None if sel is None else sel[p0:p1].to(dev)
            s_eff = s if s is not None else torch.arange(p0, p1, dtype=torch.int32, device=dev)
            _, vl_s, vr_s, c_s, alpha = self._sorted_hash(vk_l, vk_r, dg_all[p0:p1], cs, s_eff)
            codes[j0:j1] = self.ctx.verify_check(self.A, vl_s, vr_s, c_s, alpha, aggs[j0:j1], cs)
        return codes
