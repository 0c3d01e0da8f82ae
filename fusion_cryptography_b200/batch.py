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
class Tuning:
    """Knobs of the windowed aggregate / verify pipeline (measured defaults; the tests use them to reach every code path with small batches)."""
    agg_slots: Optional[int] = None   # warp-chains served per time-sliced absorb launch; None = 8 per SM (two per sub-partition)
    sign_overlap: bool = True         # keygen_sign: signature rows of chunk k on the side stream beside the challenge hash of chunk k + 1
    sliced: bool = True               # time-slice the absorb phase for batches a little above agg_slots chains
    sliced_window_div: int = 2        # sliced mode: windows are this many times shorter (the byte rings hold the backlog of the resting chains)
    sliced_budget: float = 1.01       # sliced mode: blocks per launch = expected blocks of one window x this
    sliced_extra_launches: int = 2    # sliced mode: rotating launches after the last window, before the one that closes everything
    overlap: bool = True              # per-signer work of window w+1 on a side stream beside the sponge chains of window w
    side_priority: int = 0            # CUDA priority of that side stream (-1 = scheduled ahead of the main stream's kernels; measured: no effect)
    text_stream: bool = False         # the message text of window w on a stream of its own, so that hash_ch of window w + 1 runs beside it (measured: slower)
    lookahead: int = 2                # windows the host may enqueue ahead of the sponge chains (each holds about 3.1 KB per signer position)


@dataclass
class KeyBatch:
    sk_hat: Optional[torch.Tensor]  # [n, 2, d] (None for keys from keygen_rows): the one distinct row of sk_l / sk_r (all ell rows of a seeded key are equal)
    vk_l: torch.Tensor    # [n, d]
    vk_r: torch.Tensor    # [n, d]
    sk_l: Optional[torch.Tensor] = None  # [n, ell, d] when materialised
    sk_r: Optional[torch.Tensor] = None


class Engine:
    def __init__(self, secpar: int, seed: Optional[int] = 42, device: Optional[int] = None, A: Optional[torch.Tensor] = None,
                 tuning: Optional[Tuning] = None):
        self.ctx = native.get_context(secpar, device)
        self.tuning = tuning if tuning is not None else Tuning()
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

    def keygen_sign(self, seeds, msgs: Sequence[bytes], chunk: Optional[int] = None, keep_sigs: bool = True):
        """BASELINE config[2]: keygen + sign of many one-time keys; secret keys never reach HBM as ell rows.
        Returns (vk_l, vk_r, sigs or None); with keep_sigs=False signatures of each chunk overwrite one buffer.
        chunk defaults to two full waves of the challenge-hash kernel (a thread per signer, 64-thread CTAs, ten resident per SM = 640 signers
        per SM and wave: 189,440 keys on a B200): half of the leg is that ALU-bound kernel, and a chunk that fills only part of a wave leaves the
        same part of the pipes idle.  10^6 keys on one stream: 66.9 ms in chunks of 65,536 (0.69 waves), 60.9 ms at 94,720, 58.7 ms at 189,440;
        with the two-stream pipeline below 57.8 / 54.7 / 52.9-53.9 ms (tools/sign_chunk_probe.py)."""
        import numpy as np

        if chunk is None:
            chunk = 2 * 640 * int(self.ctx.param("sm_count"))

        seeds = np.asarray(seeds, dtype=np.uint64)
        n = len(seeds)
        dev = self.ctx.device
        if not isinstance(msgs, native.PackedMessages):
            msgs = self.ctx.pack_messages(list(msgs))
        seeds_d = torch.from_numpy(seeds.view(np.int64)).to(dev)
        vk_l = torch.empty((n, self.d), dtype=torch.int32, device=dev)
        vk_r = torch.empty((n, self.d), dtype=torch.int32, device=dev)
        sigs = torch.empty((n if keep_sigs else min(chunk, n), self.ell, self.d), dtype=torch.int32, device=dev)
        ctx = self.ctx
        if not self.tuning.sign_overlap or n <= chunk:
            for lo in range(0, n, chunk):
                hi = min(n, lo + chunk)
                ctx.keygen_sign(self.A, seeds_d[lo:hi], msgs[lo:hi], vk_l[lo:hi], vk_r[lo:hi], sigs[lo:hi] if keep_sigs else sigs[: hi - lo])
            return vk_l, vk_r, (sigs if keep_sigs else None)
        # Two streams.  The signature rows of chunk k (HBM writes, 87 KB per key) run on the side stream beside the challenge hash of chunk k + 1
        # (ALU pipe) -- NOT beside its sampler: a sampler CTA owns the whole shared memory of its SM, so nothing co-resides with it and rows launched
        # then would simply run first.  Measured: 10^6 keys in 52.5 ms against 58.1 ms (tools/sign_overlap_probe.py).  With keep_sigs=False two
        # signature buffers alternate.
        main, side = torch.cuda.current_stream(dev), self._side_stream()
        sig2 = None if keep_sigs else torch.empty_like(sigs)
        sk_hat = [torch.empty((min(chunk, n), 2, self.d), dtype=torch.int32, device=dev) for _ in range(2)]
        side.wait_stream(main)
        prev = None

        def rows(p):
            k, lo, hi, c_hat = p
            out = sigs[lo:hi] if keep_sigs else (sigs if k & 1 == 0 else sig2)[: hi - lo]
            ctx.sign_mac_rows(sk_hat[k & 1][: hi - lo], c_hat, out=out)

        done = [None, None]
        for k, lo in enumerate(range(0, n, chunk)):
            hi = min(n, lo + chunk)
            if done[k & 1] is not None:
                main.wait_event(done[k & 1])          # rows of chunk k - 2 have read this secret-key buffer
            ctx.keygen(self.A, seeds_d[lo:hi], want_sk=False, out=(sk_hat[k & 1][: hi - lo], vk_l[lo:hi], vk_r[lo:hi]))
            dg = ctx.hash_msg(msgs[lo:hi])
            if prev is not None:
                sampled = torch.cuda.Event()
                sampled.record(main)
                with torch.cuda.stream(side):
                    side.wait_event(sampled)
                    rows(prev)
                    ev = torch.cuda.Event()
                    ev.record(side)
                done[prev[0] & 1] = ev
                prev[3].record_stream(side)
            c_hat = ctx.hash_ch(vk_l[lo:hi], vk_r[lo:hi], dg)
            prev = (k, lo, hi, c_hat)
        rows(prev)
        main.wait_stream(side)
        for t in (sigs, sig2, sk_hat[0], sk_hat[1]):
            if t is not None:
                t.record_stream(side)
        return vk_l, vk_r, (sigs if keep_sigs else None)

    def keygen_rows(self, row_seeds) -> KeyBatch:
        """General ell-row keygen: row_seeds [n, 2, ell] u64, one sampler seed per secret-key row (2*ell generators, 2*ell transforms and two
        ell-term products per key; nothing of quirk Q1 is used).  Keys are materialised as ell rows (166 KiB per key at secpar 256)."""
        sk_l, sk_r, vk_l, vk_r = self.ctx.keygen_rows(self.A, row_seeds)
        return KeyBatch(sk_hat=None, vk_l=vk_l, vk_r=vk_r, sk_l=sk_l, sk_r=sk_r)

    def keygen_sign_rows(self, row_seeds, msgs: Sequence[bytes], chunk: int = 8192, keep_sigs: bool = True):
        """keygen + sign through the general ell-row kernels only (fz_keygen_rows -> fz_hash_ch -> fz_sign_mac), chunked so that the secret
        keys of one chunk (2 * ell KiB each) are the only ones resident.  Returns (vk_l, vk_r, sigs or None)."""
        import numpy as np

        ctx, dev = self.ctx, self.ctx.device
        if not isinstance(row_seeds, torch.Tensor):
            row_seeds = torch.from_numpy(np.ascontiguousarray(row_seeds, dtype=np.uint64).view(np.int64))
        row_seeds = row_seeds.to(dev)
        n = row_seeds.shape[0]
        if not isinstance(msgs, native.PackedMessages):
            msgs = ctx.pack_messages(list(msgs))
        vk_l = torch.empty((n, self.d), dtype=torch.int32, device=dev)
        vk_r = torch.empty((n, self.d), dtype=torch.int32, device=dev)
        c = min(chunk, max(n, 1))
        sigs = torch.empty((n if keep_sigs else c, self.ell, self.d), dtype=torch.int32, device=dev)
        sk = (torch.empty((c, self.ell, self.d), dtype=torch.int32, device=dev), torch.empty((c, self.ell, self.d), dtype=torch.int32, device=dev))
        for lo in range(0, n, c):
            hi = min(n, lo + c)
            sk_l, sk_r, vl, vr = ctx.keygen_rows(self.A, row_seeds[lo:hi], out=sk)
            vk_l[lo:hi], vk_r[lo:hi] = vl, vr
            c_hat = ctx.hash_ch(vl, vr, ctx.hash_msg(msgs[lo:hi]))
            ctx.sign_mac(sk_l, sk_r, c_hat, out=sigs[lo:hi] if keep_sigs else sigs[: hi - lo])
        return vk_l, vk_r, (sigs if keep_sigs else None)

    # ------------------------------------------------------------------ host-buffer pipelines (what a caller with data in host memory runs)
    def keygen_sign_host(self, seeds_h: torch.Tensor, msg_buf_h: torch.Tensor, msg_offs_h: torch.Tensor, sink, chunk: int = 16384) -> dict:
        """keygen + sign (fusion.py:338-373, 534-557) END TO END through host memory: seeds_h (int64 view of the u64 seeds), the packed messages
        (msg_buf_h uint8, msg_offs_h int64 [n + 1]) live in (pinned) host memory; every chunk's verification keys and signatures are copied back
        into pinned host staging buffers and handed to ``sink(lo, hi, vk_l_h, vk_r_h, sig_h)`` (views valid until the next call but one).
        Compute of chunk k overlaps the device-to-host copy of chunk k - 1 (two device and two host slots, one copy stream).
        Returns the byte counts moved each way."""
        ctx, dev = self.ctx, self.ctx.device
        n = int(seeds_h.numel())
        c = max(1, min(chunk, n))
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_copy", None) is None:
            self._copy = torch.cuda.Stream(device=dev)
        copy = self._copy
        key = ("ksh", c)
        if getattr(self, "_ksh", None) is None or self._ksh[0] != key:
            slots = []
            for _ in range(2):
                slots.append({
                    "vk_l": torch.empty((c, self.d), dtype=torch.int32, device=dev), "vk_r": torch.empty((c, self.d), dtype=torch.int32, device=dev),
                    "sig": torch.empty((c, self.ell, self.d), dtype=torch.int32, device=dev),
                    "h_vk_l": torch.empty((c, self.d), dtype=torch.int32, pin_memory=True), "h_vk_r": torch.empty((c, self.d), dtype=torch.int32, pin_memory=True),
                    "h_sig": torch.empty((c, self.ell, self.d), dtype=torch.int32, pin_memory=True)})
            self._ksh = (key, slots)
        slots = self._ksh[1]
        total_bytes = int(msg_offs_h[-1].item())
        d_buf = torch.empty(total_bytes + 1, dtype=torch.uint8, device=dev)
        d_offs = torch.empty(n + 1, dtype=torch.int64, device=dev)
        d_seeds = torch.empty(n, dtype=torch.int64, device=dev)
        pm = native.PackedMessages(d_buf, d_offs)
        copied = [None, None]
        pending = None
        h2d = d2h = 0
        for k, lo in enumerate(range(0, n, c)):
            hi = min(n, lo + c)
            sl = slots[k & 1]
            b0, b1 = int(msg_offs_h[lo].item()), int(msg_offs_h[hi].item())
            d_seeds[lo:hi].copy_(seeds_h[lo:hi], non_blocking=True)
            d_offs[lo:hi + 1].copy_(msg_offs_h[lo:hi + 1], non_blocking=True)
            if b1 > b0:
                d_buf[b0:b1].copy_(msg_buf_h[b0:b1], non_blocking=True)
            h2d += (hi - lo) * 16 + 8 + (b1 - b0)
            if copied[k & 1] is not None:
                main.wait_event(copied[k & 1])       # the slot's previous chunk has left the device
            m = hi - lo
            ctx.keygen_sign(self.A, d_seeds[lo:hi], pm[lo:hi], sl["vk_l"][:m], sl["vk_r"][:m], sl["sig"][:m])
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(copy):
                copy.wait_event(done)
                sl["h_vk_l"][:m].copy_(sl["vk_l"][:m], non_blocking=True)
                sl["h_vk_r"][:m].copy_(sl["vk_r"][:m], non_blocking=True)
                sl["h_sig"][:m].copy_(sl["sig"][:m], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
            copied[k & 1] = ev
            d2h += m * (2 + self.ell) * self.d * 4
            if pending is not None:
                pending[0].synchronize()
                sink(*pending[1])
            pending = (ev, (lo, hi, sl["h_vk_l"][:m], sl["h_vk_r"][:m], sl["h_sig"][:m]))
        if pending is not None:
            pending[0].synchronize()
            sink(*pending[1])
        main.wait_stream(copy)
        return {"h2d_bytes": h2d, "d2h_bytes": d2h}

    def verify_host(self, vk_l_h: torch.Tensor, vk_r_h: torch.Tensor, msgs: Sequence[bytes], sel_h: torch.Tensor, msg_sel_h: torch.Tensor,
                    aggs_h: torch.Tensor, counts: Sequence[int], codes_h: torch.Tensor) -> dict:
        """verify (fusion.py:680-728) of a batch of aggregates END TO END through host memory: the key pool (vk_l_h, vk_r_h: [pool, d]), the message
        pool, the per-position key index sel_h and message index msg_sel_h (int32 [positions]) and the aggregate signatures aggs_h
        ([n_aggs, ell, d]) are copied to the device, codes_h (uint8 [n_aggs], pinned) receives the result.  Returns the byte counts."""
        ctx, dev = self.ctx, self.ctx.device
        vk_l = vk_l_h.to(dev, non_blocking=True)
        vk_r = vk_r_h.to(dev, non_blocking=True)
        sel = sel_h.to(dev, non_blocking=True)
        msel = msg_sel_h.to(dev, non_blocking=True)
        aggs = aggs_h.to(dev, non_blocking=True)
        pm = ctx.pack_messages(list(msgs)) if not isinstance(msgs, native.PackedMessages) else msgs
        dg = ctx.hash_msg(pm).index_select(0, msel.long()).contiguous()
        codes = self.verify(vk_l, vk_r, None, aggs, counts, sel=sel, digests=dg)
        codes_h.copy_(codes, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        nbytes = lambda t: t.numel() * t.element_size()
        return {"h2d_bytes": nbytes(vk_l_h) + nbytes(vk_r_h) + nbytes(sel_h) + nbytes(msg_sel_h) + nbytes(aggs_h) + int(pm.buf.numel()) + int(pm.offs.numel()) * 8,
                "d2h_bytes": nbytes(codes_h)}

    # ------------------------------------------------------------------ aggregate / verify
    # The aggregation hash is one serial sponge per aggregate (fusion.py:591), so a batch is hashed WINDOW BY WINDOW over
    # the sorted signer positions: window w holds positions [w*P, (w+1)*P) of every aggregate.  All chains advance
    # together (latency-bound chains need every aggregate in flight), memory is O(n_aggs * P) instead of O(positions),
    # and the per-signer work of window w+1 (gather, hash_ch) runs on a second stream while window w is absorbed.
    def _window_positions(self, max_positions: Optional[int]) -> int:
        return int(max_positions) if max_positions is not None else 3 * 640 * int(self.ctx.param("sm_count"))

    def _side_stream(self):
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.ctx.device, priority=self.tuning.side_priority)
        return self._side

    def _text_stream(self):
        if getattr(self, "_side2", None) is None:
            self._side2 = torch.cuda.Stream(device=self.ctx.device)
        return self._side2

    def _sorted_hash(self, vk_l, vk_r, digests_pos, counts, sel):
        """One-shot sort -> gather -> hash_ch -> hash_ag with every position resident (small batches, kernel timing)."""
        perm = self.ctx.sort_keys(vk_l, vk_r, counts, sel=sel)
        vl_s, vr_s, dg_s = self.ctx.gather_sorted(vk_l, vk_r, digests_pos, perm, sel=sel)
        c_s = self.ctx.hash_ch(vl_s, vr_s, dg_s)
        alpha = self.ctx.hash_ag(vl_s, vr_s, dg_s, c_s, counts)
        return perm, vl_s, vr_s, c_s, alpha

    def _hash_windows(self, vk_l, vk_r, dg_all, counts, sel, window_positions: int, keep_c: bool):
        """Absorb phase.  Side stream: gather -> hash_ch -> message text of window w into the per-aggregate rings; main stream:
        the sponge chains over the rings.  Returns the sponge states and the per-window index arrays for the squeeze phase."""
        import numpy as np

        ctx, dev = self.ctx, self.ctx.device
        n_aggs = len(counts)
        cnp = np.asarray(counts, dtype=np.int64)
        total, max_count = int(cnp.sum()), int(cnp.max()) if n_aggs else 0
        off_np = np.zeros(n_aggs + 1, dtype=np.int64)
        np.cumsum(cnp, out=off_np[1:])
        agg_off = torch.from_numpy(off_np).to(dev)
        sel_d = sel.to(dev, dtype=torch.int32).contiguous() if sel is not None else torch.arange(total, dtype=torch.int32, device=dev)
        perm = ctx.sort_keys(vk_l, vk_r, [int(c) for c in counts], sel=sel_d)
        P = max(1, min(max_count, window_positions // max(n_aggs, 1))) if max_count else 1
        # A little above two warp-chains per SM sub-partition (8 per SM) the absorb phase is time-sliced: three chains on one
        # sub-partition saturate its shuffle port and would set the pace of the whole launch (247 instead of 197 cycles per round), so
        # every launch serves exactly 8 * SMs chains with a block budget and the resting set rotates (fz_hash_ag_absorb_sliced).
        tn = self.tuning
        n_slots = tn.agg_slots if tn.agg_slots is not None else 8 * ctx.param("sm_count")
        sliced = n_slots < n_aggs <= (3 * n_slots) // 2 and n_aggs < 2816 and tn.sliced
        if sliced:
            P = max(1, P // tn.sliced_window_div)
            skip_n = n_aggs - n_slots
            per_signer = 71.0 if self.secpar == 256 else 22.0       # 136-byte blocks of message text per signer (9,513 B / 2,908 B typical)
            budget = int(P * per_signer * tn.sliced_budget * n_aggs / n_slots) + 2
        state = ctx.hash_ag_begin(agg_off, n_aggs)
        ring, ring_cap, text_pos = ctx.hash_ag_ring(n_aggs, 3 * P if sliced else P)   # room for the backlog of the resting chains
        agg_ids = torch.arange(n_aggs, dtype=torch.int32, device=dev)
        main = torch.cuda.current_stream(dev)
        # The per-signer work of the next window (gather, hash_ch, message text) runs beside the sponge chains on a side stream.
        # Measured warm on one B200 (FZ_AGG_OVERLAP=0 keeps everything on one stream): verify of 1,250 capacity aggregates 1.06 s
        # against 1.18 s, of 10,000 aggregates 2.80 s against 3.10 s.
        overlap = tn.overlap
        prep = self._side_stream() if overlap else main   # one side stream per engine: the caching allocator pools blocks per stream
        # Tuning.text_stream puts the message text (a warp per signer, half of the ALU pipe idle) on a stream of its own, so that hash_ch of window
        # w + 1 -- the ALU-saturating kernel -- also runs beside the text and the whole chain step of window w instead of starting after the text
        # (at 10^4 capacity aggregates a window is 17.3 ms: chains beside hash_ch 10.7, hash_ch alone 3.8, text alone 2.7).  Measured, bit-identical,
        # and SLOWER: verify 2.43 s against 2.25 s (5,000 aggregates: 1.42 against 1.27 s) -- the chain kernels, whose warps then share their
        # sub-partitions with hash_ch for their whole run, take 1.57 s instead of 1.08 s, which costs more than the 0.32 s of waiting it removes
        # (profiles/r02_agg_probe_text_stream.txt).  Off by default; a higher CUDA priority for the side stream changed nothing either.
        txt = self._text_stream() if (overlap and tn.text_stream) else prep
        if overlap:
            prep.wait_stream(main)
            if txt is not prep:
                txt.wait_stream(main)
        wins = []
        for w in range(0, max(max_count, 1), P):
            cnt = np.clip(cnp - w, 0, P)
            n_win = int(cnt.sum())
            woff_np = np.zeros(n_aggs + 1, dtype=np.int64)
            np.cumsum(cnt, out=woff_np[1:])
            if len(wins) >= 2:
                if overlap:
                    txt.wait_event(wins[-2]["absorbed"])     # ring space of window w - 2 is free once its chain step is done
            if len(wins) >= max(2, tn.lookahead):
                wins[-max(2, tn.lookahead)]["absorbed"].synchronize()   # the host stays at most `lookahead` windows ahead
            with torch.cuda.stream(prep):
                win_off = torch.from_numpy(woff_np).to(dev, non_blocking=True)
                cnt_d = torch.from_numpy(cnt).to(dev, non_blocking=True)
                base = agg_off[:-1] + w - win_off[:-1]
                idx = torch.repeat_interleave(base, cnt_d, output_size=n_win) + torch.arange(n_win, dtype=torch.int64, device=dev)
                agg_of = torch.repeat_interleave(agg_ids, cnt_d, output_size=n_win)
                pos = perm.index_select(0, idx)
                vl_w, vr_w, dg_w = ctx.gather_sorted(vk_l, vk_r, dg_all, pos, sel=sel_d)
                c_w = ctx.hash_ch(vl_w, vr_w, dg_w)
                if txt is not prep:
                    hashed = torch.cuda.Event()
                    hashed.record(prep)
            with torch.cuda.stream(txt):
                if txt is not prep:
                    txt.wait_event(hashed)
                    for t in (vl_w, vr_w, dg_w, c_w, win_off, agg_of):
                        t.record_stream(txt)                 # allocated on the side stream, read by the text kernels
                wend = ctx.hash_ag_text(vl_w, vr_w, dg_w, c_w, win_off, agg_of, agg_off, n_aggs, n_win, w, text_pos, ring, ring_cap, state=state)
                ready = torch.cuda.Event()
                ready.record(txt)
            if overlap:
                for t in (win_off, wend, pos, agg_of, c_w):
                    t.record_stream(main)                # allocated on the side stream, read on the main stream
                main.wait_event(ready)
            if sliced:
                k = len(wins)
                ctx.hash_ag_absorb_sliced(state, ring, ring_cap, wend, n_aggs, ((k + 1) * skip_n) % n_aggs, n_slots, budget)
            else:
                ctx.hash_ag_absorb_ring(state, ring, ring_cap, wend, win_off, n_aggs)
            absorbed = torch.cuda.Event()
            absorbed.record(main)
            wins.append({"win_off": win_off, "n": n_win, "pos": pos, "agg_of": agg_of, "max": int(cnt.max()), "absorbed": absorbed,
                         "c": c_w if keep_c else None})
        if sliced:
            # the text is complete: two more rotating launches, then one over every chain without a budget closes what is left
            k = len(wins)
            for extra in range(tn.sliced_extra_launches):
                ctx.hash_ag_absorb_sliced(state, ring, ring_cap, wend, n_aggs, ((k + 1 + extra) * skip_n) % n_aggs, n_slots, budget)
            ctx.hash_ag_absorb_sliced(state, ring, ring_cap, wend, n_aggs, 0, n_aggs, 1 << 62)
        # every sponge must have closed its message: one that was lapped by the text writer (ring too small for its backlog) is poisoned by the
        # kernels and shows up here -- fail loudly, never return coefficients hashed from overwritten bytes
        still_open = ctx.hash_ag_open_sponges(state, n_aggs)
        if still_open:
            raise native.FusionB200Error(f"aggregation hash: {still_open} of {n_aggs} sponges did not close their message (text ring overrun)")
        if overlap:
            for t in (ring, text_pos, agg_ids, agg_off, perm, sel_d):
                t.record_stream(prep)
                if txt is not prep:
                    t.record_stream(txt)
        return state, wins, sel_d, prep, P

    def _squeeze_windows(self, state, wins, n_aggs, P, side, consume):
        """Squeeze phase.  Main stream: the chains write squeezed blocks of window w; side stream: decode + NTT + consume(wd, alpha_w)."""
        ctx, dev = self.ctx, self.ctx.device
        main = torch.cuda.current_stream(dev)
        overlap = side is not main
        bufs = [ctx.hash_ag_xof_buffer(n_aggs, P) for _ in range(2)]
        done = [None, None]
        for k, wd in enumerate(wins):
            xof, xof_cap = bufs[k & 1]
            if done[k & 1] is not None:
                if overlap:
                    main.wait_event(done[k & 1])         # the buffer's previous window has been decoded
                done[k & 1].synchronize()
            bpos = ctx.hash_ag_squeeze_xof(state, wd["win_off"], n_aggs, xof, xof_cap)
            squeezed = torch.cuda.Event()
            squeezed.record(main)
            if overlap:
                bpos.record_stream(side)
            with torch.cuda.stream(side):
                if overlap:
                    side.wait_event(squeezed)
                alpha_w = ctx.hash_ag_decode(xof, xof_cap, bpos, wd["win_off"], wd["agg_of"], wd["n"])
                consume(wd, alpha_w)
                ev = torch.cuda.Event()
                ev.record(side)
            done[k & 1] = ev
        if overlap:
            for xof, _ in bufs:
                xof.record_stream(side)
            main.wait_stream(side)

    def aggregate(self, vk_l, vk_r, msgs: Optional[Sequence[bytes]], sigs, counts: Sequence[int], sel: Optional[torch.Tensor] = None,
                  sig_sel: Optional[torch.Tensor] = None, max_positions: Optional[int] = None, digests: Optional[torch.Tensor] = None) -> torch.Tensor:
        """counts[j] signers per aggregate; position p (0 <= p < sum(counts)) signs msgs[p] with key row sel[p] and
        signature row sig_sel[p] (both default to p).  `digests` ([positions, 32] prehashes from hash_msg) may replace msgs.
        `max_positions` bounds the signer positions resident per window (default: three full waves of the challenge-hash kernel, 284,160 on a B200:
        a window's hash_ch launch then leaves no partly filled wave).  Returns [n_aggs, ell, d]."""
        total = int(sum(counts))
        if any(c <= 0 for c in counts):
            raise IndexError("list index out of range")  # the reference indexes sorted_signatures[0] (fusion.py:670)
        if (digests.shape[0] if digests is not None else len(msgs)) != total:
            raise ValueError("one message per position is required")
        ctx, dev = self.ctx, self.ctx.device
        n_aggs = len(counts)
        dg_all = digests if digests is not None else ctx.hash_msg(msgs)
        state, wins, sel_d, side, P = self._hash_windows(vk_l, vk_r, dg_all, counts, sel, self._window_positions(max_positions), keep_c=False)
        zsel = sig_sel.to(dev, dtype=torch.int32).contiguous() if sig_sel is not None else None
        acc = {"out": None}

        def consume(wd, alpha_w):     # fusion.py:670-677 over this window's signers, added to the running sums
            rows = wd["pos"] if zsel is None else zsel.index_select(0, wd["pos"].long())
            part = ctx.aggregate_sum_off(sigs, rows.contiguous(), alpha_w, wd["win_off"], n_aggs, wd["max"])
            acc["out"] = part if acc["out"] is None else ctx.pointwise_add(acc["out"], part, out=acc["out"])

        self._squeeze_windows(state, wins, n_aggs, P, side, consume)
        out = acc["out"]
        out.record_stream(torch.cuda.current_stream(dev))
        return out.reshape(n_aggs, self.ell, self.d)

    def verify(self, vk_l, vk_r, msgs: Optional[Sequence[bytes]], aggs, counts: Sequence[int], sel: Optional[torch.Tensor] = None,
               n_msgs: Optional[Sequence[int]] = None, max_positions: Optional[int] = None, digests: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Returns uint8 codes per aggregate (0 = (True, ""), else VERIFY_REASONS[code])."""
        ctx, dev = self.ctx, self.ctx.device
        cap = ctx.capacity
        host_codes = [1 if c > cap else (2 if (n_msgs is not None and n_msgs[j] != c) else 0) for j, c in enumerate(counts)]
        if any(host_codes):  # reasons 1 / 2 are argument checks (fusion.py:686-689); those aggregates are not hashed
            raise ValueError("verify: aggregates with too many keys or a key/message count mismatch must be filtered by the caller: "
                             + ", ".join(f"{j}:{VERIFY_REASONS[c]}" for j, c in enumerate(host_codes) if c))
        total = int(sum(counts))
        if (digests.shape[0] if digests is not None else len(msgs)) != total:
            raise ValueError("one message per position is required")
        n_aggs = len(counts)
        if n_aggs == 0:
            return torch.zeros(0, dtype=torch.uint8, device=dev)
        dg_all = digests if digests is not None else ctx.hash_msg(msgs)
        state, wins, sel_d, side, P = self._hash_windows(vk_l, vk_r, dg_all, counts, sel, self._window_positions(max_positions), keep_c=True)
        acc = {"target": None}

        def consume(wd, alpha_w):     # fusion.py:706-714 over this window's signers
            vl_w, vr_w, _ = ctx.gather_sorted(vk_l, vk_r, dg_all, wd["pos"], sel=sel_d)
            part = ctx.verify_target(vl_w, vr_w, wd["c"], alpha_w, wd["win_off"], n_aggs)
            acc["target"] = part if acc["target"] is None else ctx.pointwise_add(acc["target"], part, out=acc["target"])

        self._squeeze_windows(state, wins, n_aggs, P, side, consume)
        target = acc["target"]
        target.record_stream(torch.cuda.current_stream(dev))
        return ctx.verify_finish(self.A, target, aggs)
