"""Development probe: one chunk of keygen + sign (BASELINE configs[2]) stage by stage, and the signature rows on a second stream beside the sampler /
the challenge hash.  Measured on B200 (65,536 keys): keygen 0.69 ms, hash_ch 2.62 ms, rows 0.98 ms; rows || keygen 1.60 ms (= the sum: the sampler's
CTA owns the whole shared memory of its SM, so nothing co-resides with it), rows || hash_ch 3.45-3.91 ms against 3.59 ms in sequence depending on how many
CTAs the rows kernel is given.  A two-stream Engine.keygen_sign built on this measured 66.8 ms against 67.0 ms for 10^6 keys and was not kept."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fusion_cryptography_b200.batch import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10**6
eng = Engine(256, seed=42)
ctx = eng.ctx
seeds = (10**6 + 2 * np.arange(n)).astype(np.uint64)
pm = ctx.pack_messages([f"msg-{i:07d}".encode() for i in range(n)])
# where does the overlap go?  one chunk's stages alone and together
c = 65536
main = torch.cuda.current_stream()
side = eng._side_stream()
_, _, sk_hat, vl, vr = ctx.keygen(eng.A, torch.from_numpy(seeds[:c].view(np.int64)).cuda(), want_sk=False)
dg = ctx.hash_msg(pm[:c])
c_hat = ctx.hash_ch(vl, vr, dg)
sig = torch.empty((c, eng.ell, eng.d), dtype=torch.int32, device="cuda")
sd = torch.from_numpy(seeds[:c].view(np.int64)).cuda()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
def both(first):
    def f():
        ev = torch.cuda.Event(); ev.record(main)
        with torch.cuda.stream(side):
            side.wait_event(ev)
            ctx.sign_mac_rows(sk_hat, c_hat, out=sig)
        first()
        main.wait_stream(side)
    return f
kg = lambda: ctx.keygen(eng.A, sd, want_sk=False)
hc = lambda: ctx.hash_ch(vl, vr, dg)
rows = lambda: ctx.sign_mac_rows(sk_hat, c_hat, out=sig)
print(f"one chunk of {c}: keygen {t(kg):.3f} ms, hash_ch {t(hc):.3f} ms, rows {t(rows):.3f} ms; rows || keygen {t(both(kg)):.3f} ms, rows || hash_ch {t(both(hc)):.3f} ms")

