"""Development probe: keygen + sign with the signature rows of chunk k on a second stream beside the challenge hash of chunk k + 1 (NOT beside the
sampler, whose CTA owns the shared memory of its SM), Engine.keygen_sign ships this schedule."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fusion_cryptography_b200.batch import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10**6
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 189440
eng = Engine(256, seed=42); ctx = eng.ctx
dev = ctx.device
seeds = (10**6 + 2 * np.arange(n)).astype(np.uint64)
sd = torch.from_numpy(seeds.view(np.int64)).cuda()
pm = ctx.pack_messages([f"msg-{i:07d}".encode() for i in range(n)])
vk_l = torch.empty((n, eng.d), dtype=torch.int32, device=dev); vk_r = torch.empty_like(vk_l)
sigs = [torch.empty((chunk, eng.ell, eng.d), dtype=torch.int32, device=dev) for _ in range(2)]
main, side = torch.cuda.current_stream(dev), eng._side_stream()
high = torch.cuda.Stream(device=dev, priority=-1)

def serial():
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        ctx.keygen_sign(eng.A, sd[lo:hi], pm[lo:hi], vk_l[lo:hi], vk_r[lo:hi], sigs[0][: hi - lo])

def piped():
    global side
    side.wait_stream(main)
    prev = None
    for k, lo in enumerate(range(0, n, chunk)):
        hi = min(n, lo + chunk)
        _, _, sk_hat, vl, vr = ctx.keygen(eng.A, sd[lo:hi], want_sk=False)
        dg = ctx.hash_msg(pm[lo:hi])
        if prev is not None:           # rows of the previous chunk start when this chunk's sampler is done
            e = torch.cuda.Event(); e.record(main)
            with torch.cuda.stream(side):
                side.wait_event(e)
                ctx.sign_mac_rows(prev[0], prev[1], out=sigs[(k - 1) & 1][: prev[2]])
            prev[0].record_stream(side); prev[1].record_stream(side)
        c_hat = ctx.hash_ch(vl, vr, dg)
        vk_l[lo:hi], vk_r[lo:hi] = vl, vr
        prev = (sk_hat, c_hat, hi - lo)
    ctx.sign_mac_rows(prev[0], prev[1], out=sigs[k & 1][: prev[2]])
    main.wait_stream(side)

def timeit(fn):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

base = timeit(serial)
ref = (vk_l.sum().item(), sigs[0][: (n - 1) % chunk + 1].sum().item())
print(f"serial, chunk {chunk}: {base:.2f} ms = {n / base / 1e3:.2f} M sign/s")
for hp in (0, 1):
    side = high if hp else eng._side_stream()
    tm = timeit(piped)
    last = sigs[((n - 1) // chunk) & 1][: (n - 1) % chunk + 1]
    ok = (vk_l.sum().item(), last.sum().item()) == ref
    print(f"piped, high-priority rows stream {hp}: {tm:.2f} ms = {n / tm / 1e3:.2f} M sign/s  same={ok}", flush=True)
# measured on B200, 10^6 keys, chunk 189,440: serial 58.1 ms; piped 52.5 ms; with the rows kernel capped at 1 / 2 / 4 CTAs per SM 51.6 / 52.8 / 52.8 ms;
# with the hash kernel capped at 9 / 8 / 7 CTAs per SM (dynamic shared-memory padding) 55.1 / 57.1 / 58.5 ms; a high-priority stream for the rows: no change
# the rows kernel as 1-4 CTAs of 32-128 threads per SM (so that it fits beside ten resident hash CTAs from the start): 53.5-61.5 ms, never better than 52.8 ms
