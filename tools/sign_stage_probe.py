"""Development probe: the stages of one keygen + sign chunk (fz_keygen_sign) one by one against the fused call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fusion_cryptography_b200.batch import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 378880
eng = Engine(256, seed=42); ctx = eng.ctx
seeds = (10**6 + 2 * np.arange(n)).astype(np.uint64)
sd = torch.from_numpy(seeds.view(np.int64)).cuda()
pm = ctx.pack_messages([f"msg-{i:07d}".encode() for i in range(n)])
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out
t_kg, (_, _, sk_hat, vl, vr) = t(lambda: ctx.keygen(eng.A, sd, want_sk=False))
t_pm, dg = t(lambda: ctx.hash_msg(pm))
t_hc, c_hat = t(lambda: ctx.hash_ch(vl, vr, dg))
sig = torch.empty((n, eng.ell, eng.d), dtype=torch.int32, device="cuda")
t_rows, _ = t(lambda: ctx.sign_mac_rows(sk_hat, c_hat, out=sig))
vk_l, vk_r = torch.empty_like(vl), torch.empty_like(vr)
t_all, _ = t(lambda: ctx.keygen_sign(eng.A, sd, pm, vk_l, vk_r, sig))
print(f"{n} keys: keygen {t_kg:.3f} ms, prehash {t_pm:.3f} ms, hash_ch (+ decode + NTT) {t_hc:.3f} ms, rows {t_rows:.3f} ms; sum {t_kg + t_pm + t_hc + t_rows:.3f} ms; fz_keygen_sign {t_all:.3f} ms")
# inside keygen: the sampler alone
keys = np.zeros((2 * n, 8), dtype=np.uint32); keys[0::2, 0] = (seeds & 0xffffffff).astype(np.uint32); keys[1::2, 0] = ((seeds + 1) & 0xffffffff).astype(np.uint32)
t_s, _ = t(lambda: ctx.sample_coef([int(x) for x in seeds[:4]]))  # warm
import ctypes as C
from fusion_cryptography_b200.native import _ptr
kd = torch.from_numpy(keys.view(np.int32)).cuda(); kl = torch.ones(2 * n, dtype=torch.int32, device="cuda")
out = torch.empty((2 * n, 256), dtype=torch.int32, device="cuda")
t_smp, _ = t(lambda: ctx._chk(ctx.lib.fz_sample_coef(ctx._h, _ptr(kd), C.c_int(8), _ptr(kl), C.c_size_t(2 * n), C.c_int(52), C.c_int(256), _ptr(out), ctx._stream())))
t_ntt, _ = t(lambda: ctx.ntt_fwd(out))
print(f"  sampler for {2 * n} generators {t_smp:.3f} ms, forward NTT of them {t_ntt:.3f} ms")
