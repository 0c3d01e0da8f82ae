"""Development probe: device time of the seeded secret-key sampler alone (fz_sample_coef), best of 10."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from fusion_cryptography_b200 import native
from fusion_cryptography_b200.native import _ptr
ctx = native.get_context(256)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
keys = np.zeros((n, 8), dtype=np.uint32); keys[:, 0] = 10**6 + np.arange(n)
kd = torch.from_numpy(keys.view(np.int32)).cuda(); kl = torch.ones(n, dtype=torch.int32, device="cuda")
out = torch.empty((n, 256), dtype=torch.int32, device="cuda")
def run():
    ctx._chk(ctx.lib.fz_sample_coef(ctx._h, _ptr(kd), C.c_int(8), _ptr(kl), C.c_size_t(n), C.c_int(52), C.c_int(256), _ptr(out), ctx._stream()))
impl = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx.set_option("sample_impl", impl)   # 0 general kernel, 1 dense kernel with the byte tile, 2 dense kernel with 16-byte stores
for _ in range(5): run()
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
print(f"{os.environ.get('FUSION_B200_LIB', 'default')} sample_impl {impl}: sample_coef {n} generators: {best:.3f} ms  checksum {int(out.sum().item())}")
