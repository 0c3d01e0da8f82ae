"""Development probe: fz_ntt_pass_host end to end (pinned host in -> device -> pinned host out), tapered vs uniform chunks."""
import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    os.environ["FZ_HOST_TAPER"] = sys.argv[1]
    if len(sys.argv) > 2: os.environ["FZ_HOST_CHUNK_MB"] = sys.argv[2]
    import torch
    from fusion_cryptography_b200 import native
    ctx = native.get_context(256); d = ctx.d; n = 1 << 20
    h = (native.Q - 1) // 2
    a = torch.randint(-h, h + 1, (n, d), dtype=torch.int32, device="cuda")
    bhat = ctx.ntt_fwd(torch.randint(-h, h + 1, (n, d), dtype=torch.int32, device="cuda"))
    want = ctx.ntt_pass(a, bhat)
    hin = torch.empty((n, d), dtype=torch.int32).pin_memory(); hin.copy_(a.cpu())
    hout = torch.empty((n, d), dtype=torch.int32).pin_memory()
    for _ in range(2): ctx.pass_host(hin, bhat, hout)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): ctx.pass_host(hin, bhat, hout)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"taper={sys.argv[1]} chunk={os.environ.get('FZ_HOST_CHUNK_MB','32')} MiB: {dt*1e3:.2f} ms per 2^20 polynomials, {n/dt/1e6:.2f} M polys/s, equal={bool(torch.equal(hout.cuda(), want))}")
else:
    for args in (["0"], ["1"], ["1", "64"], ["1", "16"], ["0"], ["1"]):
        subprocess.run([sys.executable, __file__] + args)
