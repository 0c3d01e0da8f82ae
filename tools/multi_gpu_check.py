"""Run under torchrun (one rank per GPU, NCCL): checks the sharded paths bit-for-bit against the single-GPU result.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from fusion_cryptography_b200 import parallel
from fusion_cryptography_b200.batch import Engine

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = Engine(256, seed=42, device=local)
ctx = eng.ctx
n = 2818 if len(sys.argv) < 2 else int(sys.argv[1])
seeds = (10**6 + 2 * np.arange(n)).astype(np.uint64)
msgs = [f"msg-{i:07d}".encode() for i in range(n)]
# every rank derives all keys (small) but keeps only ITS shard of the signatures (the large data)
lo, hi = parallel.shard_bounds(n, rank, world)
kb = eng.keygen(seeds)
from fusion_cryptography_b200.batch import KeyBatch
loc = KeyBatch(sk_hat=kb.sk_hat[lo:hi].contiguous(), vk_l=kb.vk_l[lo:hi].contiguous(), vk_r=kb.vk_r[lo:hi].contiguous())
local_sigs = eng.sign(loc, msgs[lo:hi])
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
agg = parallel.aggregate_one_sharded(eng, kb.vk_l, kb.vk_r, msgs, local_sigs)
e1.record(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
# reference: rank 0 recomputes everything alone
ok = torch.ones(1, device="cuda", dtype=torch.int32)
if rank == 0:
    sigs = eng.sign(kb, msgs)
    want = eng.aggregate(kb.vk_l, kb.vk_r, msgs, sigs, [n])
    ok[0] = int(torch.equal(want, agg))
    code = int(eng.verify(kb.vk_l, kb.vk_r, msgs, agg, [n])[0].item())
    ok[0] = ok[0] & int(code == 0)
dist.broadcast(ok, 0)
# every rank must hold the same aggregate
ref = agg.clone(); dist.broadcast(ref, 0)
same = torch.tensor([int(torch.equal(ref, agg))], device="cuda", dtype=torch.int32)
dist.all_reduce(same, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"multi_gpu_check world={world} n={n}: sharded aggregate == single-GPU aggregate: {bool(ok.item())}, identical on all ranks: {bool(same.item())}, "
          f"verify ok, {t.item():.1f} ms (max over ranks)", flush=True)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok.item() and same.item() else 1)
