"""Where does the multi-rank e2e step spend its time?  Upload of 32 MiB from pinned memory timed alone, after the
process group exists, and between sharded MSM steps.  Run under torchrun."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
os.environ.setdefault("ZKP_B200_DEVICE", str(lr))
from interactive_zkp_study_b200 import native as nat, sharded
n = 1 << 20
G1 = nat.g1_bytes((1, 2))
pinned = nat.PinnedBuffer(32 * n)
k = nat.scalars_generate(5 + rank, n)
pinned.write(nat.scalars_download(k, 0, n))
dst = nat.scalars_alloc(n)

def upload_ms(reps=10):
    best, tot = 1e9, 0.0
    for _ in range(reps):
        t0 = time.perf_counter(); nat.scalars_upload(dst, 0, pinned.addr, n); dt = (time.perf_counter() - t0) * 1e3
        best = min(best, dt); tot += dt
    return best, tot / reps
print(rank, "upload alone (best, mean) ms", upload_ms(), flush=True)
import torch, torch.distributed as dist
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dist.barrier(); torch.cuda.synchronize()
print(rank, "upload after init_process_group", upload_ms(), flush=True)
table = nat.g1_fixed_base_mul_dev(G1, k, n)
nat.table_precompute(table)
ex = sharded.PartialExchange()
for _ in range(3):
    sharded.g1_msm_sharded(ex, table, dst, n)
dist.barrier(); torch.cuda.synchronize()
up, st = [], []
for _ in range(10):
    t0 = time.perf_counter(); nat.scalars_upload(dst, 0, pinned.addr, n); t1 = time.perf_counter()
    sharded.g1_msm_sharded(ex, table, dst, n); t2 = time.perf_counter()
    up.append((t1 - t0) * 1e3); st.append((t2 - t1) * 1e3)
print(rank, "in the loop: upload mean %.2f ms (min %.2f), sharded step mean %.2f ms" % (sum(up) / 10, min(up), sum(st) / 10), flush=True)
dist.barrier()
dist.destroy_process_group()
