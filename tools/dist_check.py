"""Multi-GPU parity check (run under torchrun, one rank per GPU): the row-sharded bank + NCCL
all-gather + merge must reproduce the single-GPU result bit for bit on every rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
        --master-port 29517 tools/dist_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist

import zsaac_b200
from zsaac_b200.sharded import ShardedRelatedBank


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    device = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(device)
    dist.init_process_group("nccl", device_id=device)
    ok = True
    for (N, Q, k, excl) in [(100_003, 777, 10, False), (1_000_000, 4096, 32, False), (50_000, 300, 5, True)]:
        g = torch.Generator(device=device).manual_seed(1234)          # same data on every rank
        bank = torch.randn(N, 1024, device=device, generator=g)
        bank[N // world + 1] = bank[3]                                  # tie across a shard boundary
        queries = torch.randn(Q, 1024, device=device, generator=g)
        queries[0] = bank[3]
        self_index = torch.arange(Q, device=device) if excl else None
        whole = zsaac_b200.RelatedBank.from_tensor(bank)
        s1, i1 = whole.search(queries, k, self_index=self_index)
        sb = ShardedRelatedBank(N, 1024, device=device)
        sb.upload_global(bank)
        s2, i2 = sb.search(queries, k, self_index=self_index)
        torch.cuda.synchronize()
        same = torch.equal(s1, s2) and torch.equal(i1, i2)
        flag = torch.tensor([1 if same else 0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok &= bool(flag.item())
        if rank == 0:
            print(f"[dist_check] world={world} N={N} Q={Q} k={k} exclude_self={excl} "
                  f"shard={sb.lo}:{sb.hi} bit-exact on all ranks: {bool(flag.item())}", flush=True)
        whole.close()
        sb.local.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
