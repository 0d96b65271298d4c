"""Round-2 multi-GPU parity checks (run under torchrun, one rank per GPU):
  * SearchPipeline (three streams, extra NCCL communicators) — replicated and row_slice results,
    device- and host-fed, with and without fp32 re-scoring — against a single-GPU search
  * the generator pipeline: load_data -> process_data -> save_data_to_hdf5 sharded over the ranks
    writes the same record stream (torch.equal per tensor) as one GPU

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
        --master-port 29518 tools/dist_check_r2.py [workdir]
"""
import os
import pickle
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist

import zsaac_b200
from zsaac_b200 import related_pipeline
from zsaac_b200.sharded import SearchPipeline, ShardedRelatedBank


def all_ok(flag: bool, device) -> bool:
    t = torch.tensor([1 if flag else 0], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


def read_stream(path):
    out = []
    with open(path, "rb") as f:
        while True:
            try:
                out.append(pickle.load(f))
            except EOFError:
                return out


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    device = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(device)
    dist.init_process_group("nccl", device_id=device)
    ok = True

    # ---------------------------------------------------------------- SearchPipeline
    N, Q, k = 300_007, 3001, 10
    g = torch.Generator(device=device).manual_seed(99)
    bank = torch.nn.functional.normalize(torch.randn(N, 1024, device=device, generator=g), dim=-1)
    bank[N // world + 5] = bank[7]
    queries = torch.randn(Q, 1024, device=device, generator=g)
    queries[0] = bank[7]
    q_host = queries.cpu().pin_memory()
    whole = zsaac_b200.RelatedBank.from_tensor(bank)
    s1, i1 = whole.search(queries, k)
    r1s, r1i = zsaac_b200.retrieval.search_rescored(whole, queries, bank, k)
    sb = ShardedRelatedBank(N, 1024, device=device)
    sb.upload_global(bank)
    torch.cuda.synchronize()
    for result in ("replicated", "row_slice"):
        for from_host, how in ((False, "full"), (True, "full"), (True, "replicate")):
            for rescore in (False, True):
                pipe = SearchPipeline(sb, Q, k, depth=2, from_host=from_host, to_host=True, result=result,
                                      rescore_from=bank[sb.lo:sb.hi] if rescore else None, input=how)
                slots = [pipe.submit(q_host if from_host else queries) for _ in range(3)]   # wraps around the 2 slots
                pipe.wait_stream()
                torch.cuda.synchronize()
                lo, hi = pipe.out_rows
                ws, wi = (r1s, r1i) if rescore else (s1, i1)
                same = True
                for sl in set(slots):
                    ds, di = pipe.result_of(sl)
                    hs, hi_ = pipe.result_of(sl, host=True)
                    same &= torch.equal(ds, ws[lo:hi]) and torch.equal(di, wi[lo:hi])
                    same &= torch.equal(hs, ws[lo:hi].cpu()) and torch.equal(hi_, wi[lo:hi].cpu())
                good = all_ok(same, device)
                ok &= good
                if rank == 0:
                    print(f"[dist_check_r2] pipeline result={result} from_host={from_host} input={how} rescore={rescore}: "
                          f"bit-exact vs single GPU on all ranks: {good}", flush=True)
    sb.local.close()

    # ---------------------------------------------------------------- adaptive shard boundaries
    # every rank stores 1/4 of a shard beyond its own rows; the pipeline re-balances every 2
    # batches from measured times, a second bank gets boundaries pushed to the limits by hand
    sa = ShardedRelatedBank(N, 1024, device=device, overlap=0.25)
    sa.upload_global(bank)
    torch.cuda.synchronize()
    for result in ("replicated", "row_slice"):
        pipe = SearchPipeline(sa, Q, k, depth=2, from_host=True, to_host=True, result=result, balance_every=2)
        same, moved = True, False
        for step in range(9):
            sl = pipe.submit(q_host)
            pipe.wait_stream(sl)
            torch.cuda.synchronize()
            lo, hi = pipe.out_rows
            ds, di = pipe.result_of(sl)
            same &= torch.equal(ds, s1[lo:hi]) and torch.equal(di, i1[lo:hi])
            moved |= sa.bounds != [tuple(b) for b in sa.base_bounds]
        same &= pipe.rebalances >= 3
        good = all_ok(same, device)
        ok &= good
        if rank == 0:
            print(f"[dist_check_r2] adaptive boundaries result={result} ({pipe.rebalances} re-balances, "
                  f"boundaries moved: {moved}, sizes {[hi - lo for lo, hi in sa.bounds]}): "
                  f"bit-exact vs single GPU on all ranks: {good}", flush=True)
    same = True
    for shift in (-1, +1):                                 # every boundary at its limit, alternating direction
        cuts = [0]
        for r in range(1, world):
            lim = sa.stores[r][0] if (r % 2 == 0) == (shift > 0) else sa.stores[r - 1][1]
            cuts.append(max(lim, cuts[-1] + 64))
        cuts.append(N)
        sa.set_bounds([(cuts[r], cuts[r + 1]) for r in range(world)])
        for kk in (k, 40):
            ws, wi = whole.search(queries, kk)
            s2, i2 = sa.search(queries, kk)
            torch.cuda.synchronize()
            same &= torch.equal(s2, ws) and torch.equal(i2, wi)
    good = all_ok(same, device)
    ok &= good
    if rank == 0:
        print(f"[dist_check_r2] boundaries at the limits of the stored rows, k = {k} and 40: "
              f"bit-exact vs single GPU on all ranks: {good}", flush=True)
    whole.close()
    sa.local.close()

    # ---------------------------------------------------------------- generator, sharded vs single GPU
    workdir = sys.argv[1] if len(sys.argv) > 1 else tempfile.gettempdir()
    n_items, topk = 2500 + 37, 5
    src = os.path.join(workdir, "zs_dist_in.pkl")
    if rank == 0:
        gen = torch.Generator().manual_seed(5)
        emb = torch.randn(n_items, 1024, generator=gen)
        emb[1200] = emb[3]                                        # exact duplicate captions
        recs = [{"caption": f"caption {i}", "text_id": i, "text_embedding": emb[i:i + 1].clone()} for i in range(n_items)]
        with open(src, "wb") as f:
            pickle.dump(recs, f)
    dist.barrier()
    for exclude in (False, True):
        out_multi = os.path.join(workdir, f"zs_dist_out_multi_{int(exclude)}.pkl")
        out_single = os.path.join(workdir, f"zs_dist_out_single_{int(exclude)}.pkl")
        if rank == 0:
            for p in (out_multi, out_single):
                if os.path.exists(p):
                    os.remove(p)
        dist.barrier()
        bank_f32, all_data = related_pipeline.load_data(src)
        gen_items = related_pipeline.process_data(bank_f32, all_data, topk, exclude_self=exclude)
        related_pipeline.save_data_to_hdf5(gen_items, out_multi, len(all_data))
        same = True
        if rank == 0:
            # the same script on ONE GPU: hide the process group from the pipeline
            saved = related_pipeline._dist_info
            related_pipeline._dist_info = lambda: (None, 0, 1)
            try:
                b1, d1 = related_pipeline.load_data(src)
                related_pipeline.save_data_to_hdf5(
                    related_pipeline.process_data(b1, d1, topk, exclude_self=exclude), out_single, len(d1))
            finally:
                related_pipeline._dist_info = saved
            a, b = read_stream(out_multi), read_stream(out_single)
            same = len(a) == len(b) == n_items
            for x, y in zip(a, b):
                same &= x["text_id"] == y["text_id"] and torch.equal(x["text_embedding"], y["text_embedding"])
                same &= torch.equal(x["related_embeddings"], y["related_embeddings"])
            bank_cpu = bank_f32.cpu()
            own_first = [torch.equal(x["related_embeddings"][0], bank_cpu[x["text_id"]]) for x in a]
            if exclude:      # the item's own row is never returned (its bit-identical duplicate may be)
                same &= not any(own_first[i] for i in range(n_items) if i not in (3, 1200))
            else:            # like the reference: slot 0 is the item itself (lower index of the duplicate pair)
                same &= all(own_first[i] for i in range(n_items) if i != 1200)
            same &= not any(os.path.exists(f"{out_multi}.rank{r:03d}") for r in range(world))
        good = all_ok(same, device)
        ok &= good
        if rank == 0:
            print(f"[dist_check_r2] generator exclude_self={exclude}: {n_items} records over {world} ranks "
                  f"identical to the single-GPU stream: {good}", flush=True)
        dist.barrier()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
