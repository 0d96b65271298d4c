"""GPU (-m gpu), needs >= 2 GPUs: real NCCL run of the sharded path, bit-exact vs one GPU."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_sharded_bank_nccl_bit_exact():
    world = 2 if torch.cuda.device_count() < 4 else (4 if torch.cuda.device_count() < 8 else 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29517",
           os.path.join(ROOT, "tools", "dist_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("bit-exact on all ranks: True") == 3


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_pipeline_and_generator_sharded_equal_single_gpu(tmp_path):
    """SearchPipeline (side streams + extra communicators, replicated / row_slice, fp32 re-scoring,
    adaptive shard boundaries) and the generator pipeline sharded over the ranks reproduce the
    single-GPU results bit for bit (embeddings_related_generator.py:19-34 on N GPUs)."""
    world = 2 if torch.cuda.device_count() < 4 else (4 if torch.cuda.device_count() < 8 else 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29518",
           os.path.join(ROOT, "tools", "dist_check_r2.py"), str(tmp_path)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("on all ranks: True") == 15
    assert out.stdout.count("identical to the single-GPU stream: True") == 2


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_generator_cli_self_launches_on_n_gpus(tmp_path):
    """`--gpus N` re-launches the script under torch.distributed.run (reference CLI :41-53 + flags)."""
    import pickle
    gen = torch.Generator().manual_seed(8)
    emb = torch.randn(700, 1024, generator=gen)
    recs = [{"caption": f"c{i}", "text_embedding": emb[i:i + 1].clone()} for i in range(700)]
    src, out = tmp_path / "in.pkl", tmp_path / "out.pkl"
    with open(src, "wb") as f:
        pickle.dump(recs, f)
    cmd = [sys.executable, "-m", "zsaac_b200.data_handing.embeddings_related_generator", "--input_path", str(src),
           "--output_path", str(out), "--topnumber", "40", "--gpus", "2", "--exclude_self"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    items = []
    with open(out, "rb") as f:
        while True:
            try:
                items.append(pickle.load(f))
            except EOFError:
                break
    assert len(items) == 700 and all(it["related_embeddings"].shape == (40, 1024) for it in items)
    bank = torch.nn.functional.normalize(emb, dim=-1)
    sims = bank @ bank.T
    sims.fill_diagonal_(-1e9)
    want = sims.topk(40, dim=1).indices
    for i in (0, 350, 699):
        got = items[i]["related_embeddings"]
        assert (got - bank[want[i]]).abs().max().item() < 2e-6
