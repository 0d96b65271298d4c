"""Generate tests/golden/*.npz by running the REFERENCE's own functions (build container only).

Imports data_handing/embeddings_related_generator{,_wavcaps}.py and utils.sound_effect_choice
from /root/reference unmodified; the only shim maps the hard-coded 'cuda' device
(embeddings_related_generator.py:15,21) to 'cpu' because the build container has no GPU.
/root/reference does not exist on the GPU box, so the outputs are committed as small fixtures:
inputs are stored as seeds (numpy RandomState streams are frozen by numpy's compatibility
policy) and outputs as indices + scores + per-row float64 checksums.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz
"""
import importlib.util
import os
import pickle
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
D = 1024


def load_ref_module(rel_path, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel_path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class cuda_to_cpu:
    """Route Tensor.to('cuda') to the CPU for the duration of the block."""

    def __enter__(self):
        self._orig = torch.Tensor.to

        def to(t, *a, **kw):
            a = tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) else x for x in a)
            if isinstance(kw.get("device"), str) and kw["device"].startswith("cuda"):
                kw["device"] = "cpu"
            return self._orig(t, *a, **kw)

        torch.Tensor.to = to
        return self

    def __exit__(self, *exc):
        torch.Tensor.to = self._orig


# ---- input recipes (shared with tests/helpers.py: keep in sync) -----------------------------
def make_embeddings(case):
    rs = np.random.RandomState(case["seed"])
    n = case["n"]
    if case["dist"] == "gauss":
        x = rs.standard_normal((n, D)).astype(np.float32)
    elif case["dist"] == "clustered":
        centres = rs.standard_normal((case["centres"], D)).astype(np.float32)
        assign = rs.randint(0, case["centres"], size=n)
        x = centres[assign] + case["sigma"] * rs.standard_normal((n, D)).astype(np.float32)
    else:
        raise ValueError(case["dist"])
    x = x * (0.5 + rs.rand(n, 1).astype(np.float32) * 4.0)   # rows are NOT unit norm on input
    for (dst, src) in case.get("duplicates", []):
        x[dst] = x[src]
    return x


CASES = {
    "generator_gauss": dict(seed=1001, n=64, k=5, dist="gauss", files=[64], module="single"),
    "generator_clustered": dict(seed=1002, n=96, k=5, dist="clustered", centres=8, sigma=0.05,
                                files=[96], module="single"),
    "wavcaps_multi_dup": dict(seed=1003, n=72, k=3, dist="gauss", files=[40, 24, 8],
                              duplicates=[(50, 7), (71, 7)], module="wavcaps"),
    "generator_k1": dict(seed=1004, n=33, k=1, dist="gauss", files=[33], module="single"),
    # --topnumber beyond one pass of the fused kernel (32), duplicates straddling the pass boundary
    "generator_k40": dict(seed=1005, n=160, k=40, dist="gauss", files=[160],
                          duplicates=[(150, 3), (151, 3), (152, 90)], module="single"),
    "wavcaps_k70_clustered": dict(seed=1006, n=300, k=70, dist="clustered", centres=6, sigma=0.05,
                                  files=[100, 200], module="wavcaps"),
    # a bank of several bank tiles and a query batch of several query tiles (single-launch mode of
    # the fused kernel), --topnumber 10 as in BASELINE config 2
    "generator_n2048_k10": dict(seed=1007, n=2048, k=10, dist="gauss", files=[2048], module="single"),
}


def run_generator_case(name, case):
    mod = load_ref_module(
        "data_handing/embeddings_related_generator.py" if case["module"] == "single"
        else "data_handing/embeddings_related_generator_wavcaps.py", "ref_gen_" + name)
    x = make_embeddings(case)
    records = [{"caption": f"caption number {i} of the synthetic set", "text_id": i,
                "text_embedding": torch.from_numpy(x[i:i + 1].copy())} for i in range(case["n"])]
    tmp = tempfile.mkdtemp()
    paths, lo = [], 0
    for fi, cnt in enumerate(case["files"]):
        p = os.path.join(tmp, f"in{fi}.pkl")
        with open(p, "wb") as f:
            pickle.dump(records[lo:lo + cnt], f)
        paths.append(p)
        lo += cnt
    out_path = os.path.join(tmp, "out_related.pkl")
    with cuda_to_cpu():
        bank, all_data = mod.load_data(paths[0] if case["module"] == "single" else paths)
        gen = mod.process_data(bank, all_data, case["k"])
        mod.save_data_to_hdf5(gen, out_path, len(all_data))
    # read back with the reader loop of dataset/dataset.py:64-78 (restated; that module does not
    # import under transformers 5.x)
    items = []
    with open(out_path, "rb") as f:
        while True:
            try:
                it = pickle.load(f)
                items.extend(it) if isinstance(it, list) else items.append(it)
            except EOFError:
                break
    assert len(items) == case["n"]
    xn = torch.nn.functional.normalize(torch.from_numpy(x), dim=-1)
    rel_idx = np.zeros((case["n"], case["k"]), np.int64)
    rel_score = np.zeros((case["n"], case["k"]), np.float32)
    rel_rowsum = np.zeros((case["n"], case["k"]), np.float64)
    for i, it in enumerate(items):
        rel = it["related_embeddings"]
        assert rel.shape == (case["k"], D) and rel.dtype == torch.float32 and rel.device.type == "cpu"
        assert it["text_id"] == i and set(it.keys()) == {"caption", "text_id", "text_embedding",
                                                         "related_embeddings"}
        sim = rel @ xn.T                       # which input row is each related row?
        rel_idx[i] = sim.argmax(dim=1).numpy()
        rel_score[i] = (xn[i:i + 1] @ rel.T).numpy()[0]
        rel_rowsum[i] = rel.double().sum(dim=1).numpy()
    # bank order the reference's set() produced, expressed as input indices
    order = (bank @ xn.T).argmax(dim=1).numpy().astype(np.int64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), related_index=rel_idx,
                        related_score=rel_score, related_rowsum=rel_rowsum, bank_order=order,
                        bank_rowsum=bank.double().sum(dim=1).numpy())
    print(f"{name}: n={case['n']} k={case['k']} top1==self {float((rel_idx[:, 0] == np.arange(case['n'])).mean()):.3f}")


SEC_CASES = {
    "sound_effect_q1": dict(seed=2001, q=1, labels=527, k=3),
    "sound_effect_q4": dict(seed=2002, q=4, labels=527, k=5),
}


def make_sec_inputs(case):
    rs = np.random.RandomState(case["seed"])
    bank = rs.standard_normal((case["labels"], D)).astype(np.float32)
    bank /= np.linalg.norm(bank, axis=1, keepdims=True)
    prefix = rs.standard_normal((case["q"], D)).astype(np.float32)
    prefix /= np.linalg.norm(prefix, axis=1, keepdims=True)
    return prefix, bank


def run_sec_case(name, case):
    sys.path.insert(0, REF)
    import utils as ref_utils   # /root/reference/utils.py (imports models.caption_model)
    prefix, bank = make_sec_inputs(case)
    idx = ref_utils.sound_effect_choice(torch.from_numpy(prefix), torch.from_numpy(bank), case["k"])
    assert idx.dtype == torch.int64 and tuple(idx.shape) == (case["q"], case["k"])
    sim = torch.from_numpy(prefix) @ torch.from_numpy(bank).T
    np.savez_compressed(os.path.join(HERE, name + ".npz"), index=idx.numpy(),
                        score=sim.gather(1, idx).numpy())
    print(f"{name}: index[0]={idx[0].tolist()}")


# the METHOD form of the same retrieval (models/caption_model.py:15-21): returns bank[index].squeeze(1)
SEC_METHOD_CASES = {
    "sound_effect_method_b4": dict(seed=2003, q=4, labels=527, k=3, lead=(4,)),
    "sound_effect_method_b4x1": dict(seed=2004, q=4, labels=527, k=3, lead=(4, 1)),
    "sound_effect_method_k1": dict(seed=2005, q=6, labels=527, k=1, lead=(6,)),
}


def run_sec_method_case(name, case):
    sys.path.insert(0, REF)
    from models.caption_model import ClapCaptionModel      # /root/reference/models/caption_model.py
    prefix, bank = make_sec_inputs(case)
    prefix_t = torch.from_numpy(prefix).reshape(*case["lead"], D)
    out = ClapCaptionModel.sound_effect_choice(None, prefix_t, torch.from_numpy(bank), case["k"])
    flat = out.reshape(-1, D)
    index = (flat @ torch.from_numpy(bank).T).argmax(dim=1)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), shape=np.array(out.shape, np.int64),
                        index=index.numpy(), rowsum=flat.double().sum(dim=1).numpy())
    print(f"{name}: out shape {tuple(out.shape)} first rows {index[:3].tolist()}")


# the label retrieval at its real call site: Dataset.__getitem__ (dataset/dataset.py:365-368) per
# sample, then collate (:632-647) — what zsaac_b200.dataset.collate_with_sound_effects replaces
COLLATE_CASES = {
    "collate_train_b6": dict(seed=2101, q=6, labels=527, k=3, training=True),
    "collate_eval_b3": dict(seed=2102, q=3, labels=527, k=4, training=False),
}


class WordTokenizer:
    """Stand-in for the GPT-2 tokenizer (no network here): one deterministic id per word."""

    def encode(self, text):
        return [sum(ord(c) * (j + 1) for j, c in enumerate(w)) % 50257 for w in text.split()]


def collate_labels(case):
    rs = np.random.RandomState(case["seed"] + 7)
    words = ["dog", "bark", "rain", "engine", "speech", "music", "door", "bird", "water", "wind"]
    return [" ".join(words[j] for j in rs.randint(0, len(words), size=1 + i % 3)).title() for i in range(case["labels"])]


def collate_samples(case):
    """(prefixes [q,1,d], bank, labels, per-sample leading elements) — shared with the tests."""
    prefix, bank = make_sec_inputs(case)
    prefixes = torch.from_numpy(prefix).reshape(case["q"], 1, D)
    if case["training"]:
        lead = [(torch.arange(5) + 10 * n, torch.ones(5) * (n % 2)) for n in range(case["q"])]
    else:
        lead = [(f"audio_{n}.wav",) for n in range(case["q"])]
    return prefixes, torch.from_numpy(bank), collate_labels(case), lead


def run_collate_case(name, case):
    sys.path.insert(0, REF)
    import utils as ref_utils
    import ast
    tree = ast.parse(open(os.path.join(REF, "dataset/dataset.py")).read())
    ns = {"torch": torch, "padding_captions": ref_utils.padding_captions}
    for node in tree.body:                        # dataset.py does not import under transformers 5.x
        if isinstance(node, ast.FunctionDef) and node.name == "collate":
            exec(compile(ast.Module(body=[node], type_ignores=[]), "dataset/dataset.py", "exec"), ns)
    prefixes, bank, labels, lead = collate_samples(case)
    tok = WordTokenizer()
    batch = []
    for n in range(case["q"]):                    # dataset/dataset.py:365-368, verbatim
        prefix = prefixes[n]
        sound_effects_index = ref_utils.sound_effect_choice(prefix, bank, case["k"]).squeeze(0)
        selected_labels = [labels[i].lower() for i in list(sound_effects_index)]
        hard_prompt = ref_utils.parse_entities(tok, selected_labels, 0)
        batch.append((*lead[n], prefix, hard_prompt, len(hard_prompt)))
    out = ns["collate"](batch)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), hard_prompt=out[-2].numpy(), mask=out[-1].numpy(),
                        prefix_sum=out[-3].double().sum().item())
    print(f"{name}: padded hard prompts {tuple(out[-2].shape)}")


RETRIEVAL_CASES = {
    "retrieval_metrics": dict(seed=3001, audios=60, clusters=6, sigma_in=0.5, noise=6.0),
    "retrieval_metrics_easy": dict(seed=3002, audios=37, clusters=37, sigma_in=0.0, noise=0.9),
}


def make_retrieval_inputs(case):
    """CLAP-like pairs: clustered audio embeddings, each with 5 noisy caption embeddings, noisy
    enough that the ground truth is not always retrieved first."""
    rs = np.random.RandomState(case["seed"])
    c = rs.standard_normal((case["clusters"], D)).astype(np.float32)
    a = c[rs.randint(0, case["clusters"], size=case["audios"])] \
        + case["sigma_in"] * rs.standard_normal((case["audios"], D)).astype(np.float32)
    audio_embs = np.repeat(a, 5, axis=0)                      # reference layout: audio repeated 5x
    cap_embs = audio_embs + case["noise"] * rs.standard_normal(audio_embs.shape).astype(np.float32)
    return audio_embs, cap_embs


def run_retrieval_case(name, case):
    # sentence-transformers is not installed: provide util.cos_sim per its published definition
    # (normalise both operands, mm) so that the reference module imports and runs unmodified
    import types
    st = types.ModuleType("sentence_transformers")
    st_util = types.ModuleType("sentence_transformers.util")

    def cos_sim(a, b):
        a = torch.as_tensor(a, dtype=torch.float32)
        b = torch.as_tensor(b, dtype=torch.float32)
        if a.dim() == 1:
            a = a.unsqueeze(0)
        if b.dim() == 1:
            b = b.unsqueeze(0)
        return torch.mm(torch.nn.functional.normalize(a, p=2, dim=1),
                        torch.nn.functional.normalize(b, p=2, dim=1).transpose(0, 1))

    st_util.cos_sim = cos_sim
    st.util = st_util
    sys.modules.setdefault("sentence_transformers", st)
    sys.modules.setdefault("sentence_transformers.util", st_util)
    os.environ.setdefault("WANDB_MODE", "disabled")
    ref = load_ref_module("retrieval/tools/utils.py", "ref_retrieval_utils")
    audio_embs, cap_embs = make_retrieval_inputs(case)
    a = ref.a2t(audio_embs, cap_embs, return_ranks=True)
    t = ref.t2a(audio_embs, cap_embs, return_ranks=True)
    np.savez_compressed(os.path.join(HERE, name + ".npz"),
                        a2t_metrics=np.array(a[:7], np.float64), a2t_ranks=a[7], a2t_top1=a[8],
                        t2a_metrics=np.array(t[:7], np.float64), t2a_ranks=t[7], t2a_top1=t[8])
    print(f"{name}: a2t {np.round(a[:7], 2)} t2a {np.round(t[:7], 2)}")


MEMORY_CASES = {
    "map2memory_q1": dict(seed=4001, q=1, n=700, spread=0.35),
    "map2memory_q5": dict(seed=4002, q=5, n=1500, spread=0.6),
}


def make_memory_inputs(case):
    """Unit-norm caption memory with a few rows close to each query, so that softmax(100 * sim)
    is neither a one-hot nor uniform."""
    rs = np.random.RandomState(case["seed"])
    bank = rs.standard_normal((case["n"], D)).astype(np.float32)
    q = rs.standard_normal((case["q"], D)).astype(np.float32)
    for i in range(case["q"]):
        for j in range(12):
            bank[(37 * i + 11 * j) % case["n"]] = q[i] + case["spread"] * (1 + 0.02 * j) * rs.standard_normal(D).astype(np.float32)
    bank /= np.linalg.norm(bank, axis=1, keepdims=True)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return q, bank


def load_ref_functions(rel_path, names):
    """Execute only the named top-level functions of a reference file that cannot be imported as
    a module (predict_prompt.py pulls gpt2_prefix_eval -> a non-existent `train` module)."""
    import ast
    src = open(os.path.join(REF, rel_path)).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "pickle": pickle}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), rel_path, "exec"), ns)
    return ns


def run_memory_case(name, case):
    ns = load_ref_functions("predict_prompt.py", {"map2memory", "construct_support_memory"})
    q, bank = make_memory_inputs(case)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")            # torch.tensor(tensor) copy-construct warning (:24)
        out = ns["map2memory"](torch.from_numpy(q), torch.from_numpy(bank))
    w = (torch.from_numpy(q) @ torch.from_numpy(bank).T * 100).softmax(dim=-1)
    # construct_support_memory on a small pickle stream (dict items filtered by caption length)
    tmp = tempfile.mkdtemp()
    p = os.path.join(tmp, "mem.pkl")
    caps = ["too short", "a caption that has exactly eight words in it", " ".join(["w"] * 25),
            "another caption with nine words in it right here now"]
    with open(p, "wb") as f:
        for i, c in enumerate(caps):
            pickle.dump({"caption": c, "text_embedding": torch.from_numpy(bank[i:i + 1] * (i + 2.0))}, f)
        pickle.dump([{"caption": "listed", "text_embedding": torch.from_numpy(bank[9:10] * 3.0)}], f)
    mem = ns["construct_support_memory"]([p])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), out=out.numpy(),
                        max_weight=w.max(dim=1).values.numpy(), memory=mem.numpy())
    print(f"{name}: max softmax weight per query {np.round(w.max(dim=1).values.numpy(), 3)} memory rows {tuple(mem.shape)}")


if __name__ == "__main__":
    only = set(sys.argv[1:])          # optional: regenerate just the named fixtures
    for n, c in MEMORY_CASES.items():
        if not only or n in only:
            run_memory_case(n, c)
    for n, c in RETRIEVAL_CASES.items():
        if not only or n in only:
            run_retrieval_case(n, c)
    for n, c in CASES.items():
        if not only or n in only:
            run_generator_case(n, c)
    for n, c in COLLATE_CASES.items():
        if not only or n in only:
            run_collate_case(n, c)
    for n, c in SEC_METHOD_CASES.items():
        if not only or n in only:
            run_sec_method_case(n, c)
    for n, c in SEC_CASES.items():
        if not only or n in only:
            run_sec_case(n, c)
