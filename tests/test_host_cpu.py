"""CPU: the host side without a GPU — the C-ABI library loads and exports every declared symbol,
the product never touches the oracle, and the product path fails loudly without CUDA."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "zero-shot-aac_b200")
HEADER = os.path.join(ROOT, "include", "zsaac.h")

import zsaac_b200  # noqa: E402
from zsaac_b200 import _abi  # noqa: E402
from zsaac_b200.sharded import shard_bounds  # noqa: E402


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zs_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_in_tree():
    assert os.path.exists(_abi.library_path()), "run `make lib` (or __graft_entry__.build())"
    assert os.path.realpath(_abi.library_path()).startswith(os.path.realpath(PKG))


def test_library_exports_every_header_symbol_and_binding_matches_header():
    lib = ctypes.CDLL(_abi.library_path())
    names = declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/zsaac.h but not exported"
    assert sorted(_abi.SIGNATURES) == names, "ctypes SIGNATURES out of sync with include/zsaac.h"


def test_binding_argument_counts_match_header():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, argtypes) in _abi.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^)]*)\)", text)
        args = m.group(1).strip()
        n = 0 if args in ("", "void") else args.count(",") + 1
        assert n == len(argtypes), name


def test_no_compute_free_calls_work_without_gpu():
    lib = _abi.load_library()
    assert lib.zs_abi_version() == 1
    assert lib.zs_kernel_name() == b"zs_simtopk_kernel"
    assert lib.zs_bank_rows(None) == 0 and lib.zs_launch_count(None) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_cuda(tmp_path):
    lib = _abi.load_library()
    handle = ctypes.c_void_p()
    assert lib.zs_create(ctypes.byref(handle), 0) == _abi.ZS_ERR_NO_DEVICE
    assert b"no CPU path" in lib.zs_last_error()
    q, b = torch.randn(2, 1024), torch.randn(10, 1024)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        zsaac_b200.related_topk(q, b, 3)
    from zsaac_b200.utils import sound_effect_choice
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sound_effect_choice(q, b, 3)
    from zsaac_b200.utils import sound_effect_embeddings_choice
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sound_effect_embeddings_choice(q, b, 3)
    from zsaac_b200.data_handing import embeddings_related_generator as gen
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        gen.load_data(str(tmp_path / "missing.pkl"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        next(gen.process_data(b, [{"text_embedding": q[:1]}], 3))


def test_missing_library_is_an_error_not_a_fallback(tmp_path):
    code = ("import os,sys; sys.path.insert(0, %r); os.environ['ZSAAC_B200_LIB']=%r\n"
            "import zsaac_b200\n"
            "try:\n    zsaac_b200.load_library()\nexcept OSError as e:\n    print('OSERROR', 'no CPU fallback' in str(e))\n"
            % (ROOT, str(tmp_path / "nope.so")))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "OSERROR True" in out.stdout, out.stdout + out.stderr


def test_product_never_imports_the_oracle():
    offenders = []
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M) or "oracle/" in text:
                    offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders


def test_cli_surface_matches_reference():
    from zsaac_b200.data_handing import embeddings_related_generator as g1
    from zsaac_b200.data_handing import embeddings_related_generator_wavcaps as g2
    for mod in (g1, g2):
        for fn in ("load_data", "process_data", "save_data_to_hdf5", "main"):
            assert callable(getattr(mod, fn))
    with pytest.raises(TypeError):
        g2.load_data("a_single_string.pkl")


def test_shard_bounds_cover_the_bank_exactly():
    for n, w in [(10_000_000, 8), (400_000, 8), (19_195, 4), (49_838, 2), (7, 8), (8, 8), (1, 1)]:
        b = shard_bounds(n, w)
        assert len(b) == w and b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        per = -(-n // w)
        assert all(hi - lo <= per for lo, hi in b)
    assert shard_bounds(10_000_000, 8)[3] == (3_750_000, 5_000_000)


def test_weighted_shard_bounds():
    from zsaac_b200.sharded import SHARD_ALIGN, shard_bounds
    for n, w in [(10_000_000, [1.0, 0.97, 1.02, 1.0, 0.93, 1.01, 0.99, 1.04]), (1000, [1, 1, 1, 5]),
                 (300, [1, 1]), (5, [3, 1, 1, 1]), (70_000, [1e-3, 1.0])]:
        b = shard_bounds(n, len(w), w)
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[r][1] == b[r + 1][0] for r in range(len(w) - 1))
        assert all(lo <= hi for lo, hi in b)
        if n == 10_000_000:                                        # large banks: cuts on tile boundaries
            assert all(lo % SHARD_ALIGN == 0 for lo, _ in b[1:])
        if n >= len(w):
            assert all(hi > lo for lo, hi in b)                    # nobody is left without rows
    b = shard_bounds(10_000_000, 2, [3.0, 1.0])
    assert abs((b[0][1] - b[0][0]) - 7_500_000) <= SHARD_ALIGN
    assert shard_bounds(1000, 4, None) == shard_bounds(1000, 4)
    for bad in ([1.0], [1.0, 0.0], [1.0, -1.0], [1.0, float("nan")]):
        with pytest.raises(ValueError):
            shard_bounds(100, 2, bad)


def test_parallel_writer_emits_the_same_bytes_as_the_serial_loop(tmp_path, monkeypatch):
    """save_data_to_hdf5(workers=N) must write exactly the reference's per-record pickle stream."""
    import pickle
    from zsaac_b200 import related_pipeline as rp
    monkeypatch.setattr(rp, "WRITER_BATCH", 37)                  # several rounds + a ragged tail
    g = torch.Generator().manual_seed(3)
    items = [{"caption": f"caption {i}", "text_id": i, "text_embedding": torch.randn(1, 64, generator=g),
              "related_embeddings": torch.randn(5, 64, generator=g)} for i in range(150)]
    serial, par = tmp_path / "serial.pkl", tmp_path / "parallel.pkl"
    rp.save_data_to_hdf5(iter(items), str(serial), len(items))
    rp.save_data_to_hdf5(iter(items), str(par), len(items), workers=4)
    assert serial.read_bytes() == par.read_bytes()
    rp.save_data_to_hdf5(iter(items[:3]), str(par), 3, workers=2)            # append mode is kept
    back = []
    with open(par, "rb") as f:
        while True:
            try:
                back.append(pickle.load(f))
            except EOFError:
                break
    assert len(back) == 153 and torch.equal(back[-1]["related_embeddings"], items[2]["related_embeddings"])
    assert not list(tmp_path.glob("*.part*"))


def test_default_writer_emits_the_bytes_of_pickle_dump(tmp_path, monkeypatch):
    """The default writer fills torch's per-storage stream into a template instead of running
    torch.save per tensor; the file must hold exactly the bytes of the reference's loop
    (`pickle.dump(item, file)` per record, embeddings_related_generator.py:30-34) for every kind of
    value a record may carry — and whatever is not a plain CPU tensor must take torch's own
    reduction (also byte for byte)."""
    import io
    import pickle
    from zsaac_b200 import related_pipeline as rp
    g = torch.Generator().manual_seed(12)
    base = torch.randn(6, 8, generator=g)
    shared = torch.randn(1, 64, generator=g)                       # one audio embedding, many records
    attr = torch.randn(3, generator=g)
    attr.note = "python state"                                     # -> _rebuild_from_type_v2
    conj = torch.randn(2, generator=g, dtype=torch.complex64).conj()
    items = [{"audio_embedding": shared, "caption": f"caption {i}", "audio_id": f"Y{i:05d}.wav",
              "text_embedding": torch.randn(1, 64, generator=g),
              "related_embeddings": torch.randn(5, 64, generator=g),
              "view": base[:, ::2], "row": base[3], "tail": base[2:, 4:], "t": base.t(),
              "long": torch.arange(7) * i, "int32": torch.arange(3, dtype=torch.int32),
              "int16": torch.arange(3, dtype=torch.int16), "int8": torch.arange(3, dtype=torch.int8),
              "uint8": torch.arange(300).to(torch.uint8), "bool": torch.tensor([True, False, True]),
              "half": torch.randn(3, generator=g).half(), "bf16": torch.randn(70000, generator=g).bfloat16(),
              "f64": torch.randn(2, 2, generator=g, dtype=torch.float64),
              "empty": torch.empty(0), "empty2d": torch.empty(0, 64), "scalar": torch.tensor(2.5),
              "twice": [shared, shared, base, base[1]],              # memoised object / shared storage
              "nested": {"k": (torch.ones(2), [torch.zeros(1, dtype=torch.int64)])},
              "leaf": torch.randn(3, generator=g).requires_grad_(),  # autograd state
              "param": torch.nn.Parameter(torch.randn(2, generator=g)),   # subclass
              "attr": attr, "conj": conj, "complex": torch.randn(2, generator=g, dtype=torch.complex64),
              "fp8": torch.randn(4, generator=g).to(torch.float8_e4m3fn),  # newer storage format
              "placeholder": 0, "none": None, "float": 1.5, "bytes": b"\x00\x01", "big": 2 ** 70}
             for i in range(30)]
    want = io.BytesIO()
    for item in items:
        pickle.dump(item, want)                                    # reference :34
    path = tmp_path / "out.pkl"
    rp.save_data_to_hdf5(iter(items), str(path), len(items))
    assert path.read_bytes() == want.getvalue()
    # the template path was really taken (and only for plain tensors)
    calls = {"n": 0}
    real = torch.storage.TypedStorage.__reduce__

    def counting(self):
        calls["n"] += 1
        return real(self)

    monkeypatch.setattr(torch.storage.TypedStorage, "__reduce__", counting)
    plain_only = [{"caption": "c", "text_embedding": torch.randn(1, 64, generator=g),
                   "related_embeddings": torch.randn(5, 64, generator=g)} for _ in range(20)]
    sink = io.BytesIO()
    for item in plain_only:
        rp._dump_record(item, sink, False)
    assert calls["n"] == 0                                         # templates known: torch.save never ran
    monkeypatch.undo()
    ref = io.BytesIO()
    for item in plain_only:
        pickle.dump(item, ref)
    assert sink.getvalue() == ref.getvalue()
    # large records (> one 64 KiB pickle frame, the bytes object written outside the frames)
    big = {"caption": "big", "related_embeddings": torch.randn(100, 1024, generator=g),
           "text_embedding": torch.randn(1, 1024, generator=g)}
    a, b = io.BytesIO(), io.BytesIO()
    rp._dump_record(big, a, False)
    pickle.dump(big, b)
    assert a.getvalue() == b.getvalue()
    # storages beyond the cut-off take torch's own path (no template, no probe storage of that size)
    huge = {"caption": "huge", "related_embeddings": torch.randn(1025, 1024, generator=g)}
    known = set(rp._STORAGE_TEMPLATES)
    a, b = io.BytesIO(), io.BytesIO()
    rp._dump_record(huge, a, False)
    pickle.dump(huge, b)
    assert a.getvalue() == b.getvalue() and set(rp._STORAGE_TEMPLATES) == known
    # a template that does not reproduce torch's bytes is dropped, not used
    monkeypatch.setattr(rp, "_STORAGE_TEMPLATES", {})
    monkeypatch.setattr(rp, "_storage_template", lambda dtype, numel: (b"not", b"torch's", b"bytes"))
    a, b = io.BytesIO(), io.BytesIO()
    rp._dump_record(plain_only[0], a, False)
    pickle.dump(plain_only[0], b)
    assert a.getvalue() == b.getvalue() and set(rp._STORAGE_TEMPLATES.values()) == {False}
    monkeypatch.undo()
    # the switch restores the literal loop
    monkeypatch.setenv("ZSAAC_TEMPLATE_PICKLE", "0")
    monkeypatch.setattr(rp, "_template_dump", None)                # would raise if called
    off = tmp_path / "off.pkl"
    rp.save_data_to_hdf5(iter(items), str(off), len(items))
    assert off.read_bytes() == want.getvalue()


def test_fast_pickle_stream_loads_to_the_same_records(tmp_path, monkeypatch):
    """save_data_to_hdf5(fast_pickle=True): different bytes, but the reference's reader loop
    (dataset/dataset.py:64-78, restated in helpers.read_related_stream) gets the same dicts of
    torch tensors — dtype, shape, values, writable — serially and from the parallel writer."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    from zsaac_b200 import related_pipeline as rp
    monkeypatch.setattr(rp, "WRITER_BATCH", 41)
    g = torch.Generator().manual_seed(4)
    items = [{"caption": f"caption {i}", "text_id": i, "audio_id": f"a{i}.wav",
              "text_embedding": torch.randn(1, 64, generator=g),
              "related_embeddings": torch.randn(5, 64, generator=g),
              "half": torch.randn(3, generator=g).bfloat16(),          # no numpy dtype: torch's own path
              "view": torch.randn(4, 8, generator=g)[:, ::2]}          # non-contiguous
             for i in range(120)]
    plain, fast, fast_par = tmp_path / "plain.pkl", tmp_path / "fast.pkl", tmp_path / "fast_par.pkl"
    rp.save_data_to_hdf5(iter(items), str(plain), len(items))
    rp.save_data_to_hdf5(iter(items), str(fast), len(items), fast_pickle=True)
    rp.save_data_to_hdf5(iter(items), str(fast_par), len(items), workers=3, fast_pickle=True)
    assert fast.read_bytes() == fast_par.read_bytes() and fast.read_bytes() != plain.read_bytes()
    want = helpers.read_related_stream(str(plain))
    for path in (fast, fast_par):
        got = helpers.read_related_stream(str(path))
        assert len(got) == len(want) == 120
        for a, b in zip(got, want):
            assert a.keys() == b.keys() and a["caption"] == b["caption"] and a["text_id"] == b["text_id"]
            for key in ("text_embedding", "related_embeddings", "half", "view"):
                assert type(a[key]) is torch.Tensor and a[key].dtype == b[key].dtype
                assert a[key].shape == b[key].shape and torch.equal(a[key], b[key])
                assert not a[key].requires_grad
        got[0]["related_embeddings"][0, 0] = 1.0                             # writable
    monkeypatch.setenv("ZSAAC_FAST_PICKLE", "1")                             # env default
    env_path = tmp_path / "env.pkl"
    rp.save_data_to_hdf5(iter(items), str(env_path), len(items))
    assert env_path.read_bytes() == fast.read_bytes()


def test_fast_unpickler_returns_the_same_records(tmp_path, monkeypatch):
    """_read_records resolves torch's per-tensor `_load_from_bytes` to a direct parser of the legacy
    storage stream; the records must equal pickle.load's in every dtype / view / edge case, and
    anything it does not recognise must go through torch's own loader."""
    import pickle
    from zsaac_b200 import related_pipeline as rp
    g = torch.Generator().manual_seed(9)
    base = torch.randn(6, 8, generator=g)
    recs = [{"caption": f"c{i}", "text_id": i, "text_embedding": torch.randn(1, 64, generator=g),
             "view": base[:, ::2], "row": base[3], "long": torch.arange(5) * i,
             "half": torch.randn(3, generator=g).half(), "bf16": torch.randn(3, generator=g).bfloat16(),
             "bool": torch.tensor([True, False]), "empty": torch.empty(0), "scalar": torch.tensor(2.5),
             "f64": torch.randn(2, 2, generator=g, dtype=torch.float64), "placeholder": 0}
            for i in range(40)]
    a, b = tmp_path / "a.pkl", tmp_path / "b.pkl"
    pickle.dump(recs[:25], open(a, "wb"))
    pickle.dump(recs[25:], open(b, "wb"))
    want = pickle.load(open(a, "rb")) + pickle.load(open(b, "rb"))
    blob = pickle.dumps(torch.randn(4, generator=g))
    calls = {"slow": 0}
    slow = torch.storage._load_from_bytes

    def counting(bts):
        calls["slow"] += 1
        return slow(bts)

    monkeypatch.setattr(torch.storage, "_load_from_bytes", counting)
    got = rp._read_records([str(a), str(b)])
    assert calls["slow"] == 0                                  # everything took the direct parser
    assert len(got) == len(want) == 40
    for x, y in zip(got, want):
        assert x.keys() == y.keys() and x["caption"] == y["caption"] and x["placeholder"] == 0
        for key in ("text_embedding", "view", "row", "long", "half", "bf16", "bool", "empty", "scalar", "f64"):
            assert type(x[key]) is torch.Tensor and x[key].dtype == y[key].dtype
            assert x[key].shape == y[key].shape and x[key].stride() == y[key].stride()
            assert torch.equal(x[key], y[key]) and not x[key].requires_grad
    got[0]["text_embedding"][0, 0] = 7.0                       # owns writable memory
    # plain pickling gives every tensor its own copy of the storage (torch's loader and this one)
    for recs_ in (got, want):
        assert (recs_[1]["view"].untyped_storage().data_ptr()
                != recs_[1]["row"].untyped_storage().data_ptr())
    assert got[1]["view"].untyped_storage().nbytes() == want[1]["view"].untyped_storage().nbytes()
    # autograd state takes torch's own rebuild
    leaf = torch.randn(3, generator=g).requires_grad_()
    monkeypatch.undo()
    back = rp._FastTensorUnpickler(__import__("io").BytesIO(pickle.dumps({"p": leaf}))).load()
    assert back["p"].requires_grad and torch.equal(back["p"], leaf)
    monkeypatch.setattr(torch.storage, "_load_from_bytes", counting)
    # unknown streams fall back to torch's loader
    assert b"cpu" in blob
    broken = blob.replace(b"cpu", b"xpu", 1)                   # a location the parser does not take
    try:
        rp._FastTensorUnpickler(__import__("io").BytesIO(broken)).load()
    except Exception:
        pass
    assert calls["slow"] >= 1
    monkeypatch.setenv("ZSAAC_FAST_UNPICKLE", "0")
    calls["slow"] = 0
    again = rp._read_records([str(a)])
    assert calls["slow"] > 0 and torch.equal(again[3]["view"], want[3]["view"])


def test_bench_reference_arm_prints_our_config(tmp_path):
    """bench.py --impl reference (the CPU arm the driver runs next to ours): one JSON line with
    the contract's keys, the SAME `config` dict our arm prints for that command line, rank 0 alone
    under a multi-rank launch."""
    import json
    sys.path.insert(0, ROOT)
    import bench
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
           "--workload", "clotho_eval", "--queries", "64", "--bank-rows", "4096"]
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "queries/s" and line["higher_is_better"] is True
    assert line["steps"] == 2 and line["warmup"] == 1 and line["n_gpus"] == 1 and line["value"] > 0
    assert line["config"] == bench.workload_config("clotho_eval", 64, 4096, 5, False, 1)
    assert line["config"]["workload"].startswith("clotho_eval: 64 queries vs 4096-row bank")
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "queries/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0
    # N = 4 command line: same shard arithmetic as sharded.shard_bounds, config names the sharding
    cfg4 = bench.workload_config("synthetic_10m", 65536, 10_000_000, 32, False, 4)
    assert cfg4["bank_rows_per_gpu"] == shard_bounds(10_000_000, 4)[0][1] == 2_500_000
    assert "4 GPUs" in cfg4["parallelism"] and cfg4["l2"].startswith("inputs larger than L2")
    assert bench.workload_config("clotho_eval", 1045, 19195, 5, False, 1)["l2"].startswith("L2 flushed")
    # every rank but 0 of a multi-rank launch exits 0 without printing
    env.update(RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_read_templates_only_match_streams_they_reproduce():
    """The reader skips unpickling for a storage stream that equals an already parsed one except
    for the storage key and the data; anything else must fall through to the full parser."""
    import io
    import pickle
    from zsaac_b200 import related_pipeline as rp

    class Payloads(pickle.Unpickler):                 # the bytes handed to torch.storage._load_from_bytes
        def find_class(self, module, name):
            if (module, name) == ("torch.storage", "_load_from_bytes"):
                return lambda b: b
            if (module, name) == ("torch._utils", "_rebuild_tensor_v2"):
                return lambda *a: a[0]
            return super().find_class(module, name)

    def payload_of(t):
        return Payloads(io.BytesIO(pickle.dumps(t))).load()

    g = torch.Generator().manual_seed(5)
    rp._READ_TEMPLATES.clear()
    a, b = torch.randn(3, 8, generator=g), torch.randn(3, 8, generator=g)
    pa, pb = payload_of(a), payload_of(b)
    assert rp._match_read_template(pa) is None                          # nothing known yet
    assert torch.equal(rp._storage_from_legacy_bytes(pa).flat, a.reshape(-1))
    assert len(rp._READ_TEMPLATES) == 1
    assert torch.equal(rp._match_read_template(pb), b.reshape(-1))      # other key, other data
    flat = rp._match_read_template(pb)
    flat[0] = 1.0                                                       # owns writable memory
    assert rp._match_read_template(pb)[0] == b.reshape(-1)[0]
    # same layout, other element count / dtype: not this template
    assert rp._match_read_template(payload_of(torch.randn(25, generator=g))) is None
    assert rp._match_read_template(payload_of(torch.arange(24, dtype=torch.int32))) is None
    # damaged streams: truncated, longer, keys that differ, a count that differs, non-digit key
    head = rp._READ_TEMPLATES[0][0]
    p = len(head)
    klen = int.from_bytes(pb[p + 1:p + 5], "little")
    assert rp._match_read_template(pb[:-1]) is None and rp._match_read_template(pb + b"\0") is None
    second = pb.index(pb[p:p + 5 + klen], p + 5 + klen)
    swapped = bytearray(pb)
    swapped[second + 5] = ord("9") if swapped[second + 5] != ord("9") else ord("8")
    assert rp._match_read_template(bytes(swapped)) is None
    letters = pb.replace(pb[p + 5:p + 5 + klen], b"k" * klen)
    assert rp._match_read_template(letters) is None
    count_at = len(pb) - 24 * 4 - 8
    wrong = bytearray(pb)
    wrong[count_at] ^= 1
    assert rp._match_read_template(bytes(wrong)) is None
    # zero-element storages and the most-recently-used order
    e = torch.empty(0)
    assert rp._storage_from_legacy_bytes(payload_of(e)).flat.numel() == 0
    assert rp._match_read_template(payload_of(torch.empty(0))).numel() == 0
    assert rp._READ_TEMPLATES[0][4] == 0
    assert torch.equal(rp._match_read_template(pa), a.reshape(-1)) and rp._READ_TEMPLATES[0][4] == 24
    for n in range(1, 40):                                              # the list stays bounded
        rp._storage_from_legacy_bytes(payload_of(torch.zeros(n)))
    assert len(rp._READ_TEMPLATES) == rp._READ_TEMPLATES_MAX


def test_record_reader_matches_the_reference_loop(tmp_path):
    """dataset.read_related_records = the reader loops of dataset/dataset.py:64-78 (8-20-word
    captions kept, list objects spliced) and :401-417 (everything kept), on a stream written by the
    generator's writer with a list object appended, as the reference's datasets accept."""
    import pickle
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    from zsaac_b200 import related_pipeline as rp
    from zsaac_b200.dataset import read_related_records
    g = torch.Generator().manual_seed(21)
    words = "a dog barks far away while cars pass by on the wet road and a door slams shut twice then silence falls".split()
    items = [{"caption": " ".join(words[:3 + i % 21]), "audio_id": f"Y{i}.wav",
              "text_embedding": torch.randn(1, 64, generator=g),
              "related_embeddings": torch.randn(5, 64, generator=g)} for i in range(60)]
    a, b = tmp_path / "a.pkl", tmp_path / "b.pkl"
    rp.save_data_to_hdf5(iter(items[:40]), str(a), 40)
    rp.save_data_to_hdf5(iter(items[40:]), str(b), 20, fast_pickle=True)
    extra = [{"caption": "short one", "text_embedding": torch.randn(1, 64, generator=g)} for _ in range(3)]
    with open(a, "ab") as f:
        pickle.dump(extra, f)                                    # a list object: spliced, never filtered
    want_all = helpers.read_related_stream(str(a)) + helpers.read_related_stream(str(b))
    got_all = read_related_records([str(a), str(b)])
    assert len(got_all) == len(want_all) == 63
    want_filtered = [it for it in want_all[:40] if 8 <= len(it["caption"].split()) <= 20] + want_all[40:43] \
        + [it for it in want_all[43:] if 8 <= len(it["caption"].split()) <= 20]
    got_filtered = read_related_records([str(a), str(b)], caption_words=(8, 20))
    assert 3 < len(got_filtered) == len(want_filtered) < 63
    for got, want in ((got_all, want_all), (got_filtered, want_filtered)):
        for x, y in zip(got, want):
            assert x.keys() == y.keys() and x["caption"] == y["caption"]
            for key in x:
                if isinstance(y[key], torch.Tensor):
                    assert type(x[key]) is torch.Tensor and x[key].dtype == y[key].dtype
                    assert x[key].shape == y[key].shape and torch.equal(x[key], y[key])
    assert len(read_related_records(str(b))) == 20               # one path instead of a list


def _plan_dry(sm, n, q, k, cg=0):
    lib = zsaac_b200.load_library()
    a, b, c, w, g = (ctypes.c_int() for _ in range(5))
    _abi.check(lib.zs_plan_dry(sm, n, q, k, cg, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c),
                               ctypes.byref(w), ctypes.byref(g)))
    return a.value, b.value, c.value, w.value, g.value


def test_planner_invariants_and_baseline_plans(monkeypatch):
    """The work-unit planner (bank chunks per query tile, tiles per chunk, CTAs, lock-step window)
    runs without a GPU through zs_plan_dry: structural invariants on a grid of shapes, and the
    plans of the BASELINE configs as measured on B200 (profiles/r01/bench_*.json)."""
    import random
    for var in ("ZSAAC_CHUNKS", "ZSAAC_LOCKSTEP", "ZSAAC_SYNC_WINDOW", "ZSAAC_CTA_GROUP"):
        monkeypatch.delenv(var, raising=False)
    rng = random.Random(5)
    for _ in range(400):
        sm = rng.choice([148, 132, 16, 2])
        n = rng.choice([rng.randint(1, 3000), rng.randint(3000, 500_000), rng.randint(500_000, 20_000_000)])
        q = rng.choice([rng.randint(1, 300), rng.randint(300, 70_000), rng.randint(70_000, 500_000)])
        k = rng.randint(1, 32)
        cg = rng.choice([0, 1, 2])
        chunks, tpc, ctas, window, group = _plan_dry(sm, n, q, k, cg)
        assert group == cg if cg else group in (1, 2)
        if not cg:   # one 128-row tile: single CTAs; large batches: pairs; in between: cost model
            assert group == 1 if q <= 128 else (group == 2 if q > 2048 else True)
        n_tiles = -(-n // 256)
        m_tiles = -(-q // (128 * group))
        assert 1 <= chunks <= min(n_tiles, 256)                 # two column halves each: <= 512 lists
        assert chunks * tpc >= n_tiles > (chunks - 1) * tpc      # chunks cover the bank, none empty
        assert ctas % group == 0 and group <= ctas <= max(sm // group, 1) * group
        assert ctas // group == min(m_tiles * chunks, max(sm // group, 1))
        assert window in (0, 32) and (window == 0 or (m_tiles > 1 and tpc >= 4 * window))
    pins = {(10_000_000, 65_536, 32): (13, 3005, 148, 32, 2),   # config 4, one GPU
            (1_250_000, 65_536, 32): (2, 2442, 148, 32, 2),     # config 4, one rank of eight
            (400_000, 8_192, 10): (16, 98, 148, 0, 2),          # config 3
            (400_000, 400_000, 5): (5, 313, 148, 32, 2),        # config 5
            (49_838, 975, 10): (18, 11, 144, 0, 2),             # config 2
            (19_195, 1_045, 5): (15, 5, 135, 0, 1),             # config 1: 9 tiles of 128 rows beat 5 of 256
            (400_000, 256, 10): (72, 22, 144, 0, 2),            # just above one tile: pairs
            (400_000, 128, 10): (143, 11, 143, 0, 1)}           # HBM-bound small batch
    for (n, q, k), want in pins.items():
        assert _plan_dry(148, n, q, k) == want, (n, q, k)
    monkeypatch.setenv("ZSAAC_CHUNKS", "4")                      # tuning hook
    assert _plan_dry(148, 400_000, 8_192, 10)[0] == 4
    with pytest.raises(RuntimeError):
        _plan_dry(148, 0, 10, 5)


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md maps every C symbol of include/zsaac.h to the reference lines it replaces."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [n for n in declared_functions() if n not in doc and n.rsplit("_", 1)[0] + "_*" not in doc]
    assert not missing, missing


def test_bench_roofline_arithmetic():
    """bench.roofline_of: algorithmic FLOP = 2 Q N d and bytes = 2 N d + 2 Q d + 12 Q k (DESIGN §4.1,
    SURVEY §8d) over the kernel time, against the measured peaks — the numbers of the round-1
    driver record (1.3422 PFLOP / 878.97 ms = 1527 TFLOP/s = 1.115 of 1370)."""
    sys.path.insert(0, ROOT)
    import bench
    peaks, src = bench.load_peaks()
    assert src in ("measured", "fallback") and peaks["bf16_tflops"] >= peaks["bf16_tflops_sustained"] > 0 and peaks["hbm_gbs"] > 0
    r = bench.roofline_of(peaks, src, 65536, 10_000_000, 32, 878.97, long_step=True)
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and r["kernel"] == "zs_simtopk_kernel"
    assert abs(r["achieved"] - 2 * 65536 * 1e7 * 1024 / 0.87897 / 1e12) < 1e-6
    assert abs(r["achieved"] - 1527.0) < 0.5 and r["peak"] == peaks["bf16_tflops_sustained"]
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert abs(r["frac_of_burst"] - r["achieved"] / peaks["bf16_tflops"]) < 1e-12
    burst = bench.roofline_of(peaks, src, 65536, 10_000_000, 32, 878.97, long_step=False)
    assert burst["peak"] == peaks["bf16_tflops"]
    # one query against 400 k rows is a bank stream: HBM-bound, bytes / time against the copy bandwidth
    h = bench.roofline_of(peaks, src, 1, 400_000, 10, 0.131, long_step=False)
    want = (2.0 * 400_000 * 1024 + 1 * 1024 * 2.0 + 1 * 10 * 12.0) / 0.131e-3 / 1e9
    assert h["bound"] == "hbm" and h["unit"] == "GB/s" and abs(h["achieved"] - want) < 1e-6
    assert h["peak"] == peaks["hbm_gbs"] and abs(h["frac"] - want / peaks["hbm_gbs"]) < 1e-12
    # the ridge: compute-bound from a few hundred queries on (~252 flop/byte)
    assert bench.roofline_of(peaks, src, 128, 400_000, 10, 0.15, False)["bound"] == "hbm"
    assert bench.roofline_of(peaks, src, 512, 400_000, 10, 0.31, False)["bound"] == "tensor"


def test_bench_parity_gate_flags_wrong_results():
    """bench.parity_gate is what stands between a wrong kernel and a published number: it must
    pass the exact top-k of bf16-rounded operands (what the kernel ranks) and flag every kind of
    wrong answer — a loser returned, a clear winner missing, scores off, order broken, self kept."""
    sys.path.insert(0, ROOT)
    import bench
    dev = torch.device("cpu")
    n_rows, n_q, k, seed = 3000, 40, 5, 777
    _, rows = next(bench.bank_rows_fp32(torch, dev, seed, 0, n_rows))
    assert tuple(rows.shape) == (n_rows, bench.D)
    bank_n = torch.nn.functional.normalize(rows, dim=-1)
    # two shards regenerate the same rows (per-block seeds): what makes the gate world-size independent
    _, tail = next(bench.bank_rows_fp32(torch, dev, seed, 1000, n_rows))
    assert torch.equal(tail, rows[1000:])

    def kernel_like(q, self_index=None):
        s = torch.nn.functional.normalize(q, dim=-1).bfloat16().float() @ bank_n.bfloat16().float().T
        if self_index is not None:
            s[torch.arange(q.shape[0]), self_index] = float("-inf")
        top = torch.sort(s, dim=1, descending=True, stable=True)
        return top.values[:, :k].contiguous(), top.indices[:, :k].contiguous()

    def gate(q, res, self_index=None):
        return bench.parity_gate(torch, None, 1, dev, 0, n_rows, seed, q, self_index, k, res)

    q = bench.gen_queries(torch, n_q, 11)
    s, i = kernel_like(q)
    g = gate(q, (s, i))
    assert g["ok"] and g["sampled_queries"] == n_q and g["max_abs_score_err_vs_fp32"] < 3e-4
    assert g["returned_below_band"] == 0 and g["clear_winners_missing"] == 0 and g["sorted"]
    assert g["index_agreement_with_fp32"] > 0.9
    # a loser in the last slot: returned below the band (and the true k-th may be a missing winner)
    worst = (torch.nn.functional.normalize(q, dim=-1) @ bank_n.T).argmin(dim=1)
    bad_i = i.clone()
    bad_i[:, k - 1] = worst
    g = gate(q, (s, bad_i))
    assert not g["ok"] and g["returned_below_band"] >= n_q
    # the best row replaced by a duplicate of the second: a clear winner is missing
    bad_i = i.clone()
    bad_i[:, 0] = i[:, 1]
    g = gate(q, (s, bad_i))
    assert not g["ok"] and g["clear_winners_missing"] > 0
    # scores off by more than north_star's 1e-3
    g = gate(q, (s + 5e-3, i))
    assert not g["ok"] and g["max_abs_score_err_vs_fp32"] > 4e-3
    # order broken
    g = gate(q, (s.flip(1).contiguous(), i.flip(1).contiguous()))
    assert not g["ok"] and not g["sorted"]
    # self-exclusion (BASELINE config 5): queries are bank rows; keeping the row itself is flagged
    q_self = rows[:n_q].clone()
    me = torch.arange(n_q)
    g = gate(q_self, kernel_like(q_self, me), me)
    assert g["ok"] and g["self_excluded"]
    g = gate(q_self, kernel_like(q_self), me)
    assert not g["ok"] and not g["self_excluded"]


def test_c_abi_from_a_c_program(tmp_path):
    """include/zsaac.h is plain C99 and the library links from a C program (tests/c/abi_smoke.c):
    version, dry planner, error codes + messages, and zs_create's loud failure without a GPU."""
    import shutil
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    lib_dir = os.path.dirname(_abi.library_path())
    exe = str(tmp_path / "abi_smoke")
    build = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror",
                            "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_smoke.c"),
                            "-o", exe, "-L", lib_dir, "-lzsaac_b200", f"-Wl,-rpath,{lib_dir}"],
                           capture_output=True, text=True, timeout=300)
    assert build.returncode == 0, build.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, (run.returncode, run.stdout, run.stderr)
    assert "plan chunks=13 " in run.stdout and "cta_group=2" in run.stdout     # BASELINE config 4's plan
    if not torch.cuda.is_available():
        assert "no CPU path" in run.stdout


def test_bench_our_arm_on_standins():
    """Every line of bench.py's measurement protocol and JSON assembly for OUR arm, executed on the
    CPU with stand-ins for the CUDA pieces (tests/bench_standins.py): one JSON line carrying every
    key of the contract, a green parity gate on the stand-in's bf16 emulation, all named shapes, and
    the same `config` the reference arm prints for that command line."""
    import json
    sys.path.insert(0, ROOT)
    import bench
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "bench_standins.py"), "--steps", "3"],
                         capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    line = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["unit"] == "queries/s" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["n_gpus"] == 1 and line["steps"] == 3 and line["warmup"] == 3 and line["dtype"] == "bf16"
    assert line["value"] > 0 and line["gpu_launches"] > 0
    assert line["config"] == bench.workload_config("synthetic_10m", 300, 6000, 32, False, 1)
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel", "kernel_share_of_step"):
        assert key in line["roofline"], key
    assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert line["e2e"]["h2d_bytes_per_step"] == 300 * 1024 * 4 and line["e2e"]["d2h_bytes_per_step"] == 300 * 32 * 12
    assert set(line["clocks"]) == {"sm_mhz", "sm_max_mhz", "reasons"}
    gate = line["parity_gate"]
    assert gate["ok"] and gate["e2e_readback_identical"] and gate["e2e_timed_readback_identical"]
    named = [w["workload"].split(":")[0] for w in line["workloads"]]
    for name in ("clotho_eval", "audiocaps", "wavcaps_400k", "allpairs_400k", "wavcaps_400k_q1", "wavcaps_400k_q32",
                 "wavcaps_400k_q128", "synthetic_10m_q1", "synthetic_10m_q32", "synthetic_10m_q128"):
        assert name in named, (name, named)
    for w in line["workloads"]:
        if "parity_gate" in w:
            assert w["parity_gate"]["ok"] and "roofline" in w and w["launches_per_search"] > 0
    assert any("strawman_torch_matmul_bf16_topk_ms" in w for w in line["workloads"])
    lit = line["literal_reference_loop_clotho_eval"]
    assert lit["cpu"]["value"] > 0 and lit["torch_cuda_as_written"]["value"] > 0


@pytest.mark.parametrize("world,balance", [(2, False), (3, True)])
def test_bench_sharded_protocol_on_standins(world, balance):
    """bench.py at N > 1 as torchrun drives it, on the CPU: N ranks over gloo, the real
    ShardedRelatedBank and SearchPipeline around the stand-in bank.  The run must gate itself green
    (same bits on all ranks and as one "GPU", timed result = gated result, e2e read-back = device
    result) with fixed and with moving shard boundaries, and rank 0 alone prints the line."""
    import json
    sys.path.insert(0, ROOT)
    import bench
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    cmd = [sys.executable, os.path.join(ROOT, "tests", "bench_standins.py"), "--steps", "4", "--world", str(world)]
    out = subprocess.run(cmd + (["--balance"] if balance else []), capture_output=True, text=True, timeout=1200, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    line = json.loads(lines[0])
    assert line["n_gpus"] == world and line["scaling"] == "strong" and line["cpu_baseline"] is None
    assert line["config"] == bench.workload_config("synthetic_10m", 300, 6000, 32, False, world, balance)
    assert line["config"]["bank_rows_per_gpu"] == 6000 // world
    gate = line["parity_gate"]
    assert gate["ok"] and gate["ranks"] == {"identical_on_all_ranks": True, "single_gpu_slice_queries": 300,
                                            "bit_identical_to_single_gpu": True, "ok": True}
    assert gate["timed_result_identical_to_gated"] and gate["e2e_readback_identical"]
    assert gate["e2e_timed_readback_identical"]
    roof = line["roofline"]
    assert len(roof["kernel_ms_per_rank"]) == world and roof["traffic"] is None
    assert set(roof["slowest_kernel_per_step_ms"]) == {"mean", "min", "max", "std_of_a_rank"}
    # every rank reads back only its slice of the merged rows
    assert line["e2e"]["d2h_bytes_per_step"] == -(-300 // world) * 32 * 12 * world
    assert line["e2e"]["h2d_bytes_per_step"] == 300 * 1024 * 4 * world
    if balance:
        moved = line["details"]["shard_balance"]
        assert moved["rebalances"] >= 1 and sum(moved["rows_per_rank_now"]) == 6000
    else:
        assert line["details"]["shard_balance"] is None
    for w in line["workloads"]:
        assert w["parity_gate"]["ok"] and w["parity_gate"]["ranks"]["identical_on_all_ranks"]


def test_default_writer_bytes_fuzz():
    """Random record structures (hypothesis): the template writer's stream is pickle.dump's, byte
    for byte — tensors of every supported dtype, views with offsets and strides, storages shared
    inside a record, long / non-ASCII strings, nested containers, big ints."""
    import io
    import pickle
    from hypothesis import HealthCheck, given, settings, strategies as st
    from zsaac_b200 import related_pipeline as rp

    dtypes = [torch.float32, torch.float64, torch.float16, torch.bfloat16, torch.int64, torch.int32,
              torch.int16, torch.int8, torch.uint8, torch.bool]

    @st.composite
    def tensors(draw):
        dtype = draw(st.sampled_from(dtypes))
        shape = draw(st.lists(st.integers(0, 6), min_size=0, max_size=3))
        seed = draw(st.integers(0, 2 ** 16))
        g = torch.Generator().manual_seed(seed)
        t = (torch.randn(tuple(shape), generator=g) * 50).to(dtype)
        op = draw(st.sampled_from(["plain", "t", "slice", "step", "row", "expand"]))
        if op == "t" and t.dim() >= 2:
            t = t.transpose(0, 1)
        elif op == "slice" and t.dim() >= 1 and t.shape[0] > 1:
            t = t[1:]
        elif op == "step" and t.dim() >= 1 and t.shape[-1] > 1:
            t = t[..., ::2]
        elif op == "row" and t.dim() >= 2 and t.shape[0] > 0:
            t = t[t.shape[0] - 1]
        elif op == "expand" and t.dim() >= 1:
            t = t.unsqueeze(0).expand(3, *t.shape)
        return t

    scalars = st.one_of(st.none(), st.booleans(), st.integers(-2 ** 70, 2 ** 70), st.floats(allow_nan=False),
                        st.text(max_size=300), st.binary(max_size=40))
    values = st.recursive(st.one_of(scalars, tensors()),
                          lambda inner: st.one_of(st.lists(inner, max_size=4), st.tuples(inner, inner),
                                                  st.dictionaries(st.text(max_size=8), inner, max_size=3)),
                          max_leaves=8)
    records = st.dictionaries(st.text(min_size=1, max_size=12), values, min_size=0, max_size=6)

    @settings(max_examples=150, deadline=None, suppress_health_check=list(HealthCheck))
    @given(records, st.booleans())
    def check(record, share):
        if share:                                       # the same tensor object / storage twice
            first = next((v for v in record.values() if isinstance(v, torch.Tensor)), None)
            if first is not None:
                record = dict(record, again=first, view=first.reshape(-1)[:1] if first.is_contiguous() else first)
        a, b = io.BytesIO(), io.BytesIO()
        rp._dump_record(record, a, False)
        pickle.dump(record, b)
        assert a.getvalue() == b.getvalue()

    check()


def test_fast_reader_fuzz():
    """Random records through pickle.dump, read back by pickle.load and by the direct parser (with
    its remembered stream templates warm and cold): same structure, dtypes, shapes, strides, values."""
    import io
    import pickle
    from hypothesis import HealthCheck, given, settings, strategies as st
    from zsaac_b200 import related_pipeline as rp

    dtypes = [torch.float32, torch.float64, torch.float16, torch.bfloat16, torch.int64, torch.int32,
              torch.int16, torch.int8, torch.uint8, torch.bool]

    @st.composite
    def tensors(draw):
        dtype = draw(st.sampled_from(dtypes))
        shape = draw(st.lists(st.integers(0, 5), min_size=0, max_size=3))
        g = torch.Generator().manual_seed(draw(st.integers(0, 2 ** 16)))
        t = (torch.randn(tuple(shape), generator=g) * 50).to(dtype)
        op = draw(st.sampled_from(["plain", "t", "slice", "step"]))
        if op == "t" and t.dim() >= 2:
            t = t.transpose(0, 1)
        elif op == "slice" and t.dim() >= 1 and t.shape[0] > 1:
            t = t[1:]
        elif op == "step" and t.dim() >= 1 and t.shape[-1] > 1:
            t = t[..., ::2]
        return t

    values = st.one_of(st.none(), st.integers(-10 ** 6, 10 ** 6), st.text(max_size=20), tensors(),
                       st.lists(tensors(), max_size=3))
    records = st.lists(st.dictionaries(st.text(min_size=1, max_size=8), values, max_size=5), max_size=4)

    def same(x, y):
        if isinstance(y, torch.Tensor):
            return (type(x) is torch.Tensor and x.dtype == y.dtype and x.shape == y.shape
                    and x.stride() == y.stride() and torch.equal(x, y) and not x.requires_grad)
        if isinstance(y, dict):
            return isinstance(x, dict) and list(x) == list(y) and all(same(x[k], y[k]) for k in y)
        if isinstance(y, list):
            return isinstance(x, list) and len(x) == len(y) and all(same(a, b) for a, b in zip(x, y))
        return type(x) is type(y) and x == y

    @settings(max_examples=150, deadline=None, suppress_health_check=list(HealthCheck))
    @given(records, st.booleans())
    def check(recs, cold):
        if cold:
            rp._READ_TEMPLATES.clear()
        blob = pickle.dumps(recs)
        want = pickle.loads(blob)
        got = rp._FastTensorUnpickler(io.BytesIO(blob)).load()
        assert same(got, want)

    check()
