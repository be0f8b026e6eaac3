"""CPU-side checks: the C-ABI library loads and exports every symbol include/mmer.h declares, the
drop-in modules keep the reference's state_dict contract, the product path refuses to run without
CUDA (no fallback), and the data-parallel gradient arithmetic is right (gloo, world_size 2)."""
import ctypes
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import detgen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "mmer.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mmer_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    import mmer_b200
    lib = mmer_b200._lib.load()
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mmer.h but not exported"
    assert set(names) == set(mmer_b200._lib.SIGNATURES), "ctypes signatures and header disagree"
    assert lib.mmer_version() == 100
    # struct layout agreement between the header and ctypes for the two argument structs
    raw = ctypes.CDLL(mmer_b200._lib.LIB_PATH)
    assert raw.mmer_debug_set(99, 1) != 0 and raw.mmer_debug_get(0) == 0


def test_struct_sizes_match_the_header():
    """Compile a 3-line C program against include/mmer.h and compare sizeof with the ctypes mirrors."""
    import subprocess
    import tempfile
    import mmer_b200
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write('#include <stdio.h>\n#include "mmer.h"\nint main(){printf("%zu %zu\\n", sizeof(mmer_gemm_args), '
                           'sizeof(mmer_model));return 0;}\n')
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        a, b = map(int, subprocess.check_output([exe]).split())
    assert a == ctypes.sizeof(mmer_b200._lib.GemmArgs)
    assert b == ctypes.sizeof(mmer_b200._lib.Model)


@pytest.mark.parametrize("variant", ["v2", "v1"])
def test_state_dict_contract(variant):
    import mmer_b200 as mm
    if variant == "v2":
        model = mm.MultimodalEmotionModel(max_seq_len=57, fusion_num_layers=2, classifier_hidden_dim=512)
        spec = detgen.param_spec("v2", max_seq_len=57, hidden=512)
    else:
        model = mm.v1.MultimodalEmotionModel(max_seq_len=17)
        spec = detgen.param_spec("v1", max_seq_len=17)
    sd = model.state_dict()
    assert [k for k, _ in spec] == list(sd.keys())          # same names, same order
    for k, shape in spec:
        assert tuple(sd[k].shape) == tuple(shape), k
    # loads a reference-shaped dict strictly, both directions
    P = {k: torch.from_numpy(np.asarray(v)) for k, v in detgen.make_params(
        variant, **(dict(max_seq_len=57, hidden=512) if variant == "v2" else dict(max_seq_len=17))).items()}
    model.load_state_dict(P, strict=True)
    for k, v in model.state_dict().items():
        assert torch.equal(v, P[k]), k
    n = sum(p.numel() for p in model.parameters())
    assert n == (7_785_990 if variant == "v2" else 13_672_198)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_state_dict_equals_reference_classes():
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    ref_v1, ref_v2 = make_golden.import_reference()
    import mmer_b200 as mm
    torch.manual_seed(0)
    r2 = ref_v2.MultimodalEmotionModel(max_seq_len=17, classifier_hidden_dim=512)
    torch.manual_seed(0)
    o2 = mm.MultimodalEmotionModel(max_seq_len=17, classifier_hidden_dim=512)
    torch.manual_seed(0)
    r1 = ref_v1.MultimodalEmotionModel(max_seq_len=17)
    torch.manual_seed(0)
    o1 = mm.v1.MultimodalEmotionModel(max_seq_len=17)
    for r, o in ((r2, o2), (r1, o1)):
        rs, os_ = r.state_dict(), o.state_dict()
        assert list(rs.keys()) == list(os_.keys())
        for k in rs:
            assert rs[k].shape == os_[k].shape and torch.equal(rs[k], os_[k]), k   # same default init too
        o.load_state_dict(rs, strict=True)
        r.load_state_dict(o.state_dict(), strict=True)
    # use_layernorm=False variants of the two sub-modules (train2.py:96,104-105,121 / :208,215): Identity norms in the
    # fusion module, BatchNorm1d in the head -- same keys, shapes, buffers and default initialisation
    for ref_cls, our_cls, kw in ((ref_v2.CrossModalFusion, mm.CrossModalFusion, dict(num_layers=2, max_seq_len=17)),
                                 (ref_v2.EmotionClassifier, mm.EmotionClassifier, dict(hidden_dim=512))):
        torch.manual_seed(0)
        r = ref_cls(use_layernorm=False, **kw)
        torch.manual_seed(0)
        o = our_cls(use_layernorm=False, **kw)
        rs, os_ = r.state_dict(), o.state_dict()
        assert list(rs.keys()) == list(os_.keys())
        for k in rs:
            assert rs[k].shape == os_[k].shape and torch.equal(rs[k], os_[k]), k
        o.load_state_dict(rs, strict=True)
    # attribute reads the reference's logging performs (train2.py:536-544, train.py:262-272)
    assert o2.fusion.video_proj.in_features == 768 and o2.fusion.audio_proj.in_features == 1024
    assert o2.fusion.video_proj.out_features == 512 and o2.classifier.net[-1].out_features == 6
    assert o2.fusion.pos_embed.size(1) == 17 and o2.fusion.num_layers == 2 and o2.fusion.num_heads == 8
    assert isinstance(o2.fusion.dropout, float) and isinstance(o2.classifier.dropout, float)
    assert o1.classifier.fc2.out_features == 6 and o1.fusion.num_layers == 4


def test_product_path_has_no_cpu_fallback():
    import mmer_b200 as mm
    model = mm.MultimodalEmotionModel(max_seq_len=6)
    with pytest.raises(mm.MmerError):
        model(torch.zeros(2, 5, 768), torch.zeros(2, 1024))
    with pytest.raises(mm.MmerError):
        mm.ops.cast_bf16(torch.zeros(8))
    with pytest.raises(mm.MmerError):
        mm.FocalLoss()(torch.zeros(4, 6), torch.zeros(4, dtype=torch.int64))


def test_oracle_is_not_imported_by_the_product():
    pkg = os.path.join(ROOT, "multi-modal-emotion-recognition_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "oracle" not in src, f
    assert "oracle" not in open(os.path.join(ROOT, "mmer_b200", "__init__.py")).read()


# ----------------------------------------------------------------------------- data parallel (gloo)
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import fusion_oracle as O
    from mmer_b200.trainer import allreduce_flat_gradients
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    dims = dict(video_dim=48, audio_dim=40, fused=64, max_seq_len=5, layers=1, hidden=32, ffn=128)
    P = {k: torch.from_numpy(np.asarray(v)).double() for k, v in detgen.make_params("v2", **dims).items()}
    v, a, m, y = detgen.make_batch(8, 4, tag="dp", video_dim=48, audio_dim=40)
    video, audio, mask, labels = (torch.from_numpy(x) for x in (v, a, m, y))
    alpha = torch.tensor([1, 1, 1, 1, 1.2, 1.2], dtype=torch.float64)

    def grads(sl):
        leaf = {k: t.clone().requires_grad_(True) for k, t in P.items()}
        _, logits, _, _ = O.model_forward_v2(leaf, video[sl].double(), audio[sl].double(), mask[sl], num_heads=2)
        O.focal_loss(logits, labels[sl], 2.0, alpha).backward()
        return torch.cat([leaf[k].grad.flatten() for k in sorted(leaf)])

    shard = slice(rank * 4, (rank + 1) * 4)
    flat = grads(shard)
    scale = allreduce_flat_gradients(flat, None)          # sum across ranks, returns 1/world
    full = grads(slice(0, 8))
    err = float(((flat * scale) - full).abs().max() / full.abs().max())
    q.put((rank, err, scale))
    dist.destroy_process_group()


def _bucket_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from mmer_b200.trainer import allreduce_buckets, allreduce_flat_gradients
    import mmer_b200 as mm
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    model = mm.MultimodalEmotionModel(video_dim=48, audio_dim=40, fused_dim=64, max_seq_len=5, fusion_num_layers=2,
                                      fusion_num_heads=2, classifier_hidden_dim=32)
    ctx = model._engine.ctx
    _, total = ctx.layout()
    ranges = ctx.bucket_ranges()
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(total, generator=g, dtype=torch.float64)
    a, b = flat.clone(), flat.clone()
    s1 = allreduce_flat_gradients(a, None)
    s2 = allreduce_buckets(b, ranges, None)
    q.put((rank, bool(torch.equal(a, b)), s1, s2))
    dist.destroy_process_group()


def test_gradient_buckets_tile_the_flat_buffer_in_backward_order():
    """Flat layout = embed | layer 0 | ... | layer L-1 | head; bucket_ranges() lists them in the order backward
    finishes them (head, layers last to first, embed): contiguous, disjoint, covering every parameter."""
    import mmer_b200 as mm
    model = mm.MultimodalEmotionModel(max_seq_len=17, fusion_num_layers=3, classifier_hidden_dim=512)
    ctx = model._engine.ctx
    offsets, total = ctx.layout()
    ranges = ctx.bucket_ranges()
    assert len(ranges) == 3 + 2
    assert ranges[0][1] == total and ranges[-1][0] == 0
    for (lo, hi), (lo2, hi2) in zip(ranges[1:], ranges[:-1]):
        assert hi == lo2 and lo < hi                      # contiguous, descending: completion order
    named = dict(model.named_parameters())

    def bucket_of(name):
        o = offsets[id(named[name])]
        return next(k for k, (lo, hi) in enumerate(ranges) if lo <= o < hi)

    assert bucket_of("classifier.net.8.weight") == 0 and bucket_of("fusion.out_norm.weight") == 0
    assert bucket_of("fusion.transformer.layers.2.linear1.weight") == 1
    assert bucket_of("fusion.transformer.layers.0.self_attn.in_proj_weight") == 3
    assert bucket_of("fusion.pos_embed") == 4 and bucket_of("fusion.video_proj.weight") == 4
    for n, p in named.items():                            # every parameter lies wholly inside one bucket
        o = offsets[id(p)]
        k = bucket_of(n)
        assert o + p.numel() <= ranges[k][1]


def test_bucketed_allreduce_equals_single_allreduce_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, same, s1, s2 in res:
        assert same and s1 == s2 == 0.5


def test_data_parallel_gradient_average_equals_global_batch_gloo():
    """Equal shards + per-rank mean loss + (sum all-reduce) * 1/world == global-batch gradient
    (SURVEY.md 8e), checked with two gloo ranks on CPU."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, err, scale in res:
        assert scale == 0.5
        assert err < 1e-10, (rank, err)


def test_multicast_adam_shards_tile_the_flat_buffer():
    """The fused NVLS step gives every element to exactly one rank, in 16-byte units."""
    from mmer_b200.trainer import shard_range
    for n in (64, 7_766_720, 4 * 1000 + 64, 128):
        for world in (1, 2, 3, 4, 8, 16):
            prev = 0
            for r in range(world):
                lo, hi = shard_range(n, world, r)
                assert lo == prev and lo <= hi <= n and lo % 4 == 0 and (hi % 4 == 0 or hi == n)
                prev = hi
            assert prev == n


def test_host_logic_properties_hypothesis():
    """Property checks of the host-side pieces that decide who owns which bytes: the multicast shards, the gradient buckets
    and the evaluation metrics (against scikit-learn)."""
    from hypothesis import given, settings, strategies as st
    from mmer_b200.trainer import shard_range
    from mmer_b200.evaluation import metrics_from_confusion
    from sklearn.metrics import confusion_matrix, precision_recall_fscore_support

    @settings(max_examples=200, deadline=None)
    @given(st.integers(1, 1 << 24).map(lambda k: 64 * k), st.integers(1, 64))
    def shards(n, world):
        covered = 0
        for r in range(world):
            lo, hi = shard_range(n, world, r)
            assert lo == covered and lo % 4 == 0 and hi % 4 == 0 and hi <= n
            covered = hi
        assert covered == n

    @settings(max_examples=100, deadline=None)
    @given(st.lists(st.tuples(st.integers(0, 5), st.integers(0, 5)), min_size=1, max_size=200))
    def metrics(pairs):
        y = [a for a, _ in pairs]
        p = [b for _, b in pairs]
        got = metrics_from_confusion(confusion_matrix(y, p, labels=list(range(6))))
        for avg in ("macro", "micro"):
            pr, rc, f1, _ = precision_recall_fscore_support(y, p, average=avg, zero_division=0)
            assert abs(got[avg + "_precision"] - pr) < 1e-12 and abs(got[avg + "_recall"] - rc) < 1e-12
            assert abs(got[avg + "_f1"] - f1) < 1e-12
        assert got["correct"] == sum(int(a == b) for a, b in pairs)

    shards()
    metrics()
