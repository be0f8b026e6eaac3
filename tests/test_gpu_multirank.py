"""Multi-rank GPU tests of the data-parallel step: run only where >= 2 GPUs are visible
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu`); skipped on a 1-GPU box.

1. ``dp_selfcheck``: the default exchange (multicast reduce-scatter + Adam + all-gather kernel when the fabric offers
   it) against the NCCL all-reduce path from identical state: ranks bit-identical, weights equal to summation-order
   noise, shadow == bf16(weights); with and without gradient clipping.
2. A 2-rank step over two half batches equals the single-GPU step over the whole batch (the identity the reference's
   single-process loop, train2.py:570-579, defines for any data-parallel run of it), fp32 mode, both exchanges.
3. SyncBatchNorm: the train.py (BatchNorm) model on 2 ranks equals the single-GPU global-batch golden.
"""
import os
import socket
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALPHA = [1, 1, 1, 1, 1.2, 1.2]


def _need_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _entry(name, rank, world, port, q, *args):
    """Child process: run the named worker; a failure travels to the parent as text instead of a silent exit."""
    import traceback
    try:
        globals()[name](rank, world, port, q, *args)
    except BaseException:
        q.put(("error", rank, traceback.format_exc()))
        raise


def _spawn(worker, world, *args, timeout=240):
    import queue as _queue
    import time
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_entry, args=(worker.__name__, r, world, port, q) + args) for r in range(world)]
    for p in procs:
        p.start()
    res = []
    deadline = time.time() + timeout
    try:
        while len(res) < world:
            try:
                item = q.get(timeout=2)
            except _queue.Empty:
                dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
                assert not dead, f"a worker died without reporting (exit codes {dead})"
                assert time.time() < deadline, "workers timed out"
                continue
            assert item[0] != "error", f"rank {item[1]} failed:\n{item[2]}"
            res.append(item)
    finally:
        for p in procs:
            p.join(60 if len(res) == world else 5)
            if p.is_alive():
                p.kill()
    for p in procs:
        assert p.exitcode == 0
    return sorted(res, key=lambda r: r[0])


def _init(rank, world, port):
    import torch.distributed as dist
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    return dev, dist


def _selfcheck_worker(rank, world, port, q):
    dev, dist = _init(rank, world, port)
    import mmer_b200 as mm
    out = [mm.dp_selfcheck(dev), mm.dp_selfcheck(dev, clip=0.05), mm.dp_selfcheck(dev, dtype=torch.float32, samples=64)]
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_selfcheck_default_exchange_equals_nccl_two_ranks():
    _need_two_gpus()
    res = _spawn(_selfcheck_worker, 2)
    for rank, reports in res:
        for rep in reports:
            print(rank, rep)
            assert rep["ranks_bit_identical"], rep
            assert rep["shadow_equals_bf16_weights"], rep
            assert rep["max_abs_moved"] > 1e-6, rep
            # only the summation order of the gradient differs between the exchanges
            assert rep["max_abs_vs_nccl"] <= max(2e-4 * rep["max_abs_moved"], 8e-9), rep


def _global_batch_worker(rank, world, port, q, path, mode):
    dev, dist = _init(rank, world, port)
    import mmer_b200 as mm
    blob = torch.load(path)
    B = blob["video"].shape[0]
    sl = slice(rank * B // world, (rank + 1) * B // world)
    model = mm.MultimodalEmotionModel(max_seq_len=blob["video"].shape[1] + 1, classifier_hidden_dim=512, fusion_dropout=0.0,
                                      classifier_dropout=0.0)
    if rank != 0:                                   # replicas other than rank 0 start from DIFFERENT weights on purpose:
        torch.manual_seed(1234)                     # FusedTrainStep must broadcast rank 0's (DistributedDataParallel semantics)
        model = mm.MultimodalEmotionModel(max_seq_len=blob["video"].shape[1] + 1, classifier_hidden_dim=512,
                                          fusion_dropout=0.0, classifier_dropout=0.0)
    else:
        model.load_state_dict(blob["init"])
    model.to(dev).train()
    step = mm.FusedTrainStep(model, lr=1e-2, weight_decay=1e-4, eps=1.0, loss="focal", alpha=torch.tensor(ALPHA),
                             compute_dtype=torch.float32, dp_mode=mode, clip_grad_norm=blob["clip"])
    for _ in range(2):
        loss, _ = step.step(blob["video"][sl].to(dev), blob["audio"][sl].to(dev), blob["mask"][sl].to(dev),
                            blob["labels"][sl].to(dev))
    torch.cuda.synchronize()
    err = max(float((p.detach().cpu() - blob["after"][k]).abs().max()) for k, p in model.named_parameters())
    moved = max(float((blob["after"][k] - blob["init"][k]).abs().max()) for k in blob["after"])
    # optimizer state survives a (collective) state_dict round trip in either mode
    sd = step.opt.state_dict()
    full_m = torch.cat([sd["state"][i]["exp_avg"].flatten().cpu() for i in range(len(sd["state"]))])
    q.put((rank, err, moved, step.dp_mode, float(full_m.abs().sum())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,clip", [("auto", None), ("nccl", None), ("auto", 0.05)])
def test_two_rank_step_equals_single_gpu_global_batch_step_fp32(mode, clip):
    _need_two_gpus()
    import detgen
    import mmer_b200 as mm
    B, T = 8, 5
    v, a, m, y = detgen.make_batch(B, T, tag="dp2")
    video, audio, mask, labels = (torch.from_numpy(x) for x in (v, a, m, y))
    torch.manual_seed(0)
    model = mm.MultimodalEmotionModel(max_seq_len=T + 1, classifier_hidden_dim=512, fusion_dropout=0.0,
                                      classifier_dropout=0.0)
    init = {k: t.detach().clone() for k, t in model.state_dict().items()}
    model.cuda().train()
    step = mm.FusedTrainStep(model, lr=1e-2, weight_decay=1e-4, eps=1.0, loss="focal", alpha=torch.tensor(ALPHA),
                             compute_dtype=torch.float32, clip_grad_norm=clip)
    for _ in range(2):
        step.step(video.cuda(), audio.cuda(), mask.cuda(), labels.cuda())
    after = {k: p.detach().cpu().clone() for k, p in model.named_parameters()}
    m_single = float(torch.cat([s["exp_avg"].flatten().cpu() for s in step.opt.state_dict()["state"].values()]).abs().sum())
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "blob.pt")
        torch.save(dict(video=video, audio=audio, mask=mask, labels=labels, init=init, after=after, clip=clip), path)
        res = _spawn(_global_batch_worker, 2, path, mode)
    for rank, err, moved, used, m_sum in res:
        print(rank, used, err, moved, m_sum, m_single)
        assert moved > 1e-5
        assert err < 2e-4 * moved + 1e-7, (rank, err, moved)
        assert abs(m_sum - m_single) < 1e-3 * m_single        # gathered Adam moments == single-GPU moments


def _syncbn_worker(rank, world, port, q, name):
    dev, dist = _init(rank, world, port)
    import detgen
    import mmer_b200 as mm
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    B, T = int(g["B"]), int(g["T"])
    model = mm.v1.MultimodalEmotionModel(max_seq_len=T + 1, sync_batchnorm=True)
    model.fusion.dropout = model.classifier.dropout = 0.0
    P = {k: torch.from_numpy(np.asarray(v)) for k, v in detgen.make_params("v1", max_seq_len=T + 1).items()}
    model.load_state_dict(P, strict=True)
    model.to(dev).train()
    v, a, m, y = detgen.make_batch(B, T, tag=name)
    sl = slice(rank * B // world, (rank + 1) * B // world)
    video, audio, labels = (torch.from_numpy(x)[sl].to(dev) for x in (v, a, y))
    mask = torch.from_numpy(m)[sl].to(dev) if int(g["use_mask"]) else None
    step = mm.FusedTrainStep(model, lr=1e-4, weight_decay=1e-4, loss="focal", gamma=2.0, alpha=None,
                             compute_dtype=torch.float32, dp_mode="nccl")
    loss, _ = step.step(video, audio, mask, labels)
    torch.cuda.synchronize()
    out = {"loss_local": float(loss)}
    sd = {k: t.detach().cpu().numpy() for k, t in model.state_dict().items()}
    errs = {}
    zero_grad_keys = {"fusion.video_proj.bias", "fusion.audio_proj.bias", "classifier.fc1.bias",
                      "fusion.transformer.layers.3.norm2.bias"}
    for k, p in model.named_parameters():
        if k in zero_grad_keys:
            continue                                  # analytically zero gradient: Adam turns noise into +-lr (see test_gpu_model)
        delta = sd[k].astype(np.float64) - P[k].numpy().astype(np.float64)
        ref = g["delta1/" + k]
        head = np.pad(delta.reshape(-1)[:16], (0, max(0, 16 - delta.size)))
        errs[k] = float(np.abs(head - ref[2:]).max())
    bn = {}
    for k, val in sd.items():
        if "running" in k:
            bn[k] = float(np.abs(val - g["bn_after_fwd/" + k]).max() / (np.abs(g["bn_after_fwd/" + k]).max() + 1e-12))
        if "tracked" in k:
            bn[k] = float(abs(int(val) - int(g["bn_after_fwd/" + k])))
    q.put((rank, out, errs, bn))
    dist.barrier()
    dist.destroy_process_group()


def test_sync_batchnorm_two_ranks_equal_single_gpu_global_batch_golden():
    """N4: train.py's BatchNorm model, batch 32 split over 2 ranks, equals the reference's single-process golden
    (v1_b32_t16_cfg1: Adam update of every parameter and the BatchNorm running statistics)."""
    _need_two_gpus()
    name = "v1_b32_t16_cfg1"
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    res = _spawn(_syncbn_worker, 2, name)
    losses = [r[1]["loss_local"] for r in res]
    assert abs(np.mean(losses) - float(g["loss/focal"])) < 1e-4 * max(1.0, float(g["loss/focal"]))
    for rank, out, errs, bn in res:
        worst = max(errs.items(), key=lambda kv: kv[1])
        print(rank, out, "worst delta error", worst, "bn", max(bn.values()))
        assert worst[1] < 3e-6, worst          # Adam's first step moves each coordinate by ~lr = 1e-4
        for k, e in bn.items():
            assert e < 1e-4, (k, e)
