"""Deterministic, platform-independent test data.

Weights and inputs are produced from an integer hash (splitmix64) of the element
index, converted to float with exact arithmetic, so that the golden-vector generator
(run once in the build container against /root/reference) and the tests (run
anywhere, including the GPU box where the reference is absent) reconstruct
bit-identical tensors without storing 30 MB of weights in git.
"""
from __future__ import annotations

import zlib
from typing import Dict, List, Optional, Tuple

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def det_array(shape, tag: str, scale: float = 1.0, offset: float = 0.0) -> np.ndarray:
    """float32 array, elements uniform in offset + [-scale, scale)."""
    n = int(np.prod(shape)) if len(shape) else 1
    seed = np.uint64(zlib.crc32(tag.encode()))
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64) + (seed << np.uint64(32))
    bits = _splitmix64(idx) >> np.uint64(40)               # 24 random bits
    u = bits.astype(np.float64) / float(1 << 24)           # exact
    out = (u * 2.0 - 1.0) * scale + offset
    return out.astype(np.float32).reshape(shape)


def det_ints(n: int, tag: str, lo: int, hi: int) -> np.ndarray:
    """int64 array uniform in [lo, hi]."""
    seed = np.uint64(zlib.crc32(tag.encode()))
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64) + (seed << np.uint64(32))
    bits = _splitmix64(idx) >> np.uint64(33)
    return (bits % np.uint64(hi - lo + 1)).astype(np.int64) + lo


def param_spec(variant: str, video_dim=768, audio_dim=1024, fused=512, classes=6, max_seq_len=17,
               layers: Optional[int] = None, hidden: Optional[int] = None, ffn: Optional[int] = None
               ) -> List[Tuple[str, Tuple[int, ...]]]:
    """state_dict names and shapes of the reference models."""
    v2 = variant == "v2"
    layers = layers if layers is not None else (2 if v2 else 4)
    ffn = ffn if ffn is not None else (4 * fused if v2 else 2048)
    spec: List[Tuple[str, Tuple[int, ...]]] = []
    if v2:
        spec.append(("fusion.pos_embed", (1, max_seq_len, fused)))
    spec += [("fusion.video_proj.weight", (fused, video_dim)), ("fusion.video_proj.bias", (fused,)),
             ("fusion.audio_proj.weight", (fused, audio_dim)), ("fusion.audio_proj.bias", (fused,))]
    if v2:
        spec += [("fusion.norm_video.weight", (fused,)), ("fusion.norm_video.bias", (fused,)),
                 ("fusion.norm_audio.weight", (fused,)), ("fusion.norm_audio.bias", (fused,))]
    else:
        spec = [("fusion.pos_embed", (1, max_seq_len, fused))] + spec
        for bn in ("fusion.bn_video", "fusion.bn_audio"):
            spec += [(bn + ".weight", (fused,)), (bn + ".bias", (fused,)),
                     (bn + ".running_mean", (fused,)), (bn + ".running_var", (fused,)),
                     (bn + ".num_batches_tracked", ())]
    for l in range(layers):
        p = f"fusion.transformer.layers.{l}."
        spec += [(p + "self_attn.in_proj_weight", (3 * fused, fused)), (p + "self_attn.in_proj_bias", (3 * fused,)),
                 (p + "self_attn.out_proj.weight", (fused, fused)), (p + "self_attn.out_proj.bias", (fused,)),
                 (p + "linear1.weight", (ffn, fused)), (p + "linear1.bias", (ffn,)),
                 (p + "linear2.weight", (fused, ffn)), (p + "linear2.bias", (fused,)),
                 (p + "norm1.weight", (fused,)), (p + "norm1.bias", (fused,)),
                 (p + "norm2.weight", (fused,)), (p + "norm2.bias", (fused,))]
    if v2:
        hidden = hidden if hidden is not None else fused // 2
        spec += [("fusion.out_norm.weight", (fused,)), ("fusion.out_norm.bias", (fused,)),
                 ("classifier.net.0.weight", (hidden, fused)), ("classifier.net.0.bias", (hidden,)),
                 ("classifier.net.1.weight", (hidden,)), ("classifier.net.1.bias", (hidden,)),
                 ("classifier.net.4.weight", (hidden, hidden)), ("classifier.net.4.bias", (hidden,)),
                 ("classifier.net.5.weight", (hidden,)), ("classifier.net.5.bias", (hidden,)),
                 ("classifier.net.8.weight", (classes, hidden)), ("classifier.net.8.bias", (classes,))]
    else:
        h = fused // 2
        spec += [("classifier.fc1.weight", (h, fused)), ("classifier.fc1.bias", (h,)),
                 ("classifier.bn_fc1.weight", (h,)), ("classifier.bn_fc1.bias", (h,)),
                 ("classifier.bn_fc1.running_mean", (h,)), ("classifier.bn_fc1.running_var", (h,)),
                 ("classifier.bn_fc1.num_batches_tracked", ()),
                 ("classifier.fc2.weight", (classes, h)), ("classifier.fc2.bias", (classes,))]
    return spec


def make_params(variant: str, tag: str = "w", **dims) -> Dict[str, np.ndarray]:
    """Deterministic weights with default-init-like magnitudes."""
    out: Dict[str, np.ndarray] = {}
    for name, shape in param_spec(variant, **dims):
        t = f"{tag}/{variant}/{name}"
        if name.endswith("num_batches_tracked"):
            out[name] = np.zeros((), dtype=np.int64)
        elif name.endswith("running_mean"):
            out[name] = det_array(shape, t, 0.1)
        elif name.endswith("running_var"):
            out[name] = det_array(shape, t, 0.25, 1.0)
        elif name == "fusion.pos_embed":
            out[name] = det_array(shape, t, 0.04 if variant == "v2" else 1.0)
        elif len(shape) == 2:
            out[name] = det_array(shape, t, 1.0 / np.sqrt(shape[1]))
        elif "norm" in name or ".bn_" in name or "net.1." in name or "net.5." in name:
            out[name] = det_array(shape, t, 0.2, 1.0 if name.endswith("weight") else 0.0)
        else:  # linear biases
            out[name] = det_array(shape, t, 0.05)
    return out


def make_batch(B: int, T: int, tag: str = "x", video_dim=768, audio_dim=1024, classes=6,
               ragged: bool = True):
    """(video (B,T,Dv) f32, audio (B,Da) f32, mask (B,T) bool True=pad, labels (B,) i64)."""
    video = det_array((B, T, video_dim), f"{tag}/video/{B}x{T}", 1.7)
    audio = det_array((B, audio_dim), f"{tag}/audio/{B}", 1.7)
    labels = det_ints(B, f"{tag}/labels/{B}", 0, classes - 1)
    if ragged:
        lens = det_ints(B, f"{tag}/lens/{B}x{T}", 1, T)
        lens[0] = T
    else:
        lens = np.full(B, T, dtype=np.int64)
    mask = np.arange(T)[None, :] >= lens[:, None]
    return video, audio, mask, labels


# ------------------------------------------------------------------ element-wise gradient fixtures
FULL_GRAD_MAX = 65536          # tensors up to this many elements are stored whole in the golden files
N_PROJ, N_PICK = 256, 4096     # larger ones: 256 seeded sparse +-1 projections of 4096 elements each


def projection_plan(numel: int, tag: str):
    """(index [N_PROJ, N_PICK] int64, sign [N_PROJ, N_PICK] float64) of the seeded projections of a tensor with
    ``numel`` elements; the plan depends only on (numel, tag), so generator and tests rebuild it bit for bit."""
    idx = det_ints(N_PROJ * N_PICK, tag + "/proj_idx", 0, numel - 1).reshape(N_PROJ, N_PICK)
    sign = det_ints(N_PROJ * N_PICK, tag + "/proj_sign", 0, 1).reshape(N_PROJ, N_PICK).astype(np.float64) * 2.0 - 1.0
    return idx, sign


def project(flat: np.ndarray, tag: str) -> np.ndarray:
    """The N_PROJ projections of a flattened float array (float64 accumulation)."""
    flat = np.asarray(flat, dtype=np.float64).reshape(-1)
    idx, sign = projection_plan(flat.size, tag)
    return (flat[idx] * sign).sum(axis=1)


def check_gradient_elementwise(g, name: str, grad: np.ndarray, rel: float) -> None:
    """Compare one parameter gradient with the golden file ``g`` ELEMENT-WISE: the whole tensor (``gradfull/``) or
    its 256 seeded projections (``gradproj/``).  ``rel`` is the allowed error relative to the tensor's rms value
    (a projection of N_PICK elements carries sqrt(N_PICK) times that)."""
    grad = np.asarray(grad, dtype=np.float64)
    norm = float(g["grad/" + name][0])
    rms = norm / np.sqrt(grad.size)
    floor = 2e-6                      # analytically-zero gradients (biases in front of a BatchNorm) hold noise
    if "gradfull/" + name in g.files:
        ref = np.asarray(g["gradfull/" + name], dtype=np.float64)
        assert ref.shape == grad.shape, name
        err = np.abs(grad - ref)
        tol = rel * max(rms, np.abs(ref).max() * 0.05) + floor
        # a ReLU pre-activation that rounds to the other side of zero (fp32 reference vs fp64 oracle / other summation
        # order) moves the few elements it feeds by one token's contribution: allow 0.2 % such elements, bounded
        assert (err > tol).mean() <= 2e-3 and err.max() <= max(50 * tol, 0.1 * rms), (name, err.max(), rms)
    else:
        ref = np.asarray(g["gradproj/" + name], dtype=np.float64)
        got = project(grad, name)
        err = np.abs(got - ref).max()
        assert err <= rel * rms * np.sqrt(N_PICK) + floor, (name, err, rms)
