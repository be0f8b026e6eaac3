"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the Integrated-Gradients call the reference
makes -- ``IntegratedGradients(ModelWrapper(model)).attribute(inputs=(video, audio), baselines=(zeros, zeros),
additional_forward_args=mask, target=target, n_steps=n_steps)`` at train2.py:826-834 (served copy:
back-end/app/libs/inference.py:313-321).

The algorithm itself lives in a third-party dependency that is NOT in /root/reference and not installed in this image:
``captum>=0.6.0`` (back-end/requirements.txt:16).  This file restates Captum's published algorithm for the defaults that
call uses (captum/attr/_core/integrated_gradients.py ``_attribute``; captum/attr/_utils/approximation_methods.py
``gauss_legendre_builders``): method "gausslegendre", multiply_by_inputs True, internal_batch_size None.

Pinning: Captum's own outputs cannot be produced here ("parity unpinned" against Captum itself).  What IS pinned:
tests/golden/ig_v2_b8_t5_mask.npz holds this restatement evaluated on the UNMODIFIED reference model class
(train2.MultimodalEmotionModel through the reference's own ModelWrapper, tests/golden/make_golden_ig.py), and the
completeness axiom sum(attr) = f(x) - f(baseline) that any correct IG satisfies is checked on it to quadrature accuracy.
"""
import numpy as np
import torch


def gauss_legendre(n_steps):
    """Captum ``gauss_legendre_builders``: step_sizes(n) = 0.5 * leggauss(n)[1], alphas(n) = 0.5 * (1 + leggauss(n)[0])."""
    t, w = np.polynomial.legendre.leggauss(n_steps)
    return list(0.5 * (1.0 + t)), list(0.5 * w)


def integrated_gradients(forward_fn, inputs, baselines, additional_forward_args, target, n_steps=50):
    """forward_fn(*inputs, additional_forward_args) -> logits [B, C]; inputs / baselines: tuples of tensors; target: [B]
    long.  Follows Captum's ``_attribute`` step by step."""
    alphas, step_sizes = gauss_legendre(n_steps)
    B = inputs[0].shape[0]
    # scaled_features_tpl: torch.cat over alphas -> step-major [n_steps * B, ...], requires_grad
    scaled = tuple(torch.cat([b + a * (x - b) for a in alphas], dim=0).requires_grad_(True) for x, b in zip(inputs, baselines))
    # additional args and targets are repeated n_steps times along dim 0 (_expand_additional_forward_args / _expand_target)
    extra = additional_forward_args.repeat(n_steps, *([1] * (additional_forward_args.dim() - 1))) \
        if additional_forward_args is not None else None
    tgt = target.repeat(n_steps)
    out = forward_fn(*scaled, extra)
    selected = out.gather(1, tgt.view(-1, 1)).squeeze(1)            # _select_targets
    grads = torch.autograd.grad(torch.unbind(selected), scaled)     # gradient of each sample's own output
    attrs = []
    for g, x, b in zip(grads, inputs, baselines):
        w = torch.tensor(step_sizes, dtype=g.dtype, device=g.device).view(n_steps, 1)
        scaled_g = g.contiguous().view(n_steps, -1) * w             # scaled_grads
        total = scaled_g.view((n_steps, B) + tuple(g.shape[1:])).sum(dim=0)   # _reshape_and_sum
        attrs.append(total * (x - b))                               # multiply_by_inputs
    return tuple(attrs)


def aggregate_importances(attr_video, attr_audio, abs_sum=True):
    """train2.py:841-866."""
    if abs_sum:
        attr_video, attr_audio = attr_video.abs(), attr_audio.abs()
    return attr_video.sum(dim=1), attr_audio
