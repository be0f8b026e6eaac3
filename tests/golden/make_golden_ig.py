"""Golden Integrated-Gradients vectors from the UNMODIFIED reference model class.

    python tests/golden/make_golden_ig.py          (build container only: needs /root/reference)

captum is not installed here, so the reference's compute_attributions (train2.py:776-838) cannot run as is.  This script
evaluates oracle/ig_oracle.py (the restatement of Captum's algorithm) on the reference's own
``ModelWrapper(MultimodalEmotionModel)`` (train2.py:28-38) in float64, target = predicted class like train2.py:819-823,
and stores the attributions, the targets and f(x) - f(0) for the completeness check.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import detgen  # noqa: E402
from make_golden import import_reference, load, zero_dropout  # noqa: E402
from oracle import ig_oracle  # noqa: E402

N_STEPS = 12


def main():
    _, ref_v2 = import_reference()
    name, B, T = "v2_b8_t5_mask", 8, 5
    model = ref_v2.MultimodalEmotionModel(max_seq_len=T + 1, fusion_num_layers=2, fusion_dropout=0.0,
                                          classifier_hidden_dim=512, classifier_dropout=0.0)
    zero_dropout(model)
    load(model, detgen.make_params("v2", max_seq_len=T + 1, hidden=512))
    model = model.double().eval()
    v, a, m, _ = detgen.make_batch(B, T, tag=name)
    video, audio, mask = torch.from_numpy(v).double(), torch.from_numpy(a).double(), torch.from_numpy(m)
    wrapper = ref_v2.ModelWrapper(model)
    # the fused nn.TransformerEncoder fast path has no backward; autograd on the inputs disables it by itself
    with torch.no_grad():
        logits = wrapper(video, audio, mask)
        target = logits.argmax(dim=-1)
        base_logits = wrapper(torch.zeros_like(video), torch.zeros_like(audio), mask)
    av, aa = ig_oracle.integrated_gradients(lambda vv, aa_, mk: wrapper(vv, aa_, mk), (video, audio),
                                            (torch.zeros_like(video), torch.zeros_like(audio)), mask, target, N_STEPS)
    delta = (logits - base_logits).gather(1, target.view(-1, 1)).squeeze(1)
    out = dict(n_steps=N_STEPS, target=target.numpy(), attr_video=av.detach().numpy(), attr_audio=aa.detach().numpy(),
               delta=delta.numpy(), logits=logits.numpy())
    path = os.path.join(HERE, "ig_" + name + ".npz")
    np.savez_compressed(path, **out)
    tot = av.detach().flatten(1).sum(1) + aa.detach().sum(1)
    print("wrote", path, os.path.getsize(path), "bytes; completeness gap", float((tot - delta).abs().max()))


if __name__ == "__main__":
    main()
