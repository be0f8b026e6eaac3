"""TEST / BASELINE INFRASTRUCTURE, not product code: the reference's train2 model restated with STOCK torch.nn modules
(nn.TransformerEncoder -> cuBLAS + SDPA + elementwise kernels), i.e. what the reference itself launches on a GPU.

Only tests/ and bench.py's baseline legs may import this file.  Structure and hyper-parameters follow
train2.py:87-126 (CrossModalFusion.__init__), :128-193 (forward), :196-238 (EmotionClassifier), :241-292
(MultimodalEmotionModel); state_dict keys equal the reference's, so tests/golden parameters load with strict=True
and tests/test_oracle_golden.py pins its outputs against the reference's golden logits.
"""
import torch
from torch import nn
import torch.nn.functional as F


class EagerFusion(nn.Module):
    def __init__(self, video_dim=768, audio_dim=1024, fused_dim=512, num_layers=2, num_heads=8, dropout=0.1, max_seq_len=101):
        super().__init__()
        self.video_proj = nn.Linear(video_dim, fused_dim)            # train2.py:101
        self.audio_proj = nn.Linear(audio_dim, fused_dim)            # train2.py:102
        self.norm_video = nn.LayerNorm(fused_dim)                    # train2.py:104
        self.norm_audio = nn.LayerNorm(fused_dim)                    # train2.py:105
        self.pos_embed = nn.Parameter(torch.randn(1, max_seq_len, fused_dim) * 0.02)   # train2.py:108
        layer = nn.TransformerEncoderLayer(d_model=fused_dim, nhead=num_heads, dim_feedforward=2048, dropout=dropout,
                                           activation="relu", batch_first=False)       # train2.py:111-117
        self.transformer = nn.TransformerEncoder(layer, num_layers=num_layers)          # train2.py:118
        self.dropout_layer = nn.Dropout(dropout)
        self.out_norm = nn.LayerNorm(fused_dim)                      # train2.py:121

    def forward(self, video_feats, audio_feats, mask=None):
        b, t, _ = video_feats.shape
        video = self.norm_video(self.video_proj(video_feats))                            # train2.py:150-151
        audio = self.norm_audio(self.audio_proj(audio_feats)).unsqueeze(1)               # train2.py:153-154
        x = torch.cat([video, audio], dim=1) + self.pos_embed[:, :t + 1, :]              # train2.py:157-160
        x = self.dropout_layer(x)
        full = None
        if mask is not None:
            full = torch.cat([mask, torch.zeros(b, 1, dtype=torch.bool, device=mask.device)], dim=1)   # train2.py:164-169
        x = self.transformer(x.permute(1, 0, 2), src_key_padding_mask=full).permute(1, 0, 2)        # train2.py:172-181
        if full is not None:
            valid = (~full).unsqueeze(-1).to(x.dtype)
            pooled = (x * valid).sum(1) / valid.sum(1).clamp(min=1e-6)                   # train2.py:184-187
        else:
            pooled = x.mean(1)                                                           # train2.py:189
        return self.out_norm(pooled)                                                     # train2.py:191


class EagerClassifier(nn.Module):
    def __init__(self, input_dim=512, num_classes=6, hidden_dim=None, dropout=0.2):
        super().__init__()
        hidden_dim = hidden_dim or input_dim // 2                                        # train2.py:212-213
        self.net = nn.Sequential(nn.Linear(input_dim, hidden_dim), nn.LayerNorm(hidden_dim), nn.ReLU(inplace=True),
                                 nn.Dropout(dropout), nn.Linear(hidden_dim, hidden_dim), nn.LayerNorm(hidden_dim),
                                 nn.ReLU(inplace=True), nn.Dropout(dropout), nn.Linear(hidden_dim, num_classes))   # :217-229

    def forward(self, x):
        return self.net(x)


class EagerModel(nn.Module):
    def __init__(self, video_dim=768, audio_dim=1024, fused_dim=512, num_classes=6, max_seq_len=101, fusion_num_layers=2,
                 fusion_num_heads=8, fusion_dropout=0.1, classifier_hidden_dim=None, classifier_dropout=0.2):
        super().__init__()
        self.fusion = EagerFusion(video_dim, audio_dim, fused_dim, fusion_num_layers, fusion_num_heads, fusion_dropout,
                                  max_seq_len)
        self.classifier = EagerClassifier(fused_dim, num_classes, classifier_hidden_dim, classifier_dropout)

    def forward(self, video_feats, audio_feats, mask=None):
        logits = self.classifier(self.fusion(video_feats, audio_feats, mask))            # train2.py:288-289
        return F.softmax(logits, dim=-1), logits                                         # train2.py:290


def focal_loss(logits, targets, gamma=2.0, alpha=None):
    """train2.py:40-70 (= train.py:20-37), reduction 'mean'."""
    ce = F.cross_entropy(logits, targets, reduction="none")
    pt = torch.exp(-ce)
    fl = (1 - pt) ** gamma * ce
    if alpha is not None:
        fl = alpha[targets] * fl
    return fl.mean()
