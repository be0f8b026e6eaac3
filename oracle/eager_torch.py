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
        layer = nn.TransformerEncoderLayer(d_model=fused_dim, nhead=num_heads, dim_feedforward=4 * fused_dim, dropout=dropout,
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


# ------------------------------------------------------------------------------------------------ train.py variant
class EagerFusionV1(nn.Module):
    """train.py:47-106 on stock modules: Linear -> BatchNorm1d over (B*T) -> +pos_embed -> 4-layer post-norm encoder ->
    masked mean pooling (AdaptiveAvgPool1d without a mask)."""

    def __init__(self, video_dim=768, audio_dim=1024, fused_dim=512, num_layers=4, num_heads=8, dropout=0.01, max_seq_len=101):
        super().__init__()
        self.video_proj = nn.Linear(video_dim, fused_dim)            # train.py:49
        self.audio_proj = nn.Linear(audio_dim, fused_dim)            # train.py:50
        self.bn_video = nn.BatchNorm1d(fused_dim)                    # train.py:51
        self.bn_audio = nn.BatchNorm1d(fused_dim)                    # train.py:52
        self.pos_embed = nn.Parameter(torch.randn(1, max_seq_len, fused_dim))           # train.py:53
        layer = nn.TransformerEncoderLayer(d_model=fused_dim, nhead=num_heads, dim_feedforward=2048, dropout=dropout)
        self.transformer = nn.TransformerEncoder(layer, num_layers=num_layers)          # train.py:54-57
        self.pool = nn.AdaptiveAvgPool1d(1)                          # train.py:58

    def forward(self, video_feats, audio_feats, mask=None):
        b, t, _ = video_feats.shape
        video = self.bn_video(self.video_proj(video_feats).transpose(1, 2)).transpose(1, 2)              # train.py:66-69
        audio = self.bn_audio(self.audio_proj(audio_feats.unsqueeze(1)).transpose(1, 2)).transpose(1, 2)  # train.py:71-74
        x = torch.cat([video, audio], dim=1) + self.pos_embed[:, :t + 1, :]                              # train.py:76-77
        full = None
        if mask is not None:
            full = torch.cat([mask, torch.zeros(b, 1, dtype=torch.bool, device=mask.device)], dim=1)    # train.py:80-82
        x = self.transformer(x.transpose(0, 1), src_key_padding_mask=full).transpose(0, 1)               # train.py:86-96
        if full is not None:
            keep = (~full).float().unsqueeze(-1)
            return (x * keep).sum(1) / keep.sum(1).clamp(min=1e-6)                                       # train.py:100-102
        return self.pool(x.transpose(1, 2)).squeeze(-1)                                                  # train.py:104


class EagerClassifierV1(nn.Module):
    def __init__(self, input_dim=512, num_classes=6, dropout=0.01):
        super().__init__()
        self.fc1 = nn.Linear(input_dim, input_dim // 2)             # train.py:115
        self.bn_fc1 = nn.BatchNorm1d(input_dim // 2)                # train.py:116
        self.dropout1 = nn.Dropout(dropout)
        self.fc2 = nn.Linear(input_dim // 2, num_classes)           # train.py:118
        self.dropout2 = nn.Dropout(dropout)                         # (unused by forward, present in the state-less tree)

    def forward(self, fused):
        logits = self.fc2(self.dropout1(F.relu(self.bn_fc1(self.fc1(fused)))))         # train.py:124-128
        return F.softmax(logits, dim=-1), logits                    # train.py:129-130


class EagerModelV1(nn.Module):
    """train.py:133-142; state_dict keys equal the reference's (tests/test_oracle_golden.py pins it on the v1 goldens)."""

    def __init__(self, video_dim=768, audio_dim=1024, fused_dim=512, num_classes=6, max_seq_len=101, dropout=0.01):
        super().__init__()
        self.fusion = EagerFusionV1(video_dim, audio_dim, fused_dim, dropout=dropout, max_seq_len=max_seq_len)
        self.classifier = EagerClassifierV1(fused_dim, num_classes, dropout=dropout)

    def forward(self, video_feats, audio_feats, mask=None):
        return self.classifier(self.fusion(video_feats, audio_feats, mask))
