"""The two callers of the fusion pool that the reference documents, written against the `aecf` API.

* ``VisionLanguageModel`` -- reference README.md:162-208 (BASELINE.json configs[2]): image 2048 and text 768
  features projected to 512, stacked as M = 2 modality tokens, fused, 1000-class head.
* ``XrayFusionModel`` -- reference xrays/train_xrays_example.py:108-237 (configs[3]): two encoders to
  hidden 256, presence detection by input norm, the pool (4 heads, M = 2) on the rows where both
  modalities are present, plain projections for single-modality rows, a shared classifier, and a
  curriculum stage that is switched on at run time.

``fusion`` selects the implementation of the four public names (default: this repo's ``aecf`` package);
the parity tests pass an oracle-backed stand-in to check the callers end to end.  Where the fusion package also
offers ``project_tokens`` / ``linear`` (aecf_b200 does: the library's own GEMMs writing the encoder outputs straight into the
``[B, M, D]`` token buffer), the models use them instead of ``torch.stack`` over ``nn.Linear`` outputs; the parameters are
the same ``nn.Linear`` modules either way, so state_dicts match the reference model's.
"""
from __future__ import annotations

import torch
import torch.nn as nn


def _default_fusion():
    import aecf
    return aecf


class VisionLanguageModel(nn.Module):
    def __init__(self, img_dim: int = 2048, txt_dim: int = 768, hidden_dim: int = 512, num_classes: int = 1000,
                 mask_prob: float = 0.15, fusion=None, **pool_kwargs):
        super().__init__()
        fusion = fusion or _default_fusion()
        self.img_proj = nn.Linear(img_dim, hidden_dim)
        self.txt_proj = nn.Linear(txt_dim, hidden_dim)
        self.fusion_query, self.fusion_pool = fusion.create_fusion_pool(
            embed_dim=hidden_dim, num_modalities=2, mask_prob=mask_prob, **pool_kwargs)
        self.classifier = nn.Linear(hidden_dim, num_classes)
        self._project_tokens = getattr(fusion, "project_tokens", None)
        self._linear = getattr(fusion, "linear", None)

    def forward(self, image_feats: torch.Tensor, text_feats: torch.Tensor, return_info: bool = False):
        if self._project_tokens is not None:
            tokens = self._project_tokens([(image_feats, self.img_proj), (text_feats, self.txt_proj)])   # [B, 2, hidden], no stack
        else:
            tokens = torch.stack([self.img_proj(image_feats), self.txt_proj(text_feats)], dim=1)     # reference README.md:186
        head = (lambda t: self._linear(t, self.classifier)) if self._linear is not None else self.classifier
        query = self.fusion_query.expand(tokens.size(0), -1, -1)
        if return_info:
            fused, info = self.fusion_pool(query, tokens, return_info=True)
            return head(fused.squeeze(1)), info
        return head(self.fusion_pool(query, tokens).squeeze(1))


class XrayFusionModel(nn.Module):
    def __init__(self, image_dim: int = 512, text_dim: int = 512, num_classes: int = 80, hidden_dim: int = 256,
                 num_heads: int = 4, encoder_dropout: float = 0.1, fusion=None, **pool_kwargs):
        super().__init__()
        fusion = fusion or _default_fusion()
        self.hidden_dim = hidden_dim
        self.image_encoder = nn.Sequential(nn.Linear(image_dim, hidden_dim), nn.ReLU(), nn.Dropout(encoder_dropout))
        self.text_encoder = nn.Sequential(nn.Linear(text_dim, hidden_dim), nn.ReLU(), nn.Dropout(encoder_dropout))
        self.curriculum_masking = fusion.CurriculumMasking(base_mask_prob=0.15)
        self.attention_pool = fusion.MultimodalAttentionPool(embed_dim=hidden_dim, num_heads=num_heads,
                                                             curriculum_masking=None, batch_first=True, **pool_kwargs)
        self.fusion_query = nn.Parameter(torch.randn(1, 1, hidden_dim) * 0.02)
        self.image_proj = nn.Linear(hidden_dim, hidden_dim * 2)
        self.text_proj = nn.Linear(hidden_dim, hidden_dim * 2)
        self.fusion_proj = nn.Linear(hidden_dim, hidden_dim * 2)
        self.classifier = nn.Sequential(nn.Linear(hidden_dim * 2, hidden_dim), nn.ReLU(), nn.Dropout(encoder_dropout),
                                        nn.Linear(hidden_dim, num_classes))
        import inspect
        # aecf_b200's pool can take the both-present rows as an index list and pool them IN PLACE (no gather before, no
        # scatter after); the reference's API (and the oracle stand-in of the parity tests) cannot
        self._in_place = "sample_index" in inspect.signature(self.attention_pool.forward).parameters

    def toggle_curriculum(self, enabled: bool) -> None:
        """The reference swaps the pool's masking module at epoch 40 (xrays/train_xrays_example.py:179-187)."""
        self.attention_pool.curriculum_masking = self.curriculum_masking if enabled else None

    def forward(self, image_features: torch.Tensor, text_features: torch.Tensor, return_info: bool = False):
        info = {}
        img = self.image_encoder(image_features)
        txt = self.text_encoder(text_features)
        img_present = image_features.norm(dim=1) > 1e-6
        txt_present = text_features.norm(dim=1) > 1e-6
        fused = torch.zeros(image_features.size(0), self.hidden_dim * 2, device=image_features.device, dtype=img.dtype)

        both_mask = img_present & txt_present
        both = torch.where(both_mask)[0]
        if both.numel() and self._in_place:
            # the pool reads the listed rows of the full token buffer and writes their outputs to the same rows of a full-size
            # output: what reference xrays/train_xrays_example.py:212-222 does with img[both] / txt[both] / fused[both] = ...
            tokens = torch.stack([img, txt], dim=1)
            query = self.fusion_query.expand(tokens.size(0), -1, -1)
            pooled, pool_info = self.attention_pool(query=query, key=tokens, value=tokens, return_info=True, sample_index=both)
            fused = torch.where(both_mask.unsqueeze(1), self.fusion_proj(pooled.squeeze(1)), fused)
            if return_info:
                info.update(pool_info)
        elif both.numel():
            tokens = torch.stack([img[both], txt[both]], dim=1)                  # the variable-size both-present subset
            query = self.fusion_query.expand(both.numel(), -1, -1)
            pooled, pool_info = self.attention_pool(query=query, key=tokens, value=tokens, return_info=True)
            fused[both] = self.fusion_proj(pooled.squeeze(1))
            if return_info:
                info.update(pool_info)
        only_img = torch.where(img_present & ~txt_present)[0]
        if only_img.numel():
            fused[only_img] = self.image_proj(img[only_img])
        only_txt = torch.where(~img_present & txt_present)[0]
        if only_txt.numel():
            fused[only_txt] = self.text_proj(txt[only_txt])
        logits = self.classifier(fused)
        return (logits, info) if return_info else logits
