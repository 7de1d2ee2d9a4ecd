"""aecf_b200 -- B200-native (sm_100a) implementation of the AECF fusion hot path.

Same public names as the reference package ``aecf`` (reference ``aecf/__init__.py:9-21``):
``CurriculumMasking``, ``MultimodalAttentionPool``, ``multimodal_attention_pool``,
``create_fusion_pool``.  Compute goes through the C-ABI library ``csrc/libaecf_b200.so``
(``include/aecf_b200.h``); there is no CPU or eager-PyTorch fallback.
"""
from .layers import (CurriculumMasking, MultimodalAttentionPool, create_fusion_pool, get_rng_state,
                     multimodal_attention_pool, set_rng_state)
from .projections import linear, project_tokens
from . import graphs

__version__ = "0.1.0"
__all__ = ["CurriculumMasking", "MultimodalAttentionPool", "multimodal_attention_pool", "create_fusion_pool"]
