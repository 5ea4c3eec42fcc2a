"""splintr_b200 -- B200-native batch BPE encoder behind splintr's Python API.

Mirrors the exports of the reference package (/root/reference/python/splintr/__init__.py:110-141).
Importing this package does not touch the GPU; constructing a Tokenizer loads the CUDA
library (splintr_b200/libsplintr_b200.so) and fails loudly if it or a device is missing.
"""
from .presets import (CL100K_BASE_PATTERN, O200K_BASE_PATTERN, LLAMA3_PATTERN,
                      MISTRAL_V3_PATTERN, SENTENCEPIECE_PATTERN)
from .tokenizer import Tokenizer
from .streaming import StreamingDecoder, ByteLevelStreamingDecoder
from .agent_tokens import (CL100K_AGENT_TOKENS, O200K_AGENT_TOKENS, LLAMA3_AGENT_TOKENS,
                           DEEPSEEK_V3_AGENT_TOKENS, MISTRAL_V1_AGENT_TOKENS,
                           MISTRAL_V2_AGENT_TOKENS, MISTRAL_V3_AGENT_TOKENS)

__all__ = [
    "Tokenizer", "StreamingDecoder", "ByteLevelStreamingDecoder",
    "CL100K_BASE_PATTERN", "O200K_BASE_PATTERN", "LLAMA3_PATTERN",
    "CL100K_AGENT_TOKENS", "O200K_AGENT_TOKENS", "LLAMA3_AGENT_TOKENS", "DEEPSEEK_V3_AGENT_TOKENS",
    "MISTRAL_V1_AGENT_TOKENS", "MISTRAL_V2_AGENT_TOKENS", "MISTRAL_V3_AGENT_TOKENS",
]
__version__ = "0.1.0"
