"""Pretrained presets: name -> (vocab file, split pattern, special tokens, mode).

Host-side data mirroring the reference's preset dispatch
(/root/reference/src/python/bindings.rs:101-166, src/core/pretrained.rs:131-169,
special-token sets pretrained.rs:238-547, patterns src/core/tokenizer.rs:39-64).
"""
from __future__ import annotations

import lzma
import os
from typing import Dict, NamedTuple, Optional

# --- split patterns (tokenizer.rs:39,42,45,56,64), verbatim strings -----------------
CL100K_BASE_PATTERN = r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+"
O200K_BASE_PATTERN = r"[^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]*[\p{Ll}\p{Lm}\p{Lo}\p{M}]+(?i:'s|'t|'re|'ve|'m|'ll|'d)?|[^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]+[\p{Ll}\p{Lm}\p{Lo}\p{M}]*(?i:'s|'t|'re|'ve|'m|'ll|'d)?|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+"
LLAMA3_PATTERN = O200K_BASE_PATTERN
SENTENCEPIECE_PATTERN = r"[^\s]+|\s+"
MISTRAL_V3_PATTERN = r"[^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]*[\p{Ll}\p{Lm}\p{Lo}\p{M}]+|[^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]+[\p{Ll}\p{Lm}\p{Lo}\p{M}]*|\p{N}| ?[^\s\p{L}\p{N}]+[\r\n/]*|\s*[\r\n]+|\s+(?!\S)|\s+"

# C-ABI pattern ids (include/splintr_b200.h: SPL_PATTERN_*)
PATTERN_IDS = {
    CL100K_BASE_PATTERN: 0,
    O200K_BASE_PATTERN: 1,
    MISTRAL_V3_PATTERN: 2,
}
SPL_PATTERN_SENTENCEPIECE = 3      # SENTENCEPIECE_PATTERN, only together with SentencePiece mode

# --- agent tokens (pretrained.rs:402-476): 54 names, ids base+0 .. base+53 -----------
_AGENT_NAMES = [
    "system", "user", "assistant", "im_start", "im_end",
    "think", "/think",
    "plan", "/plan", "step", "/step", "act", "/act", "observe", "/observe",
    "function", "/function", "result", "/result", "error", "/error",
    "code", "/code", "output", "/output", "lang", "/lang",
    "context", "/context", "quote", "/quote", "cite", "/cite", "source", "/source",
    "memory", "/memory", "recall", "/recall",
    "pad", "stop", "sep",
    "image", "/image", "audio", "/audio", "video", "/video",
    "title", "/title", "section", "/section", "summary", "/summary",
]
assert len(_AGENT_NAMES) == 54


def _agent_tokens(base: int, skip_multimodal: bool = False) -> Dict[str, int]:
    out = {}
    for i, name in enumerate(_AGENT_NAMES):
        if skip_multimodal and 42 <= i <= 47:       # pretrained.rs:538 (llama3 keeps 128256+)
            continue
        out[f"<|{name}|>"] = base + i
    return out


def cl100k_base_special_tokens() -> Dict[str, int]:      # pretrained.rs:238-251
    s = {"<|endoftext|>": 100257, "<|fim_prefix|>": 100258, "<|fim_middle|>": 100259,
         "<|fim_suffix|>": 100260, "<|endofprompt|>": 100276}
    s.update(_agent_tokens(100277))
    return s


def o200k_base_special_tokens() -> Dict[str, int]:       # pretrained.rs:254-264
    s = {"<|endoftext|>": 199999, "<|endofprompt|>": 200018}
    s.update(_agent_tokens(200019))
    return s


def llama3_special_tokens() -> Dict[str, int]:           # pretrained.rs:267-295
    s = {"<|begin_of_text|>": 128000, "<|end_of_text|>": 128001,
         "<|reserved_special_token_0|>": 128002, "<|reserved_special_token_1|>": 128003,
         "<|finetune_right_pad_id|>": 128004, "<|step_id|>": 128005,
         "<|start_header_id|>": 128006, "<|end_header_id|>": 128007,
         "<|eom_id|>": 128008, "<|eot_id|>": 128009, "<|python_tag|>": 128010,
         "<|image|>": 128256, "<|/image|>": 128257, "<|audio|>": 128258,
         "<|/audio|>": 128259, "<|video|>": 128260, "<|/video|>": 128261}
    s.update(_agent_tokens(128300, skip_multimodal=True))
    return s


def deepseek_v3_special_tokens() -> Dict[str, int]:      # pretrained.rs:298-335
    bar, us = "｜", "▁"
    s = {f"<{bar}begin{us}of{us}sentence{bar}>": 0, f"<{bar}end{us}of{us}sentence{bar}>": 1,
         f"<{bar}{us}pad{us}{bar}>": 2,
         "<think>": 128798, "</think>": 128799,
         f"<{bar}fim{us}hole{bar}>": 128800, f"<{bar}fim{us}begin{bar}>": 128801,
         f"<{bar}fim{us}end{bar}>": 128802,
         f"<{bar}User{bar}>": 128803, f"<{bar}Assistant{bar}>": 128804, "<|EOT|>": 128805,
         f"<{bar}tool{us}calls{us}begin{bar}>": 128806, f"<{bar}tool{us}calls{us}end{bar}>": 128807,
         f"<{bar}tool{us}call{us}begin{bar}>": 128808, f"<{bar}tool{us}call{us}end{bar}>": 128809,
         f"<{bar}tool{us}outputs{us}begin{bar}>": 128810, f"<{bar}tool{us}outputs{us}end{bar}>": 128811,
         f"<{bar}tool{us}output{us}begin{bar}>": 128812, f"<{bar}tool{us}output{us}end{bar}>": 128813,
         f"<{bar}tool{us}sep{bar}>": 128814}
    s.update(_agent_tokens(128900))
    return s


def mistral_v1_special_tokens() -> Dict[str, int]:       # pretrained.rs:338-350
    s = {"<unk>": 0, "<s>": 1, "</s>": 2}
    s.update(_agent_tokens(32000))
    return s


def mistral_v2_special_tokens() -> Dict[str, int]:       # pretrained.rs:353-370
    s = {"[INST]": 3, "[/INST]": 4, "[TOOL_CALLS]": 5, "[AVAILABLE_TOOLS]": 6,
         "[/AVAILABLE_TOOLS]": 7, "[TOOL_RESULTS]": 8, "[/TOOL_RESULTS]": 9}
    s.update(_agent_tokens(32768))
    return s


def mistral_v3_special_tokens() -> Dict[str, int]:       # pretrained.rs:373-394
    s = {"<unk>": 0, "<s>": 1, "</s>": 2, "[INST]": 3, "[/INST]": 4,
         "[AVAILABLE_TOOLS]": 5, "[/AVAILABLE_TOOLS]": 6, "[TOOL_RESULTS]": 7,
         "[/TOOL_RESULTS]": 8, "[TOOL_CALLS]": 9}
    s.update(_agent_tokens(131072))
    return s


class Preset(NamedTuple):
    vocab_file: str
    pattern: str
    special_tokens: Dict[str, int]
    byte_level: bool
    sentencepiece: bool


def _presets() -> Dict[str, Preset]:
    """Name dispatch of bindings.rs:101-166 (the Python entry point's own table)."""
    cl = Preset("cl100k_base.tiktoken", CL100K_BASE_PATTERN, cl100k_base_special_tokens(), False, False)
    o2 = Preset("o200k_base.tiktoken", O200K_BASE_PATTERN, o200k_base_special_tokens(), False, False)
    l3 = Preset("llama3.tiktoken", LLAMA3_PATTERN, llama3_special_tokens(), False, False)
    ds = Preset("deepseek_v3.tiktoken", LLAMA3_PATTERN, deepseek_v3_special_tokens(), True, False)
    m1 = Preset("mistral.tiktoken", SENTENCEPIECE_PATTERN, mistral_v1_special_tokens(), False, True)
    m2 = Preset("mistral_v2.tiktoken", SENTENCEPIECE_PATTERN, mistral_v2_special_tokens(), False, True)
    m3 = Preset("mistral_v3_tekken.tiktoken", MISTRAL_V3_PATTERN, mistral_v3_special_tokens(), True, False)
    return {
        "cl100k_base": cl, "o200k_base": o2,
        "llama3": l3, "llama3.1": l3, "llama3.2": l3, "llama3.3": l3,
        "deepseek_v3": ds, "deepseek-v3": ds,
        "mistral": m1, "mistral_v1": m1, "mistral_v2": m2, "mistral_v3": m3,
    }


PRESETS = _presets()

_VOCAB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vocabs")


def load_vocab_bytes(vocab_file: str) -> bytes:
    """Bundled vocab data (the reference's python/splintr/vocabs/*.tiktoken, stored
    xz-compressed; content is byte-identical after decompression)."""
    path = os.path.join(_VOCAB_DIR, vocab_file)
    if os.path.exists(path):
        with open(path, "rb") as f:
            return f.read()
    with lzma.open(path + ".xz", "rb") as f:
        return f.read()


def get_preset(name: str) -> Optional[Preset]:
    return PRESETS.get(name)
