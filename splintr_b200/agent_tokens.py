"""Agent-token id constants (pure data), one frozen namespace per vocabulary.

Same names and values as the reference's generated pyclasses
(/root/reference/src/python/agent_tokens_generated.rs); derived here from the special-token
tables in presets.py instead of being spelled out.
"""
from __future__ import annotations

from . import presets as _p


def _const_name(tok: str) -> str:
    name = tok[2:-2]                       # "<|/think|>" -> "/think"
    return (name[1:] + "_END").upper() if name.startswith("/") else name.upper()


class _Frozen(type):
    def __setattr__(cls, k, v):
        raise AttributeError("agent token constants are read-only")


def _make(cls_name: str, base: int, extra: dict, skip_multimodal: bool = False):
    ns = {_const_name(t): i for t, i in _p._agent_tokens(base, skip_multimodal).items()}
    ns.update(extra)
    return _Frozen(cls_name, (), ns)


CL100K_AGENT_TOKENS = _make("CL100K_AGENT_TOKENS", 100277, {})
O200K_AGENT_TOKENS = _make("O200K_AGENT_TOKENS", 200019, {})
LLAMA3_AGENT_TOKENS = _make("LLAMA3_AGENT_TOKENS", 128300, {
    "BEGIN_OF_TEXT": 128000, "END_OF_TEXT": 128001, "FINETUNE_RIGHT_PAD_ID": 128004, "STEP_ID": 128005,
    "START_HEADER_ID": 128006, "END_HEADER_ID": 128007, "EOM_ID": 128008, "EOT_ID": 128009,
    "PYTHON_TAG": 128010, "IMAGE": 128256, "IMAGE_END": 128257, "AUDIO": 128258, "AUDIO_END": 128259,
    "VIDEO": 128260, "VIDEO_END": 128261}, skip_multimodal=True)
DEEPSEEK_V3_AGENT_TOKENS = _make("DEEPSEEK_V3_AGENT_TOKENS", 128900, {
    "BEGIN_OF_SENTENCE": 0, "END_OF_SENTENCE": 1, "PAD_NATIVE": 2, "THINK_NATIVE": 128798,
    "THINK_END_NATIVE": 128799, "FIM_HOLE": 128800, "FIM_BEGIN": 128801, "FIM_END": 128802,
    "USER_NATIVE": 128803, "ASSISTANT_NATIVE": 128804, "EOT": 128805,
    "TOOL_CALLS_BEGIN": 128806, "TOOL_CALLS_END": 128807, "TOOL_CALL_BEGIN": 128808,
    "TOOL_CALL_END": 128809, "TOOL_OUTPUTS_BEGIN": 128810, "TOOL_OUTPUTS_END": 128811,
    "TOOL_OUTPUT_BEGIN": 128812, "TOOL_OUTPUT_END": 128813, "TOOL_SEP": 128814})
MISTRAL_V1_AGENT_TOKENS = _make("MISTRAL_V1_AGENT_TOKENS", 32000, {"UNK": 0, "BOS": 1, "EOS": 2})
MISTRAL_V2_AGENT_TOKENS = _make("MISTRAL_V2_AGENT_TOKENS", 32768, {
    "UNK": 0, "BOS": 1, "EOS": 2, "INST": 3, "INST_END": 4, "TOOL_CALLS": 5, "AVAILABLE_TOOLS": 6,
    "AVAILABLE_TOOLS_END": 7, "TOOL_RESULTS": 8, "TOOL_RESULTS_END": 9})
MISTRAL_V3_AGENT_TOKENS = _make("MISTRAL_V3_AGENT_TOKENS", 131072, {
    "UNK": 0, "BOS": 1, "EOS": 2, "INST": 3, "INST_END": 4, "AVAILABLE_TOOLS": 5,
    "AVAILABLE_TOOLS_END": 6, "TOOL_RESULTS": 7, "TOOL_RESULTS_END": 8, "TOOL_CALLS": 9})
