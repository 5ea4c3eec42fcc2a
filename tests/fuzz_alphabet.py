"""Adversarial alphabet for pre-tokenizer fuzzing (shared by CPU and GPU parity tests).

Every construct the split patterns distinguish is represented: CR/LF, plain and exotic
White_Space (NBSP, U+3000, U+2028, U+0085, U+1680), apostrophes and contraction suffixes
in mixed case (including U+017F which case-folds to 's'), Lu/Ll/Lt/Lm/Lo letters, marks,
digits of several scripts, punctuation, slash, emoji (4-byte), zero-width chars.
"""
import random

ALPHABET = list("aaabbcdeiostrvmlXYZABDST   \t\n\r\n\n''.,!?;:-_=#/(){}[]\"$%0123456789") + [
    " ", "　", " ", "", " ", " ",          # exotic whitespace
    "́", "̈", "ा",                                        # marks (Mn, Mn, Mc)
    "ǅ", "ʰ", "ª",                                        # Lt, Lm, Lo
    "ſ", "K", "ı", "İ",                              # long s, Kelvin, dotless i, dotted I
    "好", "世", "界", "你",                              # CJK (Lo)
    "é", "É", "ß", "Ω", "ω",                    # Latin-1 / Greek letters
    "\U0001f30d", "\U0001f600",                                          # emoji (So, 4 bytes)
    "٣", "²", "Ⅷ", "１",                              # Nd, No, Nl, fullwidth digit
    "​", "﻿", "’", "—", "，", "。",          # ZWSP, BOM, curly quote, em dash, CJK punct
    "\x00", "\x1c", "\x1f", "\x7f",                                      # control chars
    " the", " and", "'s", "'t", "'re", "'ve", "'m", "'ll", "'d", "'S", "'LL", "'Re", "'ſ",
    "\n\n", "  ", "\r\n", "<|", "|>",
]


def random_text(rng: random.Random, max_len: int = 40) -> str:
    return "".join(rng.choice(ALPHABET) for _ in range(rng.randint(1, max_len)))
