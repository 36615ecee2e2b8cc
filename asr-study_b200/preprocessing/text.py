"""Label front end: character vocabulary <-> int ids, blank = last index (preprocessing/text.py:10-146 of the
reference).  Host-side glue of the batch contract (datasets/dataset_generator.py:237-251); the label ids it
produces are what K6 / K10 consume.

Same surface: ``CharParser(mode)`` with the modes 'space'|'accents'|'punctuation'|'digits'|'sensitive' (or their
one-letter forms s, a, p, d, S, joined by '|', or 'all'), ``map`` / ``__call__`` (with the sanitiser), ``imap``,
``is_valid``, and the two module-level instances ``simple_char_parser`` (28 classes, blank 27) and
``complex_char_parser``.  ``unidecode`` (text.py:5) is not installable here: accents are folded with the Unicode
canonical decomposition (NFKD, combining marks dropped), which is what unidecode yields for the Latin accents of the
reference's pt-br corpora; characters NFKD cannot fold to ASCII are dropped instead of transliterated.
"""
from __future__ import annotations

import string
import unicodedata

import numpy as np

PUNCTUATIONS = "'\"-,.!?:;"
ACCENTS = u"ãõçâêôáíóúàüóé"
_MODES = {"sensitive": "S", "space": "s", "accents": "a", "punctuation": "p", "digits": "d"}


def _fold_accents(text):
    out = unicodedata.normalize("NFKD", text)
    return "".join(c for c in out if not unicodedata.combining(c) and ord(c) < 128)


class BaseParser(object):
    def __call__(self, _input):
        return self.map(_input)

    def map(self, _input):
        pass

    def imap(self, _input):
        pass

    def is_valid(self, _input):
        pass


class CharParser(BaseParser):
    def __init__(self, mode="space"):
        if mode == "all":
            self.mode = list(_MODES.values())
        else:
            self.mode = []
            for m in (mode or "").split("|"):
                if not m:
                    continue
                if m in _MODES:
                    self.mode.append(_MODES[m])
                elif m in _MODES.values():
                    self.mode.append(m)
                else:
                    raise ValueError("Unknown mode %s" % m)
        self._vocab, self._inv_vocab = self._gen_vocab()

    # vocabulary order (text.py:113-143): a-z, accents, upper case, space, punctuation, digits, then the blank
    def _gen_vocab(self):
        vocab = {chr(ord("a") + i): i for i in range(26)}
        if "a" in self.mode:
            for ch in ACCENTS:
                vocab[ch] = len(vocab)
        if "S" in self.mode:
            for ch in list(vocab.keys()):
                vocab[ch.upper()] = len(vocab)
        if "s" in self.mode:
            vocab[" "] = len(vocab)
        if "p" in self.mode:
            for ch in PUNCTUATIONS:
                vocab[ch] = len(vocab)
        if "d" in self.mode:
            for d in range(10):
                vocab[str(d)] = len(vocab)
        inv = {v: k for k, v in vocab.items()}
        inv[len(inv)] = "<b>"
        return vocab, inv

    @property
    def num_classes(self):
        return len(self._inv_vocab)

    @property
    def blank(self):
        return len(self._inv_vocab) - 1

    def _sanitize(self, text):
        """text.py:83-104, in its order: collapse white space, drop digits, fold accents, '-' and "'" to spaces and
        the rest of string.punctuation out, drop spaces, lower-case — each unless the matching mode is on."""
        text = " ".join(text.split())
        if "d" not in self.mode:
            text = "".join(c for c in text if not c.isdigit())
        if "a" not in self.mode:
            text = _fold_accents(text)
        if "p" not in self.mode:
            text = text.translate(str.maketrans("-'", "  ")).translate(str.maketrans("", "", string.punctuation))
        if "s" not in self.mode:
            text = text.replace(" ", "")
        if "S" not in self.mode:
            text = text.lower()
        return text

    def map(self, txt, sanitize=True):
        if sanitize:
            txt = self._sanitize(txt)
        return np.array([self._vocab[c] for c in txt], dtype="int32")

    def imap(self, labels):
        return "".join(self._inv_vocab[int(l)] for l in labels if int(l) >= 0)

    def is_valid(self, text):
        try:
            self.map(text, sanitize=False)
            return True
        except KeyError:
            return False


simple_char_parser = CharParser()
complex_char_parser = CharParser(mode="s|p|a|d")
