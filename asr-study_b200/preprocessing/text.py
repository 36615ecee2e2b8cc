"""Label front end: character vocabulary a-z (+ space ...) <-> int ids, blank = last index
(preprocessing/text.py:68-146 of the reference).  Only what the hot path consumes is kept:
vocab size 28 / blank 27 for the default parser; the unidecode-based sanitiser is out of scope."""


class CharParser(object):
    def __init__(self, mode="s"):
        self.mode = mode or ""
        vocab = {chr(ord("a") + i): i for i in range(26)}
        if "s" in self.mode:
            vocab[" "] = len(vocab)
        if "d" in self.mode:
            for d in range(10):
                vocab[str(d)] = len(vocab)
        self._vocab = vocab
        self._inv = {v: k for k, v in vocab.items()}
        self._inv[len(self._inv)] = "<b>"            # blank label is the last index

    @property
    def num_classes(self):
        return len(self._inv)

    @property
    def blank(self):
        return len(self._inv) - 1

    def map(self, txt, sanitize=True):
        if sanitize:
            txt = txt.lower()
        return [self._vocab[c] for c in txt]

    __call__ = map

    def imap(self, labels):
        return "".join(self._inv[int(l)] for l in labels if int(l) >= 0)

    def is_valid(self, txt):
        try:
            self.map(txt, sanitize=False)
            return True
        except KeyError:
            return False


simple_char_parser = CharParser()
