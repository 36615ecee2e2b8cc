"""Host logic of the batch contract (CPU): the label parser's sanitiser / modes (preprocessing/text.py of the reference)
and the rank sharding of DatasetIterator."""
import numpy as np
import pytest

from asr_study_b200.datasets.dataset_generator import DatasetGenerator, DatasetIterator
from asr_study_b200.preprocessing.text import CharParser, complex_char_parser, simple_char_parser


def test_simple_parser_sanitises_like_the_reference():
    p = simple_char_parser
    assert p.num_classes == 28 and p.blank == 27
    assert p("ab z").tolist() == [0, 1, 26, 25] and p("ab z").dtype == np.int32
    # text.py:83-104: white space collapsed, digits dropped, accents folded, '-' and "'" become spaces, the rest of
    # the punctuation is removed, lower-cased
    assert p.imap(p("  It's   a WELL-known  fact, não é?  42 ")) == "it s a well known fact nao e "
    assert p.imap([0, 1, -1, -1]) == "ab"
    assert p.is_valid("abc d") and not p.is_valid("Abc") and not p.is_valid("a,b")


def test_modes_by_name_letter_and_all():
    assert CharParser("space").mode == ["s"] and CharParser("s|digits").mode == ["s", "d"]
    # 'digits' / 'accents' must not switch the space label on (they contain the letter s)
    d = CharParser("digits")
    assert " " not in d._vocab and d.num_classes == 26 + 10 + 1 and d("a 1").tolist() == [0, 26 + 1]
    c = complex_char_parser
    assert c.mode == ["s", "p", "a", "d"]
    # order: a-z, accents, space, punctuation, digits, blank.  ACCENTS lists 'ó' twice (text.py:10): the second visit
    # re-assigns it to len(vocab), the id 'é' receives next — a collision the reference has and this keeps
    assert c._vocab["ã"] == 26 and c._vocab["ó"] == c._vocab["é"] == 38 and c._vocab[" "] == 39
    assert c._vocab["9"] == 26 + 13 + 1 + 9 + 10 - 1
    assert c.imap(c("Não, 3!")) == "não, 3!"
    S = CharParser("S|s")
    assert S("aB").tolist() == [0, S._vocab["B"]] and S._vocab["A"] == 26
    assert sorted(CharParser("all").mode) == sorted(["S", "s", "a", "p", "d"])
    with pytest.raises(ValueError):
        CharParser("nope")


def test_rank_sharding_partitions_each_epoch():
    n, B, W = 23, 4, 4
    inputs = [np.full((3, 2), i, np.float32) for i in range(n)]
    labels = ["a"] * n
    seen = []
    its = [DatasetIterator(inputs, labels, batch_size=B, shuffle=True, seed=5, label_parser=simple_char_parser,
                           rank=r, world_size=W) for r in range(W)]
    assert all(it.len == 6 for it in its)
    for it in its:
        got = []
        while len(got) < it.len:
            (x, lab, xl), _ = next(it)
            got += [int(v) for v in x[:, 0, 0]]
        assert len(got) == it.len
        seen.append(got)
    flat = sorted(v for g in seen for v in g)
    assert set(flat) == set(range(n)) and len(flat) == 24          # one wrapped utterance pads the tail
    # second epoch: a new shared permutation, still a partition
    second = []
    for it in its:
        got = []
        while len(got) < it.len:
            (x, lab, xl), _ = next(it)
            got += [int(v) for v in x[:, 0, 0]]
        second.append(got)
    assert set(v for g in second for v in g) == set(range(n)) and second != seen
    # world of one: the reference's behaviour (len = corpus size)
    one = DatasetGenerator(None, simple_char_parser, batch_size=B, shuffle=False).flow(inputs, labels, rank=0, world_size=1)
    assert one.len == n
