"""TEST INFRASTRUCTURE -- stand-in for `spacy` + the `en_core_web_md` model (oracle only).

spaCy and its model are un-pinned dependencies of the reference
(/root/reference requirements.txt:3, search.py:43-44) and are not installed here, so this
module RESTATES the small API surface search.py uses (call sites search.py:43-44,52,62,
74-75,79,123,151,166,189,194-195,322-327) [recalled, unverifiable here => parity unpinned]:

  * tokeniser-only pipeline; here: split on ASCII whitespace, which is exact for the
    synthetic corpora (single-space separated, no punctuation) and never yields
    `is_space` tokens;
  * `Token.vector` / `Token.has_vector` keyed on the verbatim text (ORTH, case-sensitive),
    float32 rows from a lexicon file named by $FANDOM_SEARCH_LEXICON (.npz with arrays
    `keys` [n] str, `rows` [n] int32, `table` [R, d] float32);
  * `Token.orth` / `Token.lower` = MurmurHash64A(utf8, seed=1) of the text / lower-cased
    text (known answer: "coffee" -> 3197928453018144401);
  * `Doc(vocab, words=...)`: every word followed by one space, so `str(span)` is the words
    joined by single spaces; `repr(token)` is the token text, so `str(list_of_tokens)` is
    "[a, b, c]".
"""
import os

import numpy

from . import tokens  # noqa: F401  (reference uses spacy.tokens.Doc with only `import spacy`)
from .tokens import Doc, Vocab, murmurhash64a  # noqa: F401

__version__ = "0.0-shim"

_WS = b' \t\n\x0b\x0c\r'


class Language(object):
    def __init__(self, vocab, disable=()):
        self.vocab = vocab
        self.disable = list(disable)

    def __call__(self, text):
        words = [w.decode('utf-8') for w in text.encode('utf-8').split()]
        return Doc(self.vocab, words=words)


_VOCABS = {}


def load(name, disable=(), **kwargs):
    path = os.environ.get('FANDOM_SEARCH_LEXICON')
    if not path:
        raise OSError("spacy shim: set FANDOM_SEARCH_LEXICON to a lexicon .npz "
                      "(stand-in for model %r)" % (name,))
    if path not in _VOCABS:
        _VOCABS[path] = Vocab.from_npz(path)
    return Language(_VOCABS[path], disable=disable)
