"""Doc / Span / Token / Vocab of the spaCy stand-in (oracle only)."""
import numpy

_M = 0xc6a4a7935bd1e995
_MASK = 0xFFFFFFFFFFFFFFFF


def murmurhash64a(data, seed=1):
    """MurmurHash64A of `data` (bytes); spaCy string ids use seed 1 [recalled]."""
    n = len(data)
    h = (seed ^ (n * _M)) & _MASK
    nblocks = n // 8
    for i in range(nblocks):
        k = int.from_bytes(data[8 * i:8 * i + 8], 'little')
        k = (k * _M) & _MASK
        k ^= k >> 47
        k = (k * _M) & _MASK
        h ^= k
        h = (h * _M) & _MASK
    tail = data[8 * nblocks:]
    if tail:
        h ^= int.from_bytes(tail, 'little')
        h = (h * _M) & _MASK
    h ^= h >> 47
    h = (h * _M) & _MASK
    h ^= h >> 47
    return h


class Vocab(object):
    def __init__(self, keys, rows, table):
        self.table = numpy.asarray(table, dtype=numpy.float32)
        self.key_to_row = {str(k): int(r) for k, r in zip(keys, rows)}
        self.vectors_length = self.table.shape[1]
        self._ids = {}

    @classmethod
    def from_npz(cls, path):
        with numpy.load(path, allow_pickle=False) as z:
            return cls(z['keys'], z['rows'], z['table'])

    def string_id(self, text):
        h = self._ids.get(text)
        if h is None:
            h = murmurhash64a(text.encode('utf-8'), 1)
            self._ids[text] = h
        return h


class Token(object):
    __slots__ = ('vocab', 'text', 'i')

    def __init__(self, vocab, text, i):
        self.vocab = vocab
        self.text = text
        self.i = i

    @property
    def is_space(self):
        return self.text.isspace()

    @property
    def has_vector(self):
        return self.text in self.vocab.key_to_row

    @property
    def vector(self):
        row = self.vocab.key_to_row.get(self.text)
        if row is None:
            return numpy.zeros((self.vocab.vectors_length,), dtype=numpy.float32)
        return self.vocab.table[row]

    @property
    def orth_(self):
        return self.text

    @property
    def orth(self):
        return self.vocab.string_id(self.text)

    @property
    def lower_(self):
        return self.text.lower()

    @property
    def lower(self):
        return self.vocab.string_id(self.text.lower())

    def __str__(self):
        return self.text

    def __repr__(self):
        return self.text

    def __len__(self):
        return len(self.text)


class Span(object):
    def __init__(self, doc, start, end):
        self.doc = doc
        self.start = start
        self.end = end

    @property
    def text(self):
        return ' '.join(self.doc.words[self.start:self.end])

    def __str__(self):
        return self.text

    def __repr__(self):
        return self.text

    def __len__(self):
        return self.end - self.start

    def __iter__(self):
        for i in range(self.start, self.end):
            yield self.doc[i]


class Doc(object):
    def __init__(self, vocab, words=None, spaces=None):
        self.vocab = vocab
        self.words = list(words) if words is not None else []
        self._tokens = [Token(vocab, w, i) for i, w in enumerate(self.words)]

    def __len__(self):
        return len(self._tokens)

    def __getitem__(self, key):
        if isinstance(key, slice):
            start, end, step = key.indices(len(self._tokens))
            if step != 1:
                raise ValueError("Stepped slices not supported in Span objects.")
            return Span(self, start, max(start, end))
        return self._tokens[key]

    def __iter__(self):
        return iter(self._tokens)

    @property
    def text(self):
        return ' '.join(self.words)

    def __str__(self):
        return self.text
