"""TEST INFRASTRUCTURE -- stand-in for python-Levenshtein (oracle only, never shipped).

Restates `Levenshtein.distance(a, b)` (call site /root/reference search.py:14,190): classic
unit-cost insert/delete/substitute edit distance over Unicode code points [recalled]."""


def distance(a, b):
    if len(a) < len(b):
        a, b = b, a
    if not b:
        return len(a)
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            best = prev[j - 1] + (ca != cb)
            if prev[j] + 1 < best:
                best = prev[j] + 1
            if cur[j - 1] + 1 < best:
                best = cur[j - 1] + 1
            cur.append(best)
        prev = cur
    return prev[-1]
