"""Pretty-print a tools/timeline.py dump: python tools/timeline_show.py FILE [first_tile n_tiles]"""
import collections
import sys

rows = collections.defaultdict(dict)
for l in open(sys.argv[1]):
    if l.startswith('#'):
        print(l.strip())
        continue
    f = l.split()
    rows[int(f[0])][int(f[1])] = [int(x) if x != '-' else None for x in f[2:]]
t0 = int(sys.argv[2]) if len(sys.argv) > 2 else 20
n = int(sys.argv[3]) if len(sys.argv) > 3 else 5
for t in range(t0, t0 + n):
    print("tile", t)
    print("  producer", [x for x in rows[0][t] if x is not None])
    print("  issuer  ", [x for x in rows[1][t] if x is not None])
    for w in (0, 1, 5, 10, 15):
        print("  epi%02d   " % w, [x for x in rows[2 + w][t] if x is not None])
