"""Generate tests/golden/ by running the UNMODIFIED reference search.py over oracle/shims.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

For every case it spawns `python -c "import search; search.analyze(args, ...)"` with
PYTHONPATH = oracle/shims : /root/reference, PYTHONHASHSEED=0 (the OOV rule of
search.py:79-83 uses the builtin str hash) and the nearpy stand-in switched to
  * exhaustive candidates                      -> golden_exhaustive.csv
  * seeded random hyperplanes (seed 7)         -> golden_lsh_seed7.csv
and copies the aggregate match-6gram-YYYYMMDD.csv plus the per-cluster batch files
(chunk_size=16 -> 3 clusters; the reference's own pool.map chunksize = chunk_size // 16
must be >= 1, search.py:385).
The inputs (lexicon, script, fanworks) are committed next to the outputs so the GPU box,
which has no /root/reference, can replay them.
"""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from fandom_search_b200 import synth  # noqa: E402

REFERENCE = "/root/reference"
GOLDEN = os.path.join(ROOT, "tests", "golden")

DRIVER = r"""
import argparse, sys, os
import search
args = argparse.Namespace(fan_works=sys.argv[1], script=sys.argv[2], skip_works=int(sys.argv[3]),
                          num_works=int(sys.argv[4]))
search.analyze(args, chunk_size=int(sys.argv[5]))
"""


def build_inputs():
    os.makedirs(GOLDEN, exist_ok=True)
    lex = synth.SynthLexicon(vocab=300, dim=300, cluster=6, oov_frac=0.05, cased_frac=0.2, seed=11)
    lex_path = lex.save(os.path.join(GOLDEN, "lexicon.npz"))
    rng = np.random.default_rng(5)
    script_ids = lex.sample_words(rng, 520)
    # an 11-fold repeated 6-gram (top-10 truncation) and an OOV-heavy stretch
    rep = script_ids[40:46].copy()
    for k in range(11):
        script_ids[100 + 12 * k:106 + 12 * k] = rep
    oov_words = np.nonzero(lex.is_oov)[0]
    script_ids[300:304] = oov_words[:4]
    script_path = os.path.join(GOLDEN, "script.txt")
    synth.write_markup_script(lex, script_ids, script_path, seed=3)
    with open(script_path, "a", encoding="utf-8") as f:     # scene-number fallback, search.py:309-317
        f.write("SCENE_NUMBER<<A>>\nCHARACTER_NAME<<LEIA>>\nLINE<<%s>>\n" %
                " ".join(lex.words[script_ids[10:22]]))
    fan_dir = os.path.join(GOLDEN, "fanworks")
    if os.path.isdir(fan_dir):
        shutil.rmtree(fan_dir)
    os.makedirs(fan_dir)
    n = 0
    for k in range(37):
        ids, _ = synth.make_fanwork_tokens(lex, script_ids, k, mean_len=180, sd_len=60, min_len=40,
                                           max_len=400, spans_mean=2.0, seed_base=900)
        if k == 3:
            ids[20:32] = script_ids[100:112]        # hits the 11x repeated 6-gram
        if k == 4:
            ids[5:15] = script_ids[298:308]         # quote through the OOV stretch
        crng = np.random.default_rng(77 + k) if k in (2, 6, 20) else None
        text = synth.fanwork_text(lex, ids, crng, 0.15 if crng is not None else 0.0)
        with open(os.path.join(fan_dir, "%07d.txt" % k), "w", encoding="utf-8") as f:
            f.write(text)
        n += 1
    with open(os.path.join(fan_dir, "%07d.txt" % 37), "w", encoding="utf-8") as f:   # < 6 tokens
        f.write(" ".join(lex.words[script_ids[0:5]]))
    with open(os.path.join(fan_dir, "%07d.txt" % 38), "w", encoding="utf-8") as f:   # empty work
        f.write("")
    with open(os.path.join(fan_dir, "%07d.txt" % 39), "w", encoding="utf-8") as f:   # exactly 6 tokens
        f.write(" ".join(lex.words[script_ids[200:206]]))
    return lex_path, script_path, fan_dir


def run_reference(lex_path, script_path, fan_dir, env_extra, out_name, chunk_size=16):
    env = dict(os.environ)
    env.update({"PYTHONPATH": os.path.join(HERE, "shims") + os.pathsep + REFERENCE,
                "PYTHONHASHSEED": "0", "FANDOM_SEARCH_LEXICON": lex_path})
    env.update(env_extra)
    with tempfile.TemporaryDirectory() as tmp:
        # relative paths, so FAN_WORK_FILENAME in the CSV is machine independent
        subprocess.check_call([sys.executable, "-c", DRIVER, os.path.relpath(fan_dir, GOLDEN),
                               os.path.relpath(script_path, GOLDEN), "-1", "-1", str(chunk_size)],
                              env=env, cwd=GOLDEN)
        aggs = sorted(glob.glob(os.path.join(GOLDEN, "match-6gram-2*.csv")))
        assert len(aggs) == 1, aggs
        shutil.move(aggs[0], os.path.join(GOLDEN, out_name + ".csv"))
        for b in sorted(glob.glob(os.path.join(GOLDEN, "match-6gram-batch-*.csv"))):
            i = os.path.basename(b)[len("match-6gram-batch-"):-4]
            shutil.move(b, os.path.join(GOLDEN, "%s.batch%s.csv" % (out_name, i)))


def main():
    if not os.path.isdir(REFERENCE):
        raise SystemExit("make_golden.py needs /root/reference (build container only)")
    lex_path, script_path, fan_dir = build_inputs()
    run_reference(lex_path, script_path, fan_dir, {"NEARPY_SHIM_EXHAUSTIVE": "1"}, "golden_exhaustive")
    run_reference(lex_path, script_path, fan_dir, {"NEARPY_SHIM_SEED": "7"}, "golden_lsh_seed7")
    # the reference shuffles os.listdir() order (search.py:349-355); record it so a test on
    # another file system can replay the same cluster membership
    with open(os.path.join(GOLDEN, "listing.txt"), "w") as f:
        f.write("\n".join(os.listdir(fan_dir)) + "\n")
    for f in sorted(os.listdir(GOLDEN)):
        p = os.path.join(GOLDEN, f)
        if os.path.isfile(p):
            print("%8d  %s" % (os.path.getsize(p), f))


if __name__ == "__main__":
    main()
