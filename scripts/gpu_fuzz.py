"""One-off differential fuzz on the GPU: random expressions x random inputs, CUDA path vs oracle.
usage: python scripts/gpu_fuzz.py SEED COUNT"""
import os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import trre_b200 as T
from oracle_util import oracle_run
from patgen import rand_expr, rand_input

rng = random.Random(int(sys.argv[1]))
count = int(sys.argv[2])
bad = checked = 0
kinds = {}
for it in range(count):
    pat = rand_expr(rng, rng.choice([1, 2, 2, 3]))
    alphabet = rng.choice(["abcd ", "ab", "abcdXY \t", "aaab"])
    nlines = rng.choice([0, 1, 3, 30, 400])
    data = rand_input(rng, nlines=nlines, alphabet=alphabet, maxlen=rng.choice([4, 12, 60]), final_newline=rng.random() < 0.85)
    if rng.random() < 0.08:
        data = data.replace(b"a", b"\0", 1)
    for flavor in (0, 1):
        rc, want, err = oracle_run(pat.encode(), data, flavor=flavor)
        if rc == 2:
            continue
        try:
            prog = T.compile(pat, flavor)
        except T.TrreError as e:
            if e.code != 2 and (e.code != rc or (rc == 1 and e.text != err)):
                bad += 1; print("COMPILE MISMATCH", flavor, pat, e.code, rc, flush=True)
            continue
        try:
            got = prog.scan(data)
        except T.TrreError as e:
            if rc == 0 and e.code not in (2, 5):
                bad += 1; print("SCAN ERROR", flavor, pat, e.code, e.text, flush=True)
            continue
        if rc != 0:
            continue
        checked += 1
        kinds[prog.info["fast_path"]] = kinds.get(prog.info["fast_path"], 0) + 1
        if got != want:
            bad += 1; print("MISMATCH", flavor, repr(pat), data[:80], got[:80], want[:80], prog.info, flush=True)
print("gpu_fuzz seed", sys.argv[1], "checked", checked, "bad", bad, "fast paths", kinds, flush=True)
sys.exit(1 if bad else 0)
