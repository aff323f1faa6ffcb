"""Differential fuzz on the GPU: random expressions x random inputs, CUDA path vs the reference.
Two steps, so that no GPU-box time (and no GPU-box RAM) goes into the CPU side - random expressions may backtrack for
minutes in the reference:
  python scripts/gpu_fuzz.py make SEED COUNT FILE    (anywhere) expectations from oracle/_ref (or the oracle) in a child
                                                      process under CPU / memory / output limits; expensive cases are dropped
  python scripts/gpu_fuzz.py run FILE                (GPU box)  every case through trre_scan, compared byte for byte"""
import base64, json, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def make(seed, count, path):
    from oracle_util import bounded_run
    from patgen import rand_expr, rand_input, rand_word_expr
    rng = random.Random(seed)
    cases, dropped = [], 0
    for it in range(count + count // 3):
        if it < count:
            pat = rand_expr(rng, rng.choice([1, 2, 2, 3]))
        else:
            # fixed words (k_lit / k_litv / k_gram when they can overlap), outputs of the same or of other lengths, and runs
            pat = rand_word_expr(rng, "abcd")
            if rng.random() < 0.5:
                pat = "|".join(part[:part.index(":(")] + ":(" + rng.choice(["", "Q", "longer", part[part.index(":(") + 2:-1] + "x"]) + ")"
                               for part in pat.split("|"))
            if rng.random() < 0.2:
                pat = rng.choice(["a+:x", "[a-b]+:#", "( +): ", "a+:|b:YY", "a+:x|c+:zz"])
        alphabet = rng.choice(["abcd ", "ab", "abcdXY \t", "aaab"])
        nlines = rng.choice([0, 1, 3, 30, 400])
        data = rand_input(rng, nlines=nlines, alphabet=alphabet, maxlen=rng.choice([4, 12, 60]), final_newline=rng.random() < 0.85)
        if rng.random() < 0.08:
            data = data.replace(b"a", b"\0", 1)
        for flavor in (0, 1):
            ref = bounded_run(pat.encode("latin1"), data, flavor=flavor, cpu_s=1, mem_mb=512, out_mb=4)
            if ref is None:
                dropped += 1
                continue
            cases.append({"pat": pat, "flavor": flavor, "in": base64.b64encode(data).decode(), "rc": ref[0],
                          "out": base64.b64encode(ref[1]).decode()})
    with open(path, "w") as f:
        json.dump({"seed": seed, "count": count, "dropped": dropped, "cases": cases}, f, separators=(",", ":"))
    print("wrote", path, len(cases), "cases,", dropped, "dropped (limits)")


def run(path):
    import trre_b200 as T
    with open(path) as f:
        doc = json.load(f)
    bad = checked = 0
    kinds = {}
    for c in doc["cases"]:
        pat, flavor, data, rc, want = c["pat"], c["flavor"], base64.b64decode(c["in"]), c["rc"], base64.b64decode(c["out"])
        try:
            prog = T.compile(pat, flavor)
        except T.TrreError as e:
            if e.code != 2 and e.code != rc:
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
        key = "%d%s%s" % (prog.info["fast_path"], "+lit" if prog.info["literal_words"] else "", "" )
        kinds[key] = kinds.get(key, 0) + 1
        if got != want:
            bad += 1; print("MISMATCH", flavor, repr(pat), data[:80], got[:80], want[:80], prog.info, flush=True)
    print("gpu_fuzz seed", doc["seed"], "count", doc["count"], "checked", checked, "bad", bad, "dropped by the limits", doc["dropped"],
          "fast paths", kinds, flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    if sys.argv[1] == "make":
        make(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
    else:
        sys.exit(run(sys.argv[2]))
