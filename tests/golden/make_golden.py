#!/usr/bin/env python
"""Generates tests/golden/ref_vectors.json by RUNNING THE REAL REFERENCE BINARIES (oracle/_ref/trre and
oracle/_ref/trre_dft, compiled from /root/reference by oracle/Makefile).  The fixtures travel to the
GPU box where /root/reference does not exist.  Re-run:  python tests/golden/make_golden.py"""
import base64
import json
import os
import random
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from kat_vectors import KATS, PROBES, BAD_EXPRESSIONS  # noqa: E402
from oracle_util import NFT, DFT, SCAN, MATCH, ref_run, have_ref  # noqa: E402
from patgen import rand_expr, rand_input  # noqa: E402


def b64(b):
    return base64.b64encode(b).decode()


def main():
    assert have_ref(), "build oracle/_ref first: make -C oracle ref"
    import trre_b200 as T
    vec = []

    def add(expr, data, flavor, mode=SCAN, all_=False, tag=""):
        try:
            rc, out, err = ref_run(expr, data, flavor=flavor, mode=mode, all_=all_, timeout=10)
        except subprocess.TimeoutExpired:
            return
        if rc not in (0, 1):
            return                                   # crashes / aborts of the reference are not fixtures
        vec.append({"tag": tag, "expr": b64(expr), "in": b64(data), "flavor": flavor, "mode": mode, "all": int(all_),
                    "rc": rc, "out": b64(out), "err": err})

    for mode, inp, tre, _ in KATS:
        data = (inp + "\n").encode()
        if mode == "S":
            add(tre.encode(), data, NFT, tag="test.sh S")
            add(tre.encode(), data, DFT, tag="test.sh S (dft)")
        else:
            add(tre.encode(), data, NFT, MATCH, True, tag="test.sh M")
    for expr, data in PROBES:
        for fl in (NFT, DFT):
            add(expr, data, fl, tag="probe")
    for expr in BAD_EXPRESSIONS:
        for fl in (NFT, DFT):
            add(expr, b"abc\n", fl, tag="bad expression")
    # BASELINE config patterns on a slice of the synthetic corpus
    text = bytes(T.synth_text(0x72726500, 6000, ["cat", "dog", "concatenate", "dogma", "catdog", "dogcat", "ca", "do"], 0.3))
    for expr in (b"cat:dog", b"[a:A-z:Z]", b"(cat:dog|dog:cat)", b"[a:b-y:zz:a]"):
        for fl in (NFT, DFT):
            add(expr, text, fl, tag="config")
            add(expr, text[:-1], fl, tag="config, no final newline")
    words = ["lamp", "tree", "rock", "fish", "bird", "wolf", "moon", "star", "leaf", "sand"]
    dict_expr = "|".join("%s:%s" % (w, w.upper()) for w in words).encode()
    dtext = bytes(T.synth_text(0x72726505, 6000, words, 0.3, True))
    for fl in (NFT, DFT):
        add(dict_expr, dtext, fl, tag="dictionary")
    # seeded random patterns x inputs (the same generator the differential tests use)
    rng = random.Random(20260924)
    n_rand = 0
    while n_rand < 400:
        pat = rand_expr(rng, 2).encode()
        data = rand_input(rng, final_newline=rng.random() < 0.85)
        before = len(vec)
        for fl in (NFT, DFT):
            add(pat, data, fl, tag="random")
        n_rand += len(vec) - before
    # inputs with NUL bytes, CR, high bytes, empty lines
    for data in (b"", b"\n", b"\n\n", b"a\0b\nab\n", b"\0\n", b"caf\xc3\xa9 cat\r\ncat\n", b"\xff\xfe cat\n", b"cat\ncat"):
        for expr in (b"cat:dog", b"a:", b":=", b"[a:A-z:Z]", b".:x"):
            for fl in (NFT, DFT):
                add(expr, data, fl, tag="edge input")
    out = os.path.join(HERE, "ref_vectors.json")
    with open(out, "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "reference": "c0stya/trre @ c474a68 (oracle/_ref)",
                   "vectors": vec}, f, separators=(",", ":"))
    print("wrote", out, len(vec), "vectors,", os.path.getsize(out), "bytes")

    # ---- the other modes of the command line: -m, -ma, -a (trre_nft.c:635-648, 782-797; trre_dft.c:1287-1292) ----------
    vec = []
    modes = ((NFT, MATCH, False), (NFT, MATCH, True), (NFT, SCAN, True), (DFT, MATCH, False))
    for mode, inp, tre, _ in KATS:
        data = (inp + "\n").encode()
        for fl, m, a in modes:
            add(tre.encode(), data, fl, m, a, tag="test.sh " + mode)
    for expr in (b"cat:dog", b"(cat:dog|dog:cat)", b"cat:x|cat:y|ca:z", b".*cat.*", b"[a-z ]*(cat:dog|dog:cat)[a-z ]*"):
        for fl, m, a in modes:
            add(expr, text[:1500], fl, m, a, tag="config text")
            add(expr, text[:1499], fl, m, a, tag="config text, no final newline")
    rng = random.Random(20260925)
    n_rand = 0
    while n_rand < 300:
        pat = rand_expr(rng, 2).encode()
        data = rand_input(rng, maxlen=12, final_newline=rng.random() < 0.85)
        before = len(vec)
        for fl, m, a in modes:
            add(pat, data, fl, m, a, tag="random")
        vec[before:] = [v for v in vec[before:] if len(v["out"]) < 20000]
        n_rand += len(vec) - before
    for data in (b"", b"\n", b"\n\n", b"a\0b\nab\n", b"\0\n", b"cat\ncat", b"a\n\0a\na"):
        for expr in (b"cat:dog", b"a:", b":=", b".:x", b"a*", b"(a:x|a)*", b".*"):
            for fl, m, a in modes:
                add(expr, data, fl, m, a, tag="edge input")
    vec = [v for v in vec if len(v["out"]) < 40000]             # -ma on long lines enumerates: keep the fixture small
    out = os.path.join(HERE, "ref_mode_vectors.json")
    with open(out, "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "reference": "c0stya/trre @ c474a68 (oracle/_ref)",
                   "vectors": vec}, f, separators=(",", ":"))
    print("wrote", out, len(vec), "vectors,", os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
