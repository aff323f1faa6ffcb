"""The kernels added in the second half of round 2, small, for compute-sanitizer: k_lit (EQ and range forms), k_litv_count +
k_bmv_write<BMV_LIT>, the RUN form of k_bmv_*, and the DFS executor in the CLI's other modes (-a, -m, -ma; trre_dft -m)."""
import os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import trre_b200 as T
from oracle_util import oracle_run, SCAN, MATCH

rng = random.Random(4)
def text(n, alphabet=b"abcd cat dog 0123", hi=90):
    out = bytearray()
    while len(out) < n:
        out += bytes(rng.choice(alphabet) for _ in range(rng.randint(0, hi))) + b"\n"
    return bytes(out[:n - 1]) + b"\n"

cases = [("cat:dog", 0), ("(cat:dog|dog:cat)", 0), ("(catd):(DOGC)", 1), ("[c-d]og:xyz|cat:CAT", 0), ("(ca):(xy)|(do):(zw)|(ig):(ab)|(ut):(cd)", 0),
         ("cat:mouse|dog:d", 0), ("(cat):", 0), ("(ca):(xyz)|(do):|(ig):(0123456789)", 0), ("[0-9]+:#", 0), ("a+:x", 0), ("a+:xyz|b:", 0),
         ("a+:x|b+:yy", 0)]
for expr, flavor in cases:
    prog = T.compile(expr, flavor)
    for n in (1, 17, 481, 513, 15361, 70001, 300007):
        for data in (text(n), text(n)[:-1], text(n).replace(b"ab", b"a\0", 1)):
            rc, want, _ = oracle_run(expr.encode(), data, flavor=flavor)
            assert rc == 0 and prog.scan(data) == want, (expr, flavor, n)
    print("ok", expr, flavor, prog.info["fast_path"], prog.info["literal_words"], flush=True)
for expr, flavor, match, all_ in (("(cat:dog|dog:cat|cat:CAT)", 0, False, True), ("[a-d ]*(cat:dog)[a-d ]*", 0, True, False),
                                  ("[a-d ]*(cat:dog|cat)[a-d ]*", 0, True, True), ("cat:dog", 1, True, False)):
    prog = T.compile(expr, flavor).set_mode(match=match, all_=all_)
    for n in (1, 17, 513, 70001):
        data = text(n, b"abcd cat", 24)
        rc, want, _ = oracle_run(expr.encode(), data, flavor=flavor, mode=MATCH if match else SCAN, all_=all_)
        assert rc == 0 and prog.scan(data) == want, (expr, match, all_, n)
    print("ok mode", expr, match, all_, flush=True)
print("sanitize_r2b: all equal")
