"""Small pass through every kernel family, meant to run under compute-sanitizer (memcheck / racecheck)."""
import os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import trre_b200 as T
from oracle_util import oracle_run

rng = random.Random(3)
def text(n, alphabet=b"abcd cat dog", hi=90):
    out = bytearray()
    while len(out) < n:
        out += bytes(rng.choice(alphabet) for _ in range(rng.randint(0, hi))) + b"\n"
    return bytes(out[:n - 1]) + b"\n"

cases = [("[a:A-z:Z]", 1), ("(cat:dog|dog:cat)", 0), ("aba:xyz", 0), ("a:", 0), ("a:xyz", 1), ("cat:mouse|dog:d", 0),
         ("(cat):", 1), ("a+:x", 0), (":=", 0), ("(a:b)+", 0), ("(a:x)*b|a", 0), ("abcabcabcabcabca:ABCABCABCABCABCA|bc:XY", 0)]
for expr, flavor in cases:
    prog = T.compile(expr, flavor)
    for n in (1, 17, 511, 513, 4097, 70001, 300007):
        for data in (text(n), text(n)[:-1], text(n).replace(b"ab", b"a\0", 1)):
            rc, want, _ = oracle_run(expr.encode(), data, flavor=flavor)
            got = prog.scan(data)
            assert rc == 0 and got == want, (expr, flavor, n)
    print("ok", expr, flavor, prog.info["fast_path"], flush=True)
print("sanitize_small: all equal")
