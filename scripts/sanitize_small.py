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

# every kernel family: k_bytemap_swar, k_bytemap (pair table), k_gram small/REG, k_gram non-uniform (staged), k_gram with the
# bitmap (many classes), k_bmv, k_gram_var, k_match (deep / cyclic, in place and count/write), k_lines_s, k_lines<inplace>,
# the DFS and the lazy-table executors
words = ["".join(chr(97 + (k * 7 + j * 3) % 26) for j in range(4)) + "q"[:k % 2] for k in range(40)]
cases = [("[a:A-z:Z]", 1), ("a:b|c:d|e:f|g:h|i:j|k:l", 0), ("(cat:dog|dog:cat)", 0), ("aba:xyz", 0), ("ab:XY|abc:XYZ|a:Z", 0),
         ("|".join("%s:%s" % (w[:4], w[:4].upper()) for w in words), 1), ("a:", 0), ("a:xyz", 1), ("cat:mouse|dog:d", 0),
         ("(cat):", 1), ("a+:x", 0), ("[0-9a]+:#", 0), (":=", 0), ("(a:b)+", 0), ("(a:b)*", 0), ("(a:x)*b|a", 0),
         ("abcabcabcabcabca:ABCABCABCABCABCA|bc:XY", 0), ("(a|b)*a(a|b){15}", 1)]
for expr, flavor in cases:
    prog = T.compile(expr, flavor)
    for n in (1, 17, 511, 513, 4097, 70001, 300007):
        for data in (text(n), text(n)[:-1], text(n).replace(b"ab", b"a\0", 1)):
            rc, want, _ = oracle_run(expr.encode(), data, flavor=flavor)
            got = prog.scan(data)
            assert rc == 0 and got == want, (expr, flavor, n)
    print("ok", expr, flavor, prog.info["fast_path"], flush=True)
print("sanitize_small: all equal")
