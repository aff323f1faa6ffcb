"""Seeded random trre pattern / input generators shared by the differential tests.

Patterns are drawn from the subset on which the reference terminates: a starred/plussed/counted
operand is never nullable (an epsilon cycle makes the reference loop until its stack cap or OOM,
README.md:157-166 of the reference)."""
import random

ALPHA = "abcd"


def rand_atom(rng, depth):
    """-> (text, nullable)"""
    r = rng.random()
    if r < 0.40 or depth <= 0:
        a = rng.choice(ALPHA)
        k = rng.random()
        if k < 0.35:
            return a + ":" + rng.choice(ALPHA + "XY"), False
        if k < 0.45:
            return "(" + a + ":)", False              # deletion
        if k < 0.52:
            return "(:" + rng.choice("XY") + ")", True  # insertion
        return a, False
    if r < 0.50:
        lo, hi = sorted(rng.sample("abcd", 2))
        k = rng.random()
        if k < 0.4:
            return "[%s-%s]" % (lo, hi), False
        if k < 0.8:
            return "[%s:%s-%s:%s]" % (lo, rng.choice("AWa"), hi, "Z"), False
        return "[%s%s:%s]" % (lo, hi, rng.choice("XY")), False
    if r < 0.56:
        return ".", False
    s, n = rand_expr(rng, depth - 1, True)
    return "(" + s + ")", n


def rand_term(rng, depth):
    a, nullable = rand_atom(rng, depth)
    r = rng.random()
    if nullable or r >= 0.18:
        return a, nullable
    if ":" in a and a[0] not in "([":          # 'a:X+' would parse as a:(X+), a producing epsilon cycle
        a = "(" + a + ")"
    if r < 0.12:
        op = rng.choice(["*", "+", "?", "*?", "+?", "??"])
        return a + op, op[0] != "+"
    lo = rng.randint(0, 2)
    hi = lo + rng.randint(1, 2)
    op = rng.choice(["{%d}" % max(lo, 1), "{%d,%d}" % (lo, hi), "{,%d}" % hi, "{%d,%d}?" % (lo, hi)])
    return a + op, op.startswith("{,") or op.startswith("{0,")


def rand_seq(rng, depth):
    parts = [rand_term(rng, depth) for _ in range(rng.randint(1, 3))]
    return "".join(p for p, _ in parts), all(n for _, n in parts)


def rand_expr(rng, depth=2, with_flag=False):
    alts = [rand_seq(rng, depth) for _ in range(rng.randint(1, 3))]
    s = "|".join(a for a, _ in alts)
    return (s, any(n for _, n in alts)) if with_flag else s


def rand_input(rng, nlines=None, alphabet="abcd ", maxlen=12, final_newline=True):
    nlines = rng.randint(0, 4) if nlines is None else nlines
    lines = ["".join(rng.choice(alphabet) for _ in range(rng.randint(0, maxlen))) for _ in range(nlines)]
    s = "\n".join(lines)
    if nlines and final_newline:
        s += "\n"
    return s.encode("latin1")


def rand_word_expr(rng, alphabet="abcdefgh"):
    """(w1):(o1)|(w2):(o2)... : 1-4 words of one length 2-4 over a small alphabet, a position sometimes a range; the
    shapes k_lit serves when the words cannot overlap, k_gram otherwise."""
    L = rng.randint(2, 4)
    parts = []
    for _ in range(rng.randint(1, 4)):
        word = ""
        for _k in range(L):
            if rng.random() < 0.2:
                a = rng.randrange(len(alphabet) - 1)
                b = rng.randrange(a + 1, len(alphabet))
                word += "[%s-%s]" % (alphabet[a], alphabet[b])
            else:
                word += rng.choice(alphabet)
        out = "".join(rng.choice("XYZxyz01") for _k in range(L))
        parts.append("(%s):(%s)" % (word, out))
    return "|".join(parts)


def rand_finite_expr(rng, alphabet="abc"):
    """Expressions whose DFS is finite and small on short lines whatever the mode (-a / -ma enumerate EVERY accepting
    path): iteration only over bodies that consume a byte, no empty-input producers inside loops, two or three
    alternatives at most."""
    def atom():
        r = rng.random()
        c = rng.choice(alphabet)
        if r < 0.35:
            return c
        if r < 0.6:
            return "%s:%s" % (c, rng.choice("XYZ"))
        if r < 0.7:
            return "(%s:)" % c
        if r < 0.8:
            return "[%s-%s]" % (alphabet[0], alphabet[-1])
        if r < 0.9:
            return "."
        return "(%s|%s:%s)" % (c, rng.choice(alphabet), rng.choice("XYZ"))

    def piece():
        a = atom()
        r = rng.random()
        if r < 0.15:
            return "(%s)*" % a
        if r < 0.25:
            return "(%s)+" % a
        if r < 0.35:
            return "(%s)?" % a
        if r < 0.4:
            return "(%s){1,2}" % a
        return a

    def seq():
        return "".join(piece() for _ in range(rng.randint(1, 4)))

    alts = [seq() for _ in range(rng.randint(1, 3))]
    e = "|".join(alts)
    if rng.random() < 0.3:
        e = "(%s)(:%s)?" % (e, rng.choice("XYZ"))
    return e
