"""Known-answer vectors of the reference's own test-suite (test.sh:29-132) as data: (mode, input,
expression, expected).  'S' = scan mode through `./trre`, 'M' = `./trre -ma` (match, all outputs).
Expected strings are what test.sh passes to `echo -e` (so "\\n" separates outputs); test.sh compares
after command substitution, i.e. modulo trailing newlines."""

KATS = [
    # basics (test.sh:29-36)
    ("M", "a", "a:x", "x"), ("M", "ab", "ab:xy", "xy"), ("M", "ab", "(a:x)(b:y)", "xy"),
    ("M", "cat", "cat:dog", "dog"), ("M", "cat", "(cat):(dog)", "dog"), ("M", "cat", "(c:d)(a:o)(t:g)", "dog"),
    ("M", "mat", "c:da:ot:g", ""),
    # deletion (38-41)
    ("M", "xor", "(x:)or", "or"), ("S", "xor", "x:", "or"), ("S", "Mary had a little lamb", "a:", "Mry hd  little lmb"),
    # insertion (43-45)
    ("M", "or", "(:x)or", "xor"), ("S", "or", ":=", "=o=r="),
    # alternation (47-50)
    ("M", "a", "a|b|c", "a"), ("M", "b", "a|b|c", "b"), ("M", "c", "a|b|c", "c"),
    # star (52-58)
    ("S", "a", "a*", "a"), ("S", "aaa", "a*", "aaa"), ("M", "b", "a*", ""), ("M", "bbb", "a*", ""),
    ("M", "abab", "(ab)*", "abab"), ("M", "ababa", "(ab)*", ""),
    # ranges (60-64)
    ("M", "a", "[a-c]", "a"), ("M", "b", "[a-c]", "b"), ("M", "c", "[a-c]", "c"), ("M", "d", "[a-c]", ""),
    # transform ranges (66-70)
    ("M", "a", "[a:x]", "x"), ("M", "a", "[a:xb:y]", "x"), ("M", "b", "[a:xb:y]", "y"), ("M", "c", "[a:xb:y]", ""),
    # range-transform ranges (72-76)
    ("M", "a", "[a:x-c:z]", "x"), ("M", "b", "[a:x-c:z]", "y"), ("M", "c", "[a:x-c:z]", "z"), ("M", "d", "[a:x-c:z]", ""),
    # any char (78-81)
    ("M", "a", ".", "a"), ("M", "b", ".", "b"), ("M", "abc", "...", "abc"),
    # iteration (83-104)
    ("M", "aa", "a{2}", "aa"), ("M", "a", "a{2}", ""), ("M", "aaa", "a{2}", ""),
    ("M", "", "a{,2}", ""), ("M", "a", "a{,2}", "a"), ("M", "aa", "a{,2}", "aa"), ("M", "aaa", "a{,2}", ""),
    ("M", "", "a{1,2}", ""), ("M", "a", "a{1,2}", "a"), ("M", "aa", "a{1,2}", "aa"), ("M", "aaa", "a{1,2}", ""),
    ("M", "", "a{2,}", ""), ("M", "a", "a{2,}", ""), ("M", "aa", "a{2,}", "aa"), ("M", "aaa", "a{2,}", "aaa"),
    # generators (106-111)
    ("S", "", ":a{,3}", "aaa"), ("M", "", ":a{,3}", "aaa\naa\na"),
    ("S", "", ":a{,3}?", ""), ("M", "", ":a{,3}?", "\na\naa\naaa"),
    # greed modifiers (113-118)
    ("S", "aaa", "(.:x)*.*", "xxx"), ("S", "aaa", "(.:x)*?.*", "aaa"),
    ("M", "aaa", "(.:x)*.*", "xxx\nxxa\nxaa\naaa"), ("M", "aaa", "(.:x)*?.*", "aaa\nxaa\nxxa\nxxx"),
    # escapes (120-126)
    ("S", ".", "\\.", "."), ("S", "+", "\\+", "+"), ("S", "?", "\\?", "?"), ("S", ":", "\\:", ":"),
    ("S", ".a", "\\.a", ".a"), ("M", ".c", "[.]c", ".c"),
    # greed modifiers priorities (128-132)
    ("S", "<cat><dog>", "<(.:)*>", "<>"), ("S", "<cat><dog>", "<(.:)*?>", "<><>"),
    ("S", "<cat><dog>", "<(.:)+>", "<>"), ("S", "<cat><dog>", "<(.:)+?>", "<><>"),
]

# SURVEY.md s4.4 probe log: (expr, stdin bytes) - behaviours the README does not tell
PROBES = [
    (b"cat:dog", b"cat"), (b"(cat)*:dog", b"catcatcat\n"), (b":=", b"or\n"), (b"a+:x", b"aaab\n"),
    (b"ab:X|abc:Y|a:Z", b"abc ab a abcd\n"), (b"(:x)a|(:y)b", b"ab\n"), (b"(..):x", b"abcd\n"),
    (b"b:X", b"ab\0cb\nb\n"), (b"xo:", b"xor\n"), (b"a:x|a:y", b"ab\n"),
]

# expressions the reference rejects: (expr) -> stderr text + exit 1 (trre_nft.c parser / nft())
BAD_EXPRESSIONS = [b"(a", b"a)", b"[a", b"[a-", b"[]", b"[:a]", b"a{1,2,3}", b"a{x}", b"a{1", b"*a", b"|a", b"a||b",
                   b"[a-c:x]", b"[[a]]", b"a{", b"(", b")", b"+"]
