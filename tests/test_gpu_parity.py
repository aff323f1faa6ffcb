"""-m gpu: the CUDA scan path (through the C ABI) against the CPU oracle, bit-exact."""
import random

import pytest

import trre_b200 as T
from oracle_util import NFT, DFT, oracle_run
from patgen import rand_expr, rand_input

pytestmark = pytest.mark.gpu

CONFIG_PATTERNS = [
    ("cat:dog", NFT), ("cat:dog", DFT),
    ("[a:A-z:Z]", DFT), ("[a:A-z:Z]", NFT),
    ("(cat:dog|dog:cat)", NFT), ("(cat:dog|dog:cat)", DFT),
    ("[a:b-y:zz:a]", DFT), ("[a:b-y:zz:a]", NFT),
]


def check(expr, flavor, data):
    rc, want, err = oracle_run(expr.encode("latin1") if isinstance(expr, str) else expr, data, flavor=flavor)
    try:
        prog = T.compile(expr, flavor)
    except T.TrreError as e:
        assert e.code == rc or rc == 2 or e.code == 2, (expr, e.code, rc)
        return None
    if rc != 0:
        with pytest.raises(T.TrreError):
            prog.scan(data)
        return prog
    got = prog.scan(data)
    assert got == want, (expr, flavor, prog.info, data[:80], got[:80], want[:80])
    return prog


@pytest.mark.parametrize("expr,flavor", CONFIG_PATTERNS)
def test_config_patterns_small(expr, flavor):
    text = bytes(T.synth_text(0x72726501, 1 << 20, ["cat", "dog", "catdog", "dogcat", "ca", "do", "cot", "dig"], 0.25))
    check(expr, flavor, text)


@pytest.mark.parametrize("data", [b"", b"\n", b"a", b"a\n", b"\n\n\n", b"abc", b"abc\nabc", b"ab\0cb\nb\n",
                                  b"\0\n", b"\0", b"cat", b"catcat\ncat\n\ncat", b"x" * 1000 + b"\n", b"x" * 1000])
@pytest.mark.parametrize("expr,flavor", [("cat:dog", NFT), ("[a:A-z:Z]", DFT), ("b:X", NFT), ("b:X", DFT),
                                         (":=", NFT), ("a:", NFT), ("a:xy", DFT), ("(cat)*:dog", NFT),
                                         ("a+:x", NFT), ("a+:x", DFT), (":a{,3}", NFT), ("(.:x)*.*", NFT)])
def test_edge_inputs(expr, flavor, data):
    check(expr, flavor, data)


def test_scan_vectors_of_reference_test_sh():
    S = [("xor", "x:", "or"), ("Mary had a little lamb", "a:", "Mry hd  little lmb"), ("or", ":=", "=o=r="),
         ("a", "a*", "a"), ("aaa", "a*", "aaa"), ("", ":a{,3}", "aaa"), ("", ":a{,3}?", ""),
         ("aaa", "(.:x)*.*", "xxx"), ("aaa", "(.:x)*?.*", "aaa"), (".", "\\.", "."), ("+", "\\+", "+"),
         ("?", "\\?", "?"), (":", "\\:", ":"), (".a", "\\.a", ".a"), ("<cat><dog>", "<(.:)*>", "<>"),
         ("<cat><dog>", "<(.:)*?>", "<><>"), ("<cat><dog>", "<(.:)+>", "<>"), ("<cat><dog>", "<(.:)+?>", "<><>")]
    for inp, tre, exp in S:
        got = T.compile(tre, NFT).scan((inp + "\n").encode())
        assert got == (exp + "\n").encode(), (inp, tre, exp, got)


@pytest.mark.parametrize("seed", range(8))
def test_random_patterns(seed):
    rng = random.Random(1000 + seed)
    for _ in range(60):
        pat = rand_expr(rng, 2)
        data = rand_input(rng, nlines=rng.randint(0, 40), maxlen=30, final_newline=rng.random() < 0.8)
        for flavor in (NFT, DFT):
            check(pat, flavor, data)


def test_device_resident_api_and_launch_count():
    torch = pytest.importorskip("torch")
    text = bytes(T.synth_text(7, 1 << 22))
    prog = T.compile("[a:A-z:Z]", DFT)
    assert prog.exec_kind == T.EXEC_BYTEMAP
    d_in = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
    d_out = torch.empty_like(d_in)
    n = prog.scan_device(d_in.data_ptr(), d_in.numel(), d_out.data_ptr(), d_out.numel(),
                         stream=torch.cuda.current_stream().cuda_stream)
    assert n == len(text)
    assert bytes(d_out.cpu().numpy().tobytes()) == text.upper()
    assert prog.launches() >= 1


# ---- k_match: the position-parallel kernel for length-preserving acyclic programs ------------------

LP_PATTERNS = [
    ("aa:bb", NFT), ("aa:bb", DFT), ("aba:xyz", NFT), ("aba:xyz", DFT),
    ("abab:wxyz|ba:XY", NFT), ("abab:wxyz|ba:XY", DFT),
    ("ab:XY|abc:XYZ|a:Z", NFT), ("ab:XY|abc:XYZ|a:Z", DFT),
    ("(cat:dog|dog:cat)", NFT), ("b:X", NFT), ("[a:A-b:B]c:C", DFT),
    ("abcabcabcabcabca:ABCABCABCABCABCA|bc:XY", NFT),          # 16-byte matches, heavy overlap
]


def _ab_text(seed, n, alphabet=b"ab", line_lo=1, line_hi=200):
    rng = random.Random(seed)
    out = bytearray()
    while len(out) < n:
        ln = rng.randint(line_lo, line_hi)
        out += bytes(rng.choice(alphabet) for _ in range(ln)) + b"\n"
    return bytes(out[:n - 1]) + b"\n"


@pytest.mark.parametrize("expr,flavor", LP_PATTERNS)
def test_match_kernel_overlapping_matches(expr, flavor):
    prog = T.compile(expr, flavor)
    assert prog.info["length_preserving"] == 1 and prog.info["max_depth"] <= 16
    for seed, alphabet in ((1, b"ab"), (2, b"abc"), (3, b"abc  ")):
        data = _ab_text(seed, 1 << 20, alphabet)
        rc, want, _ = oracle_run(expr.encode(), data, flavor=flavor)
        assert rc == 0
        got = prog.scan(data)
        assert got == want, (expr, flavor, seed)


GRAM_PATTERNS = LP_PATTERNS[:11] + [
    ("cat:dog", NFT), ("cat:dog", DFT), ("ca:xy|cat:dog", NFT), ("a:b|cat:dog", NFT), ("(c|d)(a|o)(t|g):xyz", NFT),
    ("cat:dog|do:xy|c:z", NFT), ("[a-c][a-d]:pq", DFT), ("c(a:o)t|dog:cat", DFT), ("abcab:ABCAB|bca:xyz", NFT),
    ("abcabc:ABCABC|cab:123|b:_", NFT), ("abcabc:ABCABC|cab:123|b:_", DFT), ("(a|b)(a|b)(a|b)(a|b)(a|b):vwxyz", NFT),
]


@pytest.mark.parametrize("expr,flavor", GRAM_PATTERNS)
def test_gram_kernel_against_oracle_and_match_kernel(expr, flavor, monkeypatch):
    """k_gram (every attempt tabulated by byte classes) vs the oracle, vs k_match (TRRE_NO_GRAM) on the same inputs."""
    prog = T.compile(expr, flavor)
    if prog.info["fast_path"] != 2:
        pytest.skip("byte map")
    assert prog.info["max_depth"] <= 6, prog.info
    for seed, alphabet, hi in ((1, b"ab", 200), (2, b"abc", 60), (3, b"abcat dog", 400), (4, b"cat dog do", 9)):
        data = _ab_text(seed, 1 << 20, alphabet, 0, hi)
        rc, want, _ = oracle_run(expr.encode(), data, flavor=flavor)
        assert rc == 0
        assert prog.scan(data) == want, (expr, flavor, seed, "k_gram")
        monkeypatch.setenv("TRRE_NO_GRAM", "1")
        assert prog.scan(data) == want, (expr, flavor, seed, "k_match")
        monkeypatch.delenv("TRRE_NO_GRAM")
    for data in (b"", b"cat", b"cat\n", b"catcat\ncat", b"ab\0ab\nab\n", b"a" * 9000 + b"\n", b"abab" * 3000 + b"\n",
                 b"ca\nt\n", b"c\n" * 700, b"cat" * 171 + b"\n"):
        check(expr, flavor, data)


def test_match_kernel_equals_line_kernel(monkeypatch):
    data = _ab_text(9, 3 << 20, b"abc ", 1, 3000)
    for expr, flavor in LP_PATTERNS[:8]:
        prog = T.compile(expr, flavor)
        a = prog.scan(data)
        monkeypatch.setenv("TRRE_NO_MATCH_KERNEL", "1")
        b = prog.scan(data)
        monkeypatch.delenv("TRRE_NO_MATCH_KERNEL")
        assert a == b, (expr, flavor)


@pytest.mark.parametrize("data", [
    b"a" * 20000 + b"\n",                                  # one line longer than the warm-up window
    (b"ab" * 5000 + b"\n") * 40,
    _ab_text(4, 1 << 18, b"ab")[:-1],                      # no final newline
    _ab_text(5, 1 << 18, b"ab").replace(b"aab", b"a\0b"),  # NUL bytes
    b"aa\n" * 100000,
])
def test_match_kernel_irregular_buffers_fall_back(data):
    for expr, flavor in (("aa:bb", NFT), ("aba:xyz", DFT), ("(cat:dog|dog:cat)", NFT)):
        check(expr, flavor, data)


def test_dictionary_on_gpu():
    rng = random.Random(7)
    words = set()
    while len(words) < 2000:
        words.add("".join(rng.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(4)))
    words = sorted(words)
    rng.shuffle(words)
    expr = "|".join("%s:%s" % (w, w.upper()) for w in words)
    text = bytes(T.synth_text(0x72726505, 4 << 20, words, 0.3, True))
    for flavor in (DFT, NFT):
        prog = T.compile(expr, flavor)
        data = text if flavor == DFT else text[:text.rfind(b"\n", 0, 1 << 20) + 1]    # the oracle's DFS: ~2000 pushes per byte
        rc, want, _ = oracle_run(expr.encode(), data, flavor=flavor)
        assert rc == 0 and prog.scan(data) == want, flavor


def test_dictionary_outputs_by_delta_and_by_table(monkeypatch):
    """word:WORD prints the matched bytes moved by one constant: k_gram<.., 2> adds it under a byte mask spread from the
    real-start bits (no output table).  Same bytes as the table emission (TRRE_GRAM_NO_MAP) and as the oracle; words back to
    back, across lanes and tiles, 3- and 2-letter dictionaries (other match lengths)."""
    rng = random.Random(11)
    for L, nwords in ((4, 1500), (3, 900), (2, 300)):
        words = set()
        while len(words) < nwords:
            words.add("".join(rng.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(L)))
        words = sorted(words)
        expr = "|".join("%s:%s" % (w, w.upper()) for w in words)
        dense = ("".join(rng.choice(words) for _ in range(3000)) + "\n").encode() * 3          # nothing but matches
        text = bytes(T.synth_text(0x72726517, 2 << 20, words, 0.5, True)) + dense
        want = oracle_run(expr.encode(), text, flavor=DFT)[1]
        prog = T.compile(expr, DFT)
        assert prog.info["gram_tables"] and prog.scan(text) == want, L
        monkeypatch.setenv("TRRE_GRAM_NO_MAP", "1")
        assert T.compile(expr, DFT).scan(text) == want, L
        monkeypatch.delenv("TRRE_GRAM_NO_MAP")


# ---- the general path: shared-memory staged kernels against the plain ones and the oracle -----------

@pytest.mark.parametrize("expr,flavor", [("a:", NFT), (":=", NFT), ("a:xy", DFT), ("a+:x", NFT), ("(cat)*:dog", NFT),
                                         ("[a-c]+:<&>", NFT), ("b:", DFT), ("(.:x)*?.*", NFT), ("ab:X|abc:Y|a:Z", DFT)])
def test_general_path_large_inputs(expr, flavor, monkeypatch):
    for seed, alphabet, hi in ((11, b"abc ", 300), (12, b"ab", 40), (13, b"abcat dog", 5000)):
        data = _ab_text(seed, 3 << 20, alphabet, 0, hi)
        rc, want, _ = oracle_run(expr.encode(), data, flavor=flavor)
        assert rc == 0
        prog = T.compile(expr, flavor)
        got = prog.scan(data)
        assert got == want, (expr, flavor, seed)
        monkeypatch.setenv("TRRE_NO_STAGED_LINES", "1")
        plain = prog.scan(data)
        monkeypatch.delenv("TRRE_NO_STAGED_LINES")
        assert plain == want, (expr, flavor, seed, "plain kernels")


def test_general_path_expansion_beyond_staging():
    # every byte becomes 8: the group's output (256 KB) does not fit the 40 KB staging window
    data = _ab_text(21, 1 << 20, b"ab", 1, 100)
    for expr, flavor in (("a:AAAAAAAA|b:BBBBBBBB", NFT), (":0123456789", NFT)):
        rc, want, _ = oracle_run(expr.encode(), data, flavor=flavor)
        assert rc == 0 and T.compile(expr, flavor).scan(data) == want


# ---- BASELINE sizes: size-independent properties (lines are independent => tiles replicate) ------------

def test_full_size_cfg2_one_gib_device_resident():
    torch = pytest.importorskip("torch")
    tile = T.synth_text(0x72726502, 256 << 20)
    prog = T.compile("[a:A-z:Z]", DFT)
    d_tile = torch.frombuffer(tile, dtype=torch.uint8).cuda()
    d_in = d_tile.repeat(4)                                       # 1 GiB
    d_out = torch.empty_like(d_in)
    n = prog.scan_device(d_in.data_ptr(), d_in.numel(), d_out.data_ptr(), d_out.numel(),
                         stream=torch.cuda.current_stream().cuda_stream)
    assert n == d_in.numel() and prog.launches() == 1
    lut = torch.frombuffer(bytearray(prog.bytemap()), dtype=torch.uint8).cuda()
    for k in range(4):
        sl = slice(k * len(tile), (k + 1) * len(tile))
        assert torch.equal(d_out[sl], lut[d_in[sl].long()])
    # the first 8 MiB against the oracle itself
    head = bytes(tile[:8 << 20])
    cut = head.rfind(b"\n") + 1
    rc, want, _ = oracle_run(b"[a:A-z:Z]", head[:cut], flavor=DFT)
    assert rc == 0 and bytes(d_out[:cut].cpu().numpy().tobytes()) == want


def test_full_size_cfg3_replicated_tiles():
    torch = pytest.importorskip("torch")
    words = ["cat", "dog", "catdog", "dogcat", "ca", "do", "cot", "dig"]
    tile = T.synth_text(0x72726503, 128 << 20, words, 0.25)
    prog = T.compile("(cat:dog|dog:cat)", NFT)
    d_tile = torch.frombuffer(tile, dtype=torch.uint8).cuda()
    d_in = d_tile.repeat(8)                                       # 1 GiB
    d_out = torch.empty_like(d_in)
    n = prog.scan_device(d_in.data_ptr(), d_in.numel(), d_out.data_ptr(), d_out.numel(),
                         stream=torch.cuda.current_stream().cuda_stream)
    assert n == d_in.numel()
    for k in range(1, 8):
        assert torch.equal(d_out[:len(tile)], d_out[k * len(tile):(k + 1) * len(tile)])
    head = bytes(tile[:8 << 20])
    cut = head.rfind(b"\n") + 1
    rc, want, _ = oracle_run(b"(cat:dog|dog:cat)", head[:cut], flavor=NFT)
    assert rc == 0 and bytes(d_out[:cut].cpu().numpy().tobytes()) == want
    # length preserving + involution: cat<->dog applied twice is the identity on this corpus
    d_back = torch.empty_like(d_in)
    prog.scan_device(d_out.data_ptr(), n, d_back.data_ptr(), n, stream=torch.cuda.current_stream().cuda_stream)
    assert torch.equal(d_back, d_in)


# ---- variable-length byte maps (k_bmv_count / k_bmv_write) ---------------------------------------------

BMV_PATTERNS = [("a:", NFT), ("a:", DFT), ("a:xy", NFT), ("a:xy", DFT), ("[a-c]:", NFT), ("a:|b:XYZ|c:C", NFT),
                ("a:" + "Q" * 200, NFT), (" :__", DFT), ("[a:A-c:C]|d:", NFT)]


@pytest.mark.parametrize("expr,flavor", BMV_PATTERNS)
def test_bytemap_var_against_oracle_and_general_path(expr, flavor, monkeypatch):
    for seed, alphabet, hi, size in ((31, b"abcd ", 120, 2 << 20), (32, b"ab", 3000, 1 << 20), (33, b"a", 50, 1 << 18)):
        data = _ab_text(seed, size, alphabet, 0, hi)
        rc, want, _ = oracle_run(expr.encode(), data, flavor=flavor)
        assert rc == 0
        prog = T.compile(expr, flavor)
        assert prog.scan(data) == want, (expr, flavor, seed, "k_bmv")
        monkeypatch.setenv("TRRE_NO_BMV", "1")
        assert prog.scan(data) == want, (expr, flavor, seed, "general")
        monkeypatch.delenv("TRRE_NO_BMV")
    for data in (b"", b"a", b"a\n", b"aaa", b"b\0a\na\n", b"\n\n", b"ab" * 7 + b"\n"):
        check(expr, flavor, data)


def test_bytemap_var_device_api_unaligned_output_and_capacity():
    torch = pytest.importorskip("torch")
    data = _ab_text(41, 5 << 20, b"abc ", 1, 90)
    prog = T.compile("a:", NFT)
    rc, want, _ = oracle_run(b"a:", data, flavor=NFT)
    d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    buf = torch.zeros(len(data) + 64, dtype=torch.uint8, device="cuda")
    for shift in (0, 1, 7, 15):
        d_out = buf[shift:]
        n = prog.scan_device(d_in.data_ptr(), d_in.numel(), d_out.data_ptr(), d_out.numel(),
                             stream=torch.cuda.current_stream().cuda_stream)
        assert n == len(want) and bytes(d_out[:n].cpu().numpy().tobytes()) == want, shift
    with pytest.raises(T.TrreError) as ei:
        prog.scan_device(d_in.data_ptr(), d_in.numel(), buf.data_ptr(), 1000, stream=torch.cuda.current_stream().cuda_stream)
    assert ei.value.code == 4 and ei.value.needed == len(want)


# ---- variable-length matches (k_match<OP_COUNT/OP_WRITE>): s/cat/mouse/ and friends ------------------------

MV_PATTERNS = [("cat:mouse", NFT), ("cat:mouse", DFT), ("(cat):", NFT), ("(cat):", DFT), ("cat:mouse|dog:d", NFT),
               ("ab:X|abc:Y|a:Z", NFT), ("ab:X|abc:Y|a:Z", DFT), ("(aa):b", NFT), ("aba:xy", DFT),
               ("abab:w|ba:XYZ", NFT), ("(abcabcabcabcabca):Q|bc:XYZW", NFT), ("a(b:XY)c", NFT)]


@pytest.mark.parametrize("expr,flavor", MV_PATTERNS)
def test_variable_length_matches(expr, flavor, monkeypatch):
    prog = T.compile(expr, flavor)
    assert prog.info["max_depth"] <= 16          # (the DFT flavour of ab:X|abc:Y|a:Z always takes a:Z: length preserving)
    for seed, alphabet, hi, size in ((51, b"abc", 200, 2 << 20), (52, b"ab", 60, 1 << 20), (53, b"cat dogma", 120, 2 << 20)):
        data = _ab_text(seed, size, alphabet, 0, hi)
        rc, want, _ = oracle_run(expr.encode(), data, flavor=flavor)
        assert rc == 0
        assert prog.scan(data) == want, (expr, flavor, seed, "k_gram_var / k_match var")
        monkeypatch.setenv("TRRE_NO_GRAM", "1")
        assert prog.scan(data) == want, (expr, flavor, seed, "k_match var")
        monkeypatch.delenv("TRRE_NO_GRAM")
        monkeypatch.setenv("TRRE_NO_MATCH_KERNEL", "1")
        assert prog.scan(data) == want, (expr, flavor, seed, "general")
        monkeypatch.delenv("TRRE_NO_MATCH_KERNEL")
    for data in (b"", b"cat", b"cat\n", b"catcat\ncat", b"ab\0ab\nab\n", b"a" * 9000 + b"\n", b"abab" * 3000 + b"\n"):
        check(expr, flavor, data)


def test_variable_length_dictionary_and_device_api():
    torch = pytest.importorskip("torch")
    rng = random.Random(5)
    words = sorted({"".join(rng.choice("abcdefgh") for _ in range(rng.randint(2, 6))) for _ in range(300)})
    # prefix-free is not required: the NFT flavour takes the first alternative, the DFT flavour the shortest
    expr = "|".join("%s:%s" % (w, w.upper() * rng.randint(0, 2)) for w in words)
    text = bytes(T.synth_text(77, 3 << 20, words[:40], 0.3, True)).replace(b"z", b"a")
    for flavor in (NFT, DFT):
        rc, want, _ = oracle_run(expr.encode(), text, flavor=flavor)
        assert rc == 0
        prog = T.compile(expr, flavor)
        assert prog.scan(text) == want, flavor
    prog = T.compile("cat:mouse", NFT)
    data = _ab_text(61, 4 << 20, b"cat dog", 1, 100)
    rc, want, _ = oracle_run(b"cat:mouse", data, flavor=NFT)
    d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    buf = torch.zeros(2 * len(data) + 64, dtype=torch.uint8, device="cuda")
    for shift in (0, 3, 15):
        d_out = buf[shift:]
        n = prog.scan_device(d_in.data_ptr(), d_in.numel(), d_out.data_ptr(), d_out.numel(),
                             stream=torch.cuda.current_stream().cuda_stream)
        assert n == len(want) and bytes(d_out[:n].cpu().numpy().tobytes()) == want, shift


@pytest.mark.parametrize("seed", range(4))
def test_random_patterns_on_inputs_that_span_many_tiles(seed):
    """Random expressions whose program gets a position-parallel kernel, on inputs large enough to cross tiles,
    warp spans and CTAs (the small-input tests above never leave one tile)."""
    rng = random.Random(7000 + seed)
    seen = {}
    tried = 0
    while tried < 400 and sum(seen.values()) < 24:
        tried += 1
        pat = rand_expr(rng, 2)
        for flavor in (NFT, DFT):
            try:
                prog = T.compile(pat, flavor)
            except T.TrreError:
                continue
            fp = prog.info["fast_path"]
            if fp not in (2, 3, 4, 5) or seen.get(fp, 0) >= 8:
                continue
            data = _ab_text(rng.randrange(1 << 30), 200000, b"abcd ", 0, rng.choice([8, 60, 400]))
            rc, want, _ = oracle_run(pat.encode(), data, flavor=flavor)
            if rc != 0:
                continue
            seen[fp] = seen.get(fp, 0) + 1
            got = prog.scan(data)
            assert got == want, (pat, flavor, fp, prog.info)
    assert sum(seen.values()) >= 8, seen


# ---- TRRE_EXEC_LAZY: rows built on demand, the GPU reports what it misses (trre_dft.c:1135-1175) ----------------

@pytest.mark.parametrize("expr", ["(a|b)*a(a|b){15}", "(a|b)*(a:x)(a|b){15}", "(a|b)*(a:x|b:y)(a|b){16}c", "(a:x)*b|(a:y)*c",
                                  "((a|b)(a|b)(a|b))*(a:1)(a|b){17}c"])
def test_lazy_tables_on_the_gpu(expr):
    prog = T.compile(expr, DFT)
    assert prog.exec_kind == T.EXEC_LAZY, prog.info
    for seed, alphabet, hi, size in ((71, b"ab", 60, 1 << 18), (72, b"abc ", 400, 1 << 20), (73, b"ab", 5, 1 << 16)):
        data = _ab_text(seed, size, alphabet, 0, hi)
        rc, want, _ = oracle_run(expr.encode(), data, flavor=DFT)
        assert rc == 0
        assert prog.scan(data) == want, (expr, seed)
    for data in (b"", b"a", b"ab\n", b"ab\0ab\nab", b"a" * 3000 + b"b\n", b"\n\n"):
        check(expr, DFT, data)
    q = T.from_blob(prog.blob())                          # what a rank gets from the broadcast
    data = _ab_text(74, 1 << 16, b"abc", 0, 80)
    assert q.scan(data) == oracle_run(expr.encode(), data, flavor=DFT)[1]


def test_dfs_executor_has_the_reference_stack_capacity():
    """trre's stack holds 65536 entries (trre_nft.c:34-36, 548-556); the DFS executor starts with 8192 per thread and
    repeats the scan with the reference's capacity when a line needs more.  (One deep attempt per line: a failing
    attempt of depth d costs d steps at EVERY position of the run, and one GPU thread walks them one at a time.)"""
    expr = "(a:x)*b|a"
    prog = T.compile(expr, NFT)
    assert prog.exec_kind == T.EXEC_BACKTRACK
    for run in (9000, 60000):
        data = b"ab\n" + b"a" * run + b"b\n" + b"aaab\n" + b"a" * 300 + b"\n"
        rc, want, _ = oracle_run(expr.encode(), data, flavor=NFT)
        assert rc == 0 and prog.scan(data) == want, run
    data = b"a" * 70000 + b"\n"
    rc, _, err = oracle_run(expr.encode(), data, flavor=NFT)
    assert rc == 1 and "stack max capacity" in err
    with pytest.raises(T.TrreError) as ei:
        T.compile(expr, NFT).scan(data)
    assert ei.value.code == 5 and "stack max capacity reached" in ei.value.text


@pytest.mark.parametrize("seed", range(4))
def test_match_and_all_modes_on_the_gpu(seed):
    """-m / -ma / -a (trre_nft.c:635-648, 782-797) through trre_prog_set_mode: GPU DFS vs the reference on random
    expressions whose enumeration is finite, lines of at most 8 bytes; the expected output comes from a CHILD process
    under CPU / memory / output limits (oracle_util.bounded_run) - `-ma` is allowed to print gigabytes."""
    from oracle_util import MATCH, SCAN, bounded_run
    from patgen import rand_finite_expr
    rng = random.Random(7000 + seed)
    ran = 0
    for _ in range(40):
        pat = rand_finite_expr(rng)
        data = rand_input(rng, nlines=rng.randint(0, 30), alphabet="abc ", maxlen=8, final_newline=rng.random() < 0.8)
        for flavor, mode, all_ in ((NFT, SCAN, True), (NFT, MATCH, False), (NFT, MATCH, True), (DFT, MATCH, False)):
            ref = bounded_run(pat.encode("latin1"), data, flavor=flavor, mode=mode, all_=all_, cpu_s=2, out_mb=4)
            if ref is None or ref[0] != 0:
                continue
            try:
                prog = T.compile(pat, flavor).set_mode(match=mode == MATCH, all_=all_)
            except T.TrreError:
                continue
            got = prog.scan(data)
            assert got == ref[1], (pat, flavor, mode, all_, data[:80], got[:80], ref[1][:80])
            if flavor == NFT:
                assert prog.match(data, all_=all_, scan=mode == SCAN) == ref[1]
            ran += 1
    assert ran > 60, ran


def test_modes_on_config_text():
    """-a in scan mode and -m on 1 MiB of the synthetic corpus: many records per kernel launch, chunk boundaries."""
    from oracle_util import MATCH, SCAN
    text = bytes(T.synth_text(0x72726503, 1 << 20, ["cat", "dog", "catdog", "dogcat", "ca", "do", "cot", "dig"], 0.25))
    for expr, mode, all_ in (("(cat:dog|dog:cat)", SCAN, True), ("cat:x|cat:y|ca:z", SCAN, True), (".*cat.*", MATCH, False),
                             ("(.:x|.)*", MATCH, False), ("[a-z ]*(cat:dog|dog:cat)[a-z ]*", MATCH, True)):
        rc, want, _ = oracle_run(expr.encode(), text, flavor=NFT, mode=mode, all_=all_)
        assert rc == 0
        assert T.compile(expr, NFT).set_mode(match=mode == MATCH, all_=all_).scan(text) == want, expr
    assert T.compile("cat:dog", DFT).set_mode(match=True).scan(text) == b"\n" * text.count(b"\n")


def test_literal_words_kernel(monkeypatch):
    """k_lit (lit_kernel.cuh): fixed words compared in registers - against the oracle on random word programs, buffer
    sizes around the kernel's 16-byte vectors, 30-vector warp windows and 15 360-byte CTAs, words across every such
    boundary, high bytes, and the irregular cases (NUL, no final newline) that must take the general path."""
    from patgen import rand_word_expr
    rng = random.Random(4242)
    served = 0
    for _ in range(60):
        pat = rand_word_expr(rng)
        n = rng.choice([1, 2, 5, 15, 16, 17, 31, 33, 479, 480, 481, 500, 1000, 15359, 15360, 15361, 15376, 30720, 100000])
        body = bytes(rng.choice(b"abcdefgh  \n") for _k in range(n))
        data = body[:-1] + b"\n"
        for flavor in (NFT, DFT):
            prog = check(pat, flavor, data)
            if prog is not None and prog.info["literal_words"]:
                served += 1
                assert prog.launches() == 1
    assert served > 30, served
    text = bytes(T.synth_text(0x72726511, 3 << 20, ["cat", "dog", "catdog", "dogcat", "ca", "do", "cot", "dig"], 0.3))
    for expr in ("cat:dog", "(cat:dog|dog:cat)", "(ca):(xy)|(do):(zw)|(ig):(ab)|(ut):(cd)", "(catd):(DOGC)", "c(a:o)t|dog:cat"):
        for flavor in (NFT, DFT):
            prog = check(expr, flavor, text)
            assert prog.info["literal_words"] >= 1
            for cut in (1, 15, 16, 17, 480 * 16, 480 * 16 + 5):                # ragged ends, every one a final newline
                piece = text[:len(text) - cut]
                piece = piece[:piece.rfind(b"\n") + 1]
                check(expr, flavor, piece)
            check(expr, flavor, text[:70000] + b"c\xe1t \xff dog\0cat\n" + text[:5000])      # high bytes, a NUL
            check(expr, flavor, text[:70001].rstrip(b"\n"))                                 # no final newline
    monkeypatch.setenv("TRRE_LIT_NO_EQ", "1")              # plain strings through the range form of k_lit: same bytes
    for expr in ("cat:dog", "(cat:dog|dog:cat)", "(catd):(DOGC)"):
        q = T.compile(expr, NFT)
        assert q.info["literal_words"] and q.scan(text) == oracle_run(expr.encode(), text, flavor=NFT)[1]
    monkeypatch.delenv("TRRE_LIT_NO_EQ")
    monkeypatch.setenv("TRRE_NO_LIT", "1")                 # the same programs on k_gram: same bytes
    for expr in ("cat:dog", "(cat:dog|dog:cat)"):
        q = T.compile(expr, NFT)
        assert q.scan(text) == oracle_run(expr.encode(), text, flavor=NFT)[1]


def test_run_form_kernels(monkeypatch):
    """k_bmv_count_run / k_bmv_write<RUN> (`X+:s`: one byte of left context): runs across lanes (16 B), tiles (512 B), spans
    and pipeline chunks; runs at the start and the end of the buffer; runs longer than a tile; several run classes."""
    rng = random.Random(99)
    pieces = ["a", "aa", "a" * 15, "a" * 16, "a" * 17, "a" * 511, "a" * 513, "a" * 5000, "b", "bbb", "7", "2026", "9" * 700, " ", "    ",
              "x", "cat", "\n", "\n", "\n"]
    body = "".join(rng.choice(pieces) for _ in range(6000))
    data = ("aaa" + body + "aaaa\n").encode()
    text = bytes(T.synth_text(0x72726513, 2 << 20, ["1", "22", "333", "4444", "a", "aa"], 0.3))
    for expr in ("[0-9]+:#", "a+:x", "a+:", "( +): ", "a+:xyz|b:", "a+:x|b+:yy", "[a-b]+:Q|[0-9]+:#|c:", "(a|b)+:x",
                 "(a|7)+:x"):                                     # (a set that is not one range: counted byte by byte)
        prog = T.compile(expr, NFT)
        assert prog.info["fast_path"] == 3                         # TRRE_FAST_BYTEMAP_VAR
        for buf in (data, text, data[:1], b"a\n", b"aaaa\n", b"\n", data[:511] + b"\n", data[:512] + b"\n", data[:513] + b"\n"):
            rc, want, _ = oracle_run(expr.encode(), buf, flavor=NFT)
            assert rc == 0 and prog.scan(buf) == want, (expr, len(buf))
        check(expr, NFT, data[:3000] + b"a\0a\n" + data[:100])          # a NUL: the general path
        check(expr, NFT, data[:3000].rstrip(b"\n") + b"aa")              # no final newline
    monkeypatch.setenv("TRRE_RUN_PER_BYTE", "1")                   # the count pass without the SWAR range form: same bytes
    for expr in ("[0-9]+:#", "a+:x|b+:yy"):
        assert T.compile(expr, NFT).scan(data) == oracle_run(expr.encode(), data, flavor=NFT)[1]


def test_literal_words_with_outputs_of_any_length():
    """k_litv_count / k_bmv_write<BMV_LIT> (`cat:mouse|dog:d`): k_lit's comparison, k_bmv's placement.  Matches across lanes
    (16 B), tiles (512 B) and spans; deletions, expansions; irregular buffers go to the general path."""
    from patgen import rand_word_expr
    rng = random.Random(777)
    served = 0
    for _ in range(40):
        pat = rand_word_expr(rng)
        pat = "|".join(part[:part.index(":(")] + ":(" + rng.choice(["", "Q", "longer output", part[part.index(":(") + 2:-1] + "x"]) + ")"
                       for part in pat.split("|"))
        n = rng.choice([1, 2, 15, 16, 17, 33, 511, 512, 513, 1000, 8191, 8192, 8193, 100000, 300000])
        body = bytes(rng.choice(b"abcdefgh  \n") for _k in range(n))
        data = body[:-1] + b"\n"
        for flavor in (NFT, DFT):
            prog = check(pat, flavor, data)
            if prog is not None and prog.info["literal_words"] and prog.info["fast_path"] == 4:
                served += 1
    assert served > 15, served
    text = bytes(T.synth_text(0x72726515, 3 << 20, ["cat", "dog", "catdog", "dogcat", "ca", "do", "cot", "dig"], 0.3))
    for expr in ("cat:mouse|dog:d", "(cat):", "cat:mouse", "(ca):(xyz)|(do):|(ig):(0123456789)", "(catd):(!)", "[c-d]at:(hello)"):
        prog = check(expr, NFT, text)
        assert prog.info["literal_words"] >= 1 and prog.info["fast_path"] == 4
        check(expr, NFT, text[:70000] + b"cat\0dog\n" + text[:5000])
        check(expr, NFT, text[:70001].rstrip(b"\n"))
        dense = (b"catdog" * 200 + b"\n") * 50                    # every byte inside a match: tiles that shrink / grow 3x
        check(expr, NFT, dense)
