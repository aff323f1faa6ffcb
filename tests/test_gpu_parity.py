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
