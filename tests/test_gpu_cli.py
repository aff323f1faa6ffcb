"""-m gpu: the drop-in command lines (bin/trre, bin/trre_dft) against the real reference binaries
(oracle/_ref, shipped prebuilt to the GPU box) or, when those are absent, the oracle: same argv, same
stdin, byte-identical stdout, same exit code."""
import os
import subprocess

import pytest

import trre_b200 as T
from oracle_util import NFT, DFT, have_ref, oracle_run, REF_NFT, REF_DFT

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = {NFT: os.path.join(ROOT, "bin", "trre"), DFT: os.path.join(ROOT, "bin", "trre_dft")}
REF = {NFT: REF_NFT, DFT: REF_DFT}

TEXT = bytes(T.synth_text(0x72726501, 200000, ["cat", "dog", "concatenate", "dogma"], 0.25))
CASES = [("cat:dog", TEXT), ("[a:A-z:Z]", TEXT), ("(cat:dog|dog:cat)", TEXT), ("[a:b-y:zz:a]", TEXT),
         ("a:", TEXT), (":=", b"or\n"), ("a+:x", b"aaab\ncaaat\n"), ("cat:dog", b"cat"), ("b:X", b"ab\0cb\nb\n"),
         ("xo:", b"xor\n"), ("(cat)*:dog", b"catcatcat\n"), ("cat:dog", b""), ("[a:A-z:Z]", TEXT[:-1])]


def expected(expr, data, flavor):
    if have_ref():
        r = subprocess.run([REF[flavor], expr], input=data, capture_output=True)
        return r.returncode, r.stdout
    rc, out, _ = oracle_run(expr.encode(), data, flavor=flavor)
    return rc, out


@pytest.mark.parametrize("flavor", [NFT, DFT])
def test_cli_stdin_matches_reference(flavor):
    for expr, data in CASES:
        rc, want = expected(expr, data, flavor)
        r = subprocess.run([BIN[flavor], expr], input=data, capture_output=True)
        assert r.returncode == rc and r.stdout == want, (expr, flavor, r.stderr[:200])


def test_cli_file_argument(tmp_path):
    f = tmp_path / "in.txt"
    f.write_bytes(TEXT)
    rc, want = expected("cat:dog", TEXT, NFT)
    r = subprocess.run([BIN[NFT], "cat:dog", str(f)], capture_output=True)
    assert r.returncode == rc and r.stdout == want


@pytest.mark.parametrize("flavor", [NFT, DFT])
def test_cli_streams_inputs_larger_than_its_block_ring(flavor, monkeypatch):
    """Scan mode is a stream (trre_nft.c:776-790): with 1 MiB blocks a 12 MiB pipe crosses the 3-block ring four times;
    a 3 MiB line makes a block grow; the last record has no newline."""
    monkeypatch.setenv("TRRE_CLI_BLOCK_MB", "1")
    big = bytes(T.synth_text(0x72726507, 12 << 20, ["cat", "dog", "catdog"], 0.25))
    long_line = big[:5 << 20].replace(b"\n", b" ", 40000) + b"x" * (3 << 20) + b"\ncat dog\n" + big[:700000] + b"cat"
    for expr, data in (("(cat:dog|dog:cat)", big), ("[a:A-z:Z]", big), ("a:", big), ("cat:mouse", big[:4 << 20]),
                       ("cat:dog", long_line), ("a+:x", big[:3 << 20])):
        rc, want = expected(expr, data, flavor)
        r = subprocess.run([BIN[flavor], expr], input=data, capture_output=True)
        assert r.returncode == rc and r.stdout == want, (expr, flavor, len(r.stdout), len(want), r.stderr[:200])


def test_cli_prints_what_precedes_a_failing_block(monkeypatch):
    """A runtime error ("stack max capacity reached") ends the reference after it has printed the earlier lines; the
    streaming CLI has printed the earlier blocks."""
    monkeypatch.setenv("TRRE_CLI_BLOCK_MB", "1")
    head = bytes(T.synth_text(9, 2 << 20)).replace(b"a", b"e")
    data = head + b"a" * 70000 + b"\n"
    r = subprocess.run([BIN[NFT], "(a:x)*b|a"], input=data, capture_output=True)
    assert r.returncode == 1 and b"stack max capacity reached" in r.stderr
    assert len(r.stdout) >= 1 << 20 and head.startswith(r.stdout)


def test_cli_match_and_all_flags_run_on_the_gpu(monkeypatch):
    """-m / -ma / -a (trre_nft.c:728-741, 635-648, 782-797): same argv as the reference, GPU DFS underneath, streamed."""
    r = subprocess.run([BIN[NFT], "-ma", "(.:x)*.*"], input=b"aaa\n", capture_output=True)
    assert r.stdout == b"xxx\nxxa\nxaa\naaa\n"
    r = subprocess.run([BIN[NFT], "-a", "a:x|a:y"], input=b"ab\n", capture_output=True)
    assert r.stdout == b"xyab\n"                                  # -a in scan mode
    r = subprocess.run([BIN[DFT], "-a", "a:x"], input=b"ab\n", capture_output=True)
    assert r.returncode == 1 and r.stderr == b"Not supported yet\n"
    monkeypatch.setenv("TRRE_CLI_BLOCK_MB", "1")
    big = bytes(T.synth_text(0x72726509, 3 << 20, ["cat", "dog", "catdog"], 0.25))
    for flags, expr, flavor in ((["-m"], ".*cat.*", NFT), (["-a"], "(cat:dog|dog:cat|cat:CAT)", NFT), (["-m", "-a"], "[a-zA-Z0-9.,; ]*(cat:dog)[a-zA-Z0-9.,; ]*", NFT),
                                (["-m"], "cat:dog", DFT)):
        if have_ref():
            ref = subprocess.run([REF[flavor]] + flags + [expr], input=big, capture_output=True, timeout=120)
            rc, want = ref.returncode, ref.stdout
        else:
            from oracle_util import MATCH, SCAN
            rc, want, _ = oracle_run(expr.encode(), big, flavor=flavor, mode=MATCH if "-m" in flags else SCAN, all_="-a" in flags)
        r = subprocess.run([BIN[flavor]] + flags + [expr], input=big, capture_output=True)
        assert r.returncode == rc and r.stdout == want, (flags, expr, len(r.stdout), len(want), r.stderr[:200])


def test_multi_gpu_scan_equals_single(monkeypatch):
    torch = pytest.importorskip("torch")
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs 2 GPUs")
    data = bytes(T.synth_text(5, 8 << 20, ["cat", "dog"], 0.25))
    for expr, flavor in (("(cat:dog|dog:cat)", NFT), ("a:", NFT), ("[a:A-z:Z]", DFT), ("a+:x", NFT)):
        prog = T.compile(expr, flavor)
        assert prog.scan_multi(data, list(range(min(ndev, 4)))) == prog.scan(data), expr
    monkeypatch.setenv("TRRE_GPUS", "2")
    r = subprocess.run([BIN[NFT], "cat:dog"], input=data, capture_output=True)
    monkeypatch.delenv("TRRE_GPUS")
    assert r.returncode == 0 and r.stdout == T.compile("cat:dog", NFT).scan(data)
