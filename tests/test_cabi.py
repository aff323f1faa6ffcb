"""The drop-in boundary without a GPU: the shared library loads, exports every symbol declared in
include/trre_b200.h, the host-only entry points work, and the scan entry points FAIL LOUDLY (no CPU
fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest

import trre_b200 as T
from oracle_util import NFT, DFT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "trre_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(trre_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    names = declared_functions()
    assert len(names) >= 17, names
    lib = ctypes.CDLL(T.lib_path())
    for n in names:
        assert hasattr(lib, n), "libtrre_b200.so does not export %s" % n


def test_no_oracle_or_reference_code_in_the_product_library():
    out = os.popen("ldd %s" % T.lib_path()).read()
    assert "oracle" not in out
    syms = os.popen("nm -D --defined-only %s" % T.lib_path()).read()
    assert "trre_oracle" not in syms


def test_shard_lines_cuts_after_newlines():
    data = bytes(T.synth_text(3, 100000))
    for parts in (1, 2, 3, 8, 64):
        cuts = T.shard_lines(data, parts)
        assert cuts[0] == 0 and cuts[-1] == len(data) and cuts == sorted(cuts)
        for c in cuts[1:-1]:
            assert c == 0 or c == len(data) or data[c - 1:c] == b"\n"
    assert T.shard_lines(b"abc", 4) == [0, 0, 3, 3, 3]          # one line cannot be split
    assert T.shard_lines(b"", 2) == [0, 0, 0]
    assert T.shard_lines(b"a\nb\nc\n", 3) == [0, 2, 4, 6]


def test_synth_text_is_deterministic_and_regular():
    a = T.synth_text(42, 50000, ["cat", "dog"], 0.25)
    b = T.synth_text(42, 50000, ["cat", "dog"], 0.25)
    assert a == b and a != T.synth_text(43, 50000, ["cat", "dog"], 0.25)
    assert a.endswith(b"\n") and b"\0" not in a and b"\r" not in a
    lines = bytes(a).split(b"\n")[:-2]
    assert all(40 <= len(l) <= 140 for l in lines), (min(map(len, lines)), max(map(len, lines)))
    assert b"cat" in a and all(32 <= c < 127 or c == 10 for c in a)
    low = T.synth_text(1, 10000, None, 0.0, True)
    assert all(c in b"abcdefghijklmnopqrstuvwxyz \n" for c in low)


def test_scan_fails_loudly_without_a_gpu():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = T.compile("cat:dog", NFT)
    with pytest.raises(T.TrreError) as ei:
        p.scan(b"cat\n")
    assert ei.value.code == 3 and "no CPU fallback" in str(ei.value)


def test_set_mode_is_host_state_of_the_program_not_of_the_blob():
    """trre_prog_set_mode (the CLI's -m / -a, trre_nft.c:728-741): accepted without a GPU; trre_dft -a is the reference's
    "Not supported yet" (trre_dft.c:1226); a program rebuilt from a broadcast blob starts in scan mode and takes the call."""
    p = T.compile("cat:dog", NFT)
    assert p.set_mode(match=True, all_=True) is p and p.set_mode() is p
    q = T.from_blob(p.blob())
    q.set_mode(match=False, all_=True)
    d = T.compile("cat:dog", DFT)
    d.set_mode(match=True)
    with pytest.raises(T.TrreError) as ei:
        d.set_mode(match=False, all_=True)
    assert ei.value.code == 2 and "Not supported yet" in ei.value.text


def gpu_present():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_cli_contract_without_touching_the_scan_path():
    import subprocess
    exe = os.path.join(ROOT, "bin", "trre")
    exe_dft = os.path.join(ROOT, "bin", "trre_dft")
    r = subprocess.run([exe], capture_output=True)
    assert r.returncode == 1 and r.stderr == b"error: missing trre expression\n"
    r = subprocess.run([exe, "(a"], input=b"", capture_output=True)
    assert r.returncode == 1 and r.stderr == b"error: unmached parenthesis\n"
    r = subprocess.run([exe, "a", "/nonexistent/file"], capture_output=True)
    assert r.returncode == 1 and r.stderr == b"error: can not open file /nonexistent/file\n"
    r = subprocess.run([exe_dft, "-a", "a"], input=b"", capture_output=True)
    assert r.returncode == 1 and r.stderr == b"Not supported yet\n"
    r = subprocess.run([exe, "-x", "a"], input=b"", capture_output=True)
    assert r.returncode == 1 and b"Usage:" in r.stderr and b"[-d] [-m] expr [file]" in r.stderr
    r = subprocess.run([exe, ""], input=b"", capture_output=True)
    assert r.returncode == -6 or r.returncode == 134           # assert() abort, trre_nft.c:285
    # -m / -a run on the GPU too (trre_prog_set_mode): without a device they fail loudly, there is no host fallback
    r = subprocess.run([exe, "-ma", ":a{,3}"], input=b"\n", capture_output=True)
    if gpu_present():
        assert r.returncode == 0 and r.stdout == b"aaa\naa\na\n\n"
    else:
        assert r.returncode == 1 and r.stdout == b"" and b"no CPU fallback" in r.stderr
    if gpu_present():
        r = subprocess.run([exe, "-m", "cat:dog"], input=b"cat\ncow\n", capture_output=True)
        assert r.stdout == b"dog\n"
        r = subprocess.run([exe_dft, "-m", "cat:dog"], input=b"cat\ncow\n", capture_output=True)
        assert r.stdout == b"\n\n"                                  # trre_dft.c:1287-1293
