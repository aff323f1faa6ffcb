"""Pins the CPU oracle (oracle/trre_oracle.c) to the reference:
  * the 69 known-answer vectors of the reference's test.sh,
  * the committed golden fixtures (outputs of the real reference binaries, tests/golden/),
  * when oracle/_ref is present (it is built from /root/reference by oracle/Makefile and travels to the
    GPU box), fresh seeded differential runs against the real binaries."""
import random
import subprocess

import pytest

from golden_util import load_golden
from kat_vectors import KATS, PROBES
from oracle_util import NFT, DFT, SCAN, MATCH, oracle_run, ref_run, have_ref
from patgen import rand_expr, rand_input

GOLDEN = load_golden()


def test_test_sh_vectors():
    assert len(KATS) == 69 and sum(1 for k in KATS if k[0] == "S") == 18
    for mode, inp, tre, exp in KATS:
        rc, out, err = oracle_run(tre.encode(), (inp + "\n").encode(), flavor=NFT,
                                  mode=SCAN if mode == "S" else MATCH, all_=(mode == "M"))
        assert rc == 0, (mode, inp, tre, err)
        # test.sh compares after $(...) which strips trailing newlines
        assert out.rstrip(b"\n") == exp.encode().rstrip(b"\n"), (mode, inp, tre, exp, out)


def test_golden_fixtures():
    assert len(GOLDEN) > 500
    for v in GOLDEN:
        rc, out, err = oracle_run(v["expr"], v["data"], flavor=v["flavor"], mode=v["mode"], all_=bool(v["all"]))
        assert rc == v["rc"], (v["tag"], v["expr"], rc, v["rc"])
        assert out == v["out"], (v["tag"], v["expr"], v["data"], out, v["out"])
        if rc == 1:
            assert err == v["err"], (v["expr"], err, v["err"])


def test_golden_mode_fixtures():
    """-m, -ma, scan-mode -a and `trre_dft -m` (stdout of the real binaries, tests/golden/ref_mode_vectors.json): the
    oracle, and the product's host restatement of those modes (trre_match_host)."""
    import trre_b200 as T
    modes = load_golden("ref_mode_vectors.json")
    assert len(modes) > 700
    for v in modes:
        rc, out, err = oracle_run(v["expr"], v["data"], flavor=v["flavor"], mode=v["mode"], all_=bool(v["all"]))
        assert rc == v["rc"] and out == v["out"], (v["tag"], v["expr"], v["mode"], v["all"], v["data"], out, v["out"])
        if v["rc"] == 0:
            try:
                prog = T.compile(v["expr"], v["flavor"])
            except T.TrreError as e:
                assert e.code == 2
                continue
            assert prog.match(v["data"], all_=bool(v["all"]), scan=v["mode"] == SCAN) == v["out"], (v["expr"], v["data"])


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_probe_log_against_reference_binaries():
    for expr, data in PROBES:
        for fl in (NFT, DFT):
            assert oracle_run(expr, data, flavor=fl)[:2] == ref_run(expr, data, flavor=fl)[:2], (expr, data, fl)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed", [101, 102, 103])
def test_differential_against_reference_binaries(seed):
    rng = random.Random(seed)
    checked = 0
    for _ in range(120):
        pat = rand_expr(rng, 2).encode()
        data = rand_input(rng, final_newline=rng.random() < 0.85)
        for fl in (NFT, DFT):
            try:
                ref = ref_run(pat, data, flavor=fl, timeout=5)
            except subprocess.TimeoutExpired:
                continue
            if ref[0] not in (0, 1):
                continue                        # the reference crashed (epsilon cycles): nothing to pin
            got = oracle_run(pat, data, flavor=fl)
            if got[0] == 2:
                continue
            assert got[:2] == ref[:2], (pat, data, fl, got, ref)
            checked += 1
    assert checked > 150


def test_empty_pattern_is_the_assert():
    assert oracle_run(b"", b"abc\n")[0] == 134
