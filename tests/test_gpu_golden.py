"""-m gpu: the committed golden vectors (stdout of the REAL reference binaries, tests/golden/) through the
CUDA path and the C ABI.  /root/reference does not exist on the GPU box; the fixtures do."""
import pytest

import trre_b200 as T
from golden_util import load_golden
from oracle_util import SCAN, MATCH

pytestmark = pytest.mark.gpu
GOLDEN = [v for v in load_golden() if v["mode"] == SCAN and not v["all"]]


def test_golden_scan_vectors_on_the_gpu():
    ran = refused = 0
    kinds = set()
    for v in GOLDEN:
        try:
            prog = T.compile(v["expr"], v["flavor"])
        except T.TrreError as e:
            if e.code == 1:
                assert v["rc"] == 1 and e.text == v["err"], (v["expr"], e.text, v["err"])
            else:
                assert e.code == 2 and v["rc"] != 0, (v["expr"], e.code, e.text)   # nothing the reference answers is refused
            continue
        if v["rc"] != 0:
            with pytest.raises(T.TrreError):            # e.g. "stack max capacity reached"
                prog.scan(v["data"])
            continue
        kinds.add(prog.exec_kind)
        got = prog.scan(v["data"])
        assert got == v["out"], (v["tag"], v["expr"], v["flavor"], v["data"][:60], got[:60], v["out"][:60], prog.info)
        ran += 1
    assert ran > 470 and refused == 0, (ran, refused)
    assert kinds == {T.EXEC_BYTEMAP, T.EXEC_TABLE, T.EXEC_BACKTRACK, T.EXEC_LAZY}


MODES = [v for v in load_golden() if v["mode"] != SCAN or v["all"]] + load_golden("ref_mode_vectors.json")


def test_golden_match_and_all_vectors_on_the_gpu():
    """`-m`, `-ma` and scan-mode `-a` (trre_nft.c:635-648, 782-797; test.sh's 51 M lines and the generated ones): the
    reference's DFS on the GPU (trre_prog_set_mode), and the host restatement of it as a cross-check."""
    ran = 0
    for v in MODES:
        try:
            prog = T.compile(v["expr"], v["flavor"])
        except T.TrreError as e:
            assert v["rc"] != 0 or e.code == 2, (v["expr"], e.code, e.text)
            continue
        try:
            prog.set_mode(match=v["mode"] == MATCH, all_=bool(v["all"]))
        except T.TrreError as e:
            assert v["flavor"] == 1 and v["all"] and "Not supported yet" in e.text
            continue
        if v["rc"] != 0:
            with pytest.raises(T.TrreError):
                prog.scan(v["data"])
            continue
        got = prog.scan(v["data"])
        assert got == v["out"], (v["tag"], v["expr"], v["flavor"], v["mode"], v["all"], v["data"][:60], got[:60], v["out"][:60])
        ran += 1
    assert ran > 800, ran
