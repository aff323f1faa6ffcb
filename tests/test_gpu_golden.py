"""-m gpu: the committed golden vectors (stdout of the REAL reference binaries, tests/golden/) through the
CUDA path and the C ABI.  /root/reference does not exist on the GPU box; the fixtures do."""
import pytest

import trre_b200 as T
from golden_util import load_golden
from oracle_util import SCAN

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
