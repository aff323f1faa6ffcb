"""The N>1 host logic on CPU (gloo, world_size 2): rank 0 compiles and BROADCASTS the program blob,
every rank takes its line-aligned shard, the output sizes are all-gathered and the shards, joined in
rank order, equal the single-process result.  The scan itself is run by the test-only table
interpreter here (no GPU in this container); on the GPU box the same flow runs in bench.py --gpus N."""
import os
import socket
import sys

import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, expr, flavor, data, q, mode=None):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import trre_b200 as T
    from dsa_sim import Blob
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if rank == 0:
            blob = T.compile(expr, flavor).blob()
            size = torch.tensor([len(blob)], dtype=torch.int64)
        else:
            size = torch.zeros(1, dtype=torch.int64)
        dist.broadcast(size, 0)
        buf = torch.empty(int(size.item()), dtype=torch.uint8)
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(blob), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        prog = T.from_blob(bytes(buf.numpy().tobytes()))            # no recompilation on rank > 0
        cuts = T.shard_lines(data, world)
        shard = data[cuts[rank]:cuts[rank + 1]]
        if mode is None:
            out = Blob(prog.blob()).scan(shard)
        else:
            # -m / -a are not part of the blob: every rank sets the mode on its own copy (on the GPU box the scan then runs
            # the DFS executor; here the host restatement of the same modes stands in for it)
            prog.set_mode(match=mode[0], all_=mode[1])
            out = prog.match(shard, all_=mode[1], scan=not mode[0])
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([len(out)], dtype=torch.int64))
        offsets = [0]
        for s in sizes[:-1]:
            offsets.append(offsets[-1] + int(s.item()))
        q.put((rank, offsets[rank], out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("expr,flavor", [("(cat:dog|dog:cat)", 0), ("a:", 0), ("[a:A-z:Z]", 1)])
def test_two_ranks_equal_one(expr, flavor):
    sys.path.insert(0, ROOT)
    import trre_b200 as T
    from oracle_util import oracle_run
    data = bytes(T.synth_text(99, 20000, ["cat", "dog", "catdog"], 0.3))
    rc, want, _ = oracle_run(expr.encode(), data, flavor=flavor)
    assert rc == 0
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, expr, flavor, data, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    joined = bytearray(len(want))
    for rank, off, out in got:
        joined[off:off + len(out)] = out
    assert bytes(joined) == want


@pytest.mark.parametrize("expr,mode", [("(cat:dog|dog:cat|cat:CAT)", (False, True)), ("[a-z ]*(cat:dog)[a-z ]*", (True, False))])
def test_two_ranks_equal_one_in_the_other_modes(expr, mode):
    """-a / -m over line-aligned shards (records are independent in every mode, trre_nft.c:776-797)."""
    sys.path.insert(0, ROOT)
    import trre_b200 as T
    from oracle_util import oracle_run, MATCH, SCAN
    data = bytes(T.synth_text(98, 12000, ["cat", "dog", "catdog"], 0.3, True))
    rc, want, _ = oracle_run(expr.encode(), data, flavor=0, mode=MATCH if mode[0] else SCAN, all_=mode[1])
    assert rc == 0
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, expr, 0, data, q, mode)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    joined = bytearray(len(want))
    for rank, off, out in got:
        joined[off:off + len(out)] = out
    assert bytes(joined) == want
