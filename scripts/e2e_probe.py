"""End-to-end probe: trre_scan_into on pinned buffers for several chunk sizes, next to a bare chunked copy pipeline
(H2D chunk -> D2H chunk on rotating streams, no kernel) of the same shape.  usage: python scripts/e2e_probe.py [label]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import trre_b200 as T

n = 1 << 30
tile = T.synth_text(0x72726502, 256 << 20)
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
for r in range(4):
    h_in[r * len(tile):(r + 1) * len(tile)].copy_(torch.frombuffer(tile, dtype=torch.uint8))
h_out = torch.empty(n + 64, dtype=torch.uint8, pin_memory=True)
prog = T.compile("[a:A-z:Z]", T.FLAVOR_DFT)
label = sys.argv[1] if len(sys.argv) > 1 else ""
for mb in (8, 16, 32, 64, 128):
    os.environ["TRRE_CHUNK_MB"] = str(mb)
    prog.scan_into(h_in.data_ptr(), n, h_out.data_ptr(), h_out.numel())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        prog.scan_into(h_in.data_ptr(), n, h_out.data_ptr(), h_out.numel())
    torch.cuda.synchronize()
    print(label, "scan_into chunk %3d MiB: %.1f GB/s" % (mb, 3 * n / (time.perf_counter() - t0) / 1e9), flush=True)

# the same traffic pattern without any kernel: what the copy engines give a chunked, slot-limited pipeline
for slots in (3, 6):
    for mb in (32,):
        chunk = mb << 20
        d = [torch.empty(chunk, dtype=torch.uint8, device="cuda") for _ in range(slots)]
        st = [torch.cuda.Stream() for _ in range(slots)]
        def run():
            k = 0
            for off in range(0, n, chunk):
                s = k % slots
                with torch.cuda.stream(st[s]):
                    d[s].copy_(h_in[off:off + chunk], non_blocking=True)
                    h_out[off:off + chunk].copy_(d[s], non_blocking=True)
                k += 1
            torch.cuda.synchronize()
        run()
        t0 = time.perf_counter()
        for _ in range(3):
            run()
        print(label, "bare copies, %d slots, chunk %d MiB: %.1f GB/s" % (slots, mb, 3 * n / (time.perf_counter() - t0) / 1e9), flush=True)
