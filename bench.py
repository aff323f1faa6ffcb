#!/usr/bin/env python
"""bench.py - headline measurement of the trre scan path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg1|cfg3|cfg4|cfg5] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic input: at N=1 the workload is
BASELINE.json configs[1] - '[a:A-z:Z]' uppercase over 1 GiB synthetic ASCII, DFT flavour (trre_dft
semantics).  N>1 (torchrun, one rank per GPU): the compiled program is broadcast from rank 0, every
rank scans its own line-aligned 1 GiB shard (weak scaling), output sizes are all-gathered.

  value    input GB/s, inputs resident in HBM, CUDA-event timed on the launching stream, max over ranks
  e2e      same metric through the C ABI with HOST (pinned) buffers: H2D + kernels + D2H inside the clock
  roofline dominant kernel vs the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the reference binary (oracle/_ref/trre_dft, 1 thread - it is single threaded) on a
           bounded sample of the same workload, same box

--impl reference times the reference's own CPU implementation (oracle/_ref) on the same config.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GiB = 1 << 30
MiB = 1 << 20
TILE = 256 * MiB

WORKLOADS = {
    # name: (pattern, flavor, bytes per GPU, hit words, p_hit, lowercase_only, seed, reference binary)
    "cfg1": ("cat:dog", 0, 1_000_000, ["cat", "dog", "concatenate", "dogma"], 0.25, False, 0x72726501, "trre"),
    "cfg2": ("[a:A-z:Z]", 1, 1 * GiB, None, 0.0, False, 0x72726502, "trre_dft"),
    "cfg3": ("(cat:dog|dog:cat)", 0, 8 * GiB, ["cat", "dog", "catdog", "dogcat", "ca", "do", "cot", "dig"], 0.25, False, 0x72726503, "trre"),
    "cfg4": ("[a:b-y:zz:a]", 1, 8 * GiB, None, 0.0, False, 0x72726504, "trre_dft"),
}


def dictionary_words(k=10000, seed=7):
    """k distinct 4-letter lowercase words (prefix free), seeded, shuffled."""
    import random
    rng = random.Random(seed)
    words = set()
    while len(words) < k:
        words.add("".join(rng.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(4)))
    words = sorted(words)
    rng.shuffle(words)
    return words


_DICT = dictionary_words()
WORKLOADS["cfg5"] = ("|".join("%s:%s" % (w, w.upper()) for w in _DICT), 1, 8 * GiB, _DICT, 0.30, True, 0x72726505, "trre_dft")
WORKLOADS["del"] = ("a:", 0, 1 * GiB, None, 0.0, False, 0x72726506, "trre")
WORKLOADS["num"] = ("[0-9]+:#", 0, 1 * GiB, None, 0.0, False, 0x72726508, "trre")
WORKLOADS["sub"] = ("cat:mouse|dog:d", 0, 1 * GiB, ["cat", "dog", "catdog", "dogcat", "ca", "do", "cot", "dig"], 0.25, False, 0x72726507, "trre")

DESCR = {
    "cfg1": "'cat:dog' over 1 MB ASCII (NFT flavour)",
    "cfg2": "'[a:A-z:Z]' uppercase over 1 GiB synthetic ASCII per GPU, DFT flavour (BASELINE configs[1])",
    "cfg3": "'(cat:dog|dog:cat)' over 8 GiB synthetic text per GPU, NFT flavour (BASELINE configs[2])",
    "cfg4": "Caesar '[a:b-y:zz:a]' over 8 GiB synthetic ASCII per GPU, DFT flavour (BASELINE configs[3], 64 GiB at 8 GPUs)",
    "cfg5": "10k-alternation word:WORD dictionary (18 402-state DFT) over 8 GiB synthetic lowercase text per GPU, DFT flavour (BASELINE configs[4])",
    "num": "'[0-9]+:#' (collapse digit runs: cyclic automaton, per-line kernels) over 1 GiB synthetic ASCII, NFT flavour",
    "sub": "'cat:mouse|dog:d' (matches print more / fewer bytes than they consume) over 1 GiB synthetic text, NFT flavour",
    "del": "'a:' (delete every a; NOT length preserving: count + scan + write path) over 1 GiB synthetic ASCII, NFT flavour",
}
METRIC = "input GB/s (scan mode)"


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy read+write)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region, in-process through NVML (the timed
    region lasts milliseconds: spawning nvidia-smi per sample would miss it)."""

    def __init__(self, index):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # NVML indexes physical GPUs; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                try:
                    phys = int(vis.split(",")[index])
                except Exception:
                    phys = index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _sample(self):
        nv = self.nv
        sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
        try:
            reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        self.rows.append((sm, reasons))

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception:
                pass
            self._stop.wait(0.001)

    def __enter__(self):
        if self.nv:
            try:
                self._sample()
            except Exception:
                pass
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join(timeout=2)
        if self.nv:
            try:
                self._sample()
            except Exception:
                pass

    def summary(self):
        if not self.nv or not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        sm = sorted(r[0] for r in self.rows)
        allr = 0
        for r in self.rows:
            allr |= r[1]
        return {"sm_mhz": float(sm[len(sm) // 2]), "sm_max_mhz": float(self.max_sm),
                "reasons": sorted(k for k, bit in names.items() if allr & bit), "samples": len(sm)}


def make_text(T, spec, nbytes):
    """One tile of the synthetic corpus, replicated to nbytes; ends on a line boundary."""
    pattern, flavor, _, words, p_hit, lower, seed, _ = spec
    tile_n = min(TILE, nbytes)
    tile = T.synth_text(seed, tile_n, words, p_hit, lower)
    return tile


def run_reference(args, spec):
    """--impl reference: the UNMODIFIED reference binary (oracle/_ref, built from /root/reference by
    oracle/Makefile), single threaded because it is, on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import trre_b200 as T
    pattern, flavor, nbytes, words, p_hit, lower, seed, refbin = spec
    exe = os.path.join(ROOT, "oracle", "_ref", refbin)
    kind = "reference"
    sample_n = min(nbytes, 64 * MiB)
    text = bytes(T.synth_text(seed, sample_n, words, p_hit, lower))
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    path = os.path.join(shm, "trre_bench_ref_%d.txt" % os.getpid())
    with open(path, "wb") as f:
        f.write(text)
    try:
        if os.access(exe, os.X_OK):
            def one():
                t0 = time.perf_counter()
                with open(path, "rb") as fin, open(os.devnull, "wb") as fout:
                    subprocess.run([exe, pattern], stdin=fin, stdout=fout, check=True)
                return time.perf_counter() - t0
        else:
            kind = "port"
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from oracle_util import oracle_run

            def one():
                t0 = time.perf_counter()
                oracle_run(pattern.encode(), text, flavor=flavor)
                return time.perf_counter() - t0
        for _ in range(args.warmup):
            one()
        times = [one() for _ in range(args.steps)]
    finally:
        os.unlink(path)
    total = sum(times)
    gbs = sample_n * args.steps / total / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": DESCR[args.workload], "pattern": pattern if len(pattern) < 200 else pattern[:60] + "...(%d bytes)" % len(pattern),
                   "flavor": "dft" if flavor else "nft", "sample_bytes_per_step": sample_n},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": 1, "kind": kind,
                         "sample": "%d MiB of the same synthetic corpus per step, file in /dev/shm, stdout to /dev/null; "
                                   "single threaded (the reference has no threads)" % (sample_n // MiB)},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def cpu_baseline(T, spec):
    """The reference binary on a bounded sample of the same workload (rank 0, N=1 only)."""
    pattern, flavor, nbytes, words, p_hit, lower, seed, refbin = spec
    exe = os.path.join(ROOT, "oracle", "_ref", refbin)
    # ~10 s of CPU work: trre_dft runs near 50 MB/s whatever the pattern, trre (NFT) 12-165 MB/s
    sample_n = min(nbytes, 256 * MiB if refbin == "trre_dft" else 128 * MiB)
    text = bytes(T.synth_text(seed, sample_n, words, p_hit, lower))
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    path = os.path.join(shm, "trre_bench_cpu_%d.txt" % os.getpid())
    with open(path, "wb") as f:
        f.write(text)
    try:
        if os.access(exe, os.X_OK):
            kind = "reference"
            t0 = time.perf_counter()
            with open(path, "rb") as fin, open(os.devnull, "wb") as fout:
                subprocess.run([exe, pattern], stdin=fin, stdout=fout, check=True)
            dt = time.perf_counter() - t0
        else:
            kind = "port"
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from oracle_util import oracle_run
            small = text[: 16 * MiB]
            sample_n = len(small)
            t0 = time.perf_counter()
            oracle_run(pattern.encode(), small, flavor=flavor)
            dt = time.perf_counter() - t0
    finally:
        os.unlink(path)
    return {"value": sample_n / dt / 1e9, "unit": "GB/s", "cores": 1, "kind": kind,
            "sample": "%d MiB of the same synthetic corpus, one run, file in /dev/shm, stdout to /dev/null; "
                      "oracle/_ref/%s is single threaded (host has %d cores)" % (sample_n // MiB, refbin, os.cpu_count() or 0)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--bytes", type=int, default=0, help="override bytes per GPU (for quick checks)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    spec = list(WORKLOADS[args.workload])
    if args.bytes:
        spec[2] = args.bytes
    spec = tuple(spec)
    if args.impl == "reference":
        return run_reference(args, spec)

    import torch
    import trre_b200 as T

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the trre_b200 scan path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    pattern, flavor, nbytes, words, p_hit, lower, seed, refbin = spec

    # ---- program: compiled once on rank 0, the flat tables broadcast to every rank -----------------
    if rank == 0:
        prog = T.compile(pattern, flavor)
        blob = prog.blob()
    if world > 1:
        size = torch.tensor([len(blob) if rank == 0 else 0], dtype=torch.int64, device=dev)
        dist.broadcast(size, 0)
        buf = torch.empty(int(size.item()), dtype=torch.uint8, device=dev)
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(blob), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        if rank != 0:
            prog = T.from_blob(bytes(buf.cpu().numpy().tobytes()))
    info = prog.info

    # ---- input: one synthetic tile per rank, replicated in HBM to the shard size ----------------------
    tile = make_text(T, (pattern, flavor, nbytes, words, p_hit, lower, seed + rank, refbin), nbytes)
    tile_n = len(tile)
    reps = max(1, nbytes // tile_n)
    n = tile_n * reps
    h_tile = torch.frombuffer(tile, dtype=torch.uint8)
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    d_tile = h_tile.to(dev)
    for r in range(reps):
        d_in[r * tile_n:(r + 1) * tile_n].copy_(d_tile)
    del d_tile
    out_cap = n + 64 if info["length_preserving"] else 2 * n + 64      # expanding programs: room for 2x
    d_out = torch.empty(out_cap, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    def step():
        prog.scan_device(d_in.data_ptr(), n, d_out.data_ptr(), d_out.numel(), device=local_rank, stream=sp, sync=False)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------------------
    for _ in range(args.warmup):
        step()
    out_n = prog.scan_status(device=local_rank, stream=sp)
    launches_per_step = prog.launches(device=local_rank)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
        torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    out_n = prog.scan_status(device=local_rank, stream=sp)
    barrier()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    sizes = torch.tensor([out_n], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gathered = [torch.zeros_like(sizes) for _ in range(world)]
        dist.all_gather(gathered, sizes)          # exclusive prefix of these = global output offsets
        sizes_all = [int(g.item()) for g in gathered]
    else:
        sizes_all = [out_n]
    ms_max = float(t.item())
    ms_per_step = ms_max / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e9

    # ---- calibration outside the clock: a plain device copy of the SAME size (the roofline denominator in
    # MEASURED_PEAKS.json was taken on 2 GiB buffers; fixed launch cost weighs more on 1 GiB)
    copy_gbs = None
    try:
        for _ in range(3):
            d_out[:n].copy_(d_in)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        for _ in range(10):
            d_out[:n].copy_(d_in)
        c1.record(stream)
        torch.cuda.synchronize()
        copy_gbs = 2.0 * n * 10 / (c0.elapsed_time(c1) * 1e-3) / 1e9
        step()                      # restore d_out
        prog.scan_status(device=local_rank, stream=sp)
    except Exception:
        pass

    # ---- verification outside the clock (size independent): compare with the byte map applied by torch ---
    verified = None
    if info["exec_kind"] == T.EXEC_BYTEMAP:
        lut = torch.frombuffer(bytearray(prog.bytemap()), dtype=torch.uint8).to(dev)
        chunk = 256 * MiB
        ok = out_n == n
        for off in range(0, n, chunk):
            want = lut[d_in[off:off + chunk].long()]
            ok = ok and bool(torch.equal(want, d_out[off:off + chunk][: want.numel()]))
            del want
        verified = ok
    elif info["length_preserving"]:
        # replicated tiles must give replicated outputs (lines are independent)
        ok = out_n == n
        for r in range(1, reps):
            ok = ok and bool(torch.equal(d_out[:tile_n], d_out[r * tile_n:(r + 1) * tile_n]))
        verified = ok

    # ---- N > 1: the final gather of the output shards to rank 0 over NVLink, timed on its own (it is
    # NVLink-bound, 7 x shard / ~770 GB/s into one GPU, and would hide the scan if it sat inside the clock)
    gather = None
    if world > 1:
        offs = [0]
        for sz in sizes_all[:-1]:
            offs.append(offs[-1] + sz)
        total_out = offs[-1] + sizes_all[-1]
        full = torch.empty(total_out if rank == 0 else 1, dtype=torch.uint8, device=dev)

        def do_gather():
            ops = []
            if rank == 0:
                full[:out_n].copy_(d_out[:out_n])
                for r in range(1, world):
                    ops.append(dist.P2POp(dist.irecv, full[offs[r]:offs[r] + sizes_all[r]], r))
            else:
                ops.append(dist.P2POp(dist.isend, d_out[:out_n], 0))
            for w in dist.batch_isend_irecv(ops):
                w.wait()

        do_gather()                                   # warm-up (NCCL channel setup)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        do_gather()
        g1.record(stream)
        torch.cuda.synchronize()
        gt = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
        dist.all_reduce(gt, op=dist.ReduceOp.MAX)
        gms = float(gt.item())
        moved = total_out - sizes_all[0]
        gather = {"ms": gms, "bytes_into_rank0": moved, "GB/s_into_rank0": moved / (gms * 1e-3) / 1e9,
                  "note": "NCCL point-to-point over NVLink 5 / NVSwitch, not part of value; measured peer-copy reference 770 GB/s"}
        if rank == 0 and info["fast_path"] == 1:
            lut_g = torch.frombuffer(bytearray(prog.bytemap()), dtype=torch.uint8).to(dev)
            gather["first_shard_verified"] = bool(torch.equal(full[:out_n], lut_g[d_in.long()][:out_n]))
            del lut_g
        del full

    # ---- end to end through the C ABI with host buffers (H2D + kernels + D2H inside the clock) -------------
    e2e = None
    if args.e2e_steps > 0:
        h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        for r in range(reps):
            h_in[r * tile_n:(r + 1) * tile_n].copy_(h_tile)
        h_out = torch.empty(out_cap, dtype=torch.uint8, pin_memory=True)
        prog.scan_into(h_in.data_ptr(), n, h_out.data_ptr(), h_out.numel(), device=local_rank)   # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            got = prog.scan_into(h_in.data_ptr(), n, h_out.data_ptr(), h_out.numel(), device=local_rank)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": world * n * args.e2e_steps / dt / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": world * n, "d2h_bytes_per_step": world * got, "steps": args.e2e_steps}
        del h_in, h_out

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_hbm_peak()
    algo_bytes = float(n + out_n)            # 1 B read per input byte + 1 B written per output byte
    achieved = algo_bytes / (ms_per_step * 1e-3) / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "traffic_%s.json" % args.workload)
    if os.path.exists(prof):
        try:
            with open(prof) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
    kernel = {0: "k_lines_s<count> + k_scan_groups + k_lines_s<write>" if info["exec_kind"] != T.EXEC_BACKTRACK
                 else "k_lines<dfs> count + k_scan_groups + k_lines<dfs> write",
              1: "k_bytemap", 2: "k_match", 3: "k_bmv_count + k_scan_groups + k_bmv_write",
              4: "k_match<count> + k_scan_groups + k_match<write>", 5: "k_lines<inplace>"}[info["fast_path"]]
    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": DESCR[args.workload], "pattern": pattern if len(pattern) < 200 else pattern[:60] + "...(%d bytes)" % len(pattern),
                   "flavor": "dft" if flavor else "nft", "bytes_per_gpu": n, "exec": kernel, "dsa_states": info["dsa_states"], "nft_states": info["nft_states"],
                   "l2": "input %d MiB per step exceeds the 126 MB L2; no flush needed" % (n // MiB),
                   "parallelism": "lines sharded over %d GPU(s), program blob broadcast, sizes all-gathered" % world},
        "clocks": clocks.summary(),
        "e2e": e2e,
        "gpu_launches": launches_per_step * args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": kernel, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": algo_bytes,
                     "input_read_frac_of_peak": value / world / peak,
                     "torch_copy_same_size_gbs": copy_gbs},
        "verified": verified,
        "output_bytes_per_rank": sizes_all,
        "gather": gather,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(T, spec)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
