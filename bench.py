#!/usr/bin/env python
"""bench.py - headline measurement of the trre scan path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg1|cfg3|cfg4|cfg5|...] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic input: at N=1 the workload is
BASELINE.json configs[1] - '[a:A-z:Z]' uppercase over 1 GiB synthetic ASCII, DFT flavour (trre_dft
semantics).  N>1 (torchrun, one rank per GPU): the compiled program is broadcast from rank 0, every
rank scans its own line-aligned 1 GiB shard (weak scaling), output sizes are all-gathered.

  value    input GB/s, inputs resident in HBM, CUDA-event timed on the launching stream, max over ranks
  e2e      same metric through the C ABI with HOST (pinned) buffers: H2D + kernels + D2H inside the clock,
           next to the PCIe copy rates measured in the same process (its roofline)
  roofline dominant kernel vs the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the reference binary (oracle/_ref/trre_dft, 1 thread - it is single threaded) on a
           bounded sample of the same workload, same box
  others   measured after the headline, same process: the other BASELINE configs (1 GiB slices and their
           stated sizes), each verified against the SHA-256 of the real reference's stdout
           (tests/golden/config_hashes.json); at N>1 also Caesar 8 GiB per GPU and the 64 GiB dictionary split
           over the ranks (strong scaling) with the program broadcast and the size all-gather inside the clock

--impl reference times the reference's own CPU implementation (oracle/_ref) on the same config.
"""
import argparse
import hashlib
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
# the reference's DFS itself on the GPU (one record per thread): the executor of last resort, and the CLI's -a / -m modes
_CATS = ["cat", "dog", "catdog", "dogcat", "ca", "do", "cot", "dig"]
WORKLOADS["dfs"] = ("(a:x)*b|a", 0, 64 * MiB, None, 0.0, False, 0x72726509, "trre")
WORKLOADS["mode_a"] = ("(cat:dog|dog:cat|cat:CAT)", 0, 64 * MiB, _CATS, 0.25, False, 0x7272650A, "trre", (0, 1))
WORKLOADS["mode_m"] = ("[a-zA-Z0-9.,; ]*(cat:dog)[a-zA-Z0-9.,; ]*", 0, 16 * MiB, _CATS, 0.25, False, 0x7272650B, "trre", (1, 0))

DESCR = {
    "cfg1": "'cat:dog' over 1 MB ASCII (NFT flavour)",
    "cfg2": "'[a:A-z:Z]' uppercase over 1 GiB synthetic ASCII per GPU, DFT flavour (BASELINE configs[1])",
    "cfg3": "'(cat:dog|dog:cat)' over 8 GiB synthetic text per GPU, NFT flavour (BASELINE configs[2])",
    "cfg4": "Caesar '[a:b-y:zz:a]' over 8 GiB synthetic ASCII per GPU, DFT flavour (BASELINE configs[3], 64 GiB at 8 GPUs)",
    "cfg5": "10k-alternation word:WORD dictionary (18 402-state DFT) over 8 GiB synthetic lowercase text per GPU, DFT flavour (BASELINE configs[4])",
    "num": "'[0-9]+:#' (collapse digit runs: cyclic automaton, per-line kernels) over 1 GiB synthetic ASCII, NFT flavour",
    "sub": "'cat:mouse|dog:d' (matches print more / fewer bytes than they consume) over 1 GiB synthetic text, NFT flavour",
    "del": "'a:' (delete every a; NOT length preserving: count + scan + write path) over 1 GiB synthetic ASCII, NFT flavour",
    "dfs": "'(a:x)*b|a' (not subsequential: the reference's backtracking DFS on the GPU, one record per thread) over 64 MiB, NFT flavour",
    "mode_a": "`trre -a '(cat:dog|dog:cat|cat:CAT)'` (every accepting path at every position; DFS executor) over 64 MiB",
    "mode_m": "`trre -m '[a-zA-Z0-9.,; ]*(cat:dog)[a-zA-Z0-9.,; ]*'` (whole-record match; DFS executor) over 16 MiB",
}
METRIC = "input GB/s (scan mode)"
KERNELS = {0: "k_lines_s<count> + k_scan_groups + k_lines_s<write>", 1: "k_bytemap", 2: "k_match",
           3: "k_bmv_count + k_scan_groups + k_bmv_write", 4: "k_match<count> + k_scan_groups + k_match<write>",
           5: "k_lines<inplace>"}


def short_pattern(pattern):
    return pattern if len(pattern) < 200 else pattern[:60] + "...(%d bytes)" % len(pattern)


def config_of(name, spec, n_gpus):
    """The `config` object - the SAME in both arms (the driver compares them)."""
    pattern, flavor, nbytes = spec[0], spec[1], spec[2]
    return {"workload": DESCR[name], "pattern": short_pattern(pattern), "flavor": "dft" if flavor else "nft",
            "bytes_per_gpu": nbytes, "n_gpus": n_gpus}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy read+write)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def config_hashes():
    try:
        with open(os.path.join(ROOT, "tests", "golden", "config_hashes.json")) as f:
            return json.load(f)["configs"]
    except Exception:
        return {}


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
            self.h = pynvml.nvmlDeviceGetHandleByIndex(physical_index(index))
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


def physical_index(index):
    """NVML indexes physical GPUs; honour CUDA_VISIBLE_DEVICES when it is a plain index list."""
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[index])
        except Exception:
            pass
    return index


def bind_to_gpu_numa_node(index):
    """Pin this rank to the CPU cores next to its GPU (NVML affinity mask) BEFORE any pinned buffer is allocated:
    first touch then places the staging memory on the GPU's own NUMA node.  With 8 ranks pushing 1 GiB each way this is
    the difference between every copy crossing the socket link and none."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(physical_index(index))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"cpus": "%d-%d (%d)" % (min(cpus), max(cpus), len(cpus))}
    except Exception as e:
        return {"cpus": None, "error": str(e)[:80]}
    return {"cpus": None}


# ---------------------------------------------------------------------------------------------------------------------
# --impl reference
# ---------------------------------------------------------------------------------------------------------------------

def run_reference(args, name, spec):
    """--impl reference: the UNMODIFIED reference binary (oracle/_ref, built from /root/reference by
    oracle/Makefile), single threaded because it is, on a bounded sample per step.  The corpus comes from the oracle
    library's own generator: this arm never maps the product's library."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_util import oracle_run, oracle_synth_text
    pattern, flavor, nbytes, words, p_hit, lower, seed, refbin = spec[:8]
    exe = os.path.join(ROOT, "oracle", "_ref", refbin)
    kind = "reference"
    sample_n = min(nbytes, 64 * MiB)
    text = bytes(oracle_synth_text(seed, sample_n, words, p_hit, lower))
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
        "config": config_of(name, spec, args.gpus),
        "config_detail": {"sample_bytes_per_step": sample_n,
                          "note": "the reference is single threaded and its speed does not depend on the input size "
                                  "(~50 MB/s): every step times a %d MiB sample of the same corpus" % (sample_n // MiB)},
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
    pattern, flavor, nbytes, words, p_hit, lower, seed, refbin = spec[:8]
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


def cli_end_to_end(T, name, spec):
    """What a user of the drop-in feels: `bin/trre_dft EXPR FILE > /dev/null` on a file in /dev/shm, wall clock of the
    whole process (CUDA context creation, compilation, streaming read -> GPU -> write), next to the reference binary on a
    sample of the same file."""
    pattern, flavor, nbytes, words, p_hit, lower, seed, refbin = spec[:8]
    exe = os.path.join(ROOT, "bin", refbin)
    ref = os.path.join(ROOT, "oracle", "_ref", refbin)
    if not os.access(exe, os.X_OK):
        return None
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    path = os.path.join(shm, "trre_bench_cli_%d.txt" % os.getpid())
    tile = bytes(T.synth_text(seed, min(TILE, nbytes), words, p_hit, lower))
    reps = max(1, nbytes // len(tile))
    try:
        with open(path, "wb") as f:
            for _ in range(reps):
                f.write(tile)
        n = reps * len(tile)
        times = []
        for _ in range(2):
            t0 = time.perf_counter()
            with open(os.devnull, "wb") as fout:
                subprocess.run([exe, pattern, path], stdout=fout, check=True)
            times.append(time.perf_counter() - t0)
        # correctness of the streamed output: sha256 of the whole stdout against the reference's hash, tile by tile
        h = hashlib.sha256()
        p = subprocess.Popen([exe, pattern, path], stdout=subprocess.PIPE)
        first = p.stdout.read(len(tile))
        rest_ok = True
        for _ in range(1, reps):
            rest_ok = rest_ok and p.stdout.read(len(tile)) == first
        tail = p.stdout.read()
        p.wait()
        rec = [t for t in config_hashes().get(name, {}).get("tiles", []) if t["rank"] == 0 and t["tile_bytes"] == len(tile)]
        verified = None
        if rec and rec[0]["out_bytes"] == len(tile):
            verified = hashlib.sha256(first).hexdigest() == rec[0]["out_sha256"] and rest_ok and tail == b"" and p.returncode == 0
        out = {"seconds": min(times), "value": n / min(times) / 1e9, "unit": "GB/s", "bytes": n, "verified": verified,
               "command": "bin/%s EXPR FILE > /dev/null, file in /dev/shm, best of 2, whole process" % refbin}
        if os.access(ref, os.X_OK):
            sample = min(n, 64 * MiB)
            spath = path + ".sample"
            with open(spath, "wb") as f:
                f.write(tile[:sample])
            t0 = time.perf_counter()
            with open(os.devnull, "wb") as fout:
                subprocess.run([ref, pattern, spath], stdout=fout, check=True)
            dt = time.perf_counter() - t0
            os.unlink(spath)
            out["reference_gbs"] = sample / dt / 1e9
            out["reference_sample_bytes"] = sample
            out["speedup_vs_reference_cli"] = out["value"] / out["reference_gbs"]
        return out
    finally:
        if os.path.exists(path):
            os.unlink(path)


# ---------------------------------------------------------------------------------------------------------------------
# the B200 arm
# ---------------------------------------------------------------------------------------------------------------------

class Ctx:
    """torch / distributed handles of this rank"""

    def __init__(self, args):
        import torch
        import trre_b200 as T
        self.torch, self.T = torch, T
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device - the trre_b200 scan path has no CPU fallback")
        self.numa = bind_to_gpu_numa_node(self.local_rank)
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        self.stream = torch.cuda.current_stream()
        self.sp = self.stream.cuda_stream
        self.hashes = config_hashes()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(self, ok):
        t = self.torch.tensor([1 if ok else 0], dtype=self.torch.int32, device=self.dev)
        if self.dist:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())

    def broadcast_program(self, pattern, flavor):
        """Compiled once on rank 0, the flat tables broadcast to every rank (ONE collective)."""
        T, torch = self.T, self.torch
        prog = blob = None
        if self.rank == 0:
            prog = T.compile(pattern, flavor)
            blob = prog.blob()
        self.d_blob = None
        if self.dist:
            size = torch.tensor([len(blob) if self.rank == 0 else 0], dtype=torch.int64, device=self.dev)
            self.dist.broadcast(size, 0)
            buf = torch.empty(int(size.item()), dtype=torch.uint8, device=self.dev)
            if self.rank == 0:
                buf.copy_(torch.frombuffer(bytearray(blob), dtype=torch.uint8))
            self.dist.broadcast(buf, 0)
            if self.rank != 0:
                prog = T.from_blob(bytes(buf.cpu().numpy().tobytes()))
            self.d_blob = buf
        return prog


def measure(ctx, name, nbytes, steps, warmup, collectives=False, tile_rank=None, clocks=False):
    """One workload, device resident: build the input from the rank's synthetic tile, time `steps` scans with CUDA
    events on the launching stream, verify against the reference's hash.  Returns a dict; buffers are released."""
    T, torch = ctx.T, ctx.torch
    pattern, flavor, _, words, p_hit, lower, seed, refbin = WORKLOADS[name][:8]
    mode = WORKLOADS[name][8] if len(WORKLOADS[name]) > 8 else None
    prog = ctx.broadcast_program(pattern, flavor)
    if mode:
        prog.set_mode(match=bool(mode[0]), all_=bool(mode[1]))     # (not part of the blob: every rank sets it)
    info = dict(prog.info, mode=mode)
    rank_tile = ctx.rank if tile_rank is None else tile_rank
    tile_n = min(TILE, nbytes)
    tile = T.synth_text(seed + rank_tile, tile_n, words, p_hit, lower)
    reps = max(1, nbytes // tile_n)
    n = tile_n * reps
    d_tile = torch.frombuffer(tile, dtype=torch.uint8).to(ctx.dev)
    d_in = d_tile.repeat(reps) if reps > 1 else d_tile
    if reps > 1:
        del d_tile
    out_cap = n + 64 if info["length_preserving"] and not mode else 2 * n + 64      # expanding programs: room for 2x
    d_out = torch.empty(out_cap, dtype=torch.uint8, device=ctx.dev)
    sizes = torch.zeros(1, dtype=torch.int64, device=ctx.dev)
    gathered = [torch.zeros_like(sizes) for _ in range(ctx.world)]

    def step():
        if collectives and ctx.dist:
            ctx.dist.broadcast(ctx.d_blob, 0)                      # the compiled tables, rank 0 -> all
        prog.scan_device(d_in.data_ptr(), n, d_out.data_ptr(), d_out.numel(), device=ctx.local_rank, stream=ctx.sp, sync=False)
        if collectives and ctx.dist:
            ctx.dist.all_gather(gathered, sizes)                   # output sizes -> global offsets

    for _ in range(warmup):
        step()
    out_n = prog.scan_status(device=ctx.local_rank, stream=ctx.sp)
    sizes.fill_(out_n)
    launches_per_step = prog.launches(device=ctx.local_rank)
    ctx.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(ctx.local_rank) if clocks else None
    if sampler:
        sampler.__enter__()
    ev0.record(ctx.stream)
    for _ in range(steps):
        step()
    ev1.record(ctx.stream)
    torch.cuda.synchronize()
    if sampler:
        sampler.__exit__()
    ms = ev0.elapsed_time(ev1)
    out_n = prog.scan_status(device=ctx.local_rank, stream=ctx.sp)
    ctx.barrier()
    ms_per_step = ctx.max_over_ranks(ms) / steps

    coll = None
    if collectives and ctx.dist:
        # the two collectives on their own, same barrier-to-barrier discipline
        def timed(fn, reps_=10):
            ctx.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(ctx.stream)
            for _ in range(reps_):
                fn()
            b.record(ctx.stream)
            torch.cuda.synchronize()
            return ctx.max_over_ranks(a.elapsed_time(b)) / reps_
        coll = {"bcast_ms": timed(lambda: ctx.dist.broadcast(ctx.d_blob, 0)), "bcast_bytes": int(ctx.d_blob.numel()),
                "allgather_ms": timed(lambda: ctx.dist.all_gather(gathered, sizes)), "inside_the_clock": True}

    # ---- verification outside the clock: tile 0 against the SHA-256 of the REAL reference's stdout, the other tiles
    # against tile 0 on the device (lines are independent => replicated tiles give replicated outputs)
    verified, how = None, None
    rec = None
    for t in ctx.hashes.get(name, {}).get("tiles", []):
        if t["rank"] == rank_tile and t["tile_bytes"] == tile_n:
            rec = t
    if rec is not None:
        first = bytes(d_out[:rec["out_bytes"]].cpu().numpy().tobytes())
        ok = hashlib.sha256(first).hexdigest() == rec["out_sha256"] and out_n == reps * rec["out_bytes"]
        for r in range(1, reps):
            ok = ok and bool(torch.equal(d_out[:rec["out_bytes"]], d_out[r * rec["out_bytes"]:(r + 1) * rec["out_bytes"]]))
        verified, how = ok, "sha256 of the reference's stdout (tests/golden/config_hashes.json) for tile 0, tiles 1..%d equal tile 0" % (reps - 1)
    elif info["exec_kind"] == T.EXEC_BYTEMAP:
        lut = torch.frombuffer(bytearray(prog.bytemap()), dtype=torch.uint8).to(ctx.dev)
        ok = out_n == n
        for off in range(0, n, 256 * MiB):
            want = lut[d_in[off:off + 256 * MiB].long()]
            ok = ok and bool(torch.equal(want, d_out[off:off + 256 * MiB][: want.numel()]))
            del want
        verified, how = ok, "byte map applied by torch"
    elif info["length_preserving"]:
        ok = out_n == n
        for r in range(1, reps):
            ok = ok and bool(torch.equal(d_out[:tile_n], d_out[r * tile_n:(r + 1) * tile_n]))
        verified, how = ok, "replicated tiles equal (no reference hash for this workload)"
    verified_all = ctx.all_true(verified is not False)
    res = {"prog": prog, "info": info, "n": n, "out_n": out_n, "ms_per_step": ms_per_step, "launches_per_step": launches_per_step,
           "verified": (verified_all if verified is not None else None), "verified_how": how, "collectives": coll,
           "clocks": sampler.summary() if sampler else None, "d_in": d_in, "d_out": d_out, "tile": tile, "tile_n": tile_n, "reps": reps}
    return res


def others_plan(world):
    """(label, workload, bytes per GPU, collectives inside the clock) of the `others` block, in order."""
    plan = [("cfg1", 1_000_000, False), ("cfg1_1GiB", 1 * GiB, False), ("cfg3_1GiB", 1 * GiB, False), ("cfg4_1GiB", 1 * GiB, False),
            ("cfg5_1GiB", 1 * GiB, False)]
    if world == 1:
        plan += [("cfg3", 8 * GiB, False), ("cfg5_64GiB", 64 * GiB, False)]
        # off-BASELINE programs: what the other kernel families and the DFS executor do (sizes: WORKLOADS)
        plan += [(w, WORKLOADS[w][2], False) for w in ("del", "num", "sub", "dfs", "mode_a", "mode_m")]
    else:
        plan += [("cfg4", 8 * GiB, True), ("cfg5_64GiB_split", 64 * GiB // world, True)]
    return [(label, label if label in WORKLOADS else label.split("_")[0], per_gpu, coll) for label, per_gpu, coll in plan]


def kernel_label(T, info):
    if info.get("mode"):
        flags = ("-m" if info["mode"][0] else "") + ("-a" if info["mode"][1] else "")
        return "k_lines<dfs, %s> count + k_scan_groups + k_lines<dfs, %s> write" % (flags, flags)
    if info["exec_kind"] == T.EXEC_BACKTRACK:
        return "k_lines<dfs> count + k_scan_groups + k_lines<dfs> write"
    if info["exec_kind"] == T.EXEC_LAZY:
        return "k_lines<lazy> count (repeated while rows are missing) + k_scan_groups + k_lines<lazy> write"
    if info["exec_kind"] == T.EXEC_BYTEMAP:
        return "k_bytemap_swar" if os.environ.get("TRRE_NO_SWAR") is None else "k_bytemap"
    label = KERNELS[info["fast_path"]]
    if info.get("literal_words") and info["fast_path"] == 2 and os.environ.get("TRRE_NO_LIT") is None:
        return "k_lit"
    if info.get("literal_words") and info["fast_path"] == 4 and os.environ.get("TRRE_NO_LIT") is None:
        return "k_litv_count + k_scan_groups + k_bmv_write<lit>"
    if info["fast_path"] == 3 and info.get("run_classes"):
        return "k_bmv_count_run + k_scan_groups + k_bmv_write<run>"
    if info.get("gram_tables") and os.environ.get("TRRE_NO_GRAM") is None:
        label = label.replace("k_match", "k_gram_var") if info["fast_path"] == 4 else label.replace("k_match", "k_gram")
    return label


def summarize(ctx, name, r, total_bytes):
    peak, _ = measured_hbm_peak()
    secs = r["ms_per_step"] * 1e-3
    algo = float(r["n"] + r["out_n"])
    return {"workload": name, "pattern": short_pattern(WORKLOADS[name][0]), "bytes_total": total_bytes, "bytes_per_gpu": r["n"],
            "value": total_bytes / secs / 1e9, "unit": "GB/s", "ms_per_step": r["ms_per_step"],
            "roofline_frac": algo / secs / 1e9 / peak, "exec": kernel_label(ctx.T, r["info"]),
            "launches_per_step": r["launches_per_step"], "verified": r["verified"], "verified_how": r["verified_how"]}


def pcie_rates(ctx, n):
    """Pinned cudaMemcpyAsync of the same size, same process: each direction alone and both at once (the e2e path
    moves every byte once in each direction, concurrently: the bidirectional figure is its roofline)."""
    torch = ctx.torch
    n = min(n, 1 * GiB)
    h_a = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d_a = torch.empty(n, dtype=torch.uint8, device=ctx.dev)
    d_b = torch.empty(n, dtype=torch.uint8, device=ctx.dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d, d2h, reps=3):
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_a.copy_(h_a, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_b.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
        dt = ctx.max_over_ranks(time.perf_counter() - t0)
        return ctx.world * n * reps / dt / 1e9
    run(True, True, 1)
    res = {"h2d_gbs": run(True, False), "d2h_gbs": run(False, True), "bidir_gbs_per_direction": run(True, True),
           "bytes": n, "note": "aggregate over %d GPU(s), pinned host memory, same process" % ctx.world}
    del h_a, h_b, d_a, d_b
    return res


def end_to_end(ctx, prog, tile, tile_n, reps, n, out_cap, e2e_steps):
    """Through the C ABI with HOST buffers (H2D + kernels + D2H inside the clock)."""
    torch = ctx.torch
    h_tile = torch.frombuffer(tile, dtype=torch.uint8)
    h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    for r in range(reps):
        h_in[r * tile_n:(r + 1) * tile_n].copy_(h_tile)
    h_out = torch.empty(out_cap, dtype=torch.uint8, pin_memory=True)
    prog.scan_into(h_in.data_ptr(), n, h_out.data_ptr(), h_out.numel(), device=ctx.local_rank)   # warm-up
    sweep = {}
    best_mb, best = None, 0.0
    for mb in (8, 16, 32, 64):
        os.environ["TRRE_CHUNK_MB"] = str(mb)
        prog.scan_into(h_in.data_ptr(), n, h_out.data_ptr(), h_out.numel(), device=ctx.local_rank)
        ctx.barrier()
        t0 = time.perf_counter()
        prog.scan_into(h_in.data_ptr(), n, h_out.data_ptr(), h_out.numel(), device=ctx.local_rank)
        torch.cuda.synchronize()
        v = ctx.world * n / ctx.max_over_ranks(time.perf_counter() - t0) / 1e9
        sweep[str(mb)] = round(v, 2)
        if v > best:
            best, best_mb = v, mb
    os.environ["TRRE_CHUNK_MB"] = str(best_mb)
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        got = prog.scan_into(h_in.data_ptr(), n, h_out.data_ptr(), h_out.numel(), device=ctx.local_rank)
    torch.cuda.synchronize()
    dt = ctx.max_over_ranks(time.perf_counter() - t0)
    first_ok = None
    rec = [t for t in ctx.hashes.get("cfg2", {}).get("tiles", []) if t["rank"] == ctx.rank]
    del h_in
    res = {"value": ctx.world * n * e2e_steps / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": ctx.world * n,
           "d2h_bytes_per_step": ctx.world * got, "steps": e2e_steps, "chunk_mb": best_mb, "chunk_sweep_gbs": sweep,
           "slots": 3, "numa": ctx.numa}
    res["_h_out"] = h_out
    return res


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
    ap.add_argument("--no-others", action="store_true", help="skip the other BASELINE configs after the headline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    name = args.workload
    spec = list(WORKLOADS[name])
    if args.bytes:
        spec[2] = args.bytes
    spec = tuple(spec)
    if args.impl == "reference":
        return run_reference(args, name, spec)

    ctx = Ctx(args)
    T, torch = ctx.T, ctx.torch
    world, rank = ctx.world, ctx.rank
    pattern, flavor, nbytes, words, p_hit, lower, seed, refbin = spec[:8]

    # ---- headline: device-resident timing ------------------------------------------------------------------------------
    r = measure(ctx, name, nbytes, args.steps, args.warmup, clocks=True)
    prog, info, n, out_n, ms_per_step = r["prog"], r["info"], r["n"], r["out_n"], r["ms_per_step"]
    value = world * n / (ms_per_step * 1e-3) / 1e9
    sizes_t = torch.tensor([out_n], dtype=torch.int64, device=ctx.dev)
    if world > 1:
        gathered = [torch.zeros_like(sizes_t) for _ in range(world)]
        ctx.dist.all_gather(gathered, sizes_t)          # exclusive prefix of these = global output offsets
        sizes_all = [int(g.item()) for g in gathered]
    else:
        sizes_all = [out_n]
    d_in, d_out = r["d_in"], r["d_out"]

    # ---- calibration outside the clock: a plain device copy of the SAME size (the roofline denominator in
    # MEASURED_PEAKS.json was taken on 2 GiB buffers; fixed launch cost weighs more on 1 GiB)
    copy_gbs = sm_copy_gbs = None
    try:
        for _ in range(3):
            d_out[:n].copy_(d_in)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(ctx.stream)
        for _ in range(10):
            d_out[:n].copy_(d_in)
        c1.record(ctx.stream)
        torch.cuda.synchronize()
        copy_gbs = 2.0 * n * 10 / (c0.elapsed_time(c1) * 1e-3) / 1e9
        # the copy above is a DMA-engine memcpy; what an SM streaming kernel gets on the same buffers (torch's vectorised
        # elementwise kernel on 8-byte words) calibrates how much of the remaining gap is the SM load/store path itself
        v_in, v_out = d_in[:n - n % 8].view(torch.int64), d_out[:n - n % 8].view(torch.int64)
        for _ in range(3):
            torch.bitwise_and(v_in, 0x5F5F5F5F5F5F5F5F, out=v_out)
        torch.cuda.synchronize()
        c0.record(ctx.stream)
        for _ in range(10):
            torch.bitwise_and(v_in, 0x5F5F5F5F5F5F5F5F, out=v_out)
        c1.record(ctx.stream)
        torch.cuda.synchronize()
        sm_copy_gbs = 2.0 * (n - n % 8) * 10 / (c0.elapsed_time(c1) * 1e-3) / 1e9
        prog.scan_device(d_in.data_ptr(), n, d_out.data_ptr(), d_out.numel(), device=ctx.local_rank, stream=ctx.sp, sync=True)   # restore d_out
    except Exception:
        pass

    # ---- N > 1: the final gather of the output shards to rank 0 over NVLink, timed on its own (it is
    # NVLink-bound, 7 x shard / ~770 GB/s into one GPU, and would hide the scan if it sat inside the clock)
    gather = None
    if world > 1:
        dist = ctx.dist
        offs = [0]
        for sz in sizes_all[:-1]:
            offs.append(offs[-1] + sz)
        total_out = offs[-1] + sizes_all[-1]
        full = torch.empty(total_out if rank == 0 else 1, dtype=torch.uint8, device=ctx.dev)

        def do_gather():
            ops = []
            if rank == 0:
                full[:out_n].copy_(d_out[:out_n])
                for q in range(1, world):
                    ops.append(dist.P2POp(dist.irecv, full[offs[q]:offs[q] + sizes_all[q]], q))
            else:
                ops.append(dist.P2POp(dist.isend, d_out[:out_n], 0))
            for w in dist.batch_isend_irecv(ops):
                w.wait()

        do_gather()                                   # warm-up (NCCL channel setup)
        ctx.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(ctx.stream)
        do_gather()
        g1.record(ctx.stream)
        torch.cuda.synchronize()
        gms = ctx.max_over_ranks(g0.elapsed_time(g1))
        moved = total_out - sizes_all[0]
        gather = {"ms": gms, "bytes_into_rank0": moved, "GB/s_into_rank0": moved / (gms * 1e-3) / 1e9,
                  "note": "NCCL point-to-point over NVLink 5 / NVSwitch, not part of value; measured peer-copy reference 770 GB/s"}
        if rank == 0:
            # every shard of the gathered buffer against the reference's hash of that rank's tile
            ok = True
            for q in range(world):
                recs = [t for t in ctx.hashes.get(name, {}).get("tiles", []) if t["rank"] == q and t["tile_bytes"] == r["tile_n"]]
                if not recs:
                    ok = None
                    break
                shard = bytes(full[offs[q]:offs[q] + recs[0]["out_bytes"]].cpu().numpy().tobytes())
                ok = ok and hashlib.sha256(shard).hexdigest() == recs[0]["out_sha256"]
            gather["every_shard_matches_reference_sha256"] = ok
        del full

    # ---- end to end through the C ABI with host buffers (H2D + kernels + D2H inside the clock) -------------
    e2e = None
    if args.e2e_steps > 0:
        out_cap = d_out.numel()
        del d_in, d_out
        r["d_in"] = r["d_out"] = None
        torch.cuda.empty_cache()
        e2e = end_to_end(ctx, prog, r["tile"], r["tile_n"], r["reps"], n, out_cap, args.e2e_steps)
        h_out = e2e.pop("_h_out")
        recs = [t for t in ctx.hashes.get(name, {}).get("tiles", []) if t["rank"] == rank and t["tile_bytes"] == r["tile_n"]]
        if recs:
            ok = hashlib.sha256(bytes(h_out[:recs[0]["out_bytes"]].numpy().tobytes())).hexdigest() == recs[0]["out_sha256"]
            e2e["verified"] = ctx.all_true(ok)
        del h_out
        pc = pcie_rates(ctx, n)
        e2e["pcie"] = pc
        e2e["frac_of_pcie"] = e2e["value"] / pc["bidir_gbs_per_direction"] if pc["bidir_gbs_per_direction"] else None
        e2e["frac_of_pcie_note"] = "e2e input GB/s / measured bidirectional pinned-copy GB/s per direction (every byte crosses once each way)"
    else:
        del d_in, d_out
        r["d_in"] = r["d_out"] = None
    r["tile"] = None
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs, same process, after the headline ---------------------------------------------------
    others = None
    if not args.no_others and name == "cfg2" and not args.bytes:
        others = {}
        free = torch.cuda.mem_get_info()[0]
        for label, wl, per_gpu, coll in others_plan(world):
            if 2 * per_gpu + (1 << 30) > free:
                others[label] = {"skipped": "needs %d GiB of HBM" % ((2 * per_gpu) >> 30)}
                continue
            try:
                steps = 10 if per_gpu <= 8 * GiB else 3
                m = measure(ctx, wl, per_gpu, steps, 3, collectives=coll)
                s = summarize(ctx, wl, m, world * m["n"])
                if coll:
                    s["collectives"] = m["collectives"]
                    s["scaling"] = "strong (64 GiB split over the ranks)" if "split" in label else "weak"
                others[label] = s
                del m
            except Exception as e:                      # a failing extra must not take the headline with it
                others[label] = {"error": str(e)[:200]}
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            ctx.dist.destroy_process_group()
        return 0

    peak, peak_src = measured_hbm_peak()
    algo_bytes = float(n + out_n)            # 1 B read per input byte + 1 B written per output byte
    achieved = algo_bytes / (ms_per_step * 1e-3) / 1e9
    traffic, traffic_src = None, None
    prof = os.path.join(ROOT, "profiles", "traffic_%s.json" % name)
    if os.path.exists(prof):
        try:
            with open(prof) as f:
                tj = json.load(f)
            traffic = tj.get("dram_bytes_per_launch")
            traffic_src = tj.get("source", "ncu --set full capture committed under profiles/ (not measured in this run)")
            cap_n = tj.get("bytes_per_launch_input")
            if traffic and cap_n and cap_n != n:                  # captured on another size: the traffic is linear in it
                traffic = traffic * n / cap_n
                traffic_src += " - scaled from a %d MiB capture" % (cap_n // MiB)
        except Exception:
            pass
    kernel = kernel_label(T, info)
    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": config_of(name, (pattern, flavor, n), world),
        "config_detail": {"exec": kernel, "dsa_states": info["dsa_states"], "nft_states": info["nft_states"],
                          "l2": "input %d MiB per step exceeds the 126 MB L2; no flush needed" % (n // MiB),
                          "parallelism": "lines sharded over %d GPU(s), program blob broadcast, sizes all-gathered" % world},
        "clocks": r["clocks"],
        "e2e": e2e,
        "gpu_launches": r["launches_per_step"] * args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "kernel": kernel, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": algo_bytes,
                     "input_read_frac_of_peak": value / world / peak,
                     "torch_copy_same_size_gbs": copy_gbs, "torch_elementwise_kernel_same_size_gbs": sm_copy_gbs},
        "verified": r["verified"], "verified_how": r["verified_how"],
        "output_bytes_per_rank": sizes_all,
        "gather": gather,
        "others": others,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(T, spec)
        try:
            line["cli_e2e"] = cli_end_to_end(T, name, spec)
        except Exception as e:
            line["cli_e2e"] = {"error": str(e)[:200]}
    print(json.dumps(line))
    if world > 1:
        ctx.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
