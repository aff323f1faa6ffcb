#!/usr/bin/env python
"""Pins the BASELINE configs at their real tile size with the REAL reference binaries (oracle/_ref/trre, trre_dft,
compiled from /root/reference by oracle/Makefile): for every config and every rank seed bench.py uses, the synthetic
256 MiB tile (1 MB for cfg1) is generated exactly as bench.py generates it, piped through the reference, and the
SHA-256 of input and stdout are committed (tests/golden/config_hashes.json), plus - for the 10 000-word dictionary - a
1 MiB slice of the reference's stdout itself (stored as zlib(in XOR out), the output differs from the input only inside
matches).  /root/reference does not exist on the GPU box; these fixtures do.  Re-run (about 5 minutes on 8 cores):
    python tests/golden/make_config_hashes.py"""
import base64
import concurrent.futures
import hashlib
import json
import os
import subprocess
import sys
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

MiB = 1 << 20
PREFIXES = (1 * MiB, 16 * MiB)


def line_cut(buf, n):
    return buf.rfind(b"\n", 0, n) + 1


def one(job):
    name, rank = job
    import bench
    import trre_b200 as T
    pattern, flavor, nbytes, words, p_hit, lower, seed, refbin = bench.WORKLOADS[name]
    tile_n = min(bench.TILE, nbytes)
    tile = bytes(T.synth_text(seed + rank, tile_n, words, p_hit, lower))
    exe = os.path.join(ROOT, "oracle", "_ref", refbin)
    out = subprocess.run([exe, pattern], input=tile, stdout=subprocess.PIPE, check=True).stdout
    rec = {"rank": rank, "seed": seed + rank, "tile_bytes": tile_n, "in_sha256": hashlib.sha256(tile).hexdigest(),
           "out_bytes": len(out), "out_sha256": hashlib.sha256(out).hexdigest()}
    if rank == 0:
        rec["prefix"] = []
        for pn in PREFIXES:
            if pn >= tile_n:
                continue
            cut = line_cut(tile, pn)
            pout = subprocess.run([exe, pattern], input=tile[:cut], stdout=subprocess.PIPE, check=True).stdout
            rec["prefix"].append({"in_bytes": cut, "out_bytes": len(pout), "out_sha256": hashlib.sha256(pout).hexdigest()})
            if name == "cfg5" and pn == 1 * MiB:
                assert len(pout) == cut
                x = bytes(a ^ b for a, b in zip(tile[:cut], pout))
                rec["slice_xor_zlib_b64"] = base64.b64encode(zlib.compress(x, 9)).decode()
    return name, rec


def main():
    import bench
    jobs = [("cfg1", 0)] + [(c, r) for c in ("cfg2", "cfg3", "cfg4", "cfg5") for r in range(8)]
    res = {}
    with concurrent.futures.ProcessPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for name, rec in ex.map(one, jobs):
            res.setdefault(name, []).append(rec)
            print(name, rec["rank"], rec["out_sha256"][:16], file=sys.stderr)
    doc = {"generator": "tests/golden/make_config_hashes.py", "reference": "c0stya/trre @ c474a68 (oracle/_ref)",
           "configs": {}}
    for name, recs in res.items():
        pattern, flavor, nbytes, words, p_hit, lower, seed, refbin = bench.WORKLOADS[name]
        doc["configs"][name] = {"binary": refbin, "flavor": flavor, "pattern_sha256": hashlib.sha256(pattern.encode()).hexdigest(),
                                "pattern_bytes": len(pattern), "tiles": sorted(recs, key=lambda r: r["rank"])}
    with open(os.path.join(HERE, "config_hashes.json"), "w") as f:
        json.dump(doc, f, indent=1)
    print("wrote config_hashes.json", file=sys.stderr)


if __name__ == "__main__":
    main()
