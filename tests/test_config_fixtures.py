"""The BASELINE configs pinned at their real tile size: tests/golden/config_hashes.json holds SHA-256 of the stdout of
the REAL reference binaries (oracle/_ref) for every config's 256 MiB synthetic tile and every rank seed, and a 1 MiB
slice of trre_dft's stdout for the 10 000-word dictionary.  Here, without a GPU: the corpus generators reproduce the
hashed inputs, and the oracle restatement reproduces the hashed outputs on the 1 MiB / 16 MiB prefixes."""
import hashlib

import pytest

import bench
import trre_b200 as T
from golden_util import load_config_hashes, dictionary_slice
from oracle_util import oracle_run, oracle_synth_text

CFG = load_config_hashes()


def tile_of(name, rank, nbytes=None):
    pattern, flavor, total, words, p_hit, lower, seed, refbin = bench.WORKLOADS[name]
    n = min(bench.TILE, total) if nbytes is None else nbytes
    return bytes(T.synth_text(seed + rank, n, words, p_hit, lower))


def test_fixture_file_covers_every_config_and_rank():
    assert set(CFG) == {"cfg1", "cfg2", "cfg3", "cfg4", "cfg5"}
    for name in ("cfg2", "cfg3", "cfg4", "cfg5"):
        assert [t["rank"] for t in CFG[name]["tiles"]] == list(range(8))
        assert hashlib.sha256(bench.WORKLOADS[name][0].encode()).hexdigest() == CFG[name]["pattern_sha256"]
    assert CFG["cfg5"]["pattern_bytes"] > 90000 and CFG["cfg1"]["tiles"][0]["tile_bytes"] == 1_000_000


def test_corpus_generators_agree_and_reproduce_the_hashed_input():
    # the product's generator and the oracle library's second implementation, byte for byte
    words = ["cat", "dog", "catdog", "dogcat", "ca", "do", "cot", "dig"]
    for seed, n, w, p, low in ((1, 100000, None, 0.0, False), (2, 300001, words, 0.25, False), (3, 77777, words, 0.3, True),
                               (4, 1, None, 0.0, False), (5, 0, None, 0.0, False), (0x72726505, 200000, bench._DICT, 0.30, True)):
        assert bytes(T.synth_text(seed, n, w, p, low)) == bytes(oracle_synth_text(seed, n, w, p, low)), seed
    tile = tile_of("cfg1", 0)
    assert hashlib.sha256(tile).hexdigest() == CFG["cfg1"]["tiles"][0]["in_sha256"]
    tile = tile_of("cfg5", 3)
    assert hashlib.sha256(tile).hexdigest() == CFG["cfg5"]["tiles"][3]["in_sha256"]


def test_oracle_reproduces_reference_stdout_of_cfg1():
    tile = tile_of("cfg1", 0)
    rc, out, _ = oracle_run(b"cat:dog", tile, flavor=0)
    assert rc == 0 and hashlib.sha256(out).hexdigest() == CFG["cfg1"]["tiles"][0]["out_sha256"]


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg4", "cfg5"])
def test_oracle_reproduces_reference_stdout_on_the_prefixes(name):
    pattern, flavor = bench.WORKLOADS[name][0], bench.WORKLOADS[name][1]
    rec = CFG[name]["tiles"][0]
    tile = tile_of(name, 0, 17 << 20)             # the generator is a stream: a prefix of the tile is a prefix
    for pre in rec["prefix"]:
        data = tile[:pre["in_bytes"]]
        assert data.endswith(b"\n")
        rc, out, _ = oracle_run(pattern.encode(), data, flavor=flavor)
        assert rc == 0 and len(out) == pre["out_bytes"]
        assert hashlib.sha256(out).hexdigest() == pre["out_sha256"], (name, pre["in_bytes"])


def test_dictionary_slice_fixture_is_the_reference_stdout():
    n, xor = dictionary_slice(CFG)
    tile = tile_of("cfg5", 0, 2 << 20)[:n]
    want = bytes(a ^ b for a, b in zip(tile, xor))
    assert hashlib.sha256(want).hexdigest() == CFG["cfg5"]["tiles"][0]["prefix"][0]["out_sha256"]
    assert want != tile and want.lower() == tile                     # only matched words change, to upper case
    rc, out, _ = oracle_run(bench.WORKLOADS["cfg5"][0].encode(), tile, flavor=1)
    assert rc == 0 and out == want
