"""-m gpu: every BASELINE config at its stated size on one B200, tile by tile against what the REAL reference binaries
print (tests/golden/config_hashes.json, made by tests/golden/make_config_hashes.py from oracle/_ref in the build
container - /root/reference does not exist on the GPU box).  Lines are independent (trre_nft.c:778-781), so the
reference's stdout for k replicated tiles is its stdout for one tile, k times: tile 0 of the device output is hashed
on the host and compared with the reference's hash, every further tile is compared with tile 0 on the device."""
import hashlib

import pytest

import bench
import trre_b200 as T
from golden_util import load_config_hashes, dictionary_slice

pytestmark = pytest.mark.gpu
CFG = load_config_hashes()
GiB = 1 << 30


def run_config(name, total_bytes, rank=0):
    torch = pytest.importorskip("torch")
    pattern, flavor, _, words, p_hit, lower, seed, refbin = bench.WORKLOADS[name]
    rec = CFG[name]["tiles"][rank]
    tile = T.synth_text(seed + rank, rec["tile_bytes"], words, p_hit, lower)
    assert hashlib.sha256(tile).hexdigest() == rec["in_sha256"]
    prog = T.compile(pattern, flavor)
    reps = max(1, total_bytes // len(tile))
    d_tile = torch.frombuffer(tile, dtype=torch.uint8).cuda()
    d_in = d_tile.repeat(reps)
    del d_tile
    d_out = torch.empty_like(d_in)
    n = prog.scan_device(d_in.data_ptr(), d_in.numel(), d_out.data_ptr(), d_out.numel(),
                         stream=torch.cuda.current_stream().cuda_stream)
    assert n == d_in.numel() == reps * rec["out_bytes"]
    first = bytes(d_out[:len(tile)].cpu().numpy().tobytes())
    assert hashlib.sha256(first).hexdigest() == rec["out_sha256"], (name, rank, "tile 0 differs from the reference's stdout")
    for k in range(1, reps):
        assert torch.equal(d_out[:len(tile)], d_out[k * len(tile):(k + 1) * len(tile)]), (name, k)
    return prog, first


def test_cfg1_cat_dog_1mb():
    prog, _ = run_config("cfg1", 1_000_000)
    assert prog.info["fast_path"] == T._lib.FAST_MATCH


def test_cfg2_uppercase_1gib():
    prog, _ = run_config("cfg2", 1 * GiB)
    assert prog.exec_kind == T.EXEC_BYTEMAP


def test_cfg3_cat_dog_swap_8gib():
    run_config("cfg3", 8 * GiB)


def test_cfg4_caesar_8gib_per_gpu_and_every_rank_tile():
    run_config("cfg4", 8 * GiB)
    for rank in range(1, 8):                   # the tiles the other 7 ranks of the 64 GiB run scan
        run_config("cfg4", 256 << 20, rank)


def test_cfg5_dictionary_10k_words_8gib_and_every_rank_tile():
    prog, first = run_config("cfg5", 8 * GiB)
    assert prog.info["dsa_states"] == 18402 and prog.info["nft_states"] == 100000
    n, xor = dictionary_slice(CFG)             # the committed 1 MiB of trre_dft's stdout itself
    tile = T.synth_text(bench.WORKLOADS["cfg5"][6], 2 << 20, bench._DICT, 0.30, True)
    assert first[:n] == bytes(a ^ b for a, b in zip(bytes(tile[:n]), xor))
    for rank in range(1, 8):
        run_config("cfg5", 256 << 20, rank)


def test_cfg5_dictionary_64gib_on_one_gpu():
    """BASELINE configs[4] at N = 1: 64 GiB in + 64 GiB out resident in 180 GB of HBM."""
    torch = pytest.importorskip("torch")
    if torch.cuda.get_device_properties(0).total_memory < 150 * GiB:
        pytest.skip("needs 128 GiB of device memory")
    run_config("cfg5", 64 * GiB)


def test_nft_flavour_of_the_dictionary_agrees():
    """The dictionary is prefix free, so `trre` and `trre_dft` print the same: the NFT-flavour program
    must reproduce trre_dft's hash as well."""
    torch = pytest.importorskip("torch")
    pattern, _, _, words, p_hit, lower, seed, _ = bench.WORKLOADS["cfg5"]
    rec = CFG["cfg5"]["tiles"][0]
    tile = T.synth_text(seed, rec["tile_bytes"], words, p_hit, lower)
    head = bytes(tile[:64 << 20])
    head = head[:head.rfind(b"\n") + 1]
    nft = T.compile(pattern, T.FLAVOR_NFT)
    assert nft.exec_kind == T.EXEC_TABLE and nft.info["fast_path"] == T._lib.FAST_MATCH
    assert nft.scan(head) == T.compile(pattern, T.FLAVOR_DFT).scan(head)
