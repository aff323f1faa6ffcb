// bmv_kernel.cuh - variable-length byte maps (PF_BYTEMAP_VAR) on regular buffers: deletions (`a:`), expansions
// (`a:xy`), tr-like maps mixed with them.  Every attempt of the reference's scan loop is decided by its first
// byte, so the driver (trre_nft.c:778-789) degenerates to "every byte prints its own string" - position
// parallel, no state between bytes; only the output offsets need a prefix sum:
//
//   k_bmv_count  per span (contiguous 512-byte tiles owned by one warp): sum of output lengths, two bytes
//                per shared-memory lookup (128 KB pair-length table staged by TMA like k_bytemap's)
//   k_scan_groups  exclusive scan over the spans (kernels.cuh)
//   k_bmv_write  per tile: 16 conflict-free entry lookups per lane (length | first byte | pool offset),
//                lane-local prefix + warp scan, bytes placed into the warp's shared-memory window, which
//                leaves in aligned 16-byte vectors; the < 16 leftover bytes carry into the next tile
//
// RUN form (BlobHeader::vrun_classes: `[0-9]+:#`, `a+:x`, ` +: `): a match that starts with a byte of run class r swallows
// every following byte of class r, so a byte prints nothing iff its predecessor has its own (non-zero) run class - one
// byte of left context (the neighbour lane's last byte by shuffle, the previous tile's by one load), still no state:
//   k_bmv_count_run / k_bmv_write<true>
//
// NUL bytes / a missing final newline raise ST_IRREGULAR (the general path reruns the buffer).
#pragma once
#include <cstdint>

#include "kernels.cuh"
#include "lit_kernel.cuh"

namespace trre {

enum { BMV_PLAIN = 0, BMV_RUN = 1, BMV_LIT = 2 };    // where k_bmv_write takes a position's entry from

// BMV_LIT: the literal word form of a program whose matches print strings of any length (BlobHeader::lit_out:
// cat:mouse|dog:d): k_lit's comparison finds the word starts, a start prints its word's string, the L - 1 bytes behind it
// nothing, every other byte itself - k_bmv's placement does the rest.
struct LitvArgs {
    LitMap M;                 // M.eq: the bytes of the words (plain strings only: every position accepts exactly one byte)
    uint32_t ent[4];          // per word: printed length | first byte << 8 | offset into the staged pool << 16
};

// Word starts in one lane's vector (k_lit's EQ comparison): code[i] byte b = 1 + w where word w starts at byte b of data
// word i (0: none).  W[4] = the data word after the vector.
template <int NW, int L>
__device__ __forceinline__ void litv_starts(const uint32_t (&W)[5], const LitMap &M, uint32_t (&code)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t D[L];
#pragma unroll
        for (int k = 0; k < L; ++k) D[k] = k == 0 ? W[i] : __funnelshift_r(W[i], W[i + 1], 8 * k);
        code[i] = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            uint32_t t = D[0] ^ M.eq[w * L];
#pragma unroll
            for (int k = 1; k < L; ++k) t |= D[k] ^ M.eq[w * L + k];
            const uint32_t nz = (t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
            code[i] += ((~(nz | t) & 0x80808080u) >> 7) * (uint32_t)(w + 1);      // words cannot start at the same byte
        }
    }
}

// 0xFF over every byte of a match: low half in the start's own data word, high half in the next one
template <int L>
__device__ __forceinline__ void litv_cover(const uint32_t (&code)[4], uint32_t (&lo)[4], uint32_t (&hi)[4]) {
    constexpr uint32_t kSpan = L == 4 ? 0xFFFFFFFFu : ((1u << (8 * L)) - 1u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t any = (code[i] | (code[i] >> 1) | (code[i] >> 2)) & 0x01010101u;      // codes are 1..4
        lo[i] = any * kSpan;
        hi[i] = __umulhi(any, kSpan);
    }
}

constexpr int kBmvCountThreads = 1024;
constexpr int kBmvWriteThreads = 512;
constexpr int kBmvTile = 512;
constexpr int kBmvWindow = 2048;                     // output bytes a warp assembles per tile without falling back
constexpr int kBmvWarpSmem = 16 + kBmvWindow + 32;

struct BmvArgs {
    const uint8_t *in;
    uint8_t *out;
    size_t n;
    size_t out_cap;
    const uint32_t *vent;            // [256]
    const uint16_t *vpairlen;        // [65536] swizzled
    const uint8_t *vpool;
    uint32_t vpool_bytes;
    unsigned long long *gsum;        // [n_spans]: lengths, then exclusive offsets
    const unsigned long long *total;
    uint32_t *status;
    size_t n_spans;
    size_t tiles_per_span;
};

__device__ __forceinline__ uint32_t pairlen_word(uint32_t w, uint32_t tab) {
    uint32_t lo, hi;
    uint32_t a0 = tab + (((w << 1) & 0x1FFFEu) ^ ((w >> 6) & 0x7Cu));
    uint32_t a1 = tab + (((w >> 15) & 0x1FFFEu) ^ ((w >> 22) & 0x7Cu));
    asm("ld.shared.u16 %0, [%1];" : "=r"(lo) : "r"(a0));
    asm("ld.shared.u16 %0, [%1];" : "=r"(hi) : "r"(a1));
    return lo + hi;
}

// RUN form whose run classes are byte ranges that print strings of one length each (`[0-9]+:#`): the count needs no
// per-byte entries - the pair-length table gives the sum over all bytes as before, and the bytes a run swallows (in the
// class and so is the byte before: one SWAR range test per data word, AND-ed with itself moved up one byte) are taken back
// out with a popcount.
struct RunMap {
    uint32_t n;              // run classes (<= 4); 0: none of this form
    uint32_t add_lo[4], add_hi[4], half[4];      // the 7-bit range test of k_bytemap_swar per class
    uint32_t len[4];         // what a byte of the class would print as a run start
};

template <bool RUNS>
__global__ void __launch_bounds__(kBmvCountThreads, 1)
k_bmv_count(BmvArgs a, RunMap R) {
    extern __shared__ __align__(128) uint8_t smem_tab[];   // kPairLutBytes
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, kPairLutBytes);
#pragma unroll
        for (int k = 0; k < 8; ++k)
            tma_bulk_g2s(smem_tab + k * (kPairLutBytes / 8), reinterpret_cast<const uint8_t *>(a.vpairlen) + k * (kPairLutBytes / 8),
                         kPairLutBytes / 8, &bar);
    }
    const uint32_t tab = smem_u32(smem_tab);
    const uint32_t lane = threadIdx.x & 31;
    const size_t n = a.n;
    const size_t n_tiles = (n + kBmvTile - 1) / kBmvTile;
    const size_t warps_total = (size_t)gridDim.x * (blockDim.x / 32);
    uint32_t nul = 0;
    mbar_wait(&bar, 0);
    for (size_t span = (size_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5); span < a.n_spans; span += warps_total) {
        const size_t t0 = span * a.tiles_per_span;
        size_t t1 = t0 + a.tiles_per_span;
        if (t1 > n_tiles) t1 = n_tiles;
        unsigned long long sum = 0;
        for (size_t t = t0; t < t1; ++t) {
            const size_t pos = t * (size_t)kBmvTile + lane * 16;
            if (!RUNS) {
                if (pos + 16 <= n) {
                    const uint4 x = ld_stream(reinterpret_cast<const uint4 *>(a.in + pos));
                    sum += pairlen_word(x.x, tab) + pairlen_word(x.y, tab) + pairlen_word(x.z, tab) + pairlen_word(x.w, tab);
                    nul |= ((x.x - 0x01010101u) & ~x.x) | ((x.y - 0x01010101u) & ~x.y) | ((x.z - 0x01010101u) & ~x.z) | ((x.w - 0x01010101u) & ~x.w);
                } else {
                    for (size_t k = pos; k < n; ++k) {
                        const uint8_t c = a.in[k];
                        if (!c) nul = 0x80u;
                        sum += __ldg(a.vent + c) & 0xFFu;
                    }
                }
            } else {
                // every lane takes part (shuffles): bytes past the buffer read as '\n', which stands in no run class
                uint4 x = make_uint4(0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au);
                uint32_t mine = 0;
                if (pos + 16 <= n) {
                    x = ld_stream(reinterpret_cast<const uint4 *>(a.in + pos));
                    mine = pairlen_word(x.x, tab) + pairlen_word(x.y, tab) + pairlen_word(x.z, tab) + pairlen_word(x.w, tab);
                } else if (pos < n) {
                    uint32_t w[4] = {x.x, x.y, x.z, x.w};
                    for (size_t k = pos; k < n; ++k) {
                        const uint32_t c = a.in[k], j = (uint32_t)(k - pos);
                        w[j >> 2] = (w[j >> 2] & ~(0xFFu << ((j & 3) * 8))) | (c << ((j & 3) * 8));
                        mine += __ldg(a.vent + c) & 0xFFu;
                    }
                    x = make_uint4(w[0], w[1], w[2], w[3]);
                }
                nul |= ((x.x - 0x01010101u) & ~x.x) | ((x.y - 0x01010101u) & ~x.y) | ((x.z - 0x01010101u) & ~x.z) | ((x.w - 0x01010101u) & ~x.w);
                // the byte before the lane's vector, in the top byte of a word (lane 0: the previous tile's last byte)
                uint32_t before = __shfl_up_sync(0xFFFFFFFFu, x.w, 1);
                if (lane == 0) before = (pos ? (uint32_t)__ldg(a.in + pos - 1) : (uint32_t)'\n') << 24;
                const uint32_t W[5] = {before, x.x, x.y, x.z, x.w};
                uint32_t swallowed = 0;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if ((uint32_t)r >= R.n) break;
                    uint32_t m[5];
#pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        const uint32_t l = W[i] & 0x7F7F7F7Fu;
                        m[i] = (l + R.add_lo[r]) & ~(l + R.add_hi[r]) & (W[i] ^ R.half[r]) & 0x80808080u;   // bit 7: the byte is in the class
                    }
                    uint32_t cnt = 0;
#pragma unroll
                    for (int i = 1; i < 5; ++i) cnt += __popc(m[i] & __funnelshift_l(m[i - 1], m[i], 8));  // ... and so is the byte before
                    swallowed += cnt * R.len[r];
                }
                sum += mine - swallowed;
            }
        }
#pragma unroll
        for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
        if (lane == 0) a.gsum[span] = sum;
    }
    nul &= 0x80808080u;
    if (blockIdx.x == 0 && threadIdx.x == 0 && n && a.in[n - 1] != '\n') nul = 1;
    if (__any_sync(0xFFFFFFFFu, nul != 0)) {
        if (lane == 0) atomicOr(a.status, ST_IRREGULAR);
    }
}

// the byte before tile position `pos` (lane 0's left context); a buffer starts where a record starts
__device__ __forceinline__ uint32_t bmv_byte_before(const uint8_t *__restrict__ in, size_t pos) {
    return pos ? (uint32_t)__ldg(in + pos - 1) : (uint32_t)'\n';
}

// RUN form: per-byte entries (the pair-length table cannot see the left context), otherwise k_bmv_count
__global__ void __launch_bounds__(kBmvCountThreads, 1)
k_bmv_count_run(BmvArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t *vent_s = reinterpret_cast<uint32_t *>(smem);                       // [256][32], bank == lane
    for (uint32_t k = threadIdx.x; k < 256u * 32u; k += blockDim.x) vent_s[k] = __ldg(a.vent + (k >> 5));
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t *vent_lane = vent_s + lane;
    const size_t n = a.n;
    const size_t n_tiles = (n + kBmvTile - 1) / kBmvTile;
    const size_t warps_total = (size_t)gridDim.x * (blockDim.x / 32);
    uint32_t nul = 0;
    for (size_t span = (size_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5); span < a.n_spans; span += warps_total) {
        const size_t t0 = span * a.tiles_per_span;
        size_t t1 = t0 + a.tiles_per_span;
        if (t1 > n_tiles) t1 = n_tiles;
        unsigned long long sum = 0;
        for (size_t t = t0; t < t1; ++t) {
            const size_t pos = t * (size_t)kBmvTile + lane * 16;
            uint4 x = make_uint4(0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au);
            uint32_t valid = 0;
            if (pos + 16 <= n) { x = ld_stream(reinterpret_cast<const uint4 *>(a.in + pos)); valid = 16; }
            else if (pos < n) {
                uint32_t w[4] = {0, 0, 0, 0};
                valid = (uint32_t)(n - pos);
                for (uint32_t k = 0; k < valid; ++k) w[k >> 2] |= (uint32_t)a.in[pos + k] << ((k & 3) * 8);
                for (uint32_t k = valid; k < 16; ++k) w[k >> 2] |= 0x0Au << ((k & 3) * 8);
                x = make_uint4(w[0], w[1], w[2], w[3]);
            }
            nul |= ((x.x - 0x01010101u) & ~x.x) | ((x.y - 0x01010101u) & ~x.y) | ((x.z - 0x01010101u) & ~x.z) | ((x.w - 0x01010101u) & ~x.w);
            uint32_t before = __shfl_up_sync(0xFFFFFFFFu, x.w >> 24, 1);
            if (lane == 0) before = bmv_byte_before(a.in, pos);
            uint32_t rc_prev = vent_lane[before << 5] >> 28;
            const uint32_t w[4] = {x.x, x.y, x.z, x.w};
            uint32_t mine = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint32_t c = (w[j >> 2] >> ((j & 3) * 8)) & 0xFFu;
                const uint32_t e = vent_lane[c << 5], rc = e >> 28;
                if ((uint32_t)j < valid && !(rc && rc == rc_prev)) mine += e & 0xFFu;
                rc_prev = rc;
            }
            sum += mine;
        }
#pragma unroll
        for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
        if (lane == 0) a.gsum[span] = sum;
    }
    nul &= 0x80808080u;
    if (blockIdx.x == 0 && threadIdx.x == 0 && n && a.in[n - 1] != '\n') nul = 1;
    if (__any_sync(0xFFFFFFFFu, nul != 0)) {
        if (lane == 0) atomicOr(a.status, ST_IRREGULAR);
    }
}

// the data word after the vector at `pos` (lane 31's right context); what lies past the buffer reads as '\n'
__device__ __forceinline__ uint32_t bmv_word_after(const uint8_t *__restrict__ in, size_t pos, size_t n) {
    uint32_t w = 0x0A0A0A0Au;
    if (pos + 20 <= n) w = __ldg(reinterpret_cast<const uint32_t *>(in + pos + 16));
    else
        for (uint32_t k = 0; k < 4 && pos + 16 + k < n; ++k) w = (w & ~(0xFFu << (8 * k))) | ((uint32_t)in[pos + 16 + k] << (8 * k));
    return w;
}

// one lane's 16 bytes at `pos` (bytes past the buffer read as '\n') and how many of them exist
__device__ __forceinline__ uint4 bmv_load(const uint8_t *__restrict__ in, size_t pos, size_t n, uint32_t &valid) {
    uint4 x = make_uint4(0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au);
    valid = 0;
    if (pos + 16 <= n) { x = ld_stream(reinterpret_cast<const uint4 *>(in + pos)); valid = 16; }
    else if (pos < n) {
        uint32_t w[4] = {x.x, x.y, x.z, x.w};
        valid = (uint32_t)(n - pos);
        for (uint32_t k = 0; k < valid; ++k) w[k >> 2] = (w[k >> 2] & ~(0xFFu << ((k & 3) * 8))) | ((uint32_t)in[pos + k] << ((k & 3) * 8));
        x = make_uint4(w[0], w[1], w[2], w[3]);
    }
    return x;
}

// BMV_LIT: output size per span, position by position as k_bmv_write<BMV_LIT> will place it: a start counts its word's
// string, a byte a match covers nothing (also when the match started in the previous lane, tile or span), others 1
template <int NW, int L>
__global__ void __launch_bounds__(kBmvCountThreads, 1)
k_litv_count(BmvArgs a, LitvArgs A) {
    const uint32_t lane = threadIdx.x & 31;
    const size_t n = a.n;
    const size_t n_tiles = (n + kBmvTile - 1) / kBmvTile;
    const size_t warps_total = (size_t)gridDim.x * (blockDim.x / 32);
    uint32_t nul = 0;
    for (size_t span = (size_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5); span < a.n_spans; span += warps_total) {
        const size_t t0 = span * a.tiles_per_span;
        size_t t1 = t0 + a.tiles_per_span;
        if (t1 > n_tiles) t1 = n_tiles;
        long long sum = 0;
        uint32_t carry_cov = 0;
        if (t0 > 0 && t0 < t1) {                                                   // a match that reaches in from the tile before the span
            const size_t before = t0 * (size_t)kBmvTile - 16;
            uint32_t valid;
            const uint4 y = bmv_load(a.in, before, n, valid);
            const uint32_t W[5] = {y.x, y.y, y.z, y.w, bmv_word_after(a.in, before, n)};
            uint32_t code[4], lo[4], hi[4];
            litv_starts<NW, L>(W, A.M, code);
            litv_cover<L>(code, lo, hi);
            carry_cov = hi[3];
        }
        for (size_t t = t0; t < t1; ++t) {
            const size_t pos = t * (size_t)kBmvTile + lane * 16;
            uint32_t valid;
            const uint4 x = bmv_load(a.in, pos, n, valid);
            nul |= ((x.x - 0x01010101u) & ~x.x) | ((x.y - 0x01010101u) & ~x.y) | ((x.z - 0x01010101u) & ~x.z) | ((x.w - 0x01010101u) & ~x.w);
            uint32_t W[5] = {x.x, x.y, x.z, x.w, __shfl_down_sync(0xFFFFFFFFu, x.x, 1)};
            if (lane == 31) W[4] = bmv_word_after(a.in, pos, n);
            uint32_t code[4], lo[4], hi[4];
            litv_starts<NW, L>(W, A.M, code);
            litv_cover<L>(code, lo, hi);
            uint32_t cov_in = __shfl_up_sync(0xFFFFFFFFu, hi[3], 1);
            if (lane == 0) cov_in = carry_cov;
            carry_cov = __shfl_sync(0xFFFFFFFFu, hi[3], 31);
            // covered bytes (never padding: '\n' stands in no word)
            int mine = (int)valid - __popc((lo[0] | cov_in) & 0x01010101u) - __popc((lo[1] | hi[0]) & 0x01010101u) -
                       __popc((lo[2] | hi[1]) & 0x01010101u) - __popc((lo[3] | hi[2]) & 0x01010101u);
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                uint32_t cnt = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t z = code[i] ^ ((uint32_t)(w + 1) * 0x01010101u);           // zero bytes = starts of word w
                    cnt += __popc(~(((z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z) & 0x80808080u);
                }
                mine += (int)(cnt * (A.ent[w] & 0xFFu));
            }
            sum += mine;
        }
#pragma unroll
        for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
        if (lane == 0) a.gsum[span] = (unsigned long long)sum;
    }
    nul &= 0x80808080u;
    if (blockIdx.x == 0 && threadIdx.x == 0 && n && a.in[n - 1] != '\n') nul = 1;
    if (__any_sync(0xFFFFFFFFu, nul != 0)) {
        if (lane == 0) atomicOr(a.status, ST_IRREGULAR);
    }
}

template <int SRC, int NW = 1, int L = 2>
__global__ void __launch_bounds__(kBmvWriteThreads)
k_bmv_write(BmvArgs a, LitvArgs A) {
    constexpr bool RUN = SRC == BMV_RUN;
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *vent_s = reinterpret_cast<uint32_t *>(smem);                       // [256][32], bank == lane
    uint8_t *vpool_s = smem + 256 * 32 * 4;
    uint8_t *s_out = vpool_s + ((a.vpool_bytes + 15) & ~15u) + warp * kBmvWarpSmem;
    if (*a.status & ST_IRREGULAR) return;                                          // k_bmv_count saw a NUL / no final newline
    if (*a.total > a.out_cap) {
        if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(a.status, ST_NOSPACE);
        return;
    }
    if (SRC != BMV_LIT)
        for (uint32_t k = threadIdx.x; k < 256u * 32u; k += blockDim.x) vent_s[k] = __ldg(a.vent + (k >> 5));
    for (uint32_t k = threadIdx.x; k < a.vpool_bytes; k += blockDim.x) vpool_s[k] = __ldg(a.vpool + k);
    __syncthreads();
    const uint32_t *vent_lane = vent_s + lane;
    const size_t n = a.n;
    const size_t n_tiles = (n + kBmvTile - 1) / kBmvTile;
    const size_t warps_total = (size_t)gridDim.x * (blockDim.x / 32);

    for (size_t span = (size_t)blockIdx.x * (blockDim.x / 32) + warp; span < a.n_spans; span += warps_total) {
        const size_t t0 = span * a.tiles_per_span;
        size_t t1 = t0 + a.tiles_per_span;
        if (t1 > n_tiles) t1 = n_tiles;
        if (t0 >= t1) continue;
        const unsigned long long span_off = a.gsum[span];
        // s_out[0] always corresponds to the 16-byte aligned output address out + gpos; the first `foreign`
        // bytes of the window belong to the previous span and are never stored
        uint32_t foreign = (uint32_t)((reinterpret_cast<uintptr_t>(a.out) + span_off) & 15);
        unsigned long long gpos = span_off - foreign;
        uint32_t fill = foreign;                                                   // bytes present in the window
        uint32_t carry_cov = 0;        // BMV_LIT: bytes at the start of the next tile that a match of this tile covers
        if (SRC == BMV_LIT && t0 > 0) {                                            // ... and of the tile before the span
            const size_t before = t0 * (size_t)kBmvTile - 16;                      // its last vector
            uint32_t valid;
            const uint4 y = bmv_load(a.in, before, n, valid);
            const uint32_t W[5] = {y.x, y.y, y.z, y.w, bmv_word_after(a.in, before, n)};
            uint32_t code[4], lo[4], hi[4];
            litv_starts<NW, L>(W, A.M, code);
            litv_cover<L>(code, lo, hi);
            carry_cov = hi[3];
        }
        for (size_t t = t0; t < t1; ++t) {
            const size_t pos = t * (size_t)kBmvTile + lane * 16;
            uint32_t valid;
            const uint4 x = bmv_load(a.in, pos, n, valid);
            const uint32_t w[4] = {x.x, x.y, x.z, x.w};
            uint32_t ent[16], mine = 0;
            if (SRC == BMV_LIT) {
                uint32_t W[5] = {x.x, x.y, x.z, x.w, __shfl_down_sync(0xFFFFFFFFu, x.x, 1)};
                if (lane == 31) W[4] = bmv_word_after(a.in, pos, n);
                uint32_t code[4], lo[4], hi[4];
                litv_starts<NW, L>(W, A.M, code);
                litv_cover<L>(code, lo, hi);
                uint32_t cov_in = __shfl_up_sync(0xFFFFFFFFu, hi[3], 1);   // what the previous lane's (tile's) last match covers here
                if (lane == 0) cov_in = carry_cov;
                carry_cov = __shfl_sync(0xFFFFFFFFu, hi[3], 31);
                const uint32_t cov[4] = {lo[0] | cov_in, lo[1] | hi[0], lo[2] | hi[1], lo[3] | hi[2]};
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t c = (code[j >> 2] >> ((j & 3) * 8)) & 0xFFu;
                    uint32_t e = ((cov[j >> 2] >> ((j & 3) * 8)) & 0xFFu) ? 0u : (1u | (((w[j >> 2] >> ((j & 3) * 8)) & 0xFFu) << 8));
#pragma unroll
                    for (int k = 0; k < NW; ++k)
                        if (c == (uint32_t)(k + 1)) e = A.ent[k];       // a start prints its word's string
                    ent[j] = (uint32_t)j < valid ? e : 0u;
                    mine += ent[j] & 0xFFu;
                }
            }
            uint32_t rc_prev = 0;
            if (RUN) {
                uint32_t before = __shfl_up_sync(0xFFFFFFFFu, x.w >> 24, 1);       // (lanes past the end hold zeros: class 0)
                if (lane == 0) before = bmv_byte_before(a.in, pos);
                rc_prev = vent_lane[before << 5] >> 28;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (SRC == BMV_LIT) break;
                const uint32_t c = (w[j >> 2] >> ((j & 3) * 8)) & 0xFFu;
                ent[j] = (uint32_t)j < valid ? vent_lane[c << 5] : 0u;
                if (RUN) {
                    const uint32_t rc = ent[j] >> 28;
                    ent[j] = (rc && rc == rc_prev) ? 0u : (ent[j] & 0x0FFFFFFFu);   // swallowed by the run it continues
                    rc_prev = rc;
                }
                mine += ent[j] & 0xFFu;
            }
            uint32_t incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= (uint32_t)d) incl += y;
            }
            const uint32_t tile_total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            uint32_t at = fill + incl - mine;
            if (fill + tile_total <= (uint32_t)kBmvWindow + 16) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t len = ent[j] & 0xFFu;
                    if (len == 1) s_out[at] = (uint8_t)(ent[j] >> 8);
                    else if (len > 1) {
                        // (measured and rejected for BMV_LIT: one copy loop per word with a warp-uniform trip count and a
                        // broadcast source behind a ballot - cat:mouse|dog:d 3.4 -> 5.9 ms per GiB)
                        const uint8_t *src = vpool_s + (ent[j] >> 16);
                        for (uint32_t k = 0; k < len; ++k) s_out[at + k] = src[k];
                    }
                    at += len;
                }
                fill += tile_total;
                __syncwarp();
                // full vectors leave; vector 0 byte-wise while it still contains foreign bytes
                const uint32_t nvec = fill / 16;
                for (uint32_t v = lane; v < nvec; v += 32) {
                    const uint4 y = *reinterpret_cast<const uint4 *>(s_out + v * 16);
                    if (v == 0 && foreign) {
                        const uint32_t yy[4] = {y.x, y.y, y.z, y.w};
                        for (uint32_t k = foreign; k < 16; ++k) a.out[gpos + k] = (uint8_t)(yy[k >> 2] >> ((k & 3) * 8));
                    } else {
                        st_stream(reinterpret_cast<uint4 *>(a.out + gpos) + v, y);
                    }
                }
                const uint32_t rest = fill - nvec * 16;
                uint8_t keep = 0;
                if (lane < rest) keep = s_out[nvec * 16 + lane];
                __syncwarp();
                if (nvec) {
                    if (lane < rest) s_out[lane] = keep;
                    gpos += (unsigned long long)nvec * 16;
                    fill = rest;
                    foreign = 0;
                }
                __syncwarp();
            } else {
                // a tile that prints more than the window holds: flush what is pending, then store directly
                for (uint32_t k = foreign + lane; k < fill; k += 32) a.out[gpos + k] = s_out[k];
                unsigned long long g = gpos + at;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t len = ent[j] & 0xFFu;
                    if (len == 1) a.out[g] = (uint8_t)(ent[j] >> 8);
                    else if (len > 1) {
                        const uint8_t *src = vpool_s + (ent[j] >> 16);
                        for (uint32_t k = 0; k < len; ++k) a.out[g + k] = src[k];
                    }
                    g += len;
                }
                const unsigned long long end = gpos + fill + tile_total;
                foreign = (uint32_t)((reinterpret_cast<uintptr_t>(a.out) + end) & 15);
                gpos = end - foreign;
                fill = foreign;
                __syncwarp();
            }
        }
        for (uint32_t k = foreign + lane; k < fill; k += 32) a.out[gpos + k] = s_out[k];   // the span's last < 16 bytes
        __syncwarp();
    }
}

}  // namespace trre
