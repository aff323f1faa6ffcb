// lit_kernel.cuh - k_lit: table-free scan for programs whose matches are a few fixed WORDS (sm_100a).
//
// cat:dog, (cat:dog|dog:cat), [Cc]at:dog ... : every match consumes and prints L = 2..4 bytes, is decided by its own L
// bytes, there are at most 4 such words, and no suffix of a word is the prefix of a word (BlobHeader::lit, build_lit in
// automaton.cpp).  Then an attempt that starts inside a match can never succeed, so "an attempt at every position not
// covered by an earlier match" (trre_nft.c:778-789) is just "every position where a word stands": no chain to resolve,
// no line structure to follow ('\n' and NUL belong to no word), no table - the scan is straight-line SWAR arithmetic on
// 32-bit words, like k_bytemap_swar:
//   A  the data shifted down by k = 1..L-1 bytes (one funnel shift per data word and k)
//   B  per position k of word w one 7-bit range test on the four bytes of the k-shifted data word (two adds whose carries
//      stay inside their byte - issued as IMAD, on the FMA pipe - and one LOP3 that also ANDs it into the word's mask):
//      bit 7 of byte b = "w can start at byte b"
//   C  start bits (0x01 per start byte, by IMAD.HI) times the word's printed bytes as one 64-bit product = the printed
//      text at its place, low half in this data word, high half in the next; the same with 0xFF..FF clears what is matched
// k_gram spends 334 instructions per lane and 16 bytes on the same programs (class lookups, gram index, chain table),
// 250 of them on the ALU pipe that bounds it; this spends about 90 there for two words of three bytes.
// Layout: a warp covers 30 consecutive 16-byte vectors per iteration; lanes 0 and 31 hold the vectors before and after
// them (halo: their masks reach lanes 1 and 30 by shuffle, they store nothing), so the only traffic besides the stream is
// 1/15 of it re-read from L2.  256 threads x 4 iterations per CTA, one CTA per contiguous 15 KB, no shared memory.
#pragma once
#include <cstdint>

#include "kernels.cuh"

namespace trre {

constexpr int kLitThreads = 256;
constexpr int kLitIters = 4;
constexpr int kLitVecPerWarpIter = 30;
constexpr int kLitVecPerCta = (kLitThreads / 32) * kLitIters * kLitVecPerWarpIter;    // 960 vectors = 15 360 bytes

struct LitMap {
    uint32_t add_lo[16];     // slot = w * L + k; the 7-bit range test of k_bytemap_swar (SwarMap), ranges below 0x80 only
    uint32_t add_hi[16];
    uint32_t eq[16];         // EQ form (every position accepts exactly one byte): that byte in all four byte lanes
    uint32_t printed[4];     // word w prints these L bytes (byte k = position k)
    uint32_t one;            // 1 at run time: `l * one + c` is an IMAD - the adds of the range tests go to the FMA pipe,
                             // which this kernel leaves idle, instead of the ALU pipe, which bounds it (LOP3 / SHF)
};

// vector `vi` of the buffer; what lies outside [0, n) reads as '\n' (belongs to no word, is not a NUL)
__device__ __forceinline__ uint4 lit_load(const uint8_t *__restrict__ in, long long vi, size_t n) {
    uint4 r = make_uint4(0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au);
    if (vi < 0) return r;
    const size_t off = (size_t)vi * 16;
    if (off + 16 <= n) return ld_stream(reinterpret_cast<const uint4 *>(in + off));
    if (off < n) {
        uint32_t w[4] = {r.x, r.y, r.z, r.w};
        for (size_t k = 0; off + k < n; ++k) {
            const uint32_t sh = (uint32_t)(k & 3) * 8;
            w[k >> 2] = (w[k >> 2] & ~(0xFFu << sh)) | ((uint32_t)in[off + k] << sh);
        }
        r = make_uint4(w[0], w[1], w[2], w[3]);
    }
    return r;
}

// One lane's vector.  Position k of a word is tested on the data shifted down by k bytes (one funnel shift per data word
// and k, whatever the number of words), so the L position masks of a word line up at its start byte and one LOP3 ANDs
// them.  A start (0x01 in its byte) times the word's printed bytes, as a 64-bit product, IS the printed text at its place:
// the low half lands in this data word, the high half in the next one (a neighbour lane's for the last word: shuffle).
// EQ (every position of every word accepts exactly one byte - plain strings, the common case): the L comparisons of a word
// collapse into ONE zero-byte test: byte b of (D0 ^ c0) | (D1 ^ c1) | .. is zero iff the word stands at b (3 LOP3 for the
// chain, 3 more for an exact zero-byte test) - 6 instructions per word and data word instead of 3 per position.
template <int NW, int L, bool EQ>
__device__ __forceinline__ uint4 lit_vec(const uint4 &x, const LitMap &M, uint32_t &nul) {
    uint32_t W[5] = {x.x, x.y, x.z, x.w, 0};
    W[4] = __shfl_down_sync(0xFFFFFFFFu, x.x, 1);                    // the next lane's first data word
    uint32_t D[L][4];                                                // the data shifted down by k bytes
    uint32_t l[L][4];                                                // ... its low 7 bits per byte
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        nul |= (EQ ? W[i] * M.one - 0x01010101u : W[i] - 0x01010101u) & ~W[i];   // (EQ: the ALU pipe is the busier one)
#pragma unroll
        for (int k = 0; k < L; ++k) {
            D[k][i] = k == 0 ? W[i] : __funnelshift_r(W[i], W[i + 1], 8 * k);
            l[k][i] = D[k][i] & 0x7F7F7F7Fu;
        }
    }
    // bytes >= 0x80 stand in no word: bit 7 of byte b of ok[i] = none of the L bytes from b on has bit 7 set
    uint32_t ok[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t hi = W[i];
#pragma unroll
        for (int k = 1; k < L; ++k) hi |= __funnelshift_r(W[i], W[i + 1], 8 * k);
        ok[i] = ~hi & 0x80808080u;
    }
    uint32_t txt_lo[4] = {0, 0, 0, 0}, txt_hi[4] = {0, 0, 0, 0};      // printed bytes: low half here, high half one data word on
    uint32_t any[4] = {0, 0, 0, 0};
#pragma unroll
    for (int w = 0; w < NW; ++w) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t m;
            if (EQ) {
                uint32_t t = D[0][i] ^ M.eq[w * L];
#pragma unroll
                for (int k = 1; k < L; ++k) t |= D[k][i] ^ M.eq[w * L + k];
                const uint32_t nz = (t & 0x7F7F7F7Fu) * M.one + 0x7F7F7F7Fu;     // bit 7 of a byte: its low 7 bits are not all zero
                m = ~(nz | t) & 0x80808080u;
            } else {
                m = ok[i];
#pragma unroll
                for (int k = 0; k < L; ++k) {
                    const int s = w * L + k;
                    const uint32_t ge = l[k][i] * M.one + M.add_lo[s], gt = l[k][i] * M.one + M.add_hi[s];
                    m &= ge & ~gt;                                   // bit 7 of a byte: position k of w can stand k bytes on
                }
            }
            const uint32_t start = __umulhi(m, 1u << 25);            // 0x01 where word w starts (IMAD.HI: FMA pipe)
            any[i] |= start;
            txt_lo[i] += start * M.printed[w];                       // words are disjoint: sums never carry
            if (L > 1) txt_hi[i] += __umulhi(start, M.printed[w]);
        }
    }
    constexpr uint32_t kSpan = L == 4 ? 0xFFFFFFFFu : ((1u << (8 * L)) - 1u);
    uint32_t cov_lo[4], cov_hi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {                                    // 0xFF over the L bytes of every match
        cov_lo[i] = any[i] * kSpan;
        cov_hi[i] = __umulhi(any[i], kSpan);
    }
    const uint32_t cov_in = __shfl_up_sync(0xFFFFFFFFu, cov_hi[3], 1), txt_in = __shfl_up_sync(0xFFFFFFFFu, txt_hi[3], 1);
    uint32_t out[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t cov = cov_lo[i] | (i ? cov_hi[i - 1] : cov_in), txt = txt_lo[i] | (i ? txt_hi[i - 1] : txt_in);
        out[i] = (W[i] & ~cov) | txt;
    }
    return make_uint4(out[0], out[1], out[2], out[3]);
}

// one CTA's 960 vectors; GUARD = false when all of them (and both halo vectors) lie inside the buffer: no bounds checks
template <int NW, int L, bool EQ, bool GUARD>
__device__ __forceinline__ uint32_t lit_cta(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, size_t n, const LitMap &M,
                                            uint32_t lane, long long first) {
    const long long nvec = (long long)((n + 15) / 16);
    uint4 x[kLitIters];
#pragma unroll
    for (int u = 0; u < kLitIters; ++u) {
        const long long vi = first + u * kLitVecPerWarpIter - 1 + (long long)lane;
        x[u] = GUARD ? lit_load(in, vi, n) : ld_stream(reinterpret_cast<const uint4 *>(in) + vi);
    }
    uint32_t nul = 0;
    const bool stores = lane >= 1 && lane <= 30;
#pragma unroll
    for (int u = 0; u < kLitIters; ++u) {
        const long long vi = first + u * kLitVecPerWarpIter - 1 + (long long)lane;
        uint32_t nul_v = 0;
        const uint4 y = lit_vec<NW, L, EQ>(x[u], M, nul_v);
        if (!GUARD) {
            if (stores) { nul |= nul_v; st_stream(reinterpret_cast<uint4 *>(out) + vi, y); }
        } else if (stores && vi < nvec) {
            nul |= nul_v;
            const size_t off = (size_t)vi * 16;
            if (off + 16 <= n) st_stream(reinterpret_cast<uint4 *>(out + off), y);
            else {
                const uint32_t w[4] = {y.x, y.y, y.z, y.w};
                for (size_t k = 0; off + k < n; ++k) out[off + k] = (uint8_t)(w[k >> 2] >> ((k & 3) * 8));
            }
        }
    }
    return nul & 0x80808080u;
}

template <int NW, int L, bool EQ>
__global__ void __launch_bounds__(kLitThreads, EQ ? (NW <= 2 ? 6 : 5) : 4)
k_lit(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, size_t n, const LitMap M, uint32_t *__restrict__ status) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long first = ((long long)blockIdx.x * (kLitThreads / 32) + warp) * (kLitIters * kLitVecPerWarpIter);
    const bool interior = blockIdx.x > 0 && ((size_t)blockIdx.x + 1) * kLitVecPerCta * 16 + 16 <= n;
    uint32_t nul = interior ? lit_cta<NW, L, EQ, false>(in, out, n, M, lane, first) : lit_cta<NW, L, EQ, true>(in, out, n, M, lane, first);
    if (blockIdx.x == 0 && threadIdx.x == 0 && n && in[n - 1] != '\n') nul = 1;   // the last record would lose its last byte
    if (__any_sync(0xFFFFFFFFu, nul != 0)) {
        if (lane == 0) atomicOr(status, ST_IRREGULAR);
    }
}

}  // namespace trre
