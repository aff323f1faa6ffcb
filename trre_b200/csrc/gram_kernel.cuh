// gram_kernel.cuh - k_gram: position-parallel scan for length-preserving acyclic programs whose attempts are decided
// by at most D <= 6 bytes and whose byte classes are few enough that EVERY possible attempt fits one table (sm_100a).
//
// k_match (match_kernel.cuh) walks the automaton from every live position: a dependent chain of table lookups per
// candidate, divergent across the warp.  Here the host walks instead (automaton.cpp, build_grams): an attempt of an
// acyclic program reads at most D = max_depth bytes, so its outcome is a function of the D byte CLASSES at the position -
// n_classes^D combinations (cat:dog 4^3, (cat:dog|dog:cat) 7^3, the 10 000-word dictionary 27^4 = 531 441).  The device
// keeps
//     gbits  1 bit per combination: the attempt matches          (shared memory; up to 512 combinations: one word per
//                                                                  combination and lane, conflict free)
//     gout   u32 per combination: what the match prints          (shared memory when small, else L2)
// and a tile (512 bytes per warp, 16 positions per lane, as k_match) becomes straight-line code:
//   A  16-byte vector load per lane (next tile prefetched); the D-1 bytes past a lane's end come from its neighbour
//      by shuffle
//   B  class of every byte: one conflict-free lookup each (32x replicated class table, bank = lane)
//   C  index of the D-gram at each position by Horner steps over the classes (IMAD), one bitmap lookup -> candidate
//      mask.  With matches of one common length (every BASELINE config) their lengths need no lookup at all.
//   D  chain resolution (resolve_chain, shared with k_match): which candidates start outside an earlier match
//   E  the real matches - a few per cent of the positions - fetch their output from gout and overwrite their span of the
//      tile.  When every match prints the same number (<= 4) of bytes the whole tile stays in REGISTERS: the bytes come
//      with the hit word (small tables) or one predicated gout load per position, and are merged into the lane's four
//      output words at compile-time byte offsets; otherwise the tile is staged in shared memory
//   F  16-byte vector store per lane
// Scan-driver semantics, warm-up from the last '\n' before a warp's span and ST_IRREGULAR handling are k_match's
// (trre_nft.c:778-789 on a regular buffer: an attempt at every position not covered by an earlier match).
#pragma once
#include <cstdint>

#include "match_kernel.cuh"

namespace trre {

constexpr int kGramThreads = 1024;
constexpr int kGramWarps = kGramThreads / 32;
constexpr int kGramMaxDepth = 6;
constexpr int kGramWarpSmem = 544 + 544;             // s_in (+halo), s_out (+halo)
constexpr uint32_t kGramSmallEntries = 512;          // up to here: one u32 per combination and lane (conflict free)

enum : uint32_t {
    GRAM_DIRECT = 1u,       // gout holds the output bytes themselves (every match prints <= 4 bytes, one common length)
    GRAM_MAPDELTA = 2u,     // ... and they are the matched bytes moved by ONE constant (word:WORD dictionaries): no gout at all
};

struct GramArgs {
    const uint8_t *in;
    uint8_t *out;
    size_t n;
    uint32_t *status;
    size_t tiles_per_warp;
    const uint8_t *classmap;     // u8[256]
    const uint32_t *gbits;       // u32[(entries + 31) / 32]
    const uint32_t *gout;        // u32[entries]: GRAM_DIRECT: bytes; else pool offset(20) | printed length(8) << 20 | (consumed - 1) << 28
    const uint8_t *gpool;
    uint32_t gpool_bytes;
    uint32_t n_classes;
    uint32_t entries;            // n_classes ^ D
    uint32_t uniform_len;        // > 0: every match consumes exactly this many bytes
    uint32_t gflags;
    uint32_t gout_in_smem;
    uint32_t pool_in_smem;
    uint32_t len_mask;           // REG form: 0xFF in the low uniform_len bytes
    uint32_t map_delta;          // REG form 2 (GRAM_MAPDELTA): added mod 256 to every matched byte
};

struct GramTables {
    const uint8_t *cls_b;        // shared: class of byte c for lane l at byte offset c * 128 + l * 4 (32 KB, bank = lane)
    const uint8_t *hit_b;        // SMALL: u32 per combination and lane at idx * 128 + l * 4, bit 31 = the attempt matches
    const uint32_t *bits;        // !SMALL: bitmap, bit (31 - idx % 32) of word idx / 32 (reversed: a left shift by idx
                                 // brings the bit to position 31)
    const uint32_t *gout;        // shared or global
    const uint8_t *pool;         // shared or global
    uint32_t lane4;              // lane * 4
    const uint16_t *chain_lut;   // register form: greedy selection of non-overlapping matches of the common length, 8
                                 // candidate bits at a time: [carry-in << 8 | bits] -> real bits | carry-out << 8
};

// index of the D-gram starting at tile position p, from the staged bytes (any lane, any position)
template <int D>
__device__ __forceinline__ uint32_t gram_index_at(const GramTables &T, const uint8_t *s_in, uint32_t p, uint32_t ncls) {
    uint32_t idx = 0;
#pragma unroll
    for (int i = D - 1; i >= 0; --i)
        idx = idx * ncls + *reinterpret_cast<const uint32_t *>(T.cls_b + (((uint32_t)s_in[p + i] << 7) | T.lane4));
    return idx;
}

// class of byte k (0..3) of word w: one shift, one and-or (the lane's column is OR-ed in: the fields do not overlap),
// one conflict-free shared-memory load
template <int K>
__device__ __forceinline__ uint32_t class_of(const GramTables &T, uint32_t w) {
    // (measured and rejected: the right shifts as IMAD.HI, on the idle FMA pipe - cfg5 0.304 -> 0.283)
    const uint32_t sh = K == 0 ? (w << 7) : (w >> (8 * K - 7));
    return *reinterpret_cast<const uint32_t *>(T.cls_b + ((sh & 0x7F80u) | T.lane4));
}

// Steps A-C shared by both tile forms: classes of the lane's 16 (+ D-1) bytes, gram index per position, candidate mask.
// pay[j] = what step E needs for position j: SMALL: the hit word itself (bit 31 = match, REG: low 24 bits = the printed
// bytes); otherwise the gram index.
template <int D, bool SMALL>
__device__ __forceinline__ uint32_t gram_candidates(const GramArgs &a, const GramTables &T, uint4 x, uint4 x_next,
                                                    uint32_t lane, uint32_t first_valid, uint32_t (&pay)[16]) {
    // the D-1 bytes after this lane's 16: the neighbour's first words (lane 31: the next tile's)
    uint32_t h0 = __shfl_down_sync(0xFFFFFFFFu, x.x, 1), h1 = 0;
    const uint32_t n0 = __shfl_sync(0xFFFFFFFFu, x_next.x, 0);
    if (lane == 31) h0 = n0;
    if (D > 5) {
        h1 = __shfl_down_sync(0xFFFFFFFFu, x.y, 1);
        const uint32_t n1 = __shfl_sync(0xFFFFFFFFu, x_next.y, 0);
        if (lane == 31) h1 = n1;
    }
    // ---- B: byte classes, conflict free (bank = lane) ------------------------------------------------------------------
    const uint32_t ncls = a.n_classes;
    uint32_t cls[16 + kGramMaxDepth - 1];
    cls[0] = class_of<0>(T, x.x);  cls[1] = class_of<1>(T, x.x);  cls[2] = class_of<2>(T, x.x);  cls[3] = class_of<3>(T, x.x);
    cls[4] = class_of<0>(T, x.y);  cls[5] = class_of<1>(T, x.y);  cls[6] = class_of<2>(T, x.y);  cls[7] = class_of<3>(T, x.y);
    cls[8] = class_of<0>(T, x.z);  cls[9] = class_of<1>(T, x.z);  cls[10] = class_of<2>(T, x.z); cls[11] = class_of<3>(T, x.z);
    cls[12] = class_of<0>(T, x.w); cls[13] = class_of<1>(T, x.w); cls[14] = class_of<2>(T, x.w); cls[15] = class_of<3>(T, x.w);
    if (D >= 2) cls[16] = class_of<0>(T, h0);
    if (D >= 3) cls[17] = class_of<1>(T, h0);
    if (D >= 4) cls[18] = class_of<2>(T, h0);
    if (D >= 5) cls[19] = class_of<3>(T, h0);
    if (D >= 6) cls[20] = class_of<0>(T, h1);
    // ---- C: D-gram index per position (Horner over the classes), one lookup each -> candidate mask --------------------
    // positions are taken from 15 down to 0 and the hit bit enters the mask through a funnel shift: 1 instruction
    uint32_t cand = 0;
    const uint32_t n2 = ncls * ncls, n4 = n2 * n2;
    uint32_t q[16 + kGramMaxDepth - 1];                           // pair codes cls[t] + ncls * cls[t+1]
    if (D >= 3) {
#pragma unroll
        for (int t = 0; t < 16 + D - 2; ++t) q[t] = cls[t] + ncls * cls[t + 1];
    }
#pragma unroll
    for (int j = 15; j >= 0; --j) {
        uint32_t idx;
        if (D == 1) idx = cls[j];
        else if (D == 2) idx = cls[j] + ncls * cls[j + 1];
        else if (D == 3) idx = q[j] + n2 * cls[j + 2];
        else if (D == 4) idx = q[j] + n2 * q[j + 2];
        else if (D == 5) idx = q[j] + n2 * q[j + 2] + n4 * cls[j + 4];
        else idx = q[j] + n2 * q[j + 2] + n4 * q[j + 4];
        uint32_t hit;                                               // bit 31 = the attempt at position j matches
        if (SMALL) { hit = *reinterpret_cast<const uint32_t *>(T.hit_b + ((idx << 7) | T.lane4)); pay[j] = hit; }
        else { hit = T.bits[idx >> 5] << (idx & 31u); pay[j] = idx; }
        cand = __funnelshift_l(hit, cand, 1);
    }
    const uint32_t my0 = lane * 16;
    if (first_valid > my0) cand &= (first_valid - my0 >= 16) ? 0u : (0xFFFFu << (first_valid - my0));
    return cand;
}

// resolve_chain for matches of ONE common length: the lane-local part - "take the lowest candidate, skip its length,
// repeat" - is a function of (8 candidate bits, carry-in), tabulated once per CTA: two shared-memory lookups replace the
// divergent loop over the lane's matches (2.8 iterations of 12 instructions on the BASELINE corpora).
__device__ __forceinline__ void resolve_chain_lut(Front &f, uint32_t C, uint32_t lane, uint32_t carry_e, const uint16_t *lut) {
    const uint32_t first = C ? (uint32_t)__ffs(C) - 1 : 16u;
    uint32_t e_used = (lane == 0) ? carry_e : 0u, real = 0, e_out = 0;
    auto chain = [&](uint32_t e) {
        const uint32_t lo = lut[(e << 8) | (C & 0xFFu)];
        const uint32_t hi = lut[(lo & 0xFF00u) | (C >> 8)];
        real = (lo & 0xFFu) | ((hi & 0xFFu) << 8);
        e_out = hi >> 8;
    };
    chain(e_used);
    for (;;) {
        uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, e_out, 1);
        if (lane == 0) prev = carry_e;
        const bool differs = prev != e_used;
        const bool need = differs && (prev > first || e_used > first);
        if (differs) e_used = prev;
        if (!__any_sync(0xFFFFFFFFu, need)) break;
        if (need) chain(e_used);
    }
    f.real = real; f.e_in = e_used; f.e_out = e_out;
}

// (a & ~m) | (b & m)
__device__ __forceinline__ uint32_t bitsel(uint32_t a, uint32_t b, uint32_t m) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xD8;" : "=r"(d) : "r"(a), "r"(b), "r"(m));
    return d;
}

// The tile in registers (GRAM_DIRECT programs: one common match length L <= 4, the printed bytes fit a word): no
// shared-memory staging at all.  After the chain is resolved the real matches are merged into the lane's four output
// words position by position (compile-time byte offsets, predicated on the `real` bit); the up to three bytes a match
// at positions 13..15 prints past the lane's end travel to the next lane (next tile: carry_sp / carry_sm) by shuffle.
// MAP (GRAM_MAPDELTA): what a match prints is its own bytes plus a constant - the emission is SWAR arithmetic on the lane's
// four words under a byte mask spread from the real-start bits (no output table, no load per match: on the 10 000-word
// dictionary that load, from a 2.1 MB table in L2, was the top stall).
template <int D, bool SMALL, bool STORE, bool MAP>
__device__ __forceinline__ void gram_tile_reg(const GramArgs &a, const GramTables &T, size_t base, bool full, uint4 x, uint4 x_next,
                                              uint32_t lane, uint32_t first_valid, uint32_t &carry_e, uint32_t &carry_sp,
                                              uint32_t &carry_sm, uint32_t &nul_acc) {
    nul_acc |= ((x.x - 0x01010101u) & ~x.x) | ((x.y - 0x01010101u) & ~x.y) | ((x.z - 0x01010101u) & ~x.z) | ((x.w - 0x01010101u) & ~x.w);
    uint32_t pay[16];
    const uint32_t cand = gram_candidates<D, SMALL>(a, T, x, x_next, lane, first_valid, pay);
    Front f;
    resolve_chain_lut(f, cand, lane, carry_e, T.chain_lut);
    // ---- E in registers --------------------------------------------------------------------------------------------------
    const uint32_t M = a.len_mask;                                  // 0xFF in the low L bytes
    uint32_t o[5] = {x.x, x.y, x.z, x.w, 0u};
    uint32_t sm = 0;                                                // which bytes of o[4] (positions 16..19) are written
    if (MAP) {
        const uint32_t L = a.uniform_len;
        uint32_t cov = f.real;                                      // bit p: position p lies inside a real match
        if (L > 1) cov |= f.real << 1;
        if (L > 2) cov |= f.real << 2;
        if (L > 3) cov |= f.real << 3;
        uint32_t cov_in = __shfl_up_sync(0xFFFFFFFFu, cov >> 16, 1);    // what the previous lane's last match covers here
        if (lane == 0) cov_in = carry_sm;
        carry_sm = __shfl_sync(0xFFFFFFFFu, cov >> 16, 31);
        cov |= cov_in;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const uint32_t m01 = (((cov >> (4 * w)) & 0xFu) * 0x00204081u) & 0x01010101u;   // 4 bits -> 0x01 in 4 bytes
            const uint32_t d = m01 * a.map_delta;
            o[w] = ((o[w] & 0x7F7F7F7Fu) + (d & 0x7F7F7F7Fu)) ^ ((o[w] ^ d) & 0x80808080u);   // byte-wise add, carries cut
        }
        if (STORE) {
            const size_t my = base + (size_t)lane * 16;
            if (full || my + 16 <= a.n) {
                *reinterpret_cast<uint4 *>(a.out + my) = make_uint4(o[0], o[1], o[2], o[3]);
            } else if (my < a.n) {
                for (int k = 0; k < 16 && my + k < a.n; ++k) a.out[my + k] = (uint8_t)(o[k >> 2] >> ((k & 3) * 8));
            }
        }
        carry_e = __shfl_sync(0xFFFFFFFFu, f.e_out, 31);
    } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (f.real & (1u << j)) {
            const uint32_t g = SMALL ? pay[j] : T.gout[pay[j]];     // SMALL: the bytes came with the hit word
            const int r = 8 * (j & 3), w = j >> 2;                  // compile-time after unrolling
            // field insert = one LOP3 per word touched (spelled in PTX: left to itself the compiler masks g first and
            // spends two more ALU instructions per match position)
            o[w] = bitsel(o[w], g << r, M << r);
            if (r) {
                const uint32_t m2 = M >> (32 - r);
                o[w + 1] = bitsel(o[w + 1], g >> (32 - r), m2);
                if (w == 3) sm |= m2;
            }
        }
    }
    uint32_t sp_in = __shfl_up_sync(0xFFFFFFFFu, o[4], 1), sm_in = __shfl_up_sync(0xFFFFFFFFu, sm, 1);
    if (lane == 0) { sp_in = carry_sp; sm_in = carry_sm; }
    o[0] = (o[0] & ~sm_in) | (sp_in & sm_in);
    if (STORE) {
        const size_t my = base + (size_t)lane * 16;
        if (full || my + 16 <= a.n) {
            *reinterpret_cast<uint4 *>(a.out + my) = make_uint4(o[0], o[1], o[2], o[3]);
        } else if (my < a.n) {
            for (int k = 0; k < 16 && my + k < a.n; ++k) a.out[my + k] = (uint8_t)(o[k >> 2] >> ((k & 3) * 8));
        }
    }
    carry_e = __shfl_sync(0xFFFFFFFFu, f.e_out, 31);
    carry_sp = __shfl_sync(0xFFFFFFFFu, o[4], 31);
    carry_sm = __shfl_sync(0xFFFFFFFFu, sm, 31);
    }
}

// The general form: tile staged in shared memory, real matches write their span there (any length, pool strings).
template <int D, bool SMALL, bool STORE>
__device__ __forceinline__ void gram_tile(const GramArgs &a, const GramTables &T, uint8_t *s_in, uint8_t *s_out,
                                          size_t base, bool full, uint4 x, uint4 x_next,
                                          uint32_t lane, uint32_t first_valid, uint32_t &carry_e, uint4 &carry_halo,
                                          uint32_t &nul_acc) {
    // ---- A: stage the tile (input and default output = raw copy); lane 0 of the next tile is the halo ---------------
    *reinterpret_cast<uint4 *>(s_in + lane * 16) = x;
    *reinterpret_cast<uint4 *>(s_out + lane * 16) = x;
    if (lane == 0) {
        *reinterpret_cast<uint4 *>(s_in + kTile) = x_next;
        *reinterpret_cast<uint4 *>(s_out + kTile) = x_next;
    }
    nul_acc |= ((x.x - 0x01010101u) & ~x.x) | ((x.y - 0x01010101u) & ~x.y) | ((x.z - 0x01010101u) & ~x.z) | ((x.w - 0x01010101u) & ~x.w);
    uint32_t pay[16];
    const uint32_t cand = gram_candidates<D, SMALL>(a, T, x, x_next, lane, first_valid, pay);
    const uint32_t ncls = a.n_classes;
    __syncwarp();                                                     // the staged tile is visible

    // ---- lengths: one common length needs no lookup; otherwise one gout lookup per candidate ----------------------------
    Front f;
    if (a.uniform_len) {
        f.lens = 0x1111111111111111ull * (unsigned long long)(a.uniform_len - 1);
    } else {
        f.lens = 0;
        uint32_t c2 = cand;
        while (c2) {
            const uint32_t j = __ffs(c2) - 1;
            c2 &= c2 - 1;
            const uint32_t g = T.gout[gram_index_at<D>(T, s_in, lane * 16 + j, ncls)];
            f.lens |= (unsigned long long)(g >> 28) << (4 * j);
        }
    }
    // ---- D ---------------------------------------------------------------------------------------------------------------
    resolve_chain(f, cand, lane, carry_e);

    // ---- E: every real match overwrites its span of the output tile ----------------------------------------------------
    if (lane == 0 && carry_e) {
        const uint32_t h[4] = {carry_halo.x, carry_halo.y, carry_halo.z, carry_halo.w};
        for (uint32_t k = 0; k < carry_e; ++k) s_out[k] = (uint8_t)(h[k >> 2] >> ((k & 3) * 8));
    }
    {
        uint32_t real = f.real;
        while (real) {
            const uint32_t j = __ffs(real) - 1;
            real &= real - 1;
            const uint32_t p = lane * 16 + j;
            const uint32_t g = T.gout[gram_index_at<D>(T, s_in, p, ncls)];
            if (a.gflags & GRAM_DIRECT) {
                const uint32_t m = a.uniform_len;
                for (uint32_t k = 0; k < m; ++k) s_out[p + k] = (uint8_t)(g >> (8 * k));
            } else {
                const uint8_t *src = T.pool + (g & 0xFFFFFu);
                const uint32_t m = (g >> 28) + 1;                      // length preserving: printed == consumed
                for (uint32_t k = 0; k < m; ++k) s_out[p + k] = src[k];
            }
        }
    }
    __syncwarp();
    if (STORE) {
        const size_t my = base + (size_t)lane * 16;
        const uint4 y = *reinterpret_cast<const uint4 *>(s_out + lane * 16);
        if (full || my + 16 <= a.n) {
            *reinterpret_cast<uint4 *>(a.out + my) = y;
        } else if (my < a.n) {
            const uint32_t w[4] = {y.x, y.y, y.z, y.w};
            for (int k = 0; k < 16 && my + k < a.n; ++k) a.out[my + k] = (uint8_t)(w[k >> 2] >> ((k & 3) * 8));
        }
    }
    carry_e = __shfl_sync(0xFFFFFFFFu, f.e_out, 31);
    carry_halo = *reinterpret_cast<const uint4 *>(s_out + kTile);     // every lane reads it; only lane 0 uses it
    __syncwarp();
}

// Tables into shared memory (every thread of the CTA takes part, ends with __syncthreads); returns where the per-warp
// areas start.  layout: [class table 32 KB] [SMALL: hit words x32 | bitmap] [gout]? [pool]?
template <bool SMALL, bool REG>
__device__ __forceinline__ uint8_t *gram_setup(const GramArgs &a, uint8_t *smem, uint32_t lane, GramTables &T) {
    uint8_t *cur = smem;
    uint32_t *cls_s = reinterpret_cast<uint32_t *>(cur); cur += 256 * 32 * 4;
    uint32_t *bits_s = reinterpret_cast<uint32_t *>(cur);
    const uint32_t words = (a.entries + 31) / 32;
    cur += SMALL ? (size_t)a.entries * 128 : (size_t)((words + 3) & ~3u) * 4;
    uint32_t *gout_s = nullptr;
    if (a.gout_in_smem) { gout_s = reinterpret_cast<uint32_t *>(cur); cur += (size_t)((a.entries + 3) & ~3u) * 4; }
    uint8_t *pool_s = nullptr;
    if (a.pool_in_smem) { pool_s = cur; cur += (a.gpool_bytes + 15) & ~15u; }

    for (uint32_t k = threadIdx.x; k < 256u * 32u; k += blockDim.x) cls_s[k] = __ldg(a.classmap + (k >> 5));
    if (SMALL) {
        for (uint32_t k = threadIdx.x; k < a.entries * 32u; k += blockDim.x) {
            const uint32_t e = k >> 5;
            // REG: the (<= 3) printed bytes ride in the hit word
            bits_s[k] = (((__ldg(a.gbits + (e >> 5)) >> (e & 31u)) & 1u) << 31) | (REG ? (__ldg(a.gout + e) & 0x00FFFFFFu) : 0u);
        }
    } else {
        for (uint32_t k = threadIdx.x; k < words; k += blockDim.x) bits_s[k] = __brev(__ldg(a.gbits + k));
    }
    if (gout_s)
        for (uint32_t k = threadIdx.x; k < a.entries; k += blockDim.x) gout_s[k] = __ldg(a.gout + k);
    if (pool_s)
        for (uint32_t k = threadIdx.x; k < a.gpool_bytes; k += blockDim.x) pool_s[k] = __ldg(a.gpool + k);
    uint16_t *lut_s = reinterpret_cast<uint16_t *>(cur);
    if (REG) {
        cur += 8 * 256 * sizeof(uint16_t);
        const uint32_t L = a.uniform_len;                          // 1..4; carry-in < L (8 rows cover the index range)
        for (uint32_t k = threadIdx.x; k < 8u * 256u; k += blockDim.x) {
            const uint32_t bits = k & 0xFFu;
            uint32_t pos = k >> 8, r = 0;
            while (pos < 8) {
                if ((bits >> pos) & 1u) { r |= 1u << pos; pos += L; } else ++pos;
            }
            lut_s[k] = (uint16_t)(r | ((pos - 8) << 8));
        }
    }
    __syncthreads();

    T.cls_b = reinterpret_cast<const uint8_t *>(cls_s);
    T.hit_b = reinterpret_cast<const uint8_t *>(bits_s);
    T.bits = bits_s;
    T.gout = gout_s ? gout_s : a.gout;
    T.pool = pool_s ? pool_s : a.gpool;
    T.lane4 = lane * 4;
    T.chain_lut = lut_s;
    return cur;
}

// ---- variable-length matches through the gram tables (PF_MATCHABLE_VAR programs with gram tables: cat:mouse, (cat):) ----
// The front end of the count / write tiles of match_kernel.cuh, as a policy: candidates and their consumed lengths come
// from the tables instead of a walk; what a real match prints is gout's pool reference.
template <int D, bool SMALL>
struct GramFE {
    const GramArgs &g;
    const GramTables &T;
    __device__ __forceinline__ const uint8_t *pool() const { return T.pool; }
    __device__ __forceinline__ Front front(const uint8_t *s_in, uint4 x, uint4 halo, uint32_t lane, uint32_t first_valid,
                                           uint32_t carry_e, uint32_t &overflow) const {
        (void)overflow;
        uint32_t pay[16];
        const uint32_t cand = gram_candidates<D, SMALL>(g, T, x, halo, lane, first_valid, pay);   // every lane holds the halo
        __syncwarp();                                                 // the staged tile is visible
        Front f;
        if (g.uniform_len) {
            f.lens = 0x1111111111111111ull * (unsigned long long)(g.uniform_len - 1);
        } else {
            f.lens = 0;
            uint32_t c2 = cand;
            while (c2) {
                const uint32_t j = __ffs(c2) - 1;
                c2 &= c2 - 1;
                const uint32_t e = T.gout[gram_index_at<D>(T, s_in, lane * 16 + j, g.n_classes)];
                f.lens |= (unsigned long long)(e >> 28) << (4 * j);
            }
        }
        resolve_chain(f, cand, lane, carry_e);
        return f;
    }
    __device__ __forceinline__ uint32_t entry(const uint8_t *s_in, uint32_t p) const {
        const uint32_t e = T.gout[gram_index_at<D>(T, s_in, p, g.n_classes)];
        return (e & 0xFFFFFu) | (((e >> 20) & 0xFFu) << 24);          // offset(24) | printed length(8), as WalkFE::entry
    }
};

template <int D, bool SMALL, int OP>
__global__ void __launch_bounds__(kGramThreads, 1)
k_gram_var(MatchArgs m, GramArgs g) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    GramTables T;
    uint8_t *cur = gram_setup<SMALL, false>(g, smem, lane, T);
    span_loop<OP>(m, GramFE<D, SMALL>{g, T}, cur + warp * kVarWarpSmem, lane);
}

template <int D, bool SMALL, int REGF>       // REGF: 0 staged tile, 1 tile in registers, 2 ... and outputs by a constant delta
__global__ void __launch_bounds__(kGramThreads, 1)
k_gram(GramArgs a) {
    constexpr bool REG = REGF != 0;
    constexpr bool MAP = REGF == 2;
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    GramTables T;
    uint8_t *cur = gram_setup<SMALL, REG>(a, smem, lane, T);
    uint8_t *warp_base = cur + (REG ? 0 : warp * kGramWarpSmem);     // REG: no staging area at all

    uint8_t *s_in = warp_base, *s_out = warp_base + 544;

    const size_t n = a.n;
    const size_t n_tiles = (n + kTile - 1) / kTile;
    const size_t gwarp = (size_t)blockIdx.x * kGramWarps + warp;
    const size_t t0 = gwarp * a.tiles_per_warp;
    size_t t1 = t0 + a.tiles_per_warp;
    if (t1 > n_tiles) t1 = n_tiles;
    uint32_t status = 0, nul_acc = 0;
    if (t0 < t1) {
        uint32_t carry_e = 0, carry_sp = 0, carry_sm = 0;
        uint4 carry_halo = make_uint4(0, 0, 0, 0);
        // ---- warm-up: the chain state at t0 is only known from the last '\n' before it (as k_match) ------------------
        size_t tw = t0;
        uint32_t first_valid = 0;
        if (t0 > 0) {
            bool found = false;
            for (int k = 1; k <= kWarmTiles && !found; ++k) {
                if ((size_t)k > t0) { tw = 0; first_valid = 0; found = true; break; }
                const size_t tb = (t0 - k) * (size_t)kTile;
                uint4 v = ldg_vec(a.in + tb + lane * 16);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                int last = -1;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (((w[j >> 2] >> ((j & 3) * 8)) & 0xFFu) == '\n') last = j;
                uint32_t has = __ballot_sync(0xFFFFFFFFu, last >= 0);
                if (has) {
                    int src = 31 - __clz(has);
                    int lj = __shfl_sync(0xFFFFFFFFu, last, src);
                    uint32_t pos_in_tile = (uint32_t)src * 16 + (uint32_t)lj + 1;
                    if (pos_in_tile >= (uint32_t)kTile) { tw = t0 - k + 1; first_valid = 0; }
                    else { tw = t0 - k; first_valid = pos_in_tile; }
                    found = true;
                }
            }
            if (!found) { status |= ST_IRREGULAR; tw = t0; }
        }
        size_t base = tw * (size_t)kTile;
        uint4 x = load_padded(a.in, base + lane * 16, n);
        for (size_t t = tw; t < t1; ++t) {
            base = t * (size_t)kTile;
            // prefetch; its lane 0 is this tile's halo.  Tiles whose successor lies inside the buffer skip every bounds check
            const bool full = base + 2 * (size_t)kTile <= n;
            const uint4 x_next = full ? ldg_vec(a.in + base + kTile + lane * 16) : load_padded(a.in, base + kTile + lane * 16, n);
            if (REG) {
                if (t < t0) gram_tile_reg<D, SMALL, false, MAP>(a, T, base, full, x, x_next, lane, first_valid, carry_e, carry_sp, carry_sm, nul_acc);
                else        gram_tile_reg<D, SMALL, true, MAP>(a, T, base, full, x, x_next, lane, first_valid, carry_e, carry_sp, carry_sm, nul_acc);
            } else {
                if (t < t0) gram_tile<D, SMALL, false>(a, T, s_in, s_out, base, full, x, x_next, lane, first_valid, carry_e, carry_halo, nul_acc);
                else        gram_tile<D, SMALL, true>(a, T, s_in, s_out, base, full, x, x_next, lane, first_valid, carry_e, carry_halo, nul_acc);
            }
            first_valid = 0;
            x = x_next;
        }
    }
    if (nul_acc & 0x80808080u) status |= ST_IRREGULAR;
    if (blockIdx.x == 0 && threadIdx.x == 0 && n && a.in[n - 1] != '\n') status |= ST_IRREGULAR;
    if (status) atomicOr(a.status, status);
}

}  // namespace trre
