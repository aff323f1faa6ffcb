// match_kernel.cuh - k_match: position-parallel scan for length-preserving, acyclic programs (sm_100a).
//
// Applies to TABLE programs whose every match emits exactly what it consumes, whose automaton has no
// cycle and whose attempts read at most 16 bytes (cat:dog, (cat:dog|dog:cat), word:WORD dictionaries).
// On a regular buffer (no NUL, final '\n') output offset == input offset and the scan driver
// (trre_nft.c:778-789) degenerates to: "an attempt starts at every position not covered by an earlier
// match of the same line".  '\n' can never be consumed (dead byte class), so lines need no bookkeeping:
//
//   warp-synchronous, one 512-byte tile per warp iteration, 16 positions per lane
//   A  16-byte vector load per lane (next tile prefetched), staged to the warp's shared-memory tile
//   B  first transition of an attempt at EVERY position: row0[byte]  -> 16-bit candidate mask
//   C  candidates finish their walk (compact u32 table: hot states in shared memory, rest L2); the
//      match lengths stay in registers (16-bit mask + 4 bits per position)
//   D  chain resolution: which candidates are real (not covered by an earlier match) - speculate
//      "entry overhang 0" per lane, repair through warp shuffles until stable; carry across tiles
//   E  output tile = raw copy, real matches overwrite their span (may spill into the next tile: carried),
//      16-byte vector store per lane
//
// Each warp owns a contiguous span of tiles; it warms up from the last '\n' before its span so the
// chain state at the span start is exact.  Anything irregular (NUL byte, missing final newline, a
// line longer than the warm-up window) raises ST_IRREGULAR and the general path reruns the buffer.
#pragma once
#include <cstdint>

#include "device_walk.cuh"

namespace trre {

constexpr int kMatchThreads = 512;
constexpr int kMatchWarps = kMatchThreads / 32;
constexpr int kTile = 512;                       // bytes per warp iteration
constexpr int kWarpSmem = 544 + 544 + 1024 + 16; // s_in (+halo), s_out (+halo), list of real matches (u16 x 512), its counter
constexpr int kWarmTiles = 128;                  // warm-up looks back at most 64 KB for a '\n' (stops at the first one)

struct MatchArgs {
    DevProg P;
    const uint8_t *in;
    uint8_t *out;
    size_t n;
    uint32_t *status;
    uint32_t smem_states;        // states [0, smem_states) of ctrans live in shared memory
    uint32_t row0_replicated;    // 1: a 32x replicated live/dead table (bank = lane) serves step B
    uint32_t pool_in_smem;       // 1: the string pool is copied to shared memory
    size_t tiles_per_warp;
    // variable-length form (k_match<.., OP_COUNT / OP_WRITE>): one output-length sum / offset per warp span
    unsigned long long *gsum;
    const unsigned long long *total;
    size_t out_cap;
};

enum { OP_INPLACE = 0, OP_COUNT = 1, OP_WRITE = 2 };
constexpr int kVarWindow = 1024;                              // output bytes assembled per tile before falling back
constexpr int kVarWarpSmem = 544 + 16 + 2048 + (16 + kVarWindow + 32);   // s_in, count, match list / entries, window

__device__ __forceinline__ uint4 ldg_vec(const uint8_t *p) {
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// 16 input bytes at `pos`, '\n'-padded beyond n
__device__ __forceinline__ uint4 load_padded(const uint8_t *__restrict__ in, size_t pos, size_t n) {
    if (pos + 16 <= n) return ldg_vec(in + pos);
    uint32_t w[4] = {0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au};
    for (int k = 0; k < 16; ++k)
        if (pos + k < n) w[k >> 2] = (w[k >> 2] & ~(0xFFu << ((k & 3) * 8))) | ((uint32_t)in[pos + k] << ((k & 3) * 8));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

struct MatchTables {
    const uint32_t *live_s;      // shared: 1/0 per byte, replicated 32x (bank = lane); nullptr when not replicated
    const uint32_t *row0_s;      // shared: first transition by raw byte
    const uint8_t *classmap_s;   // shared
    const uint32_t *ctrans_s;    // shared, first smem_states rows
    const uint32_t *ctrans_g;    // global, all rows
    const uint8_t *pool;         // shared copy when small, else global
    uint32_t n_classes, smem_states, lane;
    uint32_t capped;             // the automaton has a cycle: an attempt still alive after 16 bytes is not ours to finish
};

// table layouts: TAB_PART  compact rows, the first smem_states of them in shared memory, the rest in L2
//                TAB_ALL   compact rows, all in shared memory
//                TAB_WIDE  one 256-entry row per state in shared memory, indexed by the raw byte (small automata)
enum { TAB_PART = 0, TAB_ALL = 1, TAB_WIDE = 2 };

template <int TAB>
__device__ __forceinline__ uint32_t next_entry(const MatchTables &T, uint32_t s, uint32_t c) {
    if (TAB == TAB_WIDE) return T.ctrans_s[(s << 8) | c];
    const uint32_t idx = s * T.n_classes + T.classmap_s[c];
    return (TAB == TAB_ALL || s < T.smem_states) ? T.ctrans_s[idx] : __ldg(T.ctrans_g + idx);
}

// One finished walk: m = bytes consumed by the match (0 = no match).
template <int TAB>
__device__ __forceinline__ uint32_t walk_from(const MatchTables &T, const uint8_t *s_in, uint32_t p, uint32_t &overflow) {
    uint32_t lo = T.row0_s[s_in[p]];
    uint32_t i = 1, m = (lo & DSA_E_MARK) ? 1u : 0u;
    while (!(lo & DSA_E_DECIDED)) {
        const uint32_t nx = next_entry<TAB>(T, lo & 0xFFFFFFu, s_in[p + i]);   // p + i <= 511 + 16: inside the halo
        if ((nx & 0xFFFFFFu) == DSA_DEAD) break;                     // '\n' (line end, padding) is dead
        if (i >= 16) { overflow |= T.capped; break; }                // only a cyclic automaton gets here alive (a{17,}, [0-9]+ on
        lo = nx;                                                     // a run of 17 digits): the per-line kernels take the buffer
        ++i;
        if (lo & DSA_E_MARK) m = i;
    }
    return (lo & (DSA_E_DECIDED | DSA_E_HASFINAL)) ? m : 0u;
}

// Output bytes of the real match at tile position p (length m) -> s_out[p .. p+m)
template <int TAB>
__device__ __forceinline__ void emit_match(const MatchArgs &a, const MatchTables &T, const uint8_t *s_in,
                                           uint8_t *s_out, uint32_t p, uint32_t m, uint32_t &status) {
    if (a.P.eout) {
        // PF_OUT_LATE: replay on the compact table; the last transition (or the pending state) names ONE string
        uint32_t lo = T.row0_s[s_in[p]], i = 1;
        uint32_t from = 0, by = s_in[p];                             // the last transition taken: (state, byte)
        while (!(lo & DSA_E_DECIDED)) {
            const uint32_t c = s_in[p + i];
            const uint32_t nx = next_entry<TAB>(T, lo & 0xFFFFFFu, c);
            if ((nx & 0xFFFFFFu) == DSA_DEAD) break;
            from = lo & 0xFFFFFFu; by = c;
            lo = nx; ++i;
            if (i >= 16) break;
        }
        const uint32_t slot = (lo & DSA_E_DECIDED) ? from * T.n_classes + T.classmap_s[by]
                                                   : a.P.n_states * T.n_classes + (lo & 0xFFFFFFu);
        const uint8_t *src = T.pool + (__ldg(a.P.eout + slot) & 0xFFFFFFu);
        for (uint32_t k = 0; k < m; ++k) s_out[p + k] = src[k];
        return;
    }
    // general: replay with the full table, print every transition's string, then the final string
    Sink o;
    o.p = s_out;
    o.n = p;
    uint32_t s = 0, i = 0;
    for (;;) {
        uint64_t e = __ldg(a.P.trans + (size_t)s * a.P.n_classes + T.classmap_s[s_in[p + i]]);
        uint32_t lo = (uint32_t)e;
        if ((lo & 0xFFFFFFu) == DSA_DEAD) break;
        put_ref(o, a.P.pool, (uint32_t)(e >> 32));
        s = lo & 0xFFFFFFu;
        ++i;
        if ((lo & DSA_E_DECIDED) || i >= 16) break;
    }
    put_ref(o, a.P.pool, (uint32_t)(__ldg(a.P.sinfo + s) >> 32));
    if (o.n - p != m) status |= ST_IRREGULAR;                        // cannot happen for a length-preserving program
}

// Steps B-D of a tile, shared by k_match and the variable-length kernels (matchvar_kernel.cuh): which
// matches of this lane's 16 positions are real, their lengths, and the overhangs in and out of the lane.
struct Front {
    uint32_t real;               // bit j: a real match starts at position j of this lane
    unsigned long long lens;     // (m - 1) in 4 bits per position
    uint32_t e_in, e_out;        // positions covered at the start of this lane / of the next one
};

// Step D of a tile: which candidate matches are real, i.e. start outside every earlier match of the line.  C = this
// lane's candidate mask (16 positions), lens = (m - 1) in 4 bits per position, carry_e = positions of this tile still
// covered by the last real match of the previous tile.  Speculate "nothing hangs into my lane", then repair through
// warp shuffles until stable (matches are short and sparse: one or two rounds).
__device__ __forceinline__ void resolve_chain(Front &f, uint32_t C, uint32_t lane, uint32_t carry_e) {
    const uint32_t first = C ? (uint32_t)__ffs(C) - 1 : 16u;
    uint32_t e_used = (lane == 0) ? carry_e : 0u, real = 0, e_out = 0;
    auto chain = [&](uint32_t e) {
        uint32_t pos = e, r = 0;
        uint32_t Cm = C & (0xFFFFu << pos);                          // e <= 15
        while (Cm) {
            uint32_t j = __ffs(Cm) - 1;
            r |= 1u << j;
            pos = j + 1 + ((uint32_t)(f.lens >> (4 * j)) & 15u);
            Cm = pos < 16 ? (C & (0xFFFFu << pos)) : 0u;
        }
        real = r;
        e_out = pos > 16 ? pos - 16 : 0u;
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

template <int TAB>
__device__ __forceinline__ Front tile_frontend(const MatchTables &T, const uint8_t *s_in, uint4 x, uint32_t lane,
                                               uint32_t first_valid, uint32_t carry_e, uint32_t &overflow) {
    // ---- B: can an attempt starting with this byte survive its first transition? -> candidate mask ---
    uint32_t cand = 0;
    {
        const uint32_t w[4] = {x.x, x.y, x.z, x.w};
        if (T.live_s) {
            const uint32_t *live_lane = T.live_s + lane;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                uint32_t c = (w[j >> 2] >> ((j & 3) * 8)) & 0xFFu;
                cand |= live_lane[c << 5] << j;                       // conflict free: bank == lane
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                uint32_t c = (w[j >> 2] >> ((j & 3) * 8)) & 0xFFu;
                cand |= ((T.row0_s[c] & 0xFFFFFFu) != DSA_DEAD ? 1u : 0u) << j;
            }
        }
    }
    {
        const uint32_t my0 = lane * 16;
        if (first_valid > my0) cand &= (first_valid - my0 >= 16) ? 0u : (0xFFFFu << (first_valid - my0));
    }
    __syncwarp();                                                     // the staged tile is visible

    // ---- C: candidates finish their walk; matches stay in registers: C = mask, lens = 4 bits each ----
    uint32_t C = 0;
    Front f;
    f.lens = 0;
    while (cand) {
        const uint32_t j = __ffs(cand) - 1;
        cand &= cand - 1;
        const uint32_t m = walk_from<TAB>(T, s_in, lane * 16 + j, overflow);
        if (m) { C |= 1u << j; f.lens |= (unsigned long long)(m - 1) << (4 * j); }
    }

    // ---- D: which matches are real (start outside every earlier match) ---------------------------------
    resolve_chain(f, C, lane, carry_e);
    return f;
}

template <bool STORE, int TAB>
__device__ __forceinline__ void match_tile(const MatchArgs &a, const MatchTables &T, uint8_t *s_in,
                                           uint8_t *s_out, uint16_t *s_list, uint32_t *s_count, size_t base, uint4 x, uint4 halo_in,
                                           uint32_t lane, uint32_t first_valid, uint32_t &carry_e, uint4 &carry_halo,
                                           uint32_t &nul_acc, uint32_t &status) {
    // ---- A: stage the tile (input and default output = raw copy) -----------------------------------
    *reinterpret_cast<uint4 *>(s_in + lane * 16) = x;
    *reinterpret_cast<uint4 *>(s_out + lane * 16) = x;
    if (lane == 0) {
        *reinterpret_cast<uint4 *>(s_in + kTile) = halo_in;
        *reinterpret_cast<uint4 *>(s_out + kTile) = halo_in;
        *s_count = 0;
    }
    nul_acc |= ((x.x - 0x01010101u) & ~x.x) | ((x.y - 0x01010101u) & ~x.y) | ((x.z - 0x01010101u) & ~x.z) | ((x.w - 0x01010101u) & ~x.w);

    uint32_t overflow = 0;
    const Front f = tile_frontend<TAB>(T, s_in, x, lane, first_valid, carry_e, overflow);
    if (overflow) status |= ST_IRREGULAR;
    uint32_t real = f.real;
    const unsigned long long lens = f.lens;
    const uint32_t e_out = f.e_out;

    // ---- E: real matches overwrite their span of the output tile; the work is dealt round the warp ------
    if (lane == 0 && carry_e) {
        const uint32_t h[4] = {carry_halo.x, carry_halo.y, carry_halo.z, carry_halo.w};
        for (uint32_t k = 0; k < carry_e; ++k) s_out[k] = (uint8_t)(h[k >> 2] >> ((k & 3) * 8));
    }
    {
        // order within the list does not matter: one shared-memory atomic hands out the slots
        uint32_t slot = real ? atomicAdd(s_count, (uint32_t)__popc(real)) : 0u;
        while (real) {
            const uint32_t j = __ffs(real) - 1;
            real &= real - 1;
            s_list[slot++] = (uint16_t)((lane * 16 + j) | (((uint32_t)(lens >> (4 * j)) & 15u) << 12));
        }
        __syncwarp();
        const uint32_t total = *s_count;
        for (uint32_t q = lane; q < total; q += 32) {
            const uint32_t ent = s_list[q];
            emit_match<TAB>(a, T, s_in, s_out, ent & 0xFFFu, (ent >> 12) + 1, status);
        }
    }
    __syncwarp();
    if (STORE) {
        const size_t my = base + (size_t)lane * 16;
        uint4 y = *reinterpret_cast<const uint4 *>(s_out + lane * 16);
        if (my + 16 <= a.n) {
            *reinterpret_cast<uint4 *>(a.out + my) = y;
        } else if (my < a.n) {
            const uint32_t w[4] = {y.x, y.y, y.z, y.w};
            for (int k = 0; k < 16 && my + k < a.n; ++k) a.out[my + k] = (uint8_t)(w[k >> 2] >> ((k & 3) * 8));
        }
    }
    carry_e = __shfl_sync(0xFFFFFFFFu, e_out, 31);
    carry_halo = *reinterpret_cast<const uint4 *>(s_out + kTile);     // every lane reads it; only lane 0 uses it
    __syncwarp();
}

// ---- variable-length matches (PF_MATCHABLE_VAR): s/cat/mouse/, (cat):, dictionaries of unequal words ----
// Same front end; what changes is the output side: every uncovered non-match position prints its byte,
// every real match prints its string (offset | length from the eout table), covered positions print
// nothing -> per-span length sums (OP_COUNT), scan, then placement through a shared-memory window (OP_WRITE).

// offset(24) | length(8) of what the real match at tile position p prints
template <int TAB>
__device__ __forceinline__ uint32_t match_entry(const MatchArgs &a, const MatchTables &T, const uint8_t *s_in, uint32_t p) {
    uint32_t lo = T.row0_s[s_in[p]], i = 1;
    uint32_t from = 0, by = s_in[p];
    while (!(lo & DSA_E_DECIDED)) {
        const uint32_t c = s_in[p + i];
        const uint32_t nx = next_entry<TAB>(T, lo & 0xFFFFFFu, c);
        if ((nx & 0xFFFFFFu) == DSA_DEAD) break;
        from = lo & 0xFFFFFFu; by = c;
        lo = nx; ++i;
        if (i >= 16) break;
    }
    const uint32_t slot = (lo & DSA_E_DECIDED) ? from * T.n_classes + T.classmap_s[by]
                                               : a.P.n_states * T.n_classes + (lo & 0xFFFFFFu);
    return __ldg(a.P.eout + slot);
}

// The front end as a policy, so that the variable-length tiles serve both table forms: WalkFE walks the compact
// automaton (here), GramFE looks every attempt up in the gram tables (gram_kernel.cuh).
//   front()  steps B-D of a tile: which matches of the lane's 16 positions are real
//   entry()  offset(24) | length(8) of what the real match at tile position p prints, in pool()
template <int TAB>
struct WalkFE {
    static constexpr int kTab = TAB;
    const MatchArgs &a;
    const MatchTables &T;
    __device__ __forceinline__ const uint8_t *pool() const { return T.pool; }
    __device__ __forceinline__ Front front(const uint8_t *s_in, uint4 x, uint4 halo, uint32_t lane, uint32_t first_valid,
                                           uint32_t carry_e, uint32_t &overflow) const {
        (void)halo;
        return tile_frontend<TAB>(T, s_in, x, lane, first_valid, carry_e, overflow);
    }
    __device__ __forceinline__ uint32_t entry(const uint8_t *s_in, uint32_t p) const { return match_entry<TAB>(a, T, s_in, p); }
};

// positions of this lane covered by a match (its own real ones or the one hanging in from before)
__device__ __forceinline__ uint32_t covered_mask(const Front &f) {
    uint32_t cov = (1u << f.e_in) - 1u;
    uint32_t r = f.real;
    while (r) {
        const uint32_t j = __ffs(r) - 1;
        r &= r - 1;
        const uint32_t m = 1 + ((uint32_t)(f.lens >> (4 * j)) & 15u);
        cov |= ((1u << m) - 1u) << j;
    }
    return cov & 0xFFFFu;
}

struct VarWarp {
    uint8_t *s_in; uint32_t *s_count; uint32_t *s_ent; uint8_t *s_win;   // s_ent: tile position, then its entry
};

// hand the lane's real matches to the warp: list slots [slot0, slot0 + popc(real)), ascending position
__device__ __forceinline__ uint32_t publish_matches(const Front &f, const VarWarp &W, uint32_t lane) {
    uint32_t real = f.real;
    const uint32_t slot0 = real ? atomicAdd(W.s_count, (uint32_t)__popc(real)) : 0u;
    uint32_t slot = slot0;
    while (real) {
        const uint32_t j = __ffs(real) - 1;
        real &= real - 1;
        W.s_ent[slot++] = lane * 16 + j;
    }
    return slot0;
}

template <bool ACTIVE, class FE>
__device__ __forceinline__ void var_count_tile(const MatchArgs &a, const FE &fe, const VarWarp &W, size_t base, uint4 x,
                                               uint4 halo_in, uint32_t lane, uint32_t first_valid, uint32_t &carry_e,
                                               uint32_t &nul_acc, unsigned long long &span_sum) {
    *reinterpret_cast<uint4 *>(W.s_in + lane * 16) = x;
    if (lane == 0) { *reinterpret_cast<uint4 *>(W.s_in + kTile) = halo_in; *W.s_count = 0; }
    nul_acc |= ((x.x - 0x01010101u) & ~x.x) | ((x.y - 0x01010101u) & ~x.y) | ((x.z - 0x01010101u) & ~x.z) | ((x.w - 0x01010101u) & ~x.w);
    uint32_t overflow = 0;
    const Front f = fe.front(W.s_in, x, halo_in, lane, first_valid, carry_e, overflow);
    if (overflow) nul_acc |= 0x80u;                                   // reported as ST_IRREGULAR with the NUL detector
    carry_e = __shfl_sync(0xFFFFFFFFu, f.e_out, 31);
    if (ACTIVE) {
        const size_t my = base + (size_t)lane * 16;
        const uint32_t vmask = my + 16 <= a.n ? 0xFFFFu : (my < a.n ? (1u << (a.n - my)) - 1u : 0u);
        uint32_t mine = __popc(~covered_mask(f) & vmask);            // raw bytes
        publish_matches(f, W, lane);
        __syncwarp();
        const uint32_t total = *W.s_count;
        for (uint32_t q = lane; q < total; q += 32) mine += fe.entry(W.s_in, W.s_ent[q]) >> 24;
#pragma unroll
        for (int d = 16; d; d >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, d);
        span_sum += mine;
    }
    __syncwarp();
}

// the warp's output window: s_win[0] is the 16-byte aligned output address out + gpos; the first `foreign`
// bytes belong to the previous span and are never stored (same scheme as k_bmv_write)
struct OutWindow {
    unsigned long long gpos;
    uint32_t fill, foreign;
};

template <bool ACTIVE, class FE>
__device__ __forceinline__ void var_write_tile(const MatchArgs &a, const FE &fe, const VarWarp &W, size_t base, uint4 x,
                                               uint4 halo_in, uint32_t lane, uint32_t first_valid, uint32_t &carry_e,
                                               OutWindow &ow) {
    *reinterpret_cast<uint4 *>(W.s_in + lane * 16) = x;
    if (lane == 0) { *reinterpret_cast<uint4 *>(W.s_in + kTile) = halo_in; *W.s_count = 0; }
    uint32_t overflow = 0;                                            // the count pass has already reported it
    const Front f = fe.front(W.s_in, x, halo_in, lane, first_valid, carry_e, overflow);
    carry_e = __shfl_sync(0xFFFFFFFFu, f.e_out, 31);
    if (!ACTIVE) { __syncwarp(); return; }
    const uint32_t slot0 = publish_matches(f, W, lane);
    __syncwarp();
    {
        const uint32_t total = *W.s_count;
        for (uint32_t q = lane; q < total; q += 32) W.s_ent[q] = fe.entry(W.s_in, W.s_ent[q]);
    }
    __syncwarp();
    // lengths of this lane's 16 positions
    const size_t my = base + (size_t)lane * 16;
    const uint32_t vmask = my + 16 <= a.n ? 0xFFFFu : (my < a.n ? (1u << (a.n - my)) - 1u : 0u);
    const uint32_t rawmask = ~covered_mask(f) & vmask;
    uint32_t mine = __popc(rawmask);
    {
        const uint32_t nreal = __popc(f.real);
        for (uint32_t k = 0; k < nreal; ++k) mine += W.s_ent[slot0 + k] >> 24;
    }
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= (uint32_t)d) incl += y;
    }
    const uint32_t tile_total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
    const bool fits = ow.fill + tile_total <= (uint32_t)kVarWindow + 16;
    if (!fits) {                                                      // flush what is pending, then store directly
        for (uint32_t k = ow.foreign + lane; k < ow.fill; k += 32) a.out[ow.gpos + k] = W.s_win[k];
    }
    // place: positions in order; matches come in ascending position, like their list slots
    {
        uint32_t at = ow.fill + incl - mine, k = 0;
        uint8_t *dst = fits ? W.s_win : a.out + ow.gpos;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (f.real & (1u << j)) {
                const uint32_t ent = W.s_ent[slot0 + k++];
                const uint8_t *src = fe.pool() + (ent & 0xFFFFFFu);
                const uint32_t len = ent >> 24;
                for (uint32_t q = 0; q < len; ++q) dst[at + q] = src[q];
                at += len;
            } else if (rawmask & (1u << j)) {
                dst[at++] = (uint8_t)(w[j >> 2] >> ((j & 3) * 8));
            }
        }
    }
    if (fits) {
        ow.fill += tile_total;
        __syncwarp();
        const uint32_t nvec = ow.fill / 16;
        for (uint32_t v = lane; v < nvec; v += 32) {
            const uint4 y = *reinterpret_cast<const uint4 *>(W.s_win + v * 16);
            if (v == 0 && ow.foreign) {
                const uint32_t yy[4] = {y.x, y.y, y.z, y.w};
                for (uint32_t k = ow.foreign; k < 16; ++k) a.out[ow.gpos + k] = (uint8_t)(yy[k >> 2] >> ((k & 3) * 8));
            } else {
                *reinterpret_cast<uint4 *>(a.out + ow.gpos + (size_t)v * 16) = y;
            }
        }
        const uint32_t rest = ow.fill - nvec * 16;
        uint8_t keep = 0;
        if (lane < rest) keep = W.s_win[nvec * 16 + lane];
        __syncwarp();
        if (nvec) {
            if (lane < rest) W.s_win[lane] = keep;
            ow.gpos += (unsigned long long)nvec * 16;
            ow.fill = rest;
            ow.foreign = 0;
        }
    } else {
        const unsigned long long end = ow.gpos + ow.fill + tile_total;
        ow.foreign = (uint32_t)((reinterpret_cast<uintptr_t>(a.out) + end) & 15);
        ow.gpos = end - ow.foreign;
        ow.fill = ow.foreign;
    }
    __syncwarp();
}

// One warp's span of tiles, after the kernel has set its tables up: warm-up from the last '\n' before the span, then the
// tiles (in place / count / write).  Shared by k_match (WalkFE) and k_gram_var (GramFE).
template <int OP, class FE>
__device__ __forceinline__ void span_loop(const MatchArgs &a, const FE &fe, uint8_t *warp_base, uint32_t lane) {
    uint8_t *s_in = warp_base, *s_out = warp_base + 544;
    uint16_t *s_list = reinterpret_cast<uint16_t *>(warp_base + 1088);
    uint32_t *s_count = reinterpret_cast<uint32_t *>(warp_base + 1088 + 1024);
    VarWarp W;                                                    // layout of the variable-length forms
    W.s_in = warp_base;
    W.s_count = reinterpret_cast<uint32_t *>(warp_base + 544);
    W.s_ent = reinterpret_cast<uint32_t *>(warp_base + 544 + 16);
    W.s_win = warp_base + 544 + 16 + 2048;
    if (OP == OP_WRITE) {
        if (*a.status & ST_IRREGULAR) return;                      // the count pass saw a NUL / no final newline / long line
        if (*a.total > a.out_cap) {
            if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(a.status, ST_NOSPACE);
            return;
        }
    }
    const size_t n = a.n;
    const size_t n_tiles = (n + kTile - 1) / kTile;
    const size_t gwarp = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const size_t t0 = gwarp * a.tiles_per_warp;
    size_t t1 = t0 + a.tiles_per_warp;
    if (t1 > n_tiles) t1 = n_tiles;
    uint32_t status = 0, nul_acc = 0;
    if (OP == OP_COUNT && t0 >= t1 && lane == 0) a.gsum[gwarp] = 0;   // the scan runs over every warp of the grid
    if (t0 < t1) {
        uint32_t carry_e = 0;
        uint4 carry_halo = make_uint4(0, 0, 0, 0);
        unsigned long long span_sum = 0;
        OutWindow ow;
        ow.gpos = 0; ow.fill = 0; ow.foreign = 0;
        if (OP == OP_WRITE) {
            const unsigned long long span_off = a.gsum[gwarp];
            ow.foreign = (uint32_t)((reinterpret_cast<uintptr_t>(a.out) + span_off) & 15);
            ow.gpos = span_off - ow.foreign;
            ow.fill = ow.foreign;
        }
        // ---- warm-up: the chain state at t0 is only known from the last '\n' before it -------------
        size_t tw = t0;
        uint32_t first_valid = 0;
        if (t0 > 0) {
            bool found = false;
            for (int k = 1; k <= kWarmTiles && !found; ++k) {
                if ((size_t)k > t0) { tw = 0; first_valid = 0; found = true; break; }   // reached the buffer start
                const size_t tb = (t0 - k) * (size_t)kTile;
                uint4 v = ldg_vec(a.in + tb + lane * 16);                              // full tile: tb + 512 <= t0*512 <= n
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                int last = -1;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (((w[j >> 2] >> ((j & 3) * 8)) & 0xFFu) == '\n') last = j;
                uint32_t has = __ballot_sync(0xFFFFFFFFu, last >= 0);
                if (has) {
                    int src = 31 - __clz(has);
                    int lj = __shfl_sync(0xFFFFFFFFu, last, src);
                    uint32_t pos_in_tile = (uint32_t)src * 16 + (uint32_t)lj + 1;        // first byte after the '\n'
                    if (pos_in_tile >= (uint32_t)kTile) { tw = t0 - k + 1; first_valid = 0; }
                    else { tw = t0 - k; first_valid = pos_in_tile; }
                    found = true;
                }
            }
            if (!found) { status |= ST_IRREGULAR; tw = t0; }     // line longer than the window: general path
        }
        // ---- tiles -----------------------------------------------------------------------------------
        size_t base = tw * (size_t)kTile;
        uint4 x = load_padded(a.in, base + lane * 16, n);
        for (size_t t = tw; t < t1; ++t) {
            base = t * (size_t)kTile;
            uint4 halo_in = load_padded(a.in, base + kTile, n);
            uint4 x_next = make_uint4(0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au);
            if (t + 1 < t1) x_next = load_padded(a.in, base + kTile + lane * 16, n);   // prefetch
            if constexpr (OP == OP_INPLACE) {
                if (t < t0) match_tile<false, FE::kTab>(a, fe.T, s_in, s_out, s_list, s_count, base, x, halo_in, lane, first_valid, carry_e, carry_halo, nul_acc, status);
                else        match_tile<true, FE::kTab>(a, fe.T, s_in, s_out, s_list, s_count, base, x, halo_in, lane, first_valid, carry_e, carry_halo, nul_acc, status);
            } else if constexpr (OP == OP_COUNT) {
                if (t < t0) var_count_tile<false, FE>(a, fe, W, base, x, halo_in, lane, first_valid, carry_e, nul_acc, span_sum);
                else        var_count_tile<true, FE>(a, fe, W, base, x, halo_in, lane, first_valid, carry_e, nul_acc, span_sum);
            } else {
                if (t < t0) var_write_tile<false, FE>(a, fe, W, base, x, halo_in, lane, first_valid, carry_e, ow);
                else        var_write_tile<true, FE>(a, fe, W, base, x, halo_in, lane, first_valid, carry_e, ow);
            }
            first_valid = 0;
            x = x_next;
        }
        if (OP == OP_COUNT && lane == 0) a.gsum[gwarp] = span_sum;
        if (OP == OP_WRITE) {
            for (uint32_t k = ow.foreign + lane; k < ow.fill; k += 32) a.out[ow.gpos + k] = W.s_win[k];   // the span's last < 16 bytes
        }
    }
    if (nul_acc & 0x80808080u) status |= ST_IRREGULAR;
    if (blockIdx.x == 0 && threadIdx.x == 0 && n && a.in[n - 1] != '\n') status |= ST_IRREGULAR;
    if (status) atomicOr(a.status, status);
}

template <int TAB, int OP>
__global__ void __launch_bounds__(kMatchThreads)
k_match(MatchArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // layout: [live 32 KB]? [row0 1 KB] [classmap 256] [pool copy]? [ctrans rows] [per warp: s_in, s_out, list]
    uint8_t *cur = smem;
    uint32_t *live_s = nullptr;
    if (a.row0_replicated) { live_s = reinterpret_cast<uint32_t *>(cur); cur += 256 * 32 * 4; }
    uint32_t *row0_s = reinterpret_cast<uint32_t *>(cur); cur += 1024;
    uint8_t *classmap_s = cur; cur += 256;
    uint8_t *pool_s = nullptr;
    if (a.pool_in_smem) { pool_s = cur; cur += (a.P.pool_bytes + 15) & ~15u; }
    uint32_t *ctrans_s = reinterpret_cast<uint32_t *>(cur);
    const uint32_t ct_words = TAB == TAB_WIDE ? a.P.n_states * 256u : a.smem_states * a.P.n_classes;
    cur += ((ct_words + 3) & ~3u) * 4;
    uint8_t *warp_base = cur + warp * (OP == OP_INPLACE ? kWarpSmem : kVarWarpSmem);

    if (live_s)
        for (uint32_t k = threadIdx.x; k < 256u * 32u; k += blockDim.x)
            live_s[k] = (__ldg(a.P.row0 + (k >> 5)) & 0xFFFFFFu) != DSA_DEAD ? 1u : 0u;
    for (uint32_t k = threadIdx.x; k < 256; k += blockDim.x) { row0_s[k] = __ldg(a.P.row0 + k); classmap_s[k] = __ldg(a.P.classmap + k); }
    if (pool_s)
        for (uint32_t k = threadIdx.x; k < a.P.pool_bytes; k += blockDim.x) pool_s[k] = __ldg(a.P.pool + k);
    if (TAB == TAB_WIDE) {
        for (uint32_t k = threadIdx.x; k < ct_words; k += blockDim.x)
            ctrans_s[k] = __ldg(a.P.ctrans + (k >> 8) * a.P.n_classes + __ldg(a.P.classmap + (k & 255u)));
    } else {
        for (uint32_t k = threadIdx.x; k < ct_words; k += blockDim.x) ctrans_s[k] = __ldg(a.P.ctrans + k);
    }
    __syncthreads();

    MatchTables T;
    T.live_s = live_s; T.row0_s = row0_s; T.classmap_s = classmap_s; T.ctrans_s = ctrans_s; T.ctrans_g = a.P.ctrans;
    T.pool = pool_s ? pool_s : a.P.pool;
    T.n_classes = a.P.n_classes; T.smem_states = a.smem_states; T.lane = lane;
    T.capped = (a.P.flags & PF_ACYCLIC) ? 0u : 1u;

    span_loop<OP>(a, WalkFE<TAB>{a, T}, warp_base, lane);
}

}  // namespace trre
