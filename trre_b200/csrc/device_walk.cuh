// device_walk.cuh - per-line executors shared by the line kernels (sm_100a).
//
// One "attempt" = what one call of infer_dft (trre_dft.c:1110-1196) or infer_backtrack
// (trre_nft.c:593-657) does; one "line" = the per-record loop of the reference's scan drivers
// (trre_nft.c:778-789 / trre_dft.c:1275-1285).
#pragma once
#include <cstdint>

#include "automaton.hpp"

namespace trre {

// status word bits written by kernels
enum : uint32_t {
    ST_IRREGULAR   = 1u,    // fast path assumptions broken (NUL byte, no trailing '\n'): rerun generally
    ST_DIVERGE     = 2u,    // an attempt entered an epsilon cycle (reference: stack max capacity)
    ST_STACK       = 4u,    // DFS stack of the backtracking executor overflowed
    ST_OBUF        = 8u,    // per-attempt output buffer of the backtracking executor overflowed
    ST_NOSPACE     = 16u,   // d_out too small
    ST_MISS        = 32u,   // lazy tables: a row the input needs has not been built yet (see LazyDev)
    ST_RUNAWAY     = 64u,   // lazy tables: the input leads into a closure that does not terminate (LAZY_POISON)
    ST_BUDGET      = 128u   // DFS executor: one attempt went past kDfsStepBudget steps or kDfsPrintBudget printed bytes
};

// One attempt of the reference's DFS may backtrack exponentially ((a*)*b on aaaa...) and, with -a, print without bound
// (generators): the reference then simply runs for hours.  A GPU thread gets a budget per RECORD instead; beyond it the
// scan fails with its own error text (status_to_rc), never with wrong bytes.
constexpr uint32_t kDfsStepBudget  = 1u << 26;
constexpr uint32_t kDfsPrintBudget = 1u << 24;

// TRRE_EXEC_LAZY: rows by raw byte, built on demand by the host between passes (automaton.hpp, LazyDsa)
struct LazyDev {
    const uint64_t *rows;          // [n_states * 256]
    const uint64_t *sinfo;         // [n_states]
    const uint8_t  *pool;
    uint32_t *requested;           // [state capacity]: 1 once a state's row has been asked for
    uint32_t *miss_list;           // [miss_cap] states whose row is missing
    uint32_t *miss_count;
    uint32_t  miss_cap;
    uint32_t  start_info;
};

struct DevProg {
    const uint8_t  *classmap;
    const uint8_t  *lut;
    const uint16_t *pairlut;       // TRRE_EXEC_BYTEMAP only
    const uint32_t *row0;          // k_match: first step by raw byte
    const uint32_t *ctrans;        // k_match: low words of trans
    const uint32_t *eout;          // k_match: pool offsets of what a match prints (edge table, then per-state)
    uint32_t pool_bytes;
    uint32_t n_states;
    uint32_t max_depth;
    const uint64_t *trans;
    const uint64_t *sinfo;
    const uint8_t  *pool;
    const NftFlat  *nft;
    uint32_t n_classes;
    uint32_t start_info;
    uint32_t flags;
    int32_t  nft_start;
};

struct Sink {
    uint8_t *p;         // nullptr: count only
    size_t   n;
    __device__ __forceinline__ void put(uint8_t b) { if (p) p[n] = b; ++n; }
};

__device__ __forceinline__ uint32_t ref_len(const uint8_t *pool, uint32_t hi) {
    uint32_t len = hi >> 24;
    if (len == 255u) len = *reinterpret_cast<const uint32_t *>(pool + (hi & 0xFFFFFFu));
    return len;
}

__device__ __forceinline__ void put_ref(Sink &o, const uint8_t *pool, uint32_t hi) {
    uint32_t len = hi >> 24, off = hi & 0xFFFFFFu;
    if (len == 255u) { len = *reinterpret_cast<const uint32_t *>(pool + off); off += 4; }
    if (o.p) for (uint32_t k = 0; k < len; ++k) o.p[o.n + k] = pool[off + k];
    o.n += len;
}

// ---- table executor -----------------------------------------------------------------------------
// Attempt at in[p..lim).  Returns bytes consumed (>=0) on a match, -1 on failure; output appended
// to `o`.  Bytes 0 and '\n' are in the all-dead class, so the walk stops on them by itself.
__device__ __forceinline__ long attempt_table(const DevProg &P, const uint8_t *__restrict__ in,
                                              size_t p, size_t lim, Sink &o, uint32_t &status) {
    uint32_t s = 0, i = 0, acc = 0;
    long m = (P.start_info & DSA_S_HAS_FINAL) ? 0 : -1;
    bool decided = (P.start_info & DSA_S_DECIDED) != 0;
    if (!decided) {
        while (p + i < lim) {
            uint8_t c = __ldg(in + p + i);
            uint64_t e = __ldg(P.trans + (size_t)s * P.n_classes + __ldg(P.classmap + c));
            uint32_t lo = (uint32_t)e;
            if ((lo & 0xFFFFFFu) == DSA_DEAD) break;
            acc += ref_len(P.pool, (uint32_t)(e >> 32));
            s = lo & 0xFFFFFFu;
            ++i;
            if (lo & DSA_E_MARK) m = i;
            if (lo & DSA_E_DECIDED) { decided = true; break; }
        }
    }
    uint64_t si = __ldg(P.sinfo + s);
    uint32_t sf = (uint32_t)si;
    if (sf & DSA_S_DIVERGE) {
        // the walk only stops here when nothing of higher priority is left (decided, or end of
        // payload): the reference's DFS is now inside the cycle and dies with "stack max capacity reached"
        status |= ST_DIVERGE;
        return -1;
    }
    if (!(sf & DSA_S_HAS_FINAL)) return -1;
    // emit acc (replay the i transitions) + fin
    if (o.p) {
        uint32_t t = 0;
        for (uint32_t k = 0; k < i; ++k) {
            uint8_t c = __ldg(in + p + k);
            uint64_t e = __ldg(P.trans + (size_t)t * P.n_classes + __ldg(P.classmap + c));
            put_ref(o, P.pool, (uint32_t)(e >> 32));
            t = (uint32_t)e & 0xFFFFFFu;
        }
    } else {
        o.n += acc;
    }
    put_ref(o, P.pool, (uint32_t)(si >> 32));
    return m;
}

// ---- lazy-table executor ------------------------------------------------------------------------------
// attempt_table over rows indexed by the raw byte; a row that does not exist yet is reported (once per state) and
// the attempt gives up with ST_MISS: the host builds the row and the pass is repeated.  Returns -2 on a miss.
__device__ __forceinline__ long attempt_lazy(const LazyDev &L, const uint8_t *__restrict__ in,
                                             size_t p, size_t lim, Sink &o, uint32_t &status) {
    uint32_t s = 0, i = 0, acc = 0;
    long m = (L.start_info & DSA_S_HAS_FINAL) ? 0 : -1;
    bool decided = (L.start_info & DSA_S_DECIDED) != 0;
    if (!decided) {
        while (p + i < lim) {
            const uint8_t c = __ldg(in + p + i);
            if (c == 0 || c == '\n') break;                 // never consumed (rows are built for payload bytes only)
            const uint64_t e = L.rows[(size_t)s * 256 + c];
            const uint32_t lo = (uint32_t)e;
            if ((lo & 0xFFFFFFu) == LAZY_UNBUILT) {
                if (atomicExch(L.requested + s, 1u) == 0u) {
                    const uint32_t k = atomicAdd(L.miss_count, 1u);
                    if (k < L.miss_cap) L.miss_list[k] = s;
                    else L.requested[s] = 0;                 // list full: ask again next pass
                }
                status |= ST_MISS;
                return -2;
            }
            if ((lo & 0xFFFFFFu) == LAZY_POISON) { status |= ST_RUNAWAY; return -1; }
            if ((lo & 0xFFFFFFu) == DSA_DEAD) break;
            acc += ref_len(L.pool, (uint32_t)(e >> 32));
            s = lo & 0xFFFFFFu;
            ++i;
            if (lo & DSA_E_MARK) m = i;
            if (lo & DSA_E_DECIDED) { decided = true; break; }
        }
    }
    const uint64_t si = L.sinfo[s];
    const uint32_t sf = (uint32_t)si;
    if (sf & DSA_S_DIVERGE) { status |= ST_DIVERGE; return -1; }
    if (!(sf & DSA_S_HAS_FINAL)) return -1;
    if (o.p) {
        uint32_t t = 0;
        for (uint32_t k = 0; k < i; ++k) {
            const uint64_t e = L.rows[(size_t)t * 256 + __ldg(in + p + k)];
            put_ref(o, L.pool, (uint32_t)(e >> 32));
            t = (uint32_t)e & 0xFFFFFFu;
        }
    } else {
        o.n += acc;
    }
    put_ref(o, L.pool, (uint32_t)(si >> 32));
    return m;
}

// ---- backtracking executor (NFT flavour fallback) ---------------------------------------------------
struct BtScratch {
    int32_t  *stk_s;      // [cap]
    uint32_t *stk_i;      // [cap]
    uint32_t *stk_o;      // [cap]
    uint8_t  *obuf;       // [ocap]
    uint32_t  cap, ocap;
    mutable uint32_t steps, printed;   // spent by the current RECORD (run_record resets them): see kDfsStepBudget
};

// What one attempt of the DFS does at a FINAL state (trre_nft.c:635-648):
//   AM_FIRST      scan mode: print, return the consumed length
//   AM_SCAN_ALL   scan mode with -a: print, keep searching; every attempt "fails" (-1), so the driver copies the raw byte
//   AM_MATCH      -m: only a FINAL at the end of the record counts; print + '\n', stop
//   AM_MATCH_ALL  -ma: the same, keep searching
//   AM_NEWLINES   `trre_dft -m`: one bare '\n' per record (trre_dft.c:1287-1292; infer_dft prints nothing in MATCH mode)
enum { AM_FIRST = 0, AM_SCAN_ALL = 1, AM_MATCH = 2, AM_MATCH_ALL = 3, AM_NEWLINES = 4 };

// priority DFS over the flat NFT from in[p..lim) (trre_nft.c:598-655)
template <int AMODE = AM_FIRST>
__device__ __forceinline__ long attempt_backtrack(const DevProg &P, const BtScratch &B,
                                                  const uint8_t *__restrict__ in, size_t p, size_t lim,
                                                  Sink &o, uint32_t &status) {
    uint32_t i = 0, w = 0, top = 0;
    int32_t s = P.nft_start;
    const uint32_t len = (uint32_t)(lim - p);
    while (top || s >= 0) {
        if (s < 0) {
            --top;
            s = B.stk_s[top]; i = B.stk_i[top]; w = B.stk_o[top];
            if (s < 0) continue;
        }
        if (++B.steps > kDfsStepBudget) { status |= ST_BUDGET; return -1; }
        NftFlat n = P.nft[s];
        switch (n.op) {
            case N_CONS:
                if (i < len) {
                    uint8_t c = __ldg(in + p + i);
                    if (c != 0 && c != '\n' && c == n.val) { ++i; s = n.a; break; }
                }
                s = -1;
                break;
            case N_PROD:
                if (w >= B.ocap) { status |= ST_OBUF; return -1; }
                B.obuf[w++] = n.val;
                s = n.a;
                break;
            case N_SPLIT:
            case N_SPLITNG:
                if (top >= B.cap) { status |= ST_STACK; return -1; }
                B.stk_s[top] = (n.op == N_SPLIT) ? n.a : n.b;
                B.stk_i[top] = i; B.stk_o[top] = w; ++top;
                s = (n.op == N_SPLIT) ? n.b : n.a;
                break;
            case N_JOIN:
                s = n.a;
                break;
            default: {   // N_FINAL: fputs(output) prints up to the first produced NUL (trre_nft.c:644-645)
                bool accept = true;
                if (AMODE == AM_MATCH || AMODE == AM_MATCH_ALL) {        // input[i] == '\0' (trre_nft.c:636)
                    if (i < len) { const uint8_t c = __ldg(in + p + i); accept = (c == 0 || c == '\n'); }
                }
                if (accept) {
                    if (AMODE != AM_FIRST && AMODE != AM_MATCH) {
                        B.printed += w + 1;
                        if (B.printed > kDfsPrintBudget) { status |= ST_BUDGET; return -1; }
                    }
                    for (uint32_t k = 0; k < w; ++k) {
                        uint8_t b = B.obuf[k];
                        if (!b) break;
                        o.put(b);
                    }
                    if (AMODE == AM_MATCH || AMODE == AM_MATCH_ALL) o.put('\n');
                    if (AMODE == AM_FIRST || AMODE == AM_MATCH) return (long)i;
                }
                s = -1;
                break;
            }
        }
    }
    return -1;
}

}  // namespace trre
