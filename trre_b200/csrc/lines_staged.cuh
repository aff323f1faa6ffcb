// lines_staged.cuh - k_lines_s: the general line path (count / write) with everything hot in shared memory.
//
// Same decomposition as k_lines (kernels.cuh): one 128-byte chunk per thread, the thread owns the lines
// that START in its chunk, 256 chunks = one 32 KB group per CTA iteration, counts -> scan -> write.
// What changes is where the bytes come from:
//   * the group's 32 KB of input are loaded once with coalesced 16-byte vectors into shared memory; the
//     per-thread byte walk then reads shared memory (a line that runs past the group end falls through
//     to global memory),
//   * the walk uses the compact u32 transition table (low words), its first rows cached in shared memory,
//     and only a successful attempt replays the full table for its output strings,
//   * in write mode the group's output is assembled in shared memory and leaves with coalesced vector
//     stores (if it fits; pathological expansions write straight to global memory).
// Table executor only; the DFS executor keeps k_lines.
#pragma once
#include <cstdint>

#include "kernels.cuh"

namespace trre {

constexpr int kGroupBytes = kChunk * kLineThreads;      // 32 KB
constexpr int kStageOut = 40 * 1024;                     // output staging per group (write mode), + 16 for alignment

// All positions inside one group iteration are 32-bit offsets relative to the group start (`rel`): a record
// that starts in the group may run past the staged 32 KB (then bytes come from global memory), but not by 4 GB.
struct StagedView {
    const uint8_t *s_in;         // shared copy of the group's bytes [0, tile_len)
    const uint8_t *g_tile;       // global address of rel == 0
    uint32_t tile_len;
    uint32_t lim_rel, n_rel;     // end of payload bytes / end of buffer, relative (clamped to 0xFFFFFFF0)
    const uint32_t *row0_s;      // first transition by raw byte
    const uint32_t *s_mark;      // bit per staged byte: 1 = terminator (NUL, newline) or a byte an attempt can start with
    const uint8_t *classmap_s;
    const uint32_t *ctrans_s;
    const uint32_t *ctrans_g;
    uint32_t smem_states, n_classes;
    __device__ __forceinline__ uint8_t at(uint32_t rel) const {
        return rel < tile_len ? s_in[rel] : __ldg(g_tile + rel);
    }
    __device__ __forceinline__ uint32_t step(uint32_t s, uint8_t c) const {
        const uint32_t idx = s * n_classes + classmap_s[c];
        return s < smem_states ? ctrans_s[idx] : __ldg(ctrans_g + idx);
    }
};

// where output bytes go: `n` counts from the group's first output byte; staged when it fits the window
struct StagedSink {
    uint8_t *g_out;              // global address of n == 0; nullptr: count only
    uint8_t *s_out;              // shared address of n == 0; nullptr: not staged
    uint32_t n;
    __device__ __forceinline__ void put(uint8_t b) {
        if (s_out) s_out[n] = b;
        else if (g_out) g_out[n] = b;
        ++n;
    }
    // a run of bytes no attempt can start with: copied verbatim (shared memory source)
    __device__ __forceinline__ void put_run(const uint8_t *src, uint32_t len) {
        if (s_out) { for (uint32_t k = 0; k < len; ++k) s_out[n + k] = src[k]; }
        else if (g_out) { for (uint32_t k = 0; k < len; ++k) g_out[n + k] = src[k]; }
        n += len;
    }
    __device__ __forceinline__ void put_ref(const uint8_t *pool, uint32_t hi) {
        uint32_t len = hi >> 24, off = hi & 0xFFFFFFu;
        if (len == 255u) { len = *reinterpret_cast<const uint32_t *>(pool + off); off += 4; }
        if (s_out) { for (uint32_t k = 0; k < len; ++k) s_out[n + k] = __ldg(pool + off + k); }
        else if (g_out) { for (uint32_t k = 0; k < len; ++k) g_out[n + k] = __ldg(pool + off + k); }
        n += len;
    }
};

// the rest of an attempt whose first transition `lo` (from row0) is alive; see automaton.hpp for the rule
__device__ __forceinline__ int attempt_tail(const DevProg &P, const StagedView &V, uint32_t p, uint32_t lo,
                                            StagedSink &o, uint32_t &status) {
    uint32_t i = 1;
    int m = (lo & DSA_E_MARK) ? 1 : ((P.start_info & DSA_S_HAS_FINAL) ? 0 : -1);
    while (!(lo & DSA_E_DECIDED) && p + i < V.lim_rel) {
        const uint32_t nx = V.step(lo & 0xFFFFFFu, V.at(p + i));
        if ((nx & 0xFFFFFFu) == DSA_DEAD) break;
        lo = nx;
        ++i;
        if (lo & DSA_E_MARK) m = (int)i;
    }
    if (lo & DSA_E_DIVERGE) { status |= ST_DIVERGE; return -1; }
    if (!(lo & DSA_E_HASFINAL)) return -1;
    // replay with the full table: out(e_1) .. out(e_i) + fin(state)
    uint32_t t = 0;
    for (uint32_t k = 0; k < i; ++k) {
        const uint64_t e = __ldg(P.trans + (size_t)t * P.n_classes + V.classmap_s[V.at(p + k)]);
        o.put_ref(P.pool, (uint32_t)(e >> 32));
        t = (uint32_t)e & 0xFFFFFFu;
    }
    o.put_ref(P.pool, (uint32_t)(__ldg(P.sinfo + t) >> 32));
    return m;
}

// The reference's per-record loop (trre_nft.c:778-789 / trre_dft.c:1275-1285) for the record whose payload
// starts at rel p.  Returns the rel just past the record.
__device__ __forceinline__ uint32_t run_record_staged(const DevProg &P, const StagedView &V, uint32_t p,
                                                      StagedSink &o, uint32_t &status) {
    const bool start_accepts = (P.start_info & DSA_S_HAS_FINAL) != 0;   // an attempt can succeed without consuming
    const bool can_skip = (P.start_info & (DSA_S_HAS_FINAL | DSA_S_DIVERGE)) == 0;
    for (;;) {
        if (can_skip && p < V.tile_len) {
            // every attempt between here and the next marked byte fails on its first byte: one raw run
            uint32_t q = p, word = V.s_mark[p >> 5] >> (p & 31);
            if (word) q = p + (uint32_t)__ffs(word) - 1;
            else {
                q = (p | 31u) + 1;
                while (q < V.tile_len && V.s_mark[q >> 5] == 0) q += 32;
                if (q < V.tile_len) q += (uint32_t)__ffs(V.s_mark[q >> 5]) - 1;
            }
            if (q > V.tile_len) q = V.tile_len;
            if (q > V.lim_rel) q = V.lim_rel;                         // never past the payload
            if (q > p) { o.put_run(V.s_in + p, q - p); p = q; }
        }
        uint8_t c = '\n';
        if (p < V.lim_rel) c = V.at(p);
        const bool at_end = (c == 0 || c == '\n');
        const uint32_t lo = at_end ? DSA_DEAD : V.row0_s[c];
        int k = -1;
        if ((lo & 0xFFFFFFu) != DSA_DEAD) {
            k = attempt_tail(P, V, p, lo, o, status);
        } else if (P.start_info & DSA_S_DIVERGE) {
            status |= ST_DIVERGE;                                     // nothing of higher priority survives: epsilon cycle
        } else if (start_accepts) {
            // the first byte is dead (or the payload ended) but the start state already accepts: empty match
            o.put_ref(P.pool, (uint32_t)(__ldg(P.sinfo) >> 32));
            k = 0;
        }
        if (at_end) break;                                            // that was the attempt on the empty tail
        if (k > 0) p += (uint32_t)k;
        else { o.put(c); ++p; }
    }
    o.put('\n');
    while (p < V.n_rel && V.at(p) != '\n') ++p;                       // rest of the record after a NUL
    return p + 1;
}

struct StagedArgs {
    LineArgs L;
    uint32_t smem_states;
};

template <int MODE>
__global__ void __launch_bounds__(kLineThreads)
k_lines_s(StagedArgs sa) {
    const LineArgs &a = sa.L;
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ unsigned long long warp_sums[kLineThreads / 32];
    __shared__ unsigned long long group_base_s, group_total_s;
    uint8_t *classmap_s = smem;
    uint32_t *row0_s = reinterpret_cast<uint32_t *>(smem + 256);
    uint32_t *ctrans_s = row0_s + 256;
    const uint32_t ct_words = sa.smem_states * a.P.n_classes;
    uint8_t *s_in = reinterpret_cast<uint8_t *>(ctrans_s + ((ct_words + 3) & ~3u));
    uint32_t *s_mark = reinterpret_cast<uint32_t *>(s_in + kGroupBytes + 16);     // kGroupBytes / 32 words
    uint8_t *s_out = reinterpret_cast<uint8_t *>(s_mark + kGroupBytes / 32);

    const uint8_t *__restrict__ in = a.in;
    const size_t n = a.n;
    if (n == 0) return;
    if (MODE == MODE_WRITE && *a.total > a.out_cap) {
        if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(a.status, ST_NOSPACE);
        return;
    }
    for (uint32_t k = threadIdx.x; k < 256; k += blockDim.x) { classmap_s[k] = __ldg(a.P.classmap + k); row0_s[k] = __ldg(a.P.row0 + k); }
    for (uint32_t k = threadIdx.x; k < ct_words; k += blockDim.x) ctrans_s[k] = __ldg(a.P.ctrans + k);
    const bool ends_nl = __ldg(in + n - 1) == '\n';
    const size_t lim = ends_nl ? n : n - 1;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(in) & 15) == 0;
    uint32_t status = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (size_t g = blockIdx.x; g < a.n_groups; g += gridDim.x) {
        const size_t tile_beg = g * (size_t)kGroupBytes;
        const size_t tile_len = (tile_beg + kGroupBytes <= n) ? (size_t)kGroupBytes : n - tile_beg;
        __syncthreads();                                              // the previous group is done with smem
        // ---- stage the group's input --------------------------------------------------------------
        if (vec_ok) {
            const size_t nv = tile_len / 16;
            for (size_t v = threadIdx.x; v < nv; v += blockDim.x)
                *reinterpret_cast<uint4 *>(s_in + v * 16) = ld_stream(reinterpret_cast<const uint4 *>(in + tile_beg) + v);
            for (size_t k = nv * 16 + threadIdx.x; k < tile_len; k += blockDim.x) s_in[k] = __ldg(in + tile_beg + k);
        } else {
            for (size_t k = threadIdx.x; k < tile_len; k += blockDim.x) s_in[k] = __ldg(in + tile_beg + k);
        }
        __syncthreads();
        // ---- mark the bytes where the record loop has something to decide: terminators and bytes whose first
        // transition is alive; the runs in between are copied verbatim without an attempt per byte
        {
            const uint32_t cb = threadIdx.x * kChunk;                 // this thread's 128 bytes = 4 mark words
#pragma unroll
            for (int wd = 0; wd < kChunk / 32; ++wd) {
                uint32_t bits = 0;
                const uint32_t b0 = cb + wd * 32;
                if (b0 < tile_len) {
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        const uint4 y = *reinterpret_cast<const uint4 *>(s_in + b0 + v * 16);
                        const uint32_t ww[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const uint32_t c = (ww[j >> 2] >> ((j & 3) * 8)) & 0xFFu;
                            const uint32_t hot = (c == 0u || c == '\n' || (row0_s[c] & 0xFFFFFFu) != DSA_DEAD) ? 1u : 0u;
                            bits |= hot << (v * 16 + j);
                        }
                    }
                    if (b0 + 32 > tile_len) bits |= ~0u << (tile_len - b0);   // bytes beyond the tile: decide one by one
                }
                s_mark[(cb >> 5) + wd] = bits;
            }
        }
        const size_t chunk = g * kLineThreads + threadIdx.x;
        unsigned long long my_off = 0;
        if (MODE == MODE_WRITE) {
            unsigned long long v = (chunk < a.n_chunks) ? a.cnt[chunk] : 0, x = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, x, d);
                if (lane >= d) x += y;
            }
            if (lane == 31) warp_sums[warp] = x;
            if (threadIdx.x == 0) group_base_s = a.gsum[g];
            __syncthreads();
            unsigned long long wbase = 0, tot = 0;
            for (int w = 0; w < kLineThreads / 32; ++w) { if (w < warp) wbase += warp_sums[w]; tot += warp_sums[w]; }
            if (threadIdx.x == 0) group_total_s = tot;
            my_off = group_base_s + wbase + x - v;
        }
        __syncthreads();

        StagedView V;
        V.s_in = s_in; V.g_tile = in + tile_beg; V.tile_len = (uint32_t)tile_len;
        {
            const size_t lr = lim - tile_beg, nr = n - tile_beg, cap = 0xFFFFFFF0u;
            V.lim_rel = (uint32_t)(lr < cap ? lr : cap);
            V.n_rel = (uint32_t)(nr < cap ? nr : cap);
            if (nr >= cap) status |= 0;                               // (a single record cannot span 4 GB: see ST_NOSPACE below)
        }
        V.row0_s = row0_s; V.s_mark = s_mark; V.classmap_s = classmap_s; V.ctrans_s = ctrans_s; V.ctrans_g = a.P.ctrans;
        V.smem_states = sa.smem_states; V.n_classes = a.P.n_classes;
        StagedSink o;
        o.g_out = nullptr; o.s_out = nullptr;
        bool staged = false;
        uint32_t my_rel = 0;                                           // this thread's first output byte, relative to the group
        if (MODE == MODE_WRITE) {
            staged = group_total_s <= (unsigned long long)kStageOut;
            // staged byte k sits at s_out[shift + k]: shared and global addresses are congruent mod 16
            const uint32_t shift = (uint32_t)((reinterpret_cast<uintptr_t>(a.out) + group_base_s) & 15);
            o.g_out = a.out + group_base_s;
            if (staged) o.s_out = s_out + shift;
            my_rel = (uint32_t)(my_off - group_base_s);
            if (group_total_s > 0xFFFFFFF0ull) status |= ST_NOSPACE;
        }
        o.n = my_rel;

        const uint32_t cbeg = threadIdx.x * kChunk;
        const uint32_t cend = (cbeg + kChunk < V.n_rel) ? cbeg + kChunk : V.n_rel;
        uint32_t produced = 0;
        if (cbeg < V.n_rel) {
            uint32_t p = cbeg;
            const bool starts_here = (tile_beg == 0 && cbeg == 0) || (cbeg > 0 ? s_in[cbeg - 1] == '\n' : __ldg(in + tile_beg - 1) == '\n');
            if (!starts_here) {
                while (p < cend && V.at(p) != '\n') ++p;
                ++p;
            }
            while (p < cend) {
                const uint32_t before = o.n;
                p = run_record_staged(a.P, V, p, o, status);
                produced += o.n - before;
                if (p >= 0xFFFFFF00u) { status |= ST_NOSPACE; break; }  // a record spanning 4 GB
            }
        }
        if (MODE == MODE_COUNT) {
            if (chunk < a.n_chunks) a.cnt[chunk] = produced;
            unsigned long long x = produced;
#pragma unroll
            for (int d = 16; d; d >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, d);
            if (lane == 0) warp_sums[warp] = x;
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned long long s = 0;
                for (int w = 0; w < kLineThreads / 32; ++w) s += warp_sums[w];
                a.gsum[g] = s;
            }
        }
        if (MODE == MODE_WRITE && staged) {
            __syncthreads();
            // ---- the staged output leaves with 16-byte stores where the destination is aligned -------
            uint8_t *dst = a.out + group_base_s;
            const uint8_t *src = o.s_out;
            const size_t tot = (size_t)group_total_s;
            size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
            if (head > tot) head = tot;
            for (size_t k = threadIdx.x; k < head; k += blockDim.x) dst[k] = src[k];
            const size_t nv = (tot - head) / 16;
            for (size_t v = threadIdx.x; v < nv; v += blockDim.x)
                st_stream(reinterpret_cast<uint4 *>(dst + head) + v, *reinterpret_cast<const uint4 *>(src + head + v * 16));
            for (size_t k = head + nv * 16 + threadIdx.x; k < tot; k += blockDim.x) dst[k] = src[k];
        }
    }
    if (status) atomicOr(a.status, status);
}

}  // namespace trre
