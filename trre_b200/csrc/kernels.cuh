// kernels.cuh - sm_100a kernels of the scan path (1/4).  The kernel families, by what serves a REGULAR
// buffer (no NUL byte, final newline) - see trre_prog_info.fast_path:
//
//  k_bytemap      (here)               one byte in, one byte out: 16-byte vector loads/stores, 128 KB
//                                      two-bytes-at-once table staged into shared memory by TMA
//  k_match        (match_kernel.cuh)   position-parallel matcher for acyclic programs with short attempts:
//                                      in place when length preserving, count/scan/write otherwise
//  k_bmv_*        (bmv_kernel.cuh)     one byte in, 0..255 bytes out (deletions, expansions)
//  k_lines_s      (lines_staged.cuh)   the reference's per-line loop, shared-memory staged: everything else
//                                      and every IRREGULAR buffer (the fast kernels raise ST_IRREGULAR
//                                      themselves and the engine reruns the buffer here)
//  k_lines        (here)               the same loop without staging: in-place form for length-preserving
//                                      cyclic programs, and the DFS executor (TRRE_EXEC_BACKTRACK)
//                   modes: MODE_INPLACE  out offset == in offset
//                          MODE_COUNT    output bytes per 128-byte chunk -> cnt[], per-group sums -> gsum[]
//                          MODE_WRITE    offsets = scan(gsum)[group] + block-scan(cnt); write
//  k_scan_groups  (here)               exclusive scan of per-group / per-span sums (one CTA) + grand total
#pragma once
#include <cstdint>

#include "device_walk.cuh"

namespace trre {

constexpr int kChunk = 128;            // input bytes per thread in k_lines
constexpr int kLineThreads = 256;      // threads per CTA in k_lines = chunks per group

// ------------------------------------------------------------------------------------------------
// k_bytemap
// ------------------------------------------------------------------------------------------------
// Persistent: one 1024-thread CTA per SM.  The 128 KB pair table (LUT applied to two bytes at once,
// built on the host, part of the program blob) is staged into shared memory with TMA bulk copies
// (cp.async.bulk + mbarrier) while the first input vectors are already in flight; every thread then
// streams 16-byte vectors: 2 shared-memory lookups + 1 PRMT per 4 bytes.
constexpr int kBmThreads = 1024;
constexpr int kBmUnroll = 4;           // 16-byte vectors in flight per thread
constexpr int kPairLutBytes = 65536 * 2;

__device__ __forceinline__ uint4 ld_stream(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(uint4 *p, const uint4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
// TMA bulk copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// one 32-bit word = four input bytes -> four output bytes; also accumulates the zero-byte detector
__device__ __forceinline__ uint32_t map_word(uint32_t w, uint32_t tab, uint32_t &nul) {
    uint32_t lo, hi;
    // byte offset of pair entry (c0 | c1 << 8), bank bits XOR-swizzled with c1's low 5 bits (see automaton.cpp)
    uint32_t a0 = tab + (((w << 1) & 0x1FFFEu) ^ ((w >> 6) & 0x7Cu));
    uint32_t a1 = tab + (((w >> 15) & 0x1FFFEu) ^ ((w >> 22) & 0x7Cu));
    asm("ld.shared.u16 %0, [%1];" : "=r"(lo) : "r"(a0));
    asm("ld.shared.u16 %0, [%1];" : "=r"(hi) : "r"(a1));
    nul |= (w - 0x01010101u) & ~w;                        // bit 7 of a byte set <=> that byte is 0 (masked later)
    return __byte_perm(lo, hi, 0x5410);
}

__device__ __forceinline__ uint4 map_vec(const uint4 &x, uint32_t tab, uint32_t &nul) {
    uint4 y;
    y.x = map_word(x.x, tab, nul);
    y.y = map_word(x.y, tab, nul);
    y.z = map_word(x.z, tab, nul);
    y.w = map_word(x.w, tab, nul);
    return y;
}

// Static schedule: chunk k belongs to CTA k mod grid, so at any moment the 148 CTAs work on one contiguous ~10 MB
// window of the buffer.  A dynamic 8 KB-per-warp ticket scheme was measured and rejected (0.458 ms vs 0.373 ms per
// GiB): it widens the window to ~40 MB and serialises on one atomic.
__global__ void __launch_bounds__(kBmThreads, 1)
k_bytemap(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, size_t n,
          const uint16_t *__restrict__ pairlut, const uint8_t *__restrict__ lut, uint32_t *__restrict__ status) {
    extern __shared__ __align__(128) uint8_t smem_tab[];  // kPairLutBytes
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
            tma_bulk_g2s(smem_tab + k * (kPairLutBytes / 8), reinterpret_cast<const uint8_t *>(pairlut) + k * (kPairLutBytes / 8),
                         kPairLutBytes / 8, &bar);
    }
    const uint32_t tab = smem_u32(smem_tab);
    uint32_t nul = 0;

    // the vector body needs in and out equally misaligned; the host guarantees (in ^ out) % 16 == 0
    size_t head = ((16 - (reinterpret_cast<uintptr_t>(in) & 15)) & 15);
    if (head > n) head = n;
    const size_t nvec = (n - head) / 16;
    const uint4 *vin = reinterpret_cast<const uint4 *>(in + head);
    uint4 *vout = reinterpret_cast<uint4 *>(out + head);

    // every CTA iteration maps one CONTIGUOUS chunk of kBmUnroll x 1024 vectors (64 KB); chunks are dealt round the grid.
    // (The first version strode each thread's vectors a whole grid apart - 2.4 MB: scattered 512-byte accesses that cost
    // 15 % of the DRAM bandwidth, see k_bytemap_swar.)  Software pipelined: while one chunk is mapped and stored, the
    // next chunk's loads are already in flight.
    const size_t per_cta = (size_t)kBmUnroll * blockDim.x;
    const size_t full_chunks = nvec / per_cta;
    size_t c = blockIdx.x;
    bool table_ready = false;
    if (c < full_chunks) {
        uint4 x[kBmUnroll];
#pragma unroll
        for (int u = 0; u < kBmUnroll; ++u) x[u] = ld_stream(vin + c * per_cta + u * blockDim.x + threadIdx.x);
        mbar_wait(&bar, 0);                                   // the first loads overlapped the table copy
        table_ready = true;
        for (;;) {
            const size_t cn = c + gridDim.x;
            const bool more = cn < full_chunks;
            uint4 xn[kBmUnroll];
            if (more) {
#pragma unroll
                for (int u = 0; u < kBmUnroll; ++u) xn[u] = ld_stream(vin + cn * per_cta + u * blockDim.x + threadIdx.x);
            }
#pragma unroll
            for (int u = 0; u < kBmUnroll; ++u) st_stream(vout + c * per_cta + u * blockDim.x + threadIdx.x, map_vec(x[u], tab, nul));
            c = cn;
            if (!more) break;
#pragma unroll
            for (int u = 0; u < kBmUnroll; ++u) x[u] = xn[u];
        }
    }
    if (!table_ready) mbar_wait(&bar, 0);
    if (blockIdx.x == full_chunks % gridDim.x) {               // the last, partial chunk
        for (size_t v = full_chunks * per_cta + threadIdx.x; v < nvec; v += blockDim.x) {
            uint4 x = ld_stream(vin + v);
            st_stream(vout + v, map_vec(x, tab, nul));
        }
    }
    nul &= 0x80808080u;
    // ragged head and tail: one thread, byte by byte
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (size_t k = 0; k < head; ++k) { uint8_t c = in[k]; if (!c) nul = 1; out[k] = lut[c]; }
        for (size_t k = head + nvec * 16; k < n; ++k) { uint8_t c = in[k]; if (!c) nul = 1; out[k] = lut[c]; }
        if (n && in[n - 1] != '\n') nul = 1;              // last record would lose its last byte
    }
    if (__any_sync(0xFFFFFFFFu, nul != 0)) {
        if ((threadIdx.x & 31) == 0) atomicOr(status, ST_IRREGULAR);
    }
}

// ------------------------------------------------------------------------------------------------
// k_bytemap_swar: table-free byte maps
// ------------------------------------------------------------------------------------------------
// Most byte maps people write are "add a constant on a few byte ranges" ([a:A-z:Z]: a-z -32; Caesar: a-y +1,
// z -25).  Those need no table: per 32-bit word, a 7-bit SWAR range test per range (two adds whose carries stay
// inside their byte), the selected bytes' deltas accumulated with one IMAD per range, and a carry-isolated
// byte-wise add.  No shared memory, no LSU traffic besides the stream itself, 8 x 256 threads per SM.
// The host emits the ranges (BlobHeader::swar); maps that are not of this form keep k_bytemap's pair table.
struct SwarMap {
    uint32_t n;
    uint32_t add_lo[4];      // (0x80 - (lo & 0x7F)) in every byte: bit 7 of (l + add_lo) <=> l >= lo
    uint32_t add_hi[4];      // (0x7F - (hi & 0x7F)) in every byte: bit 7 of (l + add_hi) <=> l >  hi
    uint32_t half[4];        // 0x80808080 when the range lies in 0x00-0x7F (bit 7 of the byte must be clear), else 0
    uint32_t delta[4];       // added mod 256 to the bytes in range
};

template <int NR>
__device__ __forceinline__ uint32_t swar_word(uint32_t x, const SwarMap &M, uint32_t &nul) {
    const uint32_t l = x & 0x7F7F7F7Fu;
    uint32_t d = 0;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const uint32_t ge = l + M.add_lo[r], gt = l + M.add_hi[r];
        const uint32_t m = ge & ~gt & (x ^ M.half[r]) & 0x80808080u;      // bit 7 of every byte inside the range
        d += (m >> 7) * M.delta[r];                                       // ranges are disjoint: no byte gets two deltas
    }
    nul |= (x - 0x01010101u) & ~x;
    return (l + (d & 0x7F7F7F7Fu)) ^ ((x ^ d) & 0x80808080u);              // byte-wise x + d, carries cut at bit 7
}

template <int NR>
__device__ __forceinline__ uint4 swar_vec(const uint4 &x, const SwarMap &M, uint32_t &nul) {
    uint4 y;
    y.x = swar_word<NR>(x.x, M, nul);
    y.y = swar_word<NR>(x.y, M, nul);
    y.z = swar_word<NR>(x.z, M, nul);
    y.w = swar_word<NR>(x.w, M, nul);
    return y;
}

// One CTA per 16 KB block of the buffer (256 threads x 4 vectors, the block contiguous), no persistence: the hardware
// block scheduler walks the buffer front to back.  Measured against the persistent grid-stride sweep this kernel
// started as (each thread's 4 vectors 2.4 MB apart): 0.318 vs 0.374 ms per GiB - 1.03 vs 0.87 of the measured copy peak
// (gpurun_out/bench_cfg2_{default,v2}_r2b.json) - the sweep's scattered 512-byte accesses cost DRAM efficiency.
constexpr int kSwarThreads = 256;
constexpr int kSwarVecPerCta = 1024;

template <int NR>
__global__ void __launch_bounds__(kSwarThreads, 8)
k_bytemap_swar(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, size_t n, const SwarMap M,
                  const uint8_t *__restrict__ lut, uint32_t *__restrict__ status) {
    uint32_t nul = 0;
    size_t head = ((16 - (reinterpret_cast<uintptr_t>(in) & 15)) & 15);
    if (head > n) head = n;
    const size_t nvec = (n - head) / 16;
    const uint4 *vin = reinterpret_cast<const uint4 *>(in + head);
    uint4 *vout = reinterpret_cast<uint4 *>(out + head);
    const size_t v0 = (size_t)blockIdx.x * 1024 + threadIdx.x;
    if (v0 + 768 < nvec) {
        uint4 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = ld_stream(vin + v0 + u * 256);
#pragma unroll
        for (int u = 0; u < 4; ++u) st_stream(vout + v0 + u * 256, swar_vec<NR>(x[u], M, nul));
    } else {
        for (size_t v = v0; v < nvec && v < (size_t)(blockIdx.x + 1) * 1024; v += 256) {
            uint4 x = ld_stream(vin + v);
            st_stream(vout + v, swar_vec<NR>(x, M, nul));
        }
    }
    nul &= 0x80808080u;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (size_t k = 0; k < head; ++k) { uint8_t c = in[k]; if (!c) nul = 1; out[k] = lut[c]; }
        for (size_t k = head + nvec * 16; k < n; ++k) { uint8_t c = in[k]; if (!c) nul = 1; out[k] = lut[c]; }
        if (n && in[n - 1] != '\n') nul = 1;
    }
    if (__any_sync(0xFFFFFFFFu, nul != 0)) {
        if ((threadIdx.x & 31) == 0) atomicOr(status, ST_IRREGULAR);
    }
}

// ------------------------------------------------------------------------------------------------
// k_lines
// ------------------------------------------------------------------------------------------------
enum { MODE_INPLACE = 0, MODE_COUNT = 1, MODE_WRITE = 2 };

struct LineArgs {
    DevProg P;
    const uint8_t *in;
    uint8_t *out;
    size_t n;
    size_t out_cap;
    uint32_t *cnt;               // [n_chunks]  output bytes of the lines starting in chunk
    unsigned long long *gsum;    // [n_groups]  after k_scan_groups: exclusive offsets
    unsigned long long *total;   // grand total of output bytes
    uint32_t *status;
    size_t n_chunks;
    size_t n_groups;
    // DFS executor scratch, indexed by resident thread
    int32_t *bt_s; uint32_t *bt_i; uint32_t *bt_o; uint8_t *bt_obuf;
    uint32_t bt_cap, bt_ocap;
    LazyDev L;                   // EXEC_LAZY
};

enum { EXEC_TABLE = 0, EXEC_DFS = 1, EXEC_LAZY = 2 };   // which per-attempt executor k_lines runs

// The per-record loop of the reference's scan drivers (trre_nft.c:778-789 / trre_dft.c:1275-1285)
// for the record whose payload starts at `p`; `lim` = end of all payload bytes of the buffer (n, or
// n-1 when the buffer does not end in '\n': getline's last record loses its last byte).
// Returns the position just past the record.
// AMODE != AM_FIRST (the CLI's -a / -m / -ma, trre_nft.c:635-648, 782-797) always runs the DFS executor.
template <int EXEC, int AMODE = AM_FIRST>
__device__ __forceinline__ size_t run_record(const DevProg &P, const BtScratch &B, const LazyDev &L, const uint8_t *__restrict__ in,
                                             size_t p, size_t lim, size_t n, Sink &o, uint32_t &status) {
    if (EXEC == EXEC_DFS) { B.steps = 0; B.printed = 0; }
    if (AMODE == AM_MATCH || AMODE == AM_MATCH_ALL || AMODE == AM_NEWLINES) {
        // match mode: one attempt over the whole record, nothing copied, no newline of its own (trre_nft.c:792-796);
        // trre_dft prints a bare newline per record instead (trre_dft.c:1288-1292)
        if (AMODE == AM_NEWLINES) o.put('\n');
        else attempt_backtrack<AMODE>(P, B, in, p, lim, o, status);
    } else for (;;) {
        bool at_end = true;
        uint8_t c = 0;
        if (p < lim) { c = __ldg(in + p); at_end = (c == 0 || c == '\n'); }
        long k = EXEC == EXEC_DFS    ? attempt_backtrack<AMODE>(P, B, in, p, lim, o, status)
                 : EXEC == EXEC_LAZY ? attempt_lazy(L, in, p, lim, o, status)
                                     : attempt_table(P, in, p, lim, o, status);
        if (EXEC == EXEC_LAZY && k == -2) break;          // a row is missing: this pass is void, leave the record
        if (status & (ST_STACK | ST_OBUF | ST_DIVERGE | ST_RUNAWAY | ST_BUDGET)) break;   // the scan has failed (the reference exits here):
                                                          // do not grind through the rest of the record
        if (at_end) break;                                // that was the attempt on the empty tail
        if (k > 0) p += (size_t)k;
        else { o.put(c); ++p; }                           // no match or empty match: copy one raw byte
    }
    if (AMODE == AM_FIRST || AMODE == AM_SCAN_ALL) o.put('\n');
    // skip what getline would still have consumed: the rest of the record after a NUL, and the '\n'
    while (p < n && __ldg(in + p) != '\n') ++p;
    return p + 1;
}

template <int MODE, int EXEC, int AMODE = AM_FIRST>
__global__ void __launch_bounds__(kLineThreads)
k_lines(LineArgs a) {
    __shared__ unsigned long long warp_sums[kLineThreads / 32];
    __shared__ unsigned long long group_base;
    const uint8_t *__restrict__ in = a.in;
    const size_t n = a.n;
    if (n == 0) return;
    const bool ends_nl = __ldg(in + n - 1) == '\n';
    const size_t lim = ends_nl ? n : n - 1;
    if (MODE == MODE_WRITE && *a.total > a.out_cap) {
        if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(a.status, ST_NOSPACE);
        return;
    }
    BtScratch B{};
    if (EXEC == EXEC_DFS) {
        size_t slot = (size_t)blockIdx.x * kLineThreads + threadIdx.x;
        B.stk_s = a.bt_s + slot * a.bt_cap; B.stk_i = a.bt_i + slot * a.bt_cap; B.stk_o = a.bt_o + slot * a.bt_cap;
        B.obuf = a.bt_obuf + slot * a.bt_ocap; B.cap = a.bt_cap; B.ocap = a.bt_ocap;
    }
    uint32_t status = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (size_t g = blockIdx.x; g < a.n_groups; g += gridDim.x) {
        const size_t chunk = g * kLineThreads + threadIdx.x;
        const size_t cbeg = chunk * (size_t)kChunk;
        const size_t cend = (cbeg + kChunk < n) ? cbeg + kChunk : n;
        unsigned long long my_off = 0;
        if (MODE == MODE_WRITE) {
            // exclusive scan of this group's chunk counts -> where this thread's lines go
            unsigned long long v = (chunk < a.n_chunks) ? a.cnt[chunk] : 0, x = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, x, d);
                if (lane >= d) x += y;
            }
            __syncthreads();                              // previous iteration's readers are done
            if (lane == 31) warp_sums[warp] = x;
            if (threadIdx.x == 0) group_base = a.gsum[g];
            __syncthreads();
            unsigned long long wbase = 0;
            for (int w = 0; w < warp; ++w) wbase += warp_sums[w];
            my_off = group_base + wbase + x - v;
        }
        Sink o;
        o.p = (MODE == MODE_COUNT) ? nullptr : a.out;
        o.n = (size_t)my_off;
        size_t produced = 0;
        if (cbeg < n) {
            // find the first record starting in [cbeg, cend)
            size_t p = cbeg;
            if (p != 0 && __ldg(in + p - 1) != '\n') {
                while (p < cend && __ldg(in + p) != '\n') ++p;
                ++p;                                      // first byte after that '\n' (may be >= cend)
            }
            while (p < cend) {                            // p < n here: a record starts at p
                if (MODE == MODE_INPLACE) o.n = p;
                size_t before = o.n;
                p = run_record<EXEC, AMODE>(a.P, B, a.L, in, p, lim, n, o, status);
                produced += o.n - before;
                if (MODE == MODE_INPLACE && o.n != p) status |= ST_IRREGULAR;   // a NUL cut the line short
            }
        }
        if (MODE == MODE_COUNT) {
            if (chunk < a.n_chunks) a.cnt[chunk] = (uint32_t)produced;
            if (produced > 0xFFFFFFFFull) status |= ST_NOSPACE;   // > 4 GiB from one chunk's lines
            unsigned long long x = produced;
#pragma unroll
            for (int d = 16; d; d >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, d);
            __syncthreads();
            if (lane == 0) warp_sums[warp] = x;
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned long long s = 0;
                for (int w = 0; w < kLineThreads / 32; ++w) s += warp_sums[w];
                a.gsum[g] = s;
            }
        }
    }
    if (MODE == MODE_INPLACE && !ends_nl) status |= ST_IRREGULAR;
    if (status) atomicOr(a.status, status);
}

// ------------------------------------------------------------------------------------------------
// k_scan_groups: gsum[] -> exclusive prefix sums in place, *total = sum.  One CTA.
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 1024;

__global__ void __launch_bounds__(kScanThreads)
k_scan_groups(unsigned long long *gsum, size_t n_groups, unsigned long long *total) {
    __shared__ unsigned long long part[kScanThreads];
    const size_t per = (n_groups + kScanThreads - 1) / kScanThreads;
    const size_t b = (size_t)threadIdx.x * per;
    const size_t e = (b + per < n_groups) ? b + per : n_groups;
    unsigned long long s = 0;
    for (size_t k = b; k < e; ++k) s += gsum[k];
    part[threadIdx.x] = s;
    __syncthreads();
    // Hillis-Steele over 1024 partials
    for (int d = 1; d < kScanThreads; d <<= 1) {
        unsigned long long y = (threadIdx.x >= d) ? part[threadIdx.x - d] : 0;
        __syncthreads();
        part[threadIdx.x] += y;
        __syncthreads();
    }
    unsigned long long run = part[threadIdx.x] - s;       // exclusive
    for (size_t k = b; k < e; ++k) { unsigned long long v = gsum[k]; gsum[k] = run; run += v; }
    if (threadIdx.x == kScanThreads - 1) *total = part[kScanThreads - 1];
}

}  // namespace trre
