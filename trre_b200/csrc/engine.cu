// engine.cu - C ABI of the scan engine (include/trre_b200.h): program objects, device residency,
// kernel launches.  Everything that touches the input stream runs on the GPU; there is no CPU scan
// path in this library (trre_match_host is match mode, a different reference path).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/trre_b200.h"
#include "automaton.hpp"
#include "kernels.cuh"
#include "match_kernel.cuh"
#include "gram_kernel.cuh"
#include "lit_kernel.cuh"
#include "lines_staged.cuh"
#include "bmv_kernel.cuh"
#include "pattern.hpp"

using namespace trre;

namespace {

thread_local std::string g_last_error;

void set_err(char *err, size_t cap, const std::string &s) {
    g_last_error = s;
    if (err && cap) {
        std::snprintf(err, cap, "%s", s.c_str());
    }
}

#ifndef TRRE_SLOTS
#define TRRE_SLOTS 3
#endif
constexpr int kSlots = TRRE_SLOTS;        // pipelined host path: chunks in flight

// where a count/scan/write launch keeps its sums and reports: the per-device workspace, or one pipeline slot's
struct Ws {
    unsigned long long *gsum, *total;
    uint32_t *status;
};

struct MatchPlan {
    uint32_t repl = 0, states = 0, pool = 0;
    int tab = 0;
    size_t smem = 0;
    int blocks = 0;
};

struct DeviceState {
    int device = -1;
    uint8_t *d_blob = nullptr;
    DevProg P{};
    // workspace, grown on demand
    uint32_t *d_cnt = nullptr; size_t cap_chunks = 0;
    unsigned long long *d_gsum = nullptr; size_t cap_groups = 0;
    unsigned long long *d_total = nullptr;
    uint32_t *d_status = nullptr;          // [1 + kSlots]: [0] device-resident API, [1..] pipeline slots
    // pinned mirrors
    unsigned long long *h_total = nullptr;
    uint32_t *h_status = nullptr;          // [1 + kSlots]
    // DFS scratch
    int32_t *bt_s = nullptr; uint32_t *bt_i = nullptr; uint32_t *bt_o = nullptr; uint8_t *bt_obuf = nullptr;
    size_t bt_threads = 0; uint32_t bt_cap = 0, bt_ocap = 0;
    int bt_tier = 0;               // 0: 8192-entry stacks for every resident thread; 1: the reference's 65536 entries and
                                   // 1 MiB of output per attempt, for few threads (entered when tier 0 overflows)
    int sm_count = 148;
    const uint32_t *vent = nullptr; const uint16_t *vpairlen = nullptr; const uint8_t *vpool = nullptr; uint32_t vpool_bytes = 0;
    // what the last enqueue left to do / to report
    int launches = 0;
    bool status_dirty = true;      // d_status[0] / d_total must be cleared before the next enqueue
    bool pend_fast = false;        // only a fast path ran; on ST_IRREGULAR collect() reruns the buffer generally
    bool pend_total_is_n = false;  // fast path was length preserving: total = n, else total comes from d_total
    int pend_phase = 1;            // PH_* the fast path stood for
    int counts_kind = 0;           // whose counts sit in the workspace: 1 general (cnt/gsum per group), 2 k_bmv, 3 k_match<OP_COUNT> (gsum per span)
    size_t bmv_smem = 0; bool bmv_attr = false;
    const uint8_t *pend_in = nullptr; uint8_t *pend_out = nullptr; size_t pend_n = 0, pend_cap = 0;
    // host-buffer path: cached device buffers
    uint8_t *buf_in = nullptr, *buf_out = nullptr; size_t cap_in = 0, cap_out = 0;
    uint8_t *slot_in[kSlots] = {}, *slot_out[kSlots] = {};
    cudaStream_t slot_stream[kSlots] = {};
    size_t slot_cap = 0, slot_out_cap = 0;
    unsigned long long *slot_gsum[kSlots] = {};   // per-slot span sums (k_bmv / k_match var)
    unsigned long long *d_slot_total = nullptr, *h_slot_total = nullptr;     // [kSlots]
    cudaEvent_t slot_done[kSlots] = {};
    MatchPlan mp, mpv;             // k_match launch plans: length preserving / variable length
    bool staged_attr = false;
    GramArgs gramv{};              // k_gram_var: the same tables for programs whose matches change the length
    size_t gramv_smem = 0;
    bool gramv_small = false;
    const void *gramv_fn[2] = {nullptr, nullptr};   // count, write
    GramArgs gram{};               // k_gram tables (entries == 0: none)
    size_t gram_smem = 0;
    bool gram_small = false;
    const void *gram_fn = nullptr;
    SwarMap swar{};                // table-free byte map (n == 0: not of that form)
    LitMap lit{};                  // table-free word matcher (lit_words == 0: not of that form)
    uint32_t lit_words = 0, lit_len = 0;
    bool lit_eq = false;           // every position of every word accepts exactly one byte
    RunMap runmap{};               // RUN form whose classes are ranges of one printed length: SWAR counting (n == 0: per byte)
    LitvArgs litv{};               // the same words printing strings of any length (k_litv_count / k_bmv_write<BMV_LIT>)
    bool litv_ok = false;
    // TRRE_EXEC_LAZY: device copies of the rows built so far (automaton.hpp, LazyDsa)
    uint64_t *lz_rows = nullptr, *lz_sinfo = nullptr;
    uint8_t *lz_pool = nullptr;
    uint32_t *lz_requested = nullptr, *lz_miss = nullptr;      // lz_miss: [0] count, [1..] list
    uint32_t *h_lz_miss = nullptr;                               // pinned mirror
    size_t lz_cap_states = 0, lz_cap_pool = 0, lz_up_log = 0, lz_up_pool = 0;
    int lz_rounds = 0;
};

constexpr uint32_t kMissCap = 1u << 16;

}  // namespace

struct trre_prog {
    std::vector<uint8_t> blob;
    BlobHeader hdr;
    std::mutex mu;
    std::map<int, DeviceState *> dev;
    LazyDsa *lazy = nullptr;       // TRRE_EXEC_LAZY: the on-demand table builder, shared by all devices
    std::mutex lazy_mu;
    int amode = AM_FIRST;          // trre_prog_set_mode: what the CLI's -m / -a ask for (AM_*, device_walk.cuh)
};

namespace {

// TRRE_EXEC_LAZY programs keep their table builder: the flat NFT comes back out of the blob (so a broadcast blob
// works on every rank without recompiling the expression)
void attach_lazy(trre_prog *p) {
    if (p->hdr.exec_kind != TRRE_EXEC_LAZY) return;
    Nft g;
    const NftFlat *nf = reinterpret_cast<const NftFlat *>(p->blob.data() + p->hdr.off_nft);
    g.st.resize(p->hdr.nft_states);
    for (uint32_t k = 0; k < p->hdr.nft_states; ++k) {
        g.st[k].op = nf[k].op; g.st[k].val = nf[k].val; g.st[k].a = nf[k].a; g.st[k].b = nf[k].b;
    }
    g.start = p->hdr.nft_start;
    g.has_init_join = p->hdr.flavor == TRRE_FLAVOR_DFT;
    p->lazy = new LazyDsa(g, (int)p->hdr.flavor);
}

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            g_last_error = std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #expr; \
            return TRRE_E_CUDA;                                                                \
        }                                                                                      \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// shared-memory plan of k_match: bank-replicated live table when everything fits; raw-byte rows for small
// automata; otherwise the first (shallowest: states are numbered breadth first) rows of the compact table
int plan_match(const BlobHeader &h, size_t warp_bytes, size_t budget, int sm_count, int op, MatchPlan *mp) {
    const size_t ct_full = (((size_t)h.dsa_states * h.n_classes + 3) & ~(size_t)3) * 4;
    mp->pool = h.pool_bytes <= 4096 ? 1 : 0;
    const size_t fixed = 1024 + 256 + (mp->pool ? ((h.pool_bytes + 15) & ~15u) : 0) + warp_bytes;
    mp->tab = TAB_PART;
    size_t ct_bytes;
    if (32768 + fixed + (size_t)h.dsa_states * 1024 <= budget) {              // small automaton: raw-byte rows
        mp->repl = 1; mp->states = h.dsa_states; mp->tab = TAB_WIDE;
        ct_bytes = (size_t)h.dsa_states * 1024;
    } else if (32768 + fixed + ct_full <= budget) {
        mp->repl = 1; mp->states = h.dsa_states; mp->tab = TAB_ALL;
        ct_bytes = ct_full;
    } else {
        mp->repl = 0;
        size_t rows = (budget - fixed) / ((size_t)h.n_classes * 4);
        mp->states = (uint32_t)std::min<size_t>(rows, h.dsa_states);
        if (mp->states == h.dsa_states) mp->tab = TAB_ALL;
        ct_bytes = (((size_t)mp->states * h.n_classes + 3) & ~(size_t)3) * 4;
    }
    mp->smem = (mp->repl ? 32768 : 0) + fixed + ct_bytes;
    int per_sm = 1;
    const void *fns[3][3] = {
        {(const void *)k_match<TAB_PART, OP_INPLACE>, (const void *)k_match<TAB_PART, OP_COUNT>, (const void *)k_match<TAB_PART, OP_WRITE>},
        {(const void *)k_match<TAB_ALL, OP_INPLACE>, (const void *)k_match<TAB_ALL, OP_COUNT>, (const void *)k_match<TAB_ALL, OP_WRITE>},
        {(const void *)k_match<TAB_WIDE, OP_INPLACE>, (const void *)k_match<TAB_WIDE, OP_COUNT>, (const void *)k_match<TAB_WIDE, OP_WRITE>}};
    for (int o = (op == OP_INPLACE ? OP_INPLACE : OP_COUNT); o <= (op == OP_INPLACE ? OP_INPLACE : OP_WRITE); ++o) {
        CUDA_TRY(cudaFuncSetAttribute(fns[mp->tab][o], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mp->smem));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fns[mp->tab][o], kMatchThreads, mp->smem));
    }
    mp->blocks = sm_count * (per_sm < 1 ? 1 : per_sm);
    return TRRE_OK;
}

const void *gram_var_kernel(uint32_t depth, bool small, int op) {
#define GV(D) (small ? (op == OP_COUNT ? (const void *)k_gram_var<D, true, OP_COUNT> : (const void *)k_gram_var<D, true, OP_WRITE>) \
                     : (op == OP_COUNT ? (const void *)k_gram_var<D, false, OP_COUNT> : (const void *)k_gram_var<D, false, OP_WRITE>))
    switch (depth) {
        case 1: return GV(1);
        case 2: return GV(2);
        case 3: return GV(3);
        case 4: return GV(4);
        case 5: return GV(5);
        case 6: return GV(6);
        default: return nullptr;
    }
#undef GV
}

const void *gram_kernel(uint32_t depth, bool small, int reg) {      // reg: 0 staged, 1 registers, 2 registers + delta outputs
#define GK(D) (small ? (reg ? (const void *)k_gram<D, true, 1> : (const void *)k_gram<D, true, 0>) \
                     : (reg == 2 ? (const void *)k_gram<D, false, 2> : reg ? (const void *)k_gram<D, false, 1> : (const void *)k_gram<D, false, 0>))
    switch (depth) {
        case 1: return GK(1);
        case 2: return GK(2);
        case 3: return GK(3);
        case 4: return GK(4);
        case 5: return GK(5);
        case 6: return GK(6);
        default: return nullptr;
    }
#undef GK
}

int get_state(const trre_prog *cp, int device, DeviceState **out) {
    trre_prog *p = const_cast<trre_prog *>(cp);
    std::lock_guard<std::mutex> lock(p->mu);
    auto it = p->dev.find(device);
    if (it != p->dev.end()) { *out = it->second; return TRRE_OK; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        g_last_error = "no CUDA device: the trre_b200 scan path has no CPU fallback";
        return TRRE_E_CUDA;
    }
    if (device < 0 || device >= ndev) { g_last_error = "device index out of range"; return TRRE_E_ARG; }
    DeviceState *d = new DeviceState();
    struct Undo {                                          // a CUDA_TRY that fails below must not leak the half-built state
        DeviceState *d;
        ~Undo() {
            if (!d) return;
            cudaFree(d->d_blob); cudaFree(d->d_total); cudaFree(d->d_status);
            cudaFreeHost(d->h_total); cudaFreeHost(d->h_status);
            delete d;
        }
    } undo{d};
    d->device = device;
    const BlobHeader &h = p->hdr;
    CUDA_TRY(cudaMalloc(&d->d_blob, p->blob.size()));
    CUDA_TRY(cudaMemcpy(d->d_blob, p->blob.data(), p->blob.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&d->d_total, sizeof(unsigned long long)));
    CUDA_TRY(cudaMalloc(&d->d_status, (1 + kSlots) * sizeof(uint32_t)));
    CUDA_TRY(cudaMallocHost(&d->h_total, sizeof(unsigned long long)));
    CUDA_TRY(cudaMallocHost(&d->h_status, (1 + kSlots) * sizeof(uint32_t)));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    d->sm_count = prop.multiProcessorCount;
    d->P.classmap = d->d_blob + h.off_classmap;
    d->P.lut = d->d_blob + h.off_lut;
    d->P.pairlut = h.off_pairlut ? reinterpret_cast<const uint16_t *>(d->d_blob + h.off_pairlut) : nullptr;
    d->P.row0 = h.off_row0 ? reinterpret_cast<const uint32_t *>(d->d_blob + h.off_row0) : nullptr;
    d->P.ctrans = h.off_ctrans ? reinterpret_cast<const uint32_t *>(d->d_blob + h.off_ctrans) : nullptr;
    d->vent = h.off_vent ? reinterpret_cast<const uint32_t *>(d->d_blob + h.off_vent) : nullptr;
    d->vpairlen = h.off_vpairlen ? reinterpret_cast<const uint16_t *>(d->d_blob + h.off_vpairlen) : nullptr;
    d->vpool = h.off_vpool ? d->d_blob + h.off_vpool : nullptr;
    d->vpool_bytes = h.vpool_bytes;
    d->P.eout = h.off_eout ? reinterpret_cast<const uint32_t *>(d->d_blob + h.off_eout) : nullptr;
    d->P.pool_bytes = h.pool_bytes;
    d->P.n_states = h.dsa_states;
    d->P.max_depth = h.max_depth;
    d->P.trans = reinterpret_cast<const uint64_t *>(d->d_blob + h.off_trans);
    d->P.sinfo = reinterpret_cast<const uint64_t *>(d->d_blob + h.off_sinfo);
    d->P.pool = d->d_blob + h.off_pool;
    d->P.nft = reinterpret_cast<const NftFlat *>(d->d_blob + h.off_nft);
    d->P.n_classes = h.n_classes;
    d->P.start_info = h.start_info;
    d->P.flags = h.flags;
    d->P.nft_start = h.nft_start;
    CUDA_TRY(cudaFuncSetAttribute(k_bytemap, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairLutBytes));
    if (h.exec_kind == TRRE_EXEC_BYTEMAP && h.swar_n >= 1 && h.swar_n <= 4 && !std::getenv("TRRE_NO_SWAR")) {
        d->swar.n = h.swar_n;
        for (uint32_t k = 0; k < h.swar_n; ++k) {
            const uint32_t lo = h.swar[k] & 0xFFu, hi = (h.swar[k] >> 8) & 0xFFu, delta = (h.swar[k] >> 16) & 0xFFu;
            d->swar.add_lo[k] = (0x80u - (lo & 0x7Fu)) * 0x01010101u;
            d->swar.add_hi[k] = (0x7Fu - (hi & 0x7Fu)) * 0x01010101u;
            d->swar.half[k] = lo < 0x80u ? 0x80808080u : 0u;
            d->swar.delta[k] = delta;
        }
    }
    if (h.lit_words >= 1 && h.lit_words <= 4 && h.lit_len >= 2 && h.lit_len <= 4 && (h.flags & PF_MATCHABLE_VAR) && !(h.flags & PF_MATCHABLE) &&
        h.gpool_bytes <= 16384 && !std::getenv("TRRE_NO_LIT")) {
        LitvArgs &A = d->litv;
        bool ok = true;
        for (uint32_t k = 0; k < h.lit_words * h.lit_len; ++k) {
            const uint32_t lo = h.lit[k] & 0xFFu, hi = (h.lit[k] >> 8) & 0xFFu;
            if (lo != hi) ok = false;                                  // plain strings only (a range keeps k_gram_var)
            A.M.eq[k] = lo * 0x01010101u;
        }
        A.M.one = 1;
        const uint8_t *gp = p->blob.data() + h.off_gpool;
        for (uint32_t w = 0; w < h.lit_words; ++w) {
            const uint32_t off = h.lit_out[w] & 0xFFFFFu, len = (h.lit_out[w] >> 20) & 0xFFu;
            A.ent[w] = len | ((len ? (uint32_t)gp[off] : 0u) << 8) | (off << 16);
        }
        d->litv_ok = ok;
    }
    if (h.lit_words >= 1 && h.lit_words <= 4 && h.lit_len >= 2 && h.lit_len <= 4 && (h.flags & PF_MATCHABLE) && !std::getenv("TRRE_NO_LIT")) {
        d->lit_words = h.lit_words; d->lit_len = h.lit_len;
        d->lit_eq = true;
        for (uint32_t k = 0; k < h.lit_words * h.lit_len; ++k) {
            const uint32_t lo = h.lit[k] & 0xFFu, hi = (h.lit[k] >> 8) & 0xFFu;
            if (lo != hi) d->lit_eq = false;
            d->lit.eq[k] = lo * 0x01010101u;
            d->lit.add_lo[k] = (0x80u - (lo & 0x7Fu)) * 0x01010101u;
            d->lit.add_hi[k] = (0x7Fu - (hi & 0x7Fu)) * 0x01010101u;
            d->lit.printed[k / h.lit_len] |= ((h.lit[k] >> 16) & 0xFFu) << (8 * (k % h.lit_len));
        }
        d->lit.one = 1;
        if (std::getenv("TRRE_LIT_NO_EQ")) d->lit_eq = false;     // (test switch: the range form on plain strings)
        if (!d->lit_eq)                                            // the range form tests 7-bit ranges below 0x80 only: k_gram serves
            for (uint32_t k = 0; k < h.lit_words * h.lit_len; ++k)
                if (((h.lit[k] >> 8) & 0xFFu) >= 0x80u) d->lit_words = 0;
    }
    if (h.vrun_swar_n >= 1 && h.vrun_swar_n <= 4 && h.vrun_swar_n == h.vrun_classes && !std::getenv("TRRE_RUN_PER_BYTE")) {
        d->runmap.n = h.vrun_swar_n;
        for (uint32_t k = 0; k < h.vrun_swar_n; ++k) {
            const uint32_t lo = h.vrun_swar[k] & 0xFFu, hi = (h.vrun_swar[k] >> 8) & 0xFFu;
            d->runmap.add_lo[k] = (0x80u - (lo & 0x7Fu)) * 0x01010101u;
            d->runmap.add_hi[k] = (0x7Fu - (hi & 0x7Fu)) * 0x01010101u;
            d->runmap.half[k] = lo < 0x80u ? 0x80808080u : 0u;
            d->runmap.len[k] = (h.vrun_swar[k] >> 16) & 0xFFu;
        }
    }
    if (h.flags & PF_MATCHABLE) {
        int rc2 = plan_match(h, (size_t)kMatchWarps * kWarpSmem, 108 * 1024, d->sm_count, OP_INPLACE, &d->mp);
        if (rc2 != TRRE_OK) return rc2;
    }
    if (h.flags & PF_MATCHABLE_VAR) {
        int rc2 = plan_match(h, (size_t)kMatchWarps * kVarWarpSmem, 108 * 1024, d->sm_count, OP_COUNT, &d->mpv);
        if (rc2 != TRRE_OK) return rc2;
    }
    if ((h.flags & PF_GRAM) && (h.flags & PF_MATCHABLE)) {
        GramArgs &g = d->gram;
        g.classmap = d->P.classmap;
        g.gbits = reinterpret_cast<const uint32_t *>(d->d_blob + h.off_gbits);
        g.gout = reinterpret_cast<const uint32_t *>(d->d_blob + h.off_gout);
        g.gpool = d->d_blob + h.off_gpool;
        g.gpool_bytes = h.gpool_bytes;
        g.n_classes = h.n_classes;
        g.entries = h.gram_entries;
        g.uniform_len = h.gram_len;
        g.gflags = h.gram_flags;
        const uint32_t words = (h.gram_entries + 31) / 32;
        d->gram_small = h.gram_entries <= kGramSmallEntries;
        g.gout_in_smem = h.gram_entries <= 4096 ? 1u : 0u;
        g.pool_in_smem = (h.gpool_bytes && h.gpool_bytes <= 4096) ? 1u : 0u;
        // the register form: one common match length, printed bytes in a word (with small tables: in the hit word, <= 3)
        const bool reg = (h.gram_flags & GRAM_DIRECT) && h.gram_len >= 1 && h.gram_len <= (d->gram_small ? 3u : 4u) &&
                         !std::getenv("TRRE_GRAM_NO_REG");
        g.len_mask = h.gram_len >= 4 ? 0xFFFFFFFFu : ((1u << (8 * h.gram_len)) - 1u);
        d->gram_smem = 256 * 32 * 4 + (d->gram_small ? (size_t)h.gram_entries * 128 : (size_t)((words + 3) & ~3u) * 4) +
                       (g.gout_in_smem ? (size_t)((h.gram_entries + 3) & ~3u) * 4 : 0) +
                       (g.pool_in_smem ? ((h.gpool_bytes + 15) & ~15u) : 0) + (reg ? 8 * 256 * 2 : (size_t)kGramWarps * kGramWarpSmem);
        const bool map = reg && !d->gram_small && (h.gram_flags & GRAM_MAPDELTA) && !std::getenv("TRRE_GRAM_NO_MAP");
        g.map_delta = h.gram_delta & 0xFFu;
        d->gram_fn = gram_kernel(h.gram_depth, d->gram_small, map ? 2 : reg ? 1 : 0);
        if (d->gram_fn && d->gram_smem <= 220 * 1024) {
            CUDA_TRY(cudaFuncSetAttribute(d->gram_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->gram_smem));
        } else {
            g.entries = 0;
        }
    }
    if ((h.flags & PF_GRAM) && (h.flags & PF_MATCHABLE_VAR)) {
        GramArgs &g = d->gramv;
        g.classmap = d->P.classmap;
        g.gbits = reinterpret_cast<const uint32_t *>(d->d_blob + h.off_gbits);
        g.gout = reinterpret_cast<const uint32_t *>(d->d_blob + h.off_gout);
        g.gpool = d->d_blob + h.off_gpool;
        g.gpool_bytes = h.gpool_bytes;
        g.n_classes = h.n_classes;
        g.entries = h.gram_entries;
        g.uniform_len = h.gram_len;
        g.gflags = h.gram_flags;
        const uint32_t words = (h.gram_entries + 31) / 32;
        d->gramv_small = h.gram_entries <= kGramSmallEntries;
        g.gout_in_smem = h.gram_entries <= 4096 ? 1u : 0u;
        g.pool_in_smem = (h.gpool_bytes && h.gpool_bytes <= 4096) ? 1u : 0u;
        d->gramv_smem = 256 * 32 * 4 + (d->gramv_small ? (size_t)h.gram_entries * 128 : (size_t)((words + 3) & ~3u) * 4) +
                        (g.gout_in_smem ? (size_t)((h.gram_entries + 3) & ~3u) * 4 : 0) +
                        (g.pool_in_smem ? ((h.gpool_bytes + 15) & ~15u) : 0) + (size_t)kGramWarps * kVarWarpSmem;
        d->gramv_fn[0] = gram_var_kernel(h.gram_depth, d->gramv_small, OP_COUNT);
        d->gramv_fn[1] = gram_var_kernel(h.gram_depth, d->gramv_small, OP_WRITE);
        if (d->gramv_fn[0] && d->gramv_smem <= 220 * 1024 && !(h.gram_flags & GRAM_DIRECT)) {
            CUDA_TRY(cudaFuncSetAttribute(d->gramv_fn[0], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->gramv_smem));
            CUDA_TRY(cudaFuncSetAttribute(d->gramv_fn[1], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->gramv_smem));
        } else {
            g.entries = 0;
        }
    }
    p->dev[device] = d;
    undo.d = nullptr;
    *out = d;
    return TRRE_OK;
}

int ensure_workspace(DeviceState *d, size_t n_chunks, size_t n_groups, bool backtrack, size_t resident_threads) {
    if (n_chunks > d->cap_chunks) {
        if (d->d_cnt) cudaFree(d->d_cnt);
        d->d_cnt = nullptr; d->cap_chunks = 0;
        CUDA_TRY(cudaMalloc(&d->d_cnt, n_chunks * sizeof(uint32_t)));
        d->cap_chunks = n_chunks;
    }
    if (n_groups > d->cap_groups) {
        if (d->d_gsum) cudaFree(d->d_gsum);
        d->d_gsum = nullptr; d->cap_groups = 0;
        CUDA_TRY(cudaMalloc(&d->d_gsum, n_groups * sizeof(unsigned long long)));
        d->cap_groups = n_groups;
    }
    // per-thread DFS stack and output buffer.  The reference's stack holds 65536 entries before "stack max capacity
    // reached" (32 doubled while <= 100000, trre_nft.c:34-36, 548-556) and its output buffer grows without bound;
    // almost no line needs that, so every resident thread gets 8192 / 8 KB first and a scan that overflows them is
    // repeated (collect) with the reference's capacity on a small grid.
    uint32_t cap = 8192, ocap = 8192;
    if (const char *e = std::getenv("TRRE_BT_STACK")) { long v = std::atol(e); if (v >= 64 && v <= 65536) cap = (uint32_t)v; }
    if (d->bt_tier) { cap = 65536; ocap = 1u << 20; }
    if (backtrack && (resident_threads > d->bt_threads || cap != d->bt_cap || ocap != d->bt_ocap)) {
        if (d->bt_s) { cudaFree(d->bt_s); cudaFree(d->bt_i); cudaFree(d->bt_o); cudaFree(d->bt_obuf); }
        d->bt_s = nullptr; d->bt_threads = 0;
        CUDA_TRY(cudaMalloc(&d->bt_s, resident_threads * cap * sizeof(int32_t)));
        CUDA_TRY(cudaMalloc(&d->bt_i, resident_threads * cap * sizeof(uint32_t)));
        CUDA_TRY(cudaMalloc(&d->bt_o, resident_threads * cap * sizeof(uint32_t)));
        CUDA_TRY(cudaMalloc(&d->bt_obuf, resident_threads * (size_t)ocap));
        d->bt_threads = resident_threads; d->bt_cap = cap; d->bt_ocap = ocap;
    }
    return TRRE_OK;
}

template <int MODE>
void launch_lines(int exec, int grid, cudaStream_t st, const LineArgs &a, int amode = AM_FIRST) {
    if (amode == AM_SCAN_ALL) k_lines<MODE, EXEC_DFS, AM_SCAN_ALL><<<grid, kLineThreads, 0, st>>>(a);
    else if (amode == AM_MATCH) k_lines<MODE, EXEC_DFS, AM_MATCH><<<grid, kLineThreads, 0, st>>>(a);
    else if (amode == AM_MATCH_ALL) k_lines<MODE, EXEC_DFS, AM_MATCH_ALL><<<grid, kLineThreads, 0, st>>>(a);
    else if (amode == AM_NEWLINES) k_lines<MODE, EXEC_TABLE, AM_NEWLINES><<<grid, kLineThreads, 0, st>>>(a);
    else if (exec == EXEC_DFS) k_lines<MODE, EXEC_DFS><<<grid, kLineThreads, 0, st>>>(a);
    else if (exec == EXEC_LAZY) k_lines<MODE, EXEC_LAZY><<<grid, kLineThreads, 0, st>>>(a);
    else k_lines<MODE, EXEC_TABLE><<<grid, kLineThreads, 0, st>>>(a);
}

__global__ void k_fill64(uint64_t *p, size_t n, uint64_t v) {
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) p[k] = v;
}

// TRRE_EXEC_LAZY: bring this device's copy of the tables up to date with the host builder
int lazy_sync(trre_prog *p, DeviceState *d, cudaStream_t st) {
    std::lock_guard<std::mutex> lock(p->lazy_mu);
    LazyDsa &L = *p->lazy;
    const size_t n = L.n_states();
    if (n > d->lz_cap_states) {                               // grow: everything is uploaded again
        const size_t cap = std::max<size_t>(2 * n, 4096);
        if (d->lz_rows) { cudaFree(d->lz_rows); cudaFree(d->lz_sinfo); cudaFree(d->lz_requested); }
        d->lz_rows = nullptr; d->lz_cap_states = 0; d->lz_up_log = 0;
        CUDA_TRY(cudaMalloc(&d->lz_rows, cap * 256 * sizeof(uint64_t)));
        CUDA_TRY(cudaMalloc(&d->lz_sinfo, cap * sizeof(uint64_t)));
        CUDA_TRY(cudaMalloc(&d->lz_requested, cap * sizeof(uint32_t)));
        k_fill64<<<d->sm_count * 4, 256, 0, st>>>(d->lz_rows, cap * 256, pack64(LAZY_UNBUILT, 0));
        CUDA_TRY(cudaMemsetAsync(d->lz_requested, 0, cap * sizeof(uint32_t), st));
        d->lz_cap_states = cap;
    }
    if (!d->lz_miss) {
        CUDA_TRY(cudaMalloc(&d->lz_miss, (1 + kMissCap) * sizeof(uint32_t)));
        CUDA_TRY(cudaMallocHost(&d->h_lz_miss, (1 + kMissCap) * sizeof(uint32_t)));
    }
    const std::vector<uint8_t> &pool = L.pool();
    if (pool.size() > d->lz_cap_pool) {
        const size_t cap = std::max<size_t>(2 * pool.size(), 1 << 16);
        if (d->lz_pool) cudaFree(d->lz_pool);
        d->lz_pool = nullptr; d->lz_cap_pool = 0; d->lz_up_pool = 0;
        CUDA_TRY(cudaMalloc(&d->lz_pool, cap));
        d->lz_cap_pool = cap;
    }
    if (pool.size() > d->lz_up_pool) {
        CUDA_TRY(cudaMemcpyAsync(d->lz_pool + d->lz_up_pool, pool.data() + d->lz_up_pool, pool.size() - d->lz_up_pool, cudaMemcpyHostToDevice, st));
        d->lz_up_pool = pool.size();
    }
    CUDA_TRY(cudaMemcpyAsync(d->lz_sinfo, L.sinfo().data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    const std::vector<uint32_t> &log = L.log();
    for (size_t k = d->lz_up_log; k < log.size(); ++k) {
        const size_t s = log[k];
        CUDA_TRY(cudaMemcpyAsync(d->lz_rows + s * 256, L.rows().data() + s * 256, 256 * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    }
    d->lz_up_log = log.size();
    CUDA_TRY(cudaStreamSynchronize(st));                      // the host vectors may move once the lock is dropped
    return TRRE_OK;
}

enum { PH_SIZE = 0, PH_ALL = 1, PH_WRITE = 2 };    // count+scan | count+scan+write | write (counts already there)

LineArgs make_args(DeviceState *d, const uint8_t *d_in, size_t n, uint8_t *d_out, size_t out_cap, uint32_t *status) {
    LineArgs a{};
    a.P = d->P; a.in = d_in; a.out = d_out; a.n = n; a.out_cap = out_cap;
    a.cnt = d->d_cnt; a.gsum = d->d_gsum; a.total = d->d_total; a.status = status;
    a.n_chunks = (n + kChunk - 1) / kChunk;
    a.n_groups = (a.n_chunks + kLineThreads - 1) / kLineThreads;
    a.bt_s = d->bt_s; a.bt_i = d->bt_i; a.bt_o = d->bt_o; a.bt_obuf = d->bt_obuf;
    a.bt_cap = d->bt_cap; a.bt_ocap = d->bt_ocap;
    a.L.rows = d->lz_rows; a.L.sinfo = d->lz_sinfo; a.L.pool = d->lz_pool;
    a.L.requested = d->lz_requested; a.L.miss_count = d->lz_miss; a.L.miss_list = d->lz_miss ? d->lz_miss + 1 : nullptr;
    a.L.miss_cap = kMissCap; a.L.start_info = 0;
    return a;
}

bool fast_capable(const BlobHeader &h) {
    return (h.flags & PF_LENGTH_PRESERVING) && h.exec_kind != TRRE_EXEC_BACKTRACK;
}

// The length-preserving fast path on a regular buffer: ONE kernel, out offset == in offset.
// Raises ST_IRREGULAR in *status when the buffer is not regular (NUL byte / no final newline).
int launch_fast(const trre_prog *p, DeviceState *d, const uint8_t *d_in, size_t n, uint8_t *d_out,
                uint32_t *status, cudaStream_t st) {
    const BlobHeader &h = p->hdr;
    const bool aligned = ((reinterpret_cast<uintptr_t>(d_in) ^ reinterpret_cast<uintptr_t>(d_out)) & 15) == 0;
    if (h.exec_kind == TRRE_EXEC_BYTEMAP && aligned && d->swar.n) {
        const size_t nvec = n / 16 + 1;
        const int grid = (int)((nvec + kSwarVecPerCta - 1) / kSwarVecPerCta);
        switch (d->swar.n) {
            case 1: k_bytemap_swar<1><<<grid, kSwarThreads, 0, st>>>(d_in, d_out, n, d->swar, d->P.lut, status); break;
            case 2: k_bytemap_swar<2><<<grid, kSwarThreads, 0, st>>>(d_in, d_out, n, d->swar, d->P.lut, status); break;
            case 3: k_bytemap_swar<3><<<grid, kSwarThreads, 0, st>>>(d_in, d_out, n, d->swar, d->P.lut, status); break;
            default: k_bytemap_swar<4><<<grid, kSwarThreads, 0, st>>>(d_in, d_out, n, d->swar, d->P.lut, status); break;
        }
    } else if (h.exec_kind == TRRE_EXEC_BYTEMAP && aligned && d->P.pairlut) {
        k_bytemap<<<d->sm_count, kBmThreads, kPairLutBytes, st>>>(d_in, d_out, n, d->P.pairlut, d->P.lut, status);
    } else if (d->lit_words && aligned && (reinterpret_cast<uintptr_t>(d_in) & 15) == 0) {
        const size_t nvec = (n + 15) / 16;
        const int grid = (int)((nvec + kLitVecPerCta - 1) / kLitVecPerCta);
#define LIT_CASE(NW, L) case (NW) * 8 + (L):                                                             \
        if (d->lit_eq) k_lit<NW, L, true><<<grid, kLitThreads, 0, st>>>(d_in, d_out, n, d->lit, status); \
        else k_lit<NW, L, false><<<grid, kLitThreads, 0, st>>>(d_in, d_out, n, d->lit, status);          \
        break;
        switch (d->lit_words * 8 + d->lit_len) {
            LIT_CASE(1, 2) LIT_CASE(1, 3) LIT_CASE(1, 4)
            LIT_CASE(2, 2) LIT_CASE(2, 3) LIT_CASE(2, 4)
            LIT_CASE(3, 2) LIT_CASE(3, 3) LIT_CASE(3, 4)
            LIT_CASE(4, 2) LIT_CASE(4, 3) LIT_CASE(4, 4)
            default: g_last_error = "internal: literal word form out of range"; return TRRE_E_ARG;
        }
#undef LIT_CASE
    } else if (d->gram.entries && aligned && (reinterpret_cast<uintptr_t>(d_in) & 15) == 0 && !std::getenv("TRRE_NO_GRAM")) {
        GramArgs g = d->gram;
        g.in = d_in; g.out = d_out; g.n = n; g.status = status;
        const size_t n_tiles = (n + kTile - 1) / kTile;
        size_t blocks = std::min<size_t>((size_t)d->sm_count, (n_tiles + 16 * kGramWarps - 1) / (16 * kGramWarps));
        if (blocks < 1) blocks = 1;
        g.tiles_per_warp = (n_tiles + blocks * kGramWarps - 1) / (blocks * kGramWarps);
        void *kargs[] = {&g};
        CUDA_TRY(cudaLaunchKernel(d->gram_fn, dim3((unsigned)blocks), dim3(kGramThreads), kargs, d->gram_smem, st));
    } else if ((d->P.flags & PF_MATCHABLE) && d->P.row0 && aligned && (reinterpret_cast<uintptr_t>(d_in) & 15) == 0 && !std::getenv("TRRE_NO_MATCH_KERNEL")) {
        MatchArgs m{};
        m.P = d->P; m.in = d_in; m.out = d_out; m.n = n; m.status = status;
        m.smem_states = d->mp.states; m.row0_replicated = d->mp.repl; m.pool_in_smem = d->mp.pool;
        const size_t n_tiles = (n + kTile - 1) / kTile;
        // every warp owns a contiguous span of tiles; keep spans >= 16 tiles so the warm-up stays cheap
        size_t blocks = std::min<size_t>((size_t)d->mp.blocks, (n_tiles + 16 * kMatchWarps - 1) / (16 * kMatchWarps));
        if (blocks < 1) blocks = 1;
        m.tiles_per_warp = (n_tiles + blocks * kMatchWarps - 1) / (blocks * kMatchWarps);
        if (d->mp.tab == TAB_WIDE) k_match<TAB_WIDE, OP_INPLACE><<<(int)blocks, kMatchThreads, d->mp.smem, st>>>(m);
        else if (d->mp.tab == TAB_ALL) k_match<TAB_ALL, OP_INPLACE><<<(int)blocks, kMatchThreads, d->mp.smem, st>>>(m);
        else k_match<TAB_PART, OP_INPLACE><<<(int)blocks, kMatchThreads, d->mp.smem, st>>>(m);
    } else {
        LineArgs a = make_args(d, d_in, n, d_out, n, status);
        int grid = (int)std::min<size_t>(a.n_groups, (size_t)d->sm_count * 64);
        k_lines<MODE_INPLACE, EXEC_TABLE><<<grid, kLineThreads, 0, st>>>(a);
    }
    CUDA_TRY(cudaGetLastError());
    return TRRE_OK;
}

// The general path: count -> scan -> (write).  3 kernels (2 when only sizing).
int launch_general(const trre_prog *p, DeviceState *d, const uint8_t *d_in, size_t n, uint8_t *d_out,
                   size_t out_cap, cudaStream_t st, int phase, int *launches) {
    // -a / -m / -ma run the reference's DFS whatever tables the program has (trre_nft.c:635-648)
    const int amode = p->amode;
    const bool dfs_mode = amode == AM_SCAN_ALL || amode == AM_MATCH || amode == AM_MATCH_ALL;
    const bool backtrack = p->hdr.exec_kind == TRRE_EXEC_BACKTRACK || dfs_mode;
    const bool lazy = p->hdr.exec_kind == TRRE_EXEC_LAZY && amode == AM_FIRST;
    const int exec = backtrack ? EXEC_DFS : lazy ? EXEC_LAZY : EXEC_TABLE;
    if (lazy) { int rc0 = lazy_sync(const_cast<trre_prog *>(p), d, st); if (rc0 != TRRE_OK) return rc0; }
    const size_t n_chunks = (n + kChunk - 1) / kChunk;
    const size_t n_groups = (n_chunks + kLineThreads - 1) / kLineThreads;
    // persistent-style grid: the DFS executor's scratch is per resident thread
    int grid = (int)std::min<size_t>(n_groups, backtrack ? (d->bt_tier ? (size_t)8 : (size_t)d->sm_count) : (size_t)d->sm_count * 64);
    if (grid < 1) grid = 1;
    int rc = ensure_workspace(d, n_chunks, n_groups, backtrack, (size_t)grid * kLineThreads);
    if (rc != TRRE_OK) return rc;
    if (phase != PH_WRITE) d->counts_kind = 1;
    LineArgs a = make_args(d, d_in, n, d_out, out_cap, d->d_status);
    if (lazy) a.L.start_info = p->lazy->start_info();
    const bool staged = !backtrack && !lazy && amode == AM_FIRST && d->P.ctrans && !std::getenv("TRRE_NO_STAGED_LINES");
    StagedArgs sa{};
    size_t smem_count = 0, smem_write = 0;
    int grid_s = grid;
    if (staged) {
        // shared memory: classmap + cached rows of the compact table + the group's input (+ its output)
        const size_t row_bytes = (size_t)d->P.n_classes * 4;
        size_t rows = std::min<size_t>(d->P.n_states, (24 * 1024) / row_bytes);
        sa.L = a;
        sa.smem_states = (uint32_t)rows;
        const size_t ct = ((rows * d->P.n_classes + 3) & ~(size_t)3) * 4;
        smem_count = 256 + 1024 + ct + kGroupBytes + 16 + kGroupBytes / 8;
        smem_write = smem_count + kStageOut + 16;
        if (!d->staged_attr) {
            CUDA_TRY(cudaFuncSetAttribute(k_lines_s<MODE_COUNT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_count));
            CUDA_TRY(cudaFuncSetAttribute(k_lines_s<MODE_WRITE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_write));
            d->staged_attr = true;
        }
        grid_s = (int)std::min<size_t>(n_groups, (size_t)d->sm_count * 3);
        if (grid_s < 1) grid_s = 1;
        if (phase != PH_WRITE) k_lines_s<MODE_COUNT><<<grid_s, kLineThreads, smem_count, st>>>(sa);
    } else if (lazy && phase != PH_WRITE) {
        // count passes until the input meets no missing row: each pass reports the states whose rows it needed, the
        // host builds them (and a few behind them), the device copy is refreshed.  trre_dft.c:1135-1175, batched.
        trre_prog *mp = const_cast<trre_prog *>(p);
        for (int round = 0;; ++round) {
            CUDA_TRY(cudaMemsetAsync(d->lz_miss, 0, sizeof(uint32_t), st));
            launch_lines<MODE_COUNT>(exec, grid, st, a);
            *launches += 1;
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaMemcpyAsync(d->h_status, d->d_status, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(d->h_lz_miss, d->lz_miss, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            if (!(*d->h_status & ST_MISS)) break;
            const uint32_t cnt = std::min<uint32_t>(d->h_lz_miss[0], kMissCap);
            CUDA_TRY(cudaMemcpy(d->h_lz_miss + 1, d->lz_miss + 1, cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost));
            try {
                std::lock_guard<std::mutex> lock(mp->lazy_mu);
                mp->lazy->build(d->h_lz_miss + 1, cnt, 256);
            } catch (const PatternError &pe) {
                g_last_error = pe.text;
                return pe.code;
            } catch (const std::bad_alloc &) {
                g_last_error = "error: out of memory while building tables on demand\n";
                return TRRE_E_UNSUPPORTED;
            }
            int rc2 = lazy_sync(mp, d, st);
            if (rc2 != TRRE_OK) return rc2;
            a.L.rows = d->lz_rows; a.L.sinfo = d->lz_sinfo; a.L.pool = d->lz_pool; a.L.requested = d->lz_requested;
            CUDA_TRY(cudaMemsetAsync(d->d_status, 0, sizeof(uint32_t), st));
            d->lz_rounds = round + 1;
        }
        *launches -= 1;                                       // the last pass is the one counted below
    } else if (phase != PH_WRITE) {
        launch_lines<MODE_COUNT>(exec, grid, st, a, amode);
    }
    if (phase != PH_WRITE) {
        k_scan_groups<<<1, kScanThreads, 0, st>>>(d->d_gsum, n_groups, d->d_total);
        *launches += 2;
    }
    if (phase != PH_SIZE) {
        if (staged) k_lines_s<MODE_WRITE><<<std::min(grid_s, d->sm_count * 2), kLineThreads, smem_write, st>>>(sa);
        else launch_lines<MODE_WRITE>(exec, grid, st, a, amode);
        *launches += 1;
    }
    CUDA_TRY(cudaGetLastError());
    return TRRE_OK;
}

bool bmv_capable(const trre_prog *p, const DeviceState *d, const uint8_t *d_in) {
    return (((p->hdr.flags & PF_BYTEMAP_VAR) && d->vent) || d->litv_ok) && (reinterpret_cast<uintptr_t>(d_in) & 15) == 0 &&
           !std::getenv("TRRE_NO_BMV");
}

// variable-length byte map on a regular buffer: count -> scan -> write, all position parallel (bmv_kernel.cuh)
int launch_bmv(const trre_prog *p, DeviceState *d, const uint8_t *d_in, size_t n, uint8_t *d_out, size_t out_cap,
               cudaStream_t st, int phase, int *launches, const Ws *ws = nullptr) {
    const size_t n_tiles = (n + kBmvTile - 1) / kBmvTile;
    size_t n_spans = std::min<size_t>(std::max<size_t>(n_tiles / 16, 1), (size_t)d->sm_count * 64);
    const size_t tiles_per_span = (n_tiles + n_spans - 1) / n_spans;
    n_spans = (n_tiles + tiles_per_span - 1) / tiles_per_span;
    if (!ws) { int rc = ensure_workspace(d, 0, n_spans, false, 0); if (rc != TRRE_OK) return rc; }
    const int src = d->litv_ok ? BMV_LIT : p->hdr.vrun_classes ? BMV_RUN : BMV_PLAIN;
    const uint32_t pool_bytes = src == BMV_LIT ? p->hdr.gpool_bytes : d->vpool_bytes;
    if (!d->bmv_attr) {
        d->bmv_smem = 256 * 32 * 4 + ((pool_bytes + 15) & ~15u) + (size_t)(kBmvWriteThreads / 32) * kBmvWarpSmem;
        CUDA_TRY(cudaFuncSetAttribute(k_bmv_count<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairLutBytes));
        CUDA_TRY(cudaFuncSetAttribute(k_bmv_count<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairLutBytes));
        CUDA_TRY(cudaFuncSetAttribute(k_bmv_write<BMV_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->bmv_smem));
        CUDA_TRY(cudaFuncSetAttribute(k_bmv_write<BMV_RUN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->bmv_smem));
#define LITV_ATTR(NW, L) CUDA_TRY(cudaFuncSetAttribute(k_bmv_write<BMV_LIT, NW, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->bmv_smem));
        LITV_ATTR(1, 2) LITV_ATTR(1, 3) LITV_ATTR(1, 4) LITV_ATTR(2, 2) LITV_ATTR(2, 3) LITV_ATTR(2, 4)
        LITV_ATTR(3, 2) LITV_ATTR(3, 3) LITV_ATTR(3, 4) LITV_ATTR(4, 2) LITV_ATTR(4, 3) LITV_ATTR(4, 4)
#undef LITV_ATTR
        d->bmv_attr = true;
    }
    BmvArgs a{};
    a.in = d_in; a.out = d_out; a.n = n; a.out_cap = out_cap;
    a.vent = d->vent; a.vpairlen = d->vpairlen; a.vpool = d->vpool; a.vpool_bytes = d->vpool_bytes;
    if (src == BMV_LIT) { a.vpool = d->d_blob + p->hdr.off_gpool; a.vpool_bytes = p->hdr.gpool_bytes; }
    a.gsum = ws ? ws->gsum : d->d_gsum; a.total = ws ? ws->total : d->d_total; a.status = ws ? ws->status : d->d_status;
    a.n_spans = n_spans; a.tiles_per_span = tiles_per_span;
    if (phase != PH_WRITE) {
        if (src == BMV_LIT) {
#define LITV_COUNT(NW, L) case (NW) * 8 + (L): k_litv_count<NW, L><<<d->sm_count, kBmvCountThreads, 0, st>>>(a, d->litv); break;
            switch (p->hdr.lit_words * 8 + p->hdr.lit_len) {
                LITV_COUNT(1, 2) LITV_COUNT(1, 3) LITV_COUNT(1, 4) LITV_COUNT(2, 2) LITV_COUNT(2, 3) LITV_COUNT(2, 4)
                LITV_COUNT(3, 2) LITV_COUNT(3, 3) LITV_COUNT(3, 4) LITV_COUNT(4, 2) LITV_COUNT(4, 3) LITV_COUNT(4, 4)
                default: g_last_error = "internal: literal word form out of range"; return TRRE_E_ARG;
            }
#undef LITV_COUNT
        } else if (src == BMV_RUN && d->runmap.n) k_bmv_count<true><<<d->sm_count, kBmvCountThreads, kPairLutBytes, st>>>(a, d->runmap);
        else if (src == BMV_RUN) k_bmv_count_run<<<d->sm_count, kBmvCountThreads, 256 * 32 * 4, st>>>(a);
        else k_bmv_count<false><<<d->sm_count, kBmvCountThreads, kPairLutBytes, st>>>(a, d->runmap);
        k_scan_groups<<<1, kScanThreads, 0, st>>>(a.gsum, n_spans, const_cast<unsigned long long *>(a.total));
        *launches += 2;
        if (!ws) d->counts_kind = 2;
    }
    if (phase != PH_SIZE) {
        const size_t per = kBmvWriteThreads / 32;
        int grid = (int)std::min<size_t>((n_spans + per - 1) / per, (size_t)d->sm_count * 2);
        if (src == BMV_LIT) {
#define LITV_WRITE(NW, L) case (NW) * 8 + (L): k_bmv_write<BMV_LIT, NW, L><<<grid < 1 ? 1 : grid, kBmvWriteThreads, d->bmv_smem, st>>>(a, d->litv); break;
            switch (p->hdr.lit_words * 8 + p->hdr.lit_len) {
                LITV_WRITE(1, 2) LITV_WRITE(1, 3) LITV_WRITE(1, 4) LITV_WRITE(2, 2) LITV_WRITE(2, 3) LITV_WRITE(2, 4)
                LITV_WRITE(3, 2) LITV_WRITE(3, 3) LITV_WRITE(3, 4) LITV_WRITE(4, 2) LITV_WRITE(4, 3) LITV_WRITE(4, 4)
                default: g_last_error = "internal: literal word form out of range"; return TRRE_E_ARG;
            }
#undef LITV_WRITE
        } else if (src == BMV_RUN) k_bmv_write<BMV_RUN><<<grid < 1 ? 1 : grid, kBmvWriteThreads, d->bmv_smem, st>>>(a, d->litv);
        else k_bmv_write<BMV_PLAIN><<<grid < 1 ? 1 : grid, kBmvWriteThreads, d->bmv_smem, st>>>(a, d->litv);
        *launches += 1;
    }
    CUDA_TRY(cudaGetLastError());
    return TRRE_OK;
}

bool mv_capable(const trre_prog *p, const DeviceState *d, const uint8_t *d_in) {
    return (p->hdr.flags & PF_MATCHABLE_VAR) && d->P.eout && d->mpv.smem && (reinterpret_cast<uintptr_t>(d_in) & 15) == 0 &&
           !std::getenv("TRRE_NO_MATCH_KERNEL");
}

// variable-length matches on a regular buffer: k_match<OP_COUNT> -> scan over the warp spans -> k_match<OP_WRITE>
int launch_mv(const trre_prog *p, DeviceState *d, const uint8_t *d_in, size_t n, uint8_t *d_out, size_t out_cap,
              cudaStream_t st, int phase, int *launches, const Ws *ws = nullptr) {
    const bool gram = d->gramv.entries && !std::getenv("TRRE_NO_GRAM");
    const size_t warps = gram ? kGramWarps : kMatchWarps;
    const size_t n_tiles = (n + kTile - 1) / kTile;
    size_t blocks = std::min<size_t>(gram ? (size_t)d->sm_count : (size_t)d->mpv.blocks, (n_tiles + 16 * warps - 1) / (16 * warps));
    if (blocks < 1) blocks = 1;
    const size_t n_spans = blocks * warps;
    if (!ws) { int rc = ensure_workspace(d, 0, n_spans, false, 0); if (rc != TRRE_OK) return rc; }
    MatchArgs m{};
    m.P = d->P; m.in = d_in; m.out = d_out; m.n = n; m.status = ws ? ws->status : d->d_status;
    m.smem_states = d->mpv.states; m.row0_replicated = d->mpv.repl; m.pool_in_smem = d->mpv.pool;
    m.tiles_per_warp = (n_tiles + n_spans - 1) / n_spans;
    m.gsum = ws ? ws->gsum : d->d_gsum; m.total = ws ? ws->total : d->d_total; m.out_cap = out_cap;
    const int tab = d->mpv.tab;
    GramArgs g = d->gramv;
    void *gargs[] = {&m, &g};
    if (phase != PH_WRITE) {
        if (gram) CUDA_TRY(cudaLaunchKernel(d->gramv_fn[0], dim3((unsigned)blocks), dim3(kGramThreads), gargs, d->gramv_smem, st));
        else if (tab == TAB_WIDE) k_match<TAB_WIDE, OP_COUNT><<<(int)blocks, kMatchThreads, d->mpv.smem, st>>>(m);
        else if (tab == TAB_ALL) k_match<TAB_ALL, OP_COUNT><<<(int)blocks, kMatchThreads, d->mpv.smem, st>>>(m);
        else k_match<TAB_PART, OP_COUNT><<<(int)blocks, kMatchThreads, d->mpv.smem, st>>>(m);
        k_scan_groups<<<1, kScanThreads, 0, st>>>(m.gsum, n_spans, const_cast<unsigned long long *>(m.total));
        *launches += 2;
        if (!ws) d->counts_kind = 3;
    }
    if (phase != PH_SIZE) {
        if (gram) CUDA_TRY(cudaLaunchKernel(d->gramv_fn[1], dim3((unsigned)blocks), dim3(kGramThreads), gargs, d->gramv_smem, st));
        else if (tab == TAB_WIDE) k_match<TAB_WIDE, OP_WRITE><<<(int)blocks, kMatchThreads, d->mpv.smem, st>>>(m);
        else if (tab == TAB_ALL) k_match<TAB_ALL, OP_WRITE><<<(int)blocks, kMatchThreads, d->mpv.smem, st>>>(m);
        else k_match<TAB_PART, OP_WRITE><<<(int)blocks, kMatchThreads, d->mpv.smem, st>>>(m);
        *launches += 1;
    }
    CUDA_TRY(cudaGetLastError());
    return TRRE_OK;
}

int clear_status(DeviceState *d, cudaStream_t st) {
    CUDA_TRY(cudaMemsetAsync(d->d_status, 0, sizeof(uint32_t), st));
    CUDA_TRY(cudaMemsetAsync(d->d_total, 0, sizeof(unsigned long long), st));
    d->status_dirty = false;
    return TRRE_OK;
}

// Enqueue one whole-buffer scan.  Never synchronises.  On a length-preserving program only the fast
// kernel is enqueued; should it find the buffer irregular, collect() reruns it on the general path.
int enqueue_scan(const trre_prog *p, DeviceState *d, const uint8_t *d_in, size_t n, uint8_t *d_out,
                 size_t out_cap, cudaStream_t st, bool size_only) {
    d->launches = 0;
    d->pend_fast = false;
    d->pend_in = d_in; d->pend_out = d_out; d->pend_n = n; d->pend_cap = out_cap;
    if (d->status_dirty) { int rc = clear_status(d, st); if (rc != TRRE_OK) return rc; }
    if (n == 0) return TRRE_OK;
    d->pend_phase = size_only ? PH_SIZE : PH_ALL;
    if (p->amode != AM_FIRST) {
        d->status_dirty = true;
        return launch_general(p, d, d_in, n, d_out, out_cap, st, d->pend_phase, &d->launches);
    }
    if (!size_only && fast_capable(p->hdr) && out_cap >= n) {
        int rc = launch_fast(p, d, d_in, n, d_out, d->d_status, st);
        if (rc != TRRE_OK) return rc;
        d->launches = 1;
        d->pend_fast = true;
        d->pend_total_is_n = true;
        return TRRE_OK;
    }
    d->status_dirty = true;                 // k_scan_groups writes d_total
    if (bmv_capable(p, d, d_in)) {
        d->pend_fast = true;
        d->pend_total_is_n = false;
        return launch_bmv(p, d, d_in, n, d_out, out_cap, st, d->pend_phase, &d->launches);
    }
    if (mv_capable(p, d, d_in)) {
        d->pend_fast = true;
        d->pend_total_is_n = false;
        return launch_mv(p, d, d_in, n, d_out, out_cap, st, d->pend_phase, &d->launches);
    }
    return launch_general(p, d, d_in, n, d_out, out_cap, st, d->pend_phase, &d->launches);
}

int status_to_rc(uint32_t s) {
    if (s & (ST_DIVERGE | ST_STACK)) { g_last_error = "error: stack max capacity reached\n"; return TRRE_E_STACK; }
    if (s & ST_OBUF) { g_last_error = "error: one match attempt produced more than 1 MiB of output (limit of the GPU DFS executor, not of trre)\n"; return TRRE_E_UNSUPPORTED; }
    if (s & ST_RUNAWAY) {
        g_last_error = "error: the input leads into an epsilon cycle of the transducer (trre_dft recurses for ever there)\n";
        return TRRE_E_UNSUPPORTED;
    }
    if (s & ST_BUDGET) {
        g_last_error = "error: one match attempt went past 2^26 backtracking steps or 16 MiB of printed paths (limit of the GPU DFS executor, not of trre)\n";
        return TRRE_E_UNSUPPORTED;
    }
    if (s & ST_NOSPACE) { g_last_error = "output buffer too small"; return TRRE_E_NOSPACE; }
    return TRRE_OK;
}

// Synchronise, finish what the enqueue left open, report size and status.
int collect(const trre_prog *p, DeviceState *d, cudaStream_t st, size_t *out_n) {
    CUDA_TRY(cudaMemcpyAsync(d->h_status, d->d_status, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    const bool total_is_n = d->pend_fast && d->pend_total_is_n;
    if (!total_is_n) CUDA_TRY(cudaMemcpyAsync(d->h_total, d->d_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    uint32_t s = *d->h_status;
    size_t total = total_is_n ? d->pend_n : (size_t)*d->h_total;
    if (d->pend_fast && (s & ST_IRREGULAR)) {
        // NUL bytes or no trailing newline: same buffers, general path
        int rc = clear_status(d, st);
        if (rc != TRRE_OK) return rc;
        d->pend_fast = false;
        d->status_dirty = true;
        rc = launch_general(p, d, d->pend_in, d->pend_n, d->pend_out, d->pend_cap, st, d->pend_phase, &d->launches);
        if (rc != TRRE_OK) return rc;
        CUDA_TRY(cudaMemcpyAsync(d->h_status, d->d_status, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(d->h_total, d->d_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        s = *d->h_status;
        total = (size_t)*d->h_total;
    }
    if ((s & (ST_STACK | ST_OBUF)) && (p->hdr.exec_kind == TRRE_EXEC_BACKTRACK || p->amode != AM_FIRST) && d->bt_tier == 0) {
        // the small per-thread DFS stacks overflowed: once more with the reference's capacity (see ensure_workspace)
        d->bt_tier = 1;
        int rc = clear_status(d, st);
        if (rc != TRRE_OK) return rc;
        d->status_dirty = true;
        rc = launch_general(p, d, d->pend_in, d->pend_n, d->pend_out, d->pend_cap, st, d->pend_phase, &d->launches);
        if (rc != TRRE_OK) return rc;
        CUDA_TRY(cudaMemcpyAsync(d->h_status, d->d_status, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(d->h_total, d->d_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        s = *d->h_status;
        total = (size_t)*d->h_total;
    }
    if (s) d->status_dirty = true;
    if (out_n) *out_n = total;
    return status_to_rc(s & ~ST_IRREGULAR);
}

int ensure_host_path_buffers(DeviceState *d, size_t need_in, size_t need_out) {
    if (need_in > d->cap_in) {
        if (d->buf_in) cudaFree(d->buf_in);
        d->buf_in = nullptr; d->cap_in = 0;
        CUDA_TRY(cudaMalloc(&d->buf_in, need_in));
        d->cap_in = need_in;
    }
    if (need_out > d->cap_out) {
        if (d->buf_out) cudaFree(d->buf_out);
        d->buf_out = nullptr; d->cap_out = 0;
        CUDA_TRY(cudaMalloc(&d->buf_out, need_out));
        d->cap_out = need_out;
    }
    return TRRE_OK;
}

int ensure_slots(DeviceState *d, size_t chunk, size_t out_cap) {
    if (d->slot_cap >= chunk && d->slot_out_cap >= out_cap) return TRRE_OK;
    for (int k = 0; k < kSlots; ++k) {
        if (d->slot_in[k]) cudaFree(d->slot_in[k]);
        if (d->slot_out[k]) cudaFree(d->slot_out[k]);
        d->slot_in[k] = d->slot_out[k] = nullptr;
    }
    d->slot_cap = d->slot_out_cap = 0;
    const size_t max_spans = (size_t)d->sm_count * 64;            // >= the spans of k_bmv and of k_match<var>
    for (int k = 0; k < kSlots; ++k) {
        CUDA_TRY(cudaMalloc(&d->slot_in[k], chunk));
        CUDA_TRY(cudaMalloc(&d->slot_out[k], out_cap));
        if (!d->slot_stream[k]) CUDA_TRY(cudaStreamCreateWithFlags(&d->slot_stream[k], cudaStreamNonBlocking));
        if (!d->slot_gsum[k]) CUDA_TRY(cudaMalloc(&d->slot_gsum[k], max_spans * sizeof(unsigned long long)));
        if (!d->slot_done[k]) CUDA_TRY(cudaEventCreateWithFlags(&d->slot_done[k], cudaEventDisableTiming));
    }
    if (!d->d_slot_total) {
        CUDA_TRY(cudaMalloc(&d->d_slot_total, kSlots * sizeof(unsigned long long)));
        CUDA_TRY(cudaMallocHost(&d->h_slot_total, kSlots * sizeof(unsigned long long)));
    }
    d->slot_cap = chunk;
    d->slot_out_cap = out_cap;
    return TRRE_OK;
}

// Pinned-host copies reach the full PCIe rate only from / to 256-byte aligned HOST addresses (measured with bare chunked
// copies, scripts/e2e_probe2.py: 48.2 GB/s aligned, 42.8 GB/s at any other offset - exactly what the pipeline lost; the
// device-side alignment does not matter).  Chunks start where lines start, so every copy is split into a head of < 256
// bytes and an aligned bulk.
cudaError_t copy_h2d(uint8_t *d_dst, const uint8_t *h_src, size_t n, cudaStream_t st) {
    size_t head = (256 - (reinterpret_cast<uintptr_t>(h_src) & 255)) & 255;
    if (head > n) head = n;
    cudaError_t e = cudaSuccess;
    if (n > head) e = cudaMemcpyAsync(d_dst + head, h_src + head, n - head, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && head) e = cudaMemcpyAsync(d_dst, h_src, head, cudaMemcpyHostToDevice, st);
    return e;
}
cudaError_t copy_d2h(uint8_t *h_dst, const uint8_t *d_src, size_t n, cudaStream_t st) {
    size_t head = (256 - (reinterpret_cast<uintptr_t>(h_dst) & 255)) & 255;
    if (head > n) head = n;
    cudaError_t e = cudaSuccess;
    if (n > head) e = cudaMemcpyAsync(h_dst + head, d_src + head, n - head, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && head) e = cudaMemcpyAsync(h_dst, d_src, head, cudaMemcpyDeviceToHost, st);
    return e;
}

size_t host_chunk_bytes() {
    size_t chunk = 32u << 20;
    if (const char *e = std::getenv("TRRE_CHUNK_MB")) { long v = std::atol(e); if (v >= 1 && v <= 1024) chunk = (size_t)v << 20; }
    return chunk;
}

// Host buffers, position-parallel program whose output length differs from the input's (k_bmv / k_match var):
// same 3-slot pipeline, but a chunk's D2H can only be issued once its size is known, so chunk k is finished
// (size read back, D2H enqueued at the running host offset) while chunks k+1 and k+2 are already in flight.
// Returns 0 done, 1 = fall back to the one-shot path (irregular buffer, expansion beyond the slot, no room).
int pipelined_var(const trre_prog *p, DeviceState *d, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap,
                  size_t *out_total, int *launches, bool bmv) {
    const size_t chunk = host_chunk_bytes();
    int rc = ensure_slots(d, chunk, 2 * chunk + 64);
    if (rc != TRRE_OK) return rc;
    CUDA_TRY(cudaMemsetAsync(d->d_status + 1, 0, kSlots * sizeof(uint32_t), d->slot_stream[0]));
    CUDA_TRY(cudaStreamSynchronize(d->slot_stream[0]));
    struct Piece { size_t pos, len; };
    std::vector<Piece> pieces;
    for (size_t pos = 0; pos < n;) {
        size_t end = std::min(pos + chunk, n);
        if (end < n) {
            const void *nl = memrchr(in + pos, '\n', end - pos);
            if (!nl) return 1;                                     // a line longer than a chunk
            end = (size_t)(static_cast<const uint8_t *>(nl) - in) + 1;
        }
        pieces.push_back(Piece{pos, end - pos});
        pos = end;
    }
    size_t host_off = 0;
    bool give_up = false;
    auto finish = [&](size_t k) -> int {
        const int slot = (int)(k % kSlots);
        CUDA_TRY(cudaEventSynchronize(d->slot_done[slot]));
        const size_t total = (size_t)d->h_slot_total[slot];
        if (d->h_status[1 + slot] || total > d->slot_out_cap || host_off + total > out_cap) { give_up = true; return TRRE_OK; }
        CUDA_TRY(copy_d2h(out + host_off, d->slot_out[slot], total, d->slot_stream[slot]));
        host_off += total;
        return TRRE_OK;
    };
    for (size_t k = 0; k < pieces.size() && !give_up; ++k) {
        const int slot = (int)(k % kSlots);
        cudaStream_t st = d->slot_stream[slot];
        Ws ws{d->slot_gsum[slot], d->d_slot_total + slot, d->d_status + 1 + slot};
        CUDA_TRY(copy_h2d(d->slot_in[slot], in + pieces[k].pos, pieces[k].len, st));
        rc = bmv ? launch_bmv(p, d, d->slot_in[slot], pieces[k].len, d->slot_out[slot], d->slot_out_cap, st, PH_ALL, launches, &ws)
                 : launch_mv(p, d, d->slot_in[slot], pieces[k].len, d->slot_out[slot], d->slot_out_cap, st, PH_ALL, launches, &ws);
        if (rc != TRRE_OK) return rc;
        CUDA_TRY(cudaMemcpyAsync(d->h_slot_total + slot, d->d_slot_total + slot, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(d->h_status + 1 + slot, d->d_status + 1 + slot, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaEventRecord(d->slot_done[slot], st));
        if (k + 1 >= (size_t)kSlots) { rc = finish(k + 1 - kSlots); if (rc != TRRE_OK) return rc; }
    }
    for (size_t k = pieces.size() >= (size_t)kSlots ? pieces.size() - kSlots + 1 : 0; k < pieces.size() && !give_up; ++k) {
        rc = finish(k);
        if (rc != TRRE_OK) return rc;
    }
    for (int s2 = 0; s2 < kSlots; ++s2) CUDA_TRY(cudaStreamSynchronize(d->slot_stream[s2]));
    if (give_up) return 1;
    *out_total = host_off;
    return 0;
}

// Host buffers, length-preserving program: the input is cut at line boundaries into chunks that flow
// through kSlots (H2D -> fast kernel -> D2H) on separate streams, so the two copy engines and the SMs
// overlap.  Returns 1 when the buffer turned out irregular (caller falls back), 0 when done.
int pipelined_lp(const trre_prog *p, DeviceState *d, const uint8_t *in, size_t n, uint8_t *out, int *launches) {
    const size_t chunk = host_chunk_bytes();
    int rc = ensure_slots(d, chunk, std::max(d->slot_out_cap, chunk));
    if (rc != TRRE_OK) return rc;
    CUDA_TRY(cudaMemsetAsync(d->d_status + 1, 0, kSlots * sizeof(uint32_t), d->slot_stream[0]));
    CUDA_TRY(cudaStreamSynchronize(d->slot_stream[0]));
    size_t pos = 0;
    int k = 0;
    bool irregular = false;
    while (pos < n) {
        size_t end = std::min(pos + chunk, n);
        if (end < n) {
            const void *nl = memrchr(in + pos, '\n', end - pos);
            if (!nl) { irregular = true; break; }            // a line longer than a chunk
            end = (size_t)(static_cast<const uint8_t *>(nl) - in) + 1;
        }
        const int slot = k % kSlots;
        cudaStream_t st = d->slot_stream[slot];               // stream order = reuse order of the slot
        CUDA_TRY(copy_h2d(d->slot_in[slot], in + pos, end - pos, st));
        rc = launch_fast(p, d, d->slot_in[slot], end - pos, d->slot_out[slot], d->d_status + 1 + slot, st);
        if (rc != TRRE_OK) return rc;
        CUDA_TRY(copy_d2h(out + pos, d->slot_out[slot], end - pos, st));
        *launches += 1;
        pos = end;
        ++k;
    }
    for (int s2 = 0; s2 < kSlots; ++s2) CUDA_TRY(cudaStreamSynchronize(d->slot_stream[s2]));
    if (irregular) return 1;
    CUDA_TRY(cudaMemcpy(d->h_status + 1, d->d_status + 1, kSlots * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (int s2 = 0; s2 < kSlots; ++s2)
        if (d->h_status[1 + s2]) return 1;
    return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
extern "C" {

const char *trre_version(void) { return "trre_b200 0.1 (sm_100a)"; }
const char *trre_last_error(void) { return g_last_error.c_str(); }
void trre_free(void *p) { std::free(p); }

void *trre_host_alloc(size_t nbytes) {
    void *p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, nbytes ? nbytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        g_last_error = std::string("no CUDA device: the trre_b200 scan path has no CPU fallback (") + cudaGetErrorString(e) + ")";
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void trre_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

int trre_compile(const char *expr, int flavor, trre_prog **out, char *err, size_t errcap) {
    if (!expr || !out) { set_err(err, errcap, "null argument"); return TRRE_E_ARG; }
    *out = nullptr;
    try {
        std::vector<uint8_t> blob = compile_to_blob(expr, flavor);
        trre_prog *p = new trre_prog();
        p->blob = std::move(blob);
        std::memcpy(&p->hdr, p->blob.data(), sizeof(BlobHeader));
        attach_lazy(p);
        *out = p;
        if (err && errcap) err[0] = 0;
        return TRRE_OK;
    } catch (const PatternError &e) {
        set_err(err, errcap, e.text);
        return e.code;
    } catch (const std::bad_alloc &) {
        set_err(err, errcap, "error: node memory allocation failed\n");
        return TRRE_E_PATTERN;
    }
}

int trre_prog_from_blob(const void *blob, size_t nbytes, trre_prog **out, char *err, size_t errcap) {
    if (!out) return TRRE_E_ARG;
    *out = nullptr;
    std::string why;
    if (!blob_is_valid(blob, nbytes, &why)) { set_err(err, errcap, "invalid program blob: " + why); return TRRE_E_ARG; }
    trre_prog *p = new trre_prog();
    p->blob.assign(static_cast<const uint8_t *>(blob), static_cast<const uint8_t *>(blob) + nbytes);
    std::memcpy(&p->hdr, p->blob.data(), sizeof(BlobHeader));
    try {
        attach_lazy(p);
    } catch (const PatternError &e) {
        set_err(err, errcap, e.text);
        delete p;
        return e.code;
    }
    *out = p;
    return TRRE_OK;
}

int trre_lazy_build(trre_prog *p, const uint32_t *states, size_t n, uint32_t extra) {
    if (!p || !p->lazy || (n && !states)) return TRRE_E_ARG;
    try {
        std::lock_guard<std::mutex> lock(p->lazy_mu);
        for (size_t k = 0; k < n; ++k)
            if (states[k] >= p->lazy->n_states()) return TRRE_E_ARG;
        p->lazy->build(states, n, extra);
    } catch (const PatternError &e) {
        g_last_error = e.text;
        return e.code;
    }
    return TRRE_OK;
}

int trre_lazy_view(const trre_prog *p, uint32_t *n_states, const uint64_t **rows, const uint64_t **sinfo,
                   const uint8_t **pool, size_t *pool_bytes) {
    if (!p || !p->lazy) return TRRE_E_ARG;
    if (n_states) *n_states = p->lazy->n_states();
    if (rows) *rows = p->lazy->rows().data();
    if (sinfo) *sinfo = p->lazy->sinfo().data();
    if (pool) *pool = p->lazy->pool().data();
    if (pool_bytes) *pool_bytes = p->lazy->pool().size();
    return TRRE_OK;
}

void trre_prog_free(trre_prog *p) {
    if (!p) return;
    for (auto &kv : p->dev) {
        DeviceState *d = kv.second;
        DeviceGuard g(d->device);
        cudaFree(d->d_blob); cudaFree(d->d_cnt); cudaFree(d->d_gsum); cudaFree(d->d_total); cudaFree(d->d_status);
        cudaFreeHost(d->h_total); cudaFreeHost(d->h_status);
        if (d->bt_s) { cudaFree(d->bt_s); cudaFree(d->bt_i); cudaFree(d->bt_o); cudaFree(d->bt_obuf); }
        if (d->d_slot_total) { cudaFree(d->d_slot_total); cudaFreeHost(d->h_slot_total); }
        if (d->buf_in) cudaFree(d->buf_in);
        if (d->buf_out) cudaFree(d->buf_out);
        if (d->lz_rows) { cudaFree(d->lz_rows); cudaFree(d->lz_sinfo); cudaFree(d->lz_requested); }
        if (d->lz_pool) cudaFree(d->lz_pool);
        if (d->lz_miss) { cudaFree(d->lz_miss); cudaFreeHost(d->h_lz_miss); }
        for (int k = 0; k < kSlots; ++k) {
            if (d->slot_in[k]) cudaFree(d->slot_in[k]);
            if (d->slot_out[k]) cudaFree(d->slot_out[k]);
            if (d->slot_stream[k]) cudaStreamDestroy(d->slot_stream[k]);
            if (d->slot_gsum[k]) cudaFree(d->slot_gsum[k]);
            if (d->slot_done[k]) cudaEventDestroy(d->slot_done[k]);
        }
        delete d;
    }
    delete p->lazy;
    delete p;
}

const void *trre_prog_blob(const trre_prog *p, size_t *nbytes) {
    if (!p) return nullptr;
    if (nbytes) *nbytes = p->blob.size();
    return p->blob.data();
}

int trre_prog_get_info(const trre_prog *p, trre_prog_info *info) {
    if (!p || !info) return TRRE_E_ARG;
    const BlobHeader &h = p->hdr;
    info->abi_version = h.abi;
    info->flavor = h.flavor;
    info->exec_kind = h.exec_kind;
    info->nft_states = h.nft_states;
    info->dsa_states = h.dsa_states;
    info->byte_classes = h.n_classes;
    info->max_depth = h.max_depth;
    info->length_preserving = (h.flags & PF_LENGTH_PRESERVING) ? 1 : 0;
    info->can_match_empty = (h.flags & PF_CAN_MATCH_EMPTY) ? 1 : 0;
    info->table_bytes = h.total_bytes;
    info->fast_path = h.exec_kind == TRRE_EXEC_BYTEMAP    ? TRRE_FAST_BYTEMAP
                      : (h.flags & PF_MATCHABLE)        ? TRRE_FAST_MATCH
                      : (h.flags & PF_BYTEMAP_VAR)      ? TRRE_FAST_BYTEMAP_VAR
                      : (h.flags & PF_MATCHABLE_VAR)    ? TRRE_FAST_MATCH_VAR
                      : ((h.flags & PF_LENGTH_PRESERVING) && h.exec_kind == TRRE_EXEC_TABLE) ? TRRE_FAST_INPLACE
                                                        : TRRE_FAST_NONE;
    info->gram_tables = (h.flags & PF_GRAM) ? 1 : 0;
    info->literal_words = h.lit_words;
    return TRRE_OK;
}

int trre_prog_set_mode(trre_prog *p, int match, int all) {
    if (!p) return TRRE_E_ARG;
    if (all && p->hdr.flavor == TRRE_FLAVOR_DFT) { g_last_error = "Not supported yet\n"; return TRRE_E_UNSUPPORTED; }   // trre_dft.c:1226
    p->amode = match ? (p->hdr.flavor == TRRE_FLAVOR_DFT ? AM_NEWLINES : all ? AM_MATCH_ALL : AM_MATCH)
                     : (all ? AM_SCAN_ALL : AM_FIRST);
    return TRRE_OK;
}

int trre_prog_bytemap(const trre_prog *p, uint8_t lut[256]) {
    if (!p || !lut || p->hdr.exec_kind != TRRE_EXEC_BYTEMAP) return TRRE_E_ARG;
    std::memcpy(lut, p->blob.data() + p->hdr.off_lut, 256);
    return TRRE_OK;
}

int trre_scan_device(const trre_prog *p, int device, const uint8_t *d_in, size_t n, uint8_t *d_out,
                     size_t out_cap, size_t *out_n, void *stream) {
    if (!p || (n && (!d_in || !d_out))) { g_last_error = "null argument"; return TRRE_E_ARG; }
    DeviceGuard guard(device);
    if (!guard.ok) { g_last_error = "no CUDA device: the trre_b200 scan path has no CPU fallback"; return TRRE_E_CUDA; }
    DeviceState *d = nullptr;
    int rc = get_state(p, device, &d);
    if (rc != TRRE_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rc = enqueue_scan(p, d, d_in, n, d_out, out_cap, st, false);
    if (rc != TRRE_OK) return rc;
    if (!out_n) return TRRE_OK;
    return collect(p, d, st, out_n);
}

int trre_scan_status(const trre_prog *p, int device, size_t *out_n, void *stream) {
    if (!p) return TRRE_E_ARG;
    DeviceGuard guard(device);
    if (!guard.ok) { g_last_error = "no CUDA device: the trre_b200 scan path has no CPU fallback"; return TRRE_E_CUDA; }
    DeviceState *d = nullptr;
    int rc = get_state(p, device, &d);
    if (rc != TRRE_OK) return rc;
    return collect(p, d, static_cast<cudaStream_t>(stream), out_n);
}

int trre_scan_launches(const trre_prog *p, int device) {
    if (!p) return -1;
    DeviceGuard guard(device);
    if (!guard.ok) return -1;
    DeviceState *d = nullptr;
    if (get_state(p, device, &d) != TRRE_OK) return -1;
    return d->launches;
}

int trre_scan_into(const trre_prog *p, int device, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap,
                   size_t *out_n) {
    if (!p || (n && !in)) { g_last_error = "null argument"; return TRRE_E_ARG; }
    if (out_n) *out_n = 0;
    if (n == 0) return TRRE_OK;                                   // empty input: empty output (trre_nft.c:776: no record)
    DeviceGuard guard(device);
    if (!guard.ok) { g_last_error = "no CUDA device: the trre_b200 scan path has no CPU fallback"; return TRRE_E_CUDA; }
    DeviceState *d = nullptr;
    int rc = get_state(p, device, &d);
    if (rc != TRRE_OK) return rc;
    d->launches = 0;
    // A buffer whose last record has no '\n' (a file without a final newline) is irregular only in that record - getline's
    // last record loses its last byte (trre_nft.c:776-777).  The position-parallel kernels take everything up to the last
    // '\n'; only the tail goes to the per-line kernels, instead of the whole buffer.
    if (in[n - 1] != '\n' && p->hdr.exec_kind != TRRE_EXEC_BACKTRACK && p->hdr.exec_kind != TRRE_EXEC_LAZY && p->amode == AM_FIRST) {
        if (const void *nl = memrchr(in, '\n', n)) {
            const size_t cut = (size_t)(static_cast<const uint8_t *>(nl) - in) + 1;
            size_t n1 = 0, n2 = 0;
            int rc1 = trre_scan_into(p, device, in, cut, out, out_cap, &n1);
            if (rc1 != TRRE_OK && rc1 != TRRE_E_NOSPACE) return rc1;
            const int l1 = d->launches;
            const bool room = rc1 == TRRE_OK && out && out_cap >= n1;
            int rc2 = trre_scan_into(p, device, in + cut, n - cut, room ? out + n1 : nullptr, room ? out_cap - n1 : 0, &n2);
            d->launches += l1;
            if (out_n) *out_n = n1 + n2;
            if (rc2 != TRRE_OK && rc2 != TRRE_E_NOSPACE) return rc2;
            if (rc1 == TRRE_E_NOSPACE || rc2 == TRRE_E_NOSPACE) { g_last_error = "output buffer too small"; return TRRE_E_NOSPACE; }
            return TRRE_OK;
        }
    }
    cudaStream_t st = nullptr;
    // length-preserving program and room for n bytes: chunked, copy/compute overlapped
    if (fast_capable(p->hdr) && p->amode == AM_FIRST && out && out_cap >= n) {
        int launches = 0;
        rc = pipelined_lp(p, d, in, n, out, &launches);
        d->launches = launches;
        if (rc == 0) { if (out_n) *out_n = n; return TRRE_OK; }
        if (rc != 1) return rc;                                  // 1 = irregular buffer: general path below
    }
    if (out && out_cap && p->amode == AM_FIRST && ((p->hdr.flags & PF_BYTEMAP_VAR) || (p->hdr.flags & PF_MATCHABLE_VAR)) &&
        (bmv_capable(p, d, nullptr) || mv_capable(p, d, nullptr))) {
        int launches = 0;
        size_t total = 0;
        rc = pipelined_var(p, d, in, n, out, out_cap, &total, &launches, bmv_capable(p, d, nullptr));
        d->launches = launches;
        if (rc == 0) { if (out_n) *out_n = total; return TRRE_OK; }
        if (rc != 1) return rc;                                  // 1: irregular / too large for the slots: one-shot path below
    }
    rc = ensure_host_path_buffers(d, n, 0);
    if (rc != TRRE_OK) return rc;
    CUDA_TRY(copy_h2d(d->buf_in, in, n, st));
    size_t total = 0;
    rc = enqueue_scan(p, d, d->buf_in, n, nullptr, 0, st, true);      // count + scan: exact output size
    if (rc != TRRE_OK) return rc;
    int launches = d->launches;
    rc = collect(p, d, st, &total);
    if (rc != TRRE_OK && rc != TRRE_E_NOSPACE) return rc;
    if (out_n) *out_n = total;
    if (total > out_cap || (total && !out)) { g_last_error = "output buffer too small"; return TRRE_E_NOSPACE; }
    rc = ensure_host_path_buffers(d, n, total ? total : 1);
    if (rc != TRRE_OK) return rc;
    // the counts are still in the workspace: only the write pass is left
    {
        if (d->status_dirty) { CUDA_TRY(cudaMemsetAsync(d->d_status, 0, sizeof(uint32_t), st)); }
        int wl = 0;
        rc = d->counts_kind == 2   ? launch_bmv(p, d, d->buf_in, n, d->buf_out, total, st, PH_WRITE, &wl)
             : d->counts_kind == 3 ? launch_mv(p, d, d->buf_in, n, d->buf_out, total, st, PH_WRITE, &wl)
                                   : launch_general(p, d, d->buf_in, n, d->buf_out, total, st, PH_WRITE, &wl);
        if (rc != TRRE_OK) return rc;
        d->pend_fast = false;
        d->status_dirty = true;
        size_t t2 = 0;
        rc = collect(p, d, st, &t2);
        d->launches = launches + wl;
        if (rc != TRRE_OK) return rc;
    }
    CUDA_TRY(copy_d2h(out, d->buf_out, total, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return TRRE_OK;
}

int trre_scan(const trre_prog *p, int device, const uint8_t *in, size_t n, uint8_t **out, size_t *out_n) {
    if (!out || !out_n) return TRRE_E_ARG;
    *out = nullptr; *out_n = 0;
    // size first (cheap when length preserving: n), then fill
    size_t need = 0;
    const bool lp = p && (p->hdr.flags & PF_LENGTH_PRESERVING) && p->hdr.exec_kind != TRRE_EXEC_BACKTRACK && p->amode == AM_FIRST;
    // length preserving: exactly n; otherwise guess a little more than n (one pipelined pass when it fits,
    // the exact size comes back with TRRE_E_NOSPACE when it does not)
    size_t cap = lp ? n : n + n / 4 + 4096;
    uint8_t *buf = static_cast<uint8_t *>(std::malloc(cap ? cap : 1));
    if (!buf) return TRRE_E_ARG;
    int rc = trre_scan_into(p, device, in, n, buf, cap, &need);
    if (rc == TRRE_E_NOSPACE) {
        std::free(buf);
        cap = need;
        buf = static_cast<uint8_t *>(std::malloc(cap ? cap : 1));
        if (!buf) return TRRE_E_ARG;
        rc = trre_scan_into(p, device, in, n, buf, cap, &need);
    }
    if (rc != TRRE_OK) { std::free(buf); return rc; }
    *out = buf; *out_n = need;
    return TRRE_OK;
}

int trre_scan_multi(const trre_prog *p, const int *devices, int ndev, const uint8_t *in, size_t n, uint8_t **out,
                    size_t *out_n) {
    if (!p || !devices || ndev < 1 || !out || !out_n || (n && !in)) { g_last_error = "bad argument"; return TRRE_E_ARG; }
    *out = nullptr; *out_n = 0;
    if (ndev == 1) return trre_scan(p, devices[0], in, n, out, out_n);
    for (int a = 0; a < ndev; ++a)
        for (int b = a + 1; b < ndev; ++b)
            if (devices[a] == devices[b]) { g_last_error = "trre_scan_multi: the same device is listed twice"; return TRRE_E_ARG; }
    std::vector<size_t> cuts((size_t)ndev + 1);
    int rc = trre_shard_lines(in, n, ndev, cuts.data());
    if (rc != TRRE_OK) return rc;
    std::vector<uint8_t *> part((size_t)ndev, nullptr);
    std::vector<size_t> part_n((size_t)ndev, 0);
    std::vector<int> part_rc((size_t)ndev, TRRE_OK);
    std::vector<std::string> part_err((size_t)ndev);
    std::vector<std::thread> workers;
    for (int k = 0; k < ndev; ++k)
        workers.emplace_back([&, k]() {
            part_rc[k] = trre_scan(p, devices[k], in + cuts[k], cuts[k + 1] - cuts[k], &part[k], &part_n[k]);
            if (part_rc[k] != TRRE_OK) part_err[k] = trre_last_error();     // thread-local text
        });
    for (auto &w : workers) w.join();
    size_t total = 0;
    for (int k = 0; k < ndev; ++k) {
        if (part_rc[k] != TRRE_OK && rc == TRRE_OK) { rc = part_rc[k]; g_last_error = part_err[k]; }
        total += part_n[k];
    }
    if (rc == TRRE_OK) {
        uint8_t *buf = static_cast<uint8_t *>(std::malloc(total ? total : 1));
        if (!buf) rc = TRRE_E_ARG;
        else {
            size_t off = 0;
            for (int k = 0; k < ndev; ++k) { if (part_n[k]) std::memcpy(buf + off, part[k], part_n[k]); off += part_n[k]; }
            *out = buf; *out_n = total;
        }
    }
    for (int k = 0; k < ndev; ++k) std::free(part[k]);
    return rc;
}

int trre_shard_lines(const uint8_t *in, size_t n, int parts, size_t *cuts) {
    if (parts < 1 || !cuts || (n && !in)) return TRRE_E_ARG;
    cuts[0] = 0;
    for (int k = 1; k < parts; ++k) {
        size_t pos = (size_t)((unsigned __int128)n * k / parts);
        if (pos < cuts[k - 1]) pos = cuts[k - 1];
        // move forward to just after the next '\n' (a cut at 0 or right after a '\n' stays)
        if (pos > 0 && pos < n && in[pos - 1] != '\n') {
            const void *nl = std::memchr(in + pos, '\n', n - pos);
            pos = nl ? (size_t)(static_cast<const uint8_t *>(nl) - in) + 1 : n;
        }
        cuts[k] = pos;
    }
    cuts[parts] = n;
    return TRRE_OK;
}

}  // extern "C"
