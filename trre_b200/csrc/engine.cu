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
#include <vector>

#include "../../include/trre_b200.h"
#include "automaton.hpp"
#include "kernels.cuh"
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

struct DeviceState {
    int device = -1;
    uint8_t *d_blob = nullptr;
    DevProg P{};
    // workspace, grown on demand
    uint32_t *d_cnt = nullptr; size_t cap_chunks = 0;
    unsigned long long *d_gsum = nullptr; size_t cap_groups = 0;
    unsigned long long *d_total = nullptr;
    uint32_t *d_status = nullptr;
    // pinned mirrors
    unsigned long long *h_total = nullptr;
    uint32_t *h_status = nullptr;
    // DFS scratch
    int32_t *bt_s = nullptr; uint32_t *bt_i = nullptr; uint32_t *bt_o = nullptr; uint8_t *bt_obuf = nullptr;
    size_t bt_threads = 0; uint32_t bt_cap = 0, bt_ocap = 0;
    int sm_count = 148;
    // bookkeeping of the last enqueue
    int launches = 0;
    size_t last_n = 0;
    bool last_fast = false;        // a fast path was tried; total = n unless ST_IRREGULAR
};

}  // namespace

struct trre_prog {
    std::vector<uint8_t> blob;
    BlobHeader hdr;
    std::mutex mu;
    std::map<int, DeviceState *> dev;
};

namespace {

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
    d->device = device;
    const BlobHeader &h = p->hdr;
    CUDA_TRY(cudaMalloc(&d->d_blob, p->blob.size()));
    CUDA_TRY(cudaMemcpy(d->d_blob, p->blob.data(), p->blob.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&d->d_total, sizeof(unsigned long long)));
    CUDA_TRY(cudaMalloc(&d->d_status, sizeof(uint32_t)));
    CUDA_TRY(cudaMallocHost(&d->h_total, sizeof(unsigned long long)));
    CUDA_TRY(cudaMallocHost(&d->h_status, sizeof(uint32_t)));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    d->sm_count = prop.multiProcessorCount;
    d->P.classmap = d->d_blob + h.off_classmap;
    d->P.lut = d->d_blob + h.off_lut;
    d->P.trans = reinterpret_cast<const uint64_t *>(d->d_blob + h.off_trans);
    d->P.sinfo = reinterpret_cast<const uint64_t *>(d->d_blob + h.off_sinfo);
    d->P.pool = d->d_blob + h.off_pool;
    d->P.nft = reinterpret_cast<const NftFlat *>(d->d_blob + h.off_nft);
    d->P.n_classes = h.n_classes;
    d->P.start_info = h.start_info;
    d->P.flags = h.flags;
    d->P.nft_start = h.nft_start;
    CUDA_TRY(cudaFuncSetAttribute(k_bytemap, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024));
    p->dev[device] = d;
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
    if (backtrack && resident_threads > d->bt_threads) {
        // per-thread DFS stack (the reference allows 65536 entries, trre_nft.c:36/551) and output buffer
        uint32_t cap = 8192, ocap = 8192;
        if (const char *e = std::getenv("TRRE_BT_STACK")) { long v = std::atol(e); if (v >= 64 && v <= 65536) cap = (uint32_t)v; }
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
void launch_lines(bool backtrack, int grid, cudaStream_t st, const LineArgs &a) {
    if (backtrack) k_lines<MODE, true><<<grid, kLineThreads, 0, st>>>(a);
    else k_lines<MODE, false><<<grid, kLineThreads, 0, st>>>(a);
}

// Enqueue one whole-buffer scan.  Never synchronises.
int enqueue_scan(const trre_prog *p, DeviceState *d, const uint8_t *d_in, size_t n, uint8_t *d_out,
                 size_t out_cap, cudaStream_t st, bool count_only_general) {
    const BlobHeader &h = p->hdr;
    d->launches = 0;
    d->last_n = n;
    d->last_fast = false;
    CUDA_TRY(cudaMemsetAsync(d->d_status, 0, sizeof(uint32_t), st));
    CUDA_TRY(cudaMemsetAsync(d->d_total, 0, sizeof(unsigned long long), st));
    if (n == 0) return TRRE_OK;

    const bool backtrack = h.exec_kind == TRRE_EXEC_BACKTRACK;
    const size_t n_chunks = (n + kChunk - 1) / kChunk;
    const size_t n_groups = (n_chunks + kLineThreads - 1) / kLineThreads;
    // persistent-style grid: the DFS executor's scratch is per resident thread
    int grid_lines = (int)std::min<size_t>(n_groups, backtrack ? (size_t)d->sm_count * 1 : (size_t)d->sm_count * 64);
    if (grid_lines < 1) grid_lines = 1;
    int rc = ensure_workspace(d, n_chunks, n_groups, backtrack, (size_t)grid_lines * kLineThreads);
    if (rc != TRRE_OK) return rc;

    LineArgs a{};
    a.P = d->P; a.in = d_in; a.out = d_out; a.n = n; a.out_cap = out_cap;
    a.cnt = d->d_cnt; a.gsum = d->d_gsum; a.total = d->d_total; a.status = d->d_status;
    a.n_chunks = n_chunks; a.n_groups = n_groups; a.gate = 0;
    a.bt_s = d->bt_s; a.bt_i = d->bt_i; a.bt_o = d->bt_o; a.bt_obuf = d->bt_obuf;
    a.bt_cap = d->bt_cap; a.bt_ocap = d->bt_ocap;

    uint32_t gate = 0;
    const bool lp = (h.flags & PF_LENGTH_PRESERVING) && !backtrack;
    const bool aligned = ((reinterpret_cast<uintptr_t>(d_in) ^ reinterpret_cast<uintptr_t>(d_out)) & 15) == 0;
    if (!count_only_general && lp && out_cap >= n) {
        // fast paths assume a regular buffer (no NUL, ends in '\n'); they raise ST_IRREGULAR otherwise
        if (h.exec_kind == TRRE_EXEC_BYTEMAP && aligned) {
            size_t nvec = n / 16;
            size_t want = (nvec + (size_t)kBmThreads * kBmUnroll - 1) / ((size_t)kBmThreads * kBmUnroll);
            int grid = (int)std::min<size_t>(std::max<size_t>(want, 1), (size_t)d->sm_count * 4);
            k_bytemap<<<grid, kBmThreads, 32 * 1024, st>>>(d_in, d_out, n, d->P.lut, d->d_status);
        } else {
            launch_lines<MODE_INPLACE>(false, grid_lines, st, a);
        }
        d->launches += 1;
        d->last_fast = true;
        gate = ST_IRREGULAR;
    }
    // general path: count -> scan -> write (gated behind ST_IRREGULAR after a fast path)
    a.gate = gate;
    launch_lines<MODE_COUNT>(backtrack, grid_lines, st, a);
    k_scan_groups<<<1, kScanThreads, 0, st>>>(d->d_gsum, n_groups, d->d_total, d->d_status, gate);
    d->launches += 2;
    if (!count_only_general) {
        launch_lines<MODE_WRITE>(backtrack, grid_lines, st, a);
        d->launches += 1;
    }
    CUDA_TRY(cudaGetLastError());
    return TRRE_OK;
}

int status_to_rc(uint32_t s) {
    if (s & (ST_DIVERGE | ST_STACK)) { g_last_error = "error: stack max capacity reached\n"; return TRRE_E_STACK; }
    if (s & ST_OBUF) { g_last_error = "error: per-attempt output exceeds the DFS executor's buffer\n"; return TRRE_E_UNSUPPORTED; }
    if (s & ST_NOSPACE) { g_last_error = "output buffer too small"; return TRRE_E_NOSPACE; }
    return TRRE_OK;
}

int collect(DeviceState *d, cudaStream_t st, size_t *out_n) {
    CUDA_TRY(cudaMemcpyAsync(d->h_status, d->d_status, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(d->h_total, d->d_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    uint32_t s = *d->h_status;
    size_t total = (d->last_fast && !(s & ST_IRREGULAR)) ? d->last_n : (size_t)*d->h_total;
    if (out_n) *out_n = total;
    return status_to_rc(s);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
extern "C" {

const char *trre_version(void) { return "trre_b200 0.1 (sm_100a)"; }
const char *trre_last_error(void) { return g_last_error.c_str(); }
void trre_free(void *p) { std::free(p); }

int trre_compile(const char *expr, int flavor, trre_prog **out, char *err, size_t errcap) {
    if (!expr || !out) { set_err(err, errcap, "null argument"); return TRRE_E_ARG; }
    *out = nullptr;
    try {
        std::vector<uint8_t> blob = compile_to_blob(expr, flavor);
        trre_prog *p = new trre_prog();
        p->blob = std::move(blob);
        std::memcpy(&p->hdr, p->blob.data(), sizeof(BlobHeader));
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
    *out = p;
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
        delete d;
    }
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
    return collect(d, st, out_n);
}

int trre_scan_status(const trre_prog *p, int device, size_t *out_n, void *stream) {
    if (!p) return TRRE_E_ARG;
    DeviceGuard guard(device);
    DeviceState *d = nullptr;
    int rc = get_state(p, device, &d);
    if (rc != TRRE_OK) return rc;
    return collect(d, static_cast<cudaStream_t>(stream), out_n);
}

int trre_scan_launches(const trre_prog *p, int device) {
    if (!p) return -1;
    DeviceState *d = nullptr;
    if (get_state(p, device, &d) != TRRE_OK) return -1;
    return d->launches;
}

int trre_scan_into(const trre_prog *p, int device, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap,
                   size_t *out_n) {
    if (!p || (n && !in)) { g_last_error = "null argument"; return TRRE_E_ARG; }
    if (out_n) *out_n = 0;
    DeviceGuard guard(device);
    if (!guard.ok) { g_last_error = "no CUDA device: the trre_b200 scan path has no CPU fallback"; return TRRE_E_CUDA; }
    DeviceState *d = nullptr;
    int rc = get_state(p, device, &d);
    if (rc != TRRE_OK) return rc;
    if (n == 0) return TRRE_OK;
    cudaStream_t st = nullptr;
    uint8_t *d_in = nullptr, *d_out = nullptr;
    CUDA_TRY(cudaMalloc(&d_in, n));
    struct Free { uint8_t *&a, *&b; ~Free() { if (a) cudaFree(a); if (b) cudaFree(b); } } fr{d_in, d_out};
    CUDA_TRY(cudaMemcpyAsync(d_in, in, n, cudaMemcpyHostToDevice, st));
    const BlobHeader &h = p->hdr;
    size_t total = 0;
    const bool lp = (h.flags & PF_LENGTH_PRESERVING) && h.exec_kind != TRRE_EXEC_BACKTRACK;
    bool done = false;
    if (lp) {
        CUDA_TRY(cudaMalloc(&d_out, n));
        rc = enqueue_scan(p, d, d_in, n, d_out, n, st, false);
        if (rc != TRRE_OK) return rc;
        rc = collect(d, st, &total);
        if (rc == TRRE_OK) done = true;
        else if (rc != TRRE_E_NOSPACE) return rc;
        else { cudaFree(d_out); d_out = nullptr; }           // irregular input that grew: size it exactly
    }
    if (!done) {
        rc = enqueue_scan(p, d, d_in, n, nullptr, 0, st, true);   // count + scan only
        if (rc != TRRE_OK) return rc;
        rc = collect(d, st, &total);
        if (rc != TRRE_OK && rc != TRRE_E_NOSPACE) return rc;
        CUDA_TRY(cudaMalloc(&d_out, total ? total : 1));
        rc = enqueue_scan(p, d, d_in, n, d_out, total, st, false);
        if (rc != TRRE_OK) return rc;
        rc = collect(d, st, &total);
        if (rc != TRRE_OK) return rc;
    }
    if (out_n) *out_n = total;
    if (total > out_cap) { g_last_error = "output buffer too small"; return TRRE_E_NOSPACE; }
    CUDA_TRY(cudaMemcpyAsync(out, d_out, total, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return TRRE_OK;
}

int trre_scan(const trre_prog *p, int device, const uint8_t *in, size_t n, uint8_t **out, size_t *out_n) {
    if (!out || !out_n) return TRRE_E_ARG;
    *out = nullptr; *out_n = 0;
    // size first (cheap when length preserving: n), then fill
    size_t need = 0;
    const bool lp = p && (p->hdr.flags & PF_LENGTH_PRESERVING) && p->hdr.exec_kind != TRRE_EXEC_BACKTRACK;
    size_t cap = lp ? n : 0;
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
