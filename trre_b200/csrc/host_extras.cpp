// host_extras.cpp - host-only parts of the C ABI that are NOT the scan path:
//   trre_match_host  : `-m`, `-ma` and scan-mode `-a` on the HOST (trre_nft.c:635-642, 791-797; the DFT binary prints a
//                      bare newline per record, trre_dft.c:1287-1293).  The CLI runs these modes on the GPU
//                      (trre_prog_set_mode, the DFS executor); this restatement is what the tests cross-check that path
//                      with when no reference binary is at hand.
//   trre_synth_text  : the deterministic synthetic corpus of BASELINE.md s3 / SURVEY s8d.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/trre_b200.h"
#include "automaton.hpp"

using namespace trre;

namespace {

struct Saved { int32_t s; uint32_t i, o; };

struct HostDfs {
    const NftFlat *nft;
    int32_t start;
    std::vector<Saved> stack;
    std::string obuf;
    std::string *sink;
    bool overflow = false;

    // fputs(): up to the first NUL
    void print_output(size_t o) {
        size_t len = 0;
        while (len < o && obuf[len]) ++len;
        sink->append(obuf, 0, len);
    }

    long run(const uint8_t *p, size_t len, bool match_mode, bool all) {
        uint32_t i = 0, o = 0;
        int32_t s = start;
        stack.clear();
        while (!stack.empty() || s >= 0) {
            if (s < 0) {
                Saved sv = stack.back();
                stack.pop_back();
                s = sv.s; i = sv.i; o = sv.o;
                if (s < 0) continue;
            }
            const NftFlat &n = nft[s];
            switch (n.op) {
                case N_CONS:
                    if (i < len && p[i] == n.val) { ++i; s = n.a; } else s = -1;
                    break;
                case N_PROD:
                    if (obuf.size() <= o) obuf.resize(o * 2 + 16);
                    obuf[o++] = (char)n.val;
                    s = n.a;
                    break;
                case N_SPLIT:
                case N_SPLITNG:
                    if (stack.size() >= 65536) { overflow = true; return -1; }   // trre_nft.c:551-554
                    stack.push_back(Saved{n.op == N_SPLIT ? n.a : n.b, i, o});
                    s = n.op == N_SPLIT ? n.b : n.a;
                    break;
                case N_JOIN:
                    s = n.a;
                    break;
                default:
                    if (match_mode) {
                        if (i == len) {
                            print_output(o);
                            sink->push_back('\n');
                            if (!all) return (long)i;
                        }
                    } else {
                        print_output(o);
                        if (!all) return (long)i;
                    }
                    s = -1;
                    break;
            }
        }
        return -1;
    }
};

uint64_t splitmix(uint64_t &x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

}  // namespace

extern "C" {

int trre_match_host(const trre_prog *p, int scan, int all, const uint8_t *in, size_t n, uint8_t **out,
                    size_t *out_n, char *err, size_t errcap) {
    if (!p || !out || !out_n) return TRRE_E_ARG;
    size_t nbytes = 0;
    const uint8_t *blob = static_cast<const uint8_t *>(trre_prog_blob(p, &nbytes));
    BlobHeader h;
    std::memcpy(&h, blob, sizeof h);
    std::string sink;
    int rc = TRRE_OK;
    HostDfs dfs;
    dfs.nft = reinterpret_cast<const NftFlat *>(blob + h.off_nft);
    dfs.start = h.nft_start;
    dfs.sink = &sink;
    size_t pos = 0;
    while (pos < n) {
        const uint8_t *nl = static_cast<const uint8_t *>(std::memchr(in + pos, '\n', n - pos));
        size_t reclen = nl ? (size_t)(nl - (in + pos)) + 1 : n - pos;
        size_t payload = reclen - 1;                          // line[read-1] = '\0'
        const uint8_t *z = static_cast<const uint8_t *>(std::memchr(in + pos, 0, payload));
        size_t len = z ? (size_t)(z - (in + pos)) : payload;
        const uint8_t *line = in + pos;
        pos += reclen;
        if (h.flavor == TRRE_FLAVOR_DFT) { sink.push_back('\n'); continue; }
        if (!scan) {
            dfs.run(line, len, true, all != 0);
        } else {
            size_t ch = 0;
            while (ch < len) {
                long k = dfs.run(line + ch, len - ch, false, all != 0);
                if (dfs.overflow) break;
                if (k > 0) ch += (size_t)k; else sink.push_back((char)line[ch++]);
            }
            if (!dfs.overflow) dfs.run(line + len, 0, false, all != 0);
            sink.push_back('\n');
        }
        if (dfs.overflow) {
            if (err && errcap) std::snprintf(err, errcap, "error: stack max capacity reached\n");
            rc = TRRE_E_STACK;
            if (scan) sink.pop_back();
            break;
        }
    }
    uint8_t *buf = static_cast<uint8_t *>(std::malloc(sink.size() ? sink.size() : 1));
    if (!buf) return TRRE_E_ARG;
    std::memcpy(buf, sink.data(), sink.size());
    *out = buf;
    *out_n = sink.size();
    return rc;
}

int trre_synth_text(uint64_t seed, uint8_t *buf, size_t n, const char *words, double p_hit, int lowercase_only) {
    if (!buf && n) return TRRE_E_ARG;
    std::vector<std::string> hits;
    if (words) {
        const char *w = words;
        while (*w) {
            const char *e = std::strchr(w, '\n');
            size_t len = e ? (size_t)(e - w) : std::strlen(w);
            if (len) hits.emplace_back(w, len);
            w += len + (e ? 1 : 0);
        }
    }
    static const char mixed[] = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789.,;";
    static const char lower[] = "abcdefghijklmnopqrstuvwxyz";
    const char *alpha = lowercase_only ? lower : mixed;
    const uint64_t nalpha = lowercase_only ? 26 : 65;
    const uint64_t maxtok = lowercase_only ? 8 : 10;
    const uint64_t hit_thresh = (p_hit <= 0 || hits.empty()) ? 0 : (uint64_t)(p_hit * 4294967296.0);
    uint64_t st = seed;
    size_t pos = 0, line_len = 0;
    size_t target = 40 + splitmix(st) % 81;
    while (pos < n) {
        uint64_t r = splitmix(st);
        if ((r & 0xFFFFFFFFull) < hit_thresh) {
            const std::string &w = hits[(r >> 32) % hits.size()];
            for (size_t k = 0; k < w.size() && pos < n; ++k, ++line_len) buf[pos++] = (uint8_t)w[k];
        } else {
            uint64_t len = 1 + (r >> 32) % maxtok;
            uint64_t bits = splitmix(st), bits2 = splitmix(st);
            for (uint64_t k = 0; k < len && pos < n; ++k, ++line_len) {
                uint64_t pick = (k < 5 ? (bits >> (k * 12)) : (bits2 >> ((k - 5) * 12))) & 0xFFF;
                buf[pos++] = (uint8_t)alpha[pick % nalpha];
            }
        }
        if (pos >= n) break;
        if (line_len >= target) {
            buf[pos++] = '\n';
            line_len = 0;
            target = 40 + splitmix(st) % 81;
        } else {
            buf[pos++] = ' ';
            ++line_len;
        }
    }
    if (n) buf[n - 1] = '\n';
    return TRRE_OK;
}

}  // extern "C"
