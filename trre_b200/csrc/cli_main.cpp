// cli_main.cpp - the `trre` / `trre_dft` command lines, unchanged contract (trre.1; trre_nft.c:713-803,
// trre_dft.c:1199-1302): `trre [-d] [-m] [-a] EXPR [FILE]`, stdin when no FILE, same stderr texts and
// exit codes.  Built twice: -DTRRE_CLI_FLAVOR=0 -> bin/trre, =1 -> bin/trre_dft.
//
// Scan mode STREAMS, like the reference's getline loop (trre_nft.c:776-790, trre_dft.c:1273-1286), only in blocks:
// a reader thread fills pinned blocks from stdin / the file and cuts them at the last '\n' (the tail moves to the next
// block; a line longer than a block grows the block), the main thread hands each block to the GPU engine through the
// C ABI (trre_scan_into: H2D, kernels and D2H pipelined inside), a writer thread writes the finished blocks to stdout
// in order.  Memory is O(block), the first output appears after the first block, and input larger than RAM or an
// endless pipe work.  Inputs that end inside the first block take the one-shot path (no pinned ring to set up).
// -m / -a run on the GPU as well (trre_prog_set_mode: the reference's DFS, one record per thread) and stream the same way.
// GPU knobs come from the environment so the argv surface stays the reference's: TRRE_DEVICE (default 0),
// TRRE_GPUS (default 1: number of devices, starting at TRRE_DEVICE, the lines of each block are sharded over),
// TRRE_CLI_BLOCK_MB (default 64).
#include <unistd.h>

#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/trre_b200.h"

#ifndef TRRE_CLI_FLAVOR
#define TRRE_CLI_FLAVOR 0
#endif

static void dump_nft(const trre_prog *prog) {
    // -d: Graphviz of the flat NFT on stdout (the reference prints pointers as node ids, so the text
    // cannot be identical; shape and labels follow plot_nft, trre_nft.c:667-710)
    size_t n = 0;
    const unsigned char *blob = static_cast<const unsigned char *>(trre_prog_blob(prog, &n));
    struct Hdr { uint32_t w[24]; } h;
    std::memcpy(&h, blob, sizeof h);
    const uint32_t nst = h.w[6], off = h.w[19];
    std::printf("digraph G {\n\tsplines=true; rankdir=LR;\n");
    for (uint32_t s = 0; s < nst; ++s) {
        const unsigned char *e = blob + off + 12 * s;
        int32_t a, b;
        std::memcpy(&a, e + 4, 4);
        std::memcpy(&b, e + 8, 4);
        static const char *marks[] = {"+", "-", "S ", "Sn", "J ", ""};
        if (e[0] == 5) std::printf("\t\"s%u\" [peripheries=2, label=\"\"];\n", s);
        else if (e[0] <= 1) std::printf("\t\"s%u\" [label=\"%c%s\"];\n", s, e[1], marks[e[0]]);
        else std::printf("\t\"s%u\" [label=\"%s\"];\n", s, marks[e[0]]);
        if (a >= 0) std::printf("\t\"s%u\" -> \"s%d\";\n", s, a);
        if (b >= 0) std::printf("\t\"s%u\" -> \"s%d\" [label=\"*\"];\n", s, b);
    }
    std::printf("}\n");
}

namespace {

size_t read_fully(FILE *fp, unsigned char *dst, size_t want) {
    size_t got = 0;
    while (got < want) {
        size_t k = std::fread(dst + got, 1, want - got, fp);
        if (k == 0) break;
        got += k;
    }
    return got;
}

int fail_scan(int rc) {
    (void)rc;
    const char *msg = trre_last_error();
    std::fputs(msg, stderr);
    if (!*msg || msg[std::strlen(msg) - 1] != '\n') std::fputc('\n', stderr);
    return 1;
}

int scan_block(const trre_prog *prog, int device, int gpus, const unsigned char *in, size_t n, unsigned char **out,
               size_t *out_cap, size_t *out_n, bool pinned) {
    if (gpus > 1) {                                       // lines sharded over TRRE_GPUS devices of this box
        std::vector<int> devs;
        for (int k = 0; k < gpus; ++k) devs.push_back(device + k);
        uint8_t *o = nullptr;
        int rc = trre_scan_multi(prog, devs.data(), gpus, in, n, &o, out_n);
        if (rc != TRRE_OK) return rc;
        if (*out_n > *out_cap) {
            if (pinned) trre_host_free(*out); else std::free(*out);
            *out = static_cast<unsigned char *>(pinned ? trre_host_alloc(*out_n) : std::malloc(*out_n));
            *out_cap = *out ? *out_n : 0;
            if (!*out) { trre_free(o); return TRRE_E_ARG; }
        }
        std::memcpy(*out, o, *out_n);
        trre_free(o);
        return TRRE_OK;
    }
    int rc = trre_scan_into(prog, device, in, n, *out, *out_cap, out_n);
    if (rc == TRRE_E_NOSPACE) {                           // the program prints more than it reads: the exact size came back
        const size_t need = *out_n + *out_n / 8 + 4096;
        if (pinned) trre_host_free(*out); else std::free(*out);
        *out = static_cast<unsigned char *>(pinned ? trre_host_alloc(need) : std::malloc(need));
        *out_cap = *out ? need : 0;
        if (!*out) return TRRE_E_ARG;
        rc = trre_scan_into(prog, device, in, n, *out, *out_cap, out_n);
    }
    return rc;
}

struct Block {
    unsigned char *in = nullptr, *out = nullptr;
    size_t in_cap = 0, out_cap = 0, n = 0, out_n = 0;
    bool last = false;
};

template <class T>
class Channel {                       // tiny blocking queue
  public:
    void push(T v) { { std::lock_guard<std::mutex> l(mu_); q_.push_back(v); } cv_.notify_one(); }
    T pop() { std::unique_lock<std::mutex> l(mu_); cv_.wait(l, [&] { return !q_.empty(); }); T v = q_.front(); q_.pop_front(); return v; }
  private:
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<T> q_;
};

}  // namespace

int main(int argc, char **argv) {
    const int flavor = TRRE_CLI_FLAVOR;
    bool match = false, all = false, debug = false;
    int opt;
    while ((opt = getopt(argc, argv, "dma")) != -1) {
        switch (opt) {
            case 'd': debug = true; break;
            case 'm': match = true; break;
            case 'a':
                if (flavor == TRRE_FLAVOR_DFT) { std::fprintf(stderr, "Not supported yet\n"); return 1; }
                all = true;
                break;
            default:
                if (flavor == TRRE_FLAVOR_DFT) std::fprintf(stderr, "Usage: %s [-dma] expr [file]\n", argv[0]);
                else std::fprintf(stderr, "Usage: %s [-d] [-m] expr [file]\n", argv[0]);
                return 1;
        }
    }
    if (optind >= argc) { std::fprintf(stderr, "error: missing trre expression\n"); return 1; }

    char err[512];
    trre_prog *prog = nullptr;
    int rc = trre_compile(argv[optind], flavor, &prog, err, sizeof err);
    if (rc == TRRE_E_ASSERT) {
        std::fprintf(stderr, "%s: %s:285: parse: Assertion `operands != opd' failed.\n", argv[0],
                     flavor == TRRE_FLAVOR_DFT ? "trre_dft.c" : "trre_nft.c");
        std::abort();
    }
    if (rc != TRRE_OK) { std::fputs(err, stderr); return 1; }
    if (debug) dump_nft(prog);

    FILE *fp = stdin;
    if (optind == argc - 2) {
        fp = std::fopen(argv[optind + 1], "r");
        if (!fp) { std::fprintf(stderr, "error: can not open file %s\n", argv[optind + 1]); return 1; }
    }

    if (match || all) {                                    // trre_nft.c:635-648, 782-797: same pipeline, other DFS mode
        rc = trre_prog_set_mode(prog, match ? 1 : 0, all ? 1 : 0);
        if (rc != TRRE_OK) return fail_scan(rc);
    }

    int device = 0, gpus = 1;
    if (const char *e = std::getenv("TRRE_DEVICE")) device = std::atoi(e);
    if (const char *e = std::getenv("TRRE_GPUS")) gpus = std::atoi(e);
    size_t block = 64u << 20;
    if (const char *e = std::getenv("TRRE_CLI_BLOCK_MB")) { long v = std::atol(e); if (v >= 1 && v <= 4096) block = (size_t)v << 20; }

    // ---- small inputs: one shot, nothing pinned ------------------------------------------------------------------------
    const size_t first_cap = 1u << 20;
    std::vector<unsigned char> first(first_cap);
    const size_t first_n = read_fully(fp, first.data(), first_cap);
    if (first_n < first_cap) {
        if (fp != stdin) std::fclose(fp);
        uint8_t *out = nullptr;
        size_t out_n = 0;
        if (gpus > 1) {
            std::vector<int> devs;
            for (int k = 0; k < gpus; ++k) devs.push_back(device + k);
            rc = trre_scan_multi(prog, devs.data(), gpus, first.data(), first_n, &out, &out_n);
        } else {
            rc = trre_scan(prog, device, first.data(), first_n, &out, &out_n);
        }
        if (rc != TRRE_OK) return fail_scan(rc);
        if (out_n) std::fwrite(out, 1, out_n, stdout);
        trre_free(out);
        trre_prog_free(prog);
        return 0;
    }

    // ---- streaming: reader -> scan (this thread) -> writer, three blocks in flight ---------------------------------------
    constexpr int kBlocks = 3;
    Block blocks[kBlocks];
    for (Block &b : blocks) {
        b.in = static_cast<unsigned char *>(trre_host_alloc(block));
        b.out = static_cast<unsigned char *>(trre_host_alloc(block + 4096));
        b.in_cap = block; b.out_cap = block + 4096;
        if (!b.in || !b.out) return fail_scan(TRRE_E_CUDA);
    }
    Channel<Block *> free_q, filled_q, done_q;
    for (Block &b : blocks) free_q.push(&b);
    bool io_error = false;

    std::thread reader([&] {
        std::vector<unsigned char> tail(first.begin(), first.begin() + first_n);       // bytes not yet part of a block
        bool eof = false;
        while (true) {
            Block *b = free_q.pop();
            if (!b) break;
            // block = carried tail + fresh bytes, cut after the last '\n'; no '\n' at all: grow until one shows up / EOF
            size_t n = 0;
            if (tail.size() > b->in_cap) {
                trre_host_free(b->in);
                b->in_cap = tail.size() + block;
                b->in = static_cast<unsigned char *>(trre_host_alloc(b->in_cap));
                if (!b->in) { io_error = true; b->n = 0; b->last = true; filled_q.push(b); break; }
            }
            std::memcpy(b->in, tail.data(), tail.size());
            n = tail.size();
            tail.clear();
            for (;;) {
                if (!eof && n < b->in_cap) {
                    const size_t got = read_fully(fp, b->in + n, b->in_cap - n);
                    if (got < b->in_cap - n) eof = true;
                    n += got;
                }
                if (eof) break;
                const void *nl = memrchr(b->in, '\n', n);
                if (nl) {
                    const size_t cut = (size_t)(static_cast<const unsigned char *>(nl) - b->in) + 1;
                    tail.assign(b->in + cut, b->in + n);
                    n = cut;
                    break;
                }
                // one line fills the whole block: double it and keep reading
                unsigned char *bigger = static_cast<unsigned char *>(trre_host_alloc(b->in_cap * 2));
                if (!bigger) { io_error = true; eof = true; break; }
                std::memcpy(bigger, b->in, n);
                trre_host_free(b->in);
                b->in = bigger;
                b->in_cap *= 2;
            }
            b->n = n;
            b->last = eof;
            filled_q.push(b);
            if (eof) break;
        }
    });
    std::thread writer([&] {
        while (true) {
            Block *b = done_q.pop();
            if (!b) break;
            if (b->out_n && std::fwrite(b->out, 1, b->out_n, stdout) != b->out_n) io_error = true;
            const bool last = b->last;
            free_q.push(b);
            if (last) break;
        }
        std::fflush(stdout);
    });

    int scan_rc = TRRE_OK;
    while (true) {
        Block *b = filled_q.pop();
        if (b->n) {
            scan_rc = scan_block(prog, device, gpus, b->in, b->n, &b->out, &b->out_cap, &b->out_n, true);
        } else {
            b->out_n = 0;
        }
        if (scan_rc != TRRE_OK) {                          // what the earlier blocks printed stays printed, like the reference
            b->out_n = 0;
            b->last = true;
            done_q.push(b);
            free_q.push(nullptr);                          // release the reader if it waits for a block
            break;
        }
        const bool last = b->last;
        done_q.push(b);
        if (last) break;
    }
    writer.join();
    if (scan_rc != TRRE_OK) { reader.detach(); std::fflush(stdout); fail_scan(scan_rc); _exit(1); }
    reader.join();
    if (fp != stdin) std::fclose(fp);
    for (Block &b : blocks) { trre_host_free(b.in); trre_host_free(b.out); }
    trre_prog_free(prog);
    return io_error ? 1 : 0;
}
