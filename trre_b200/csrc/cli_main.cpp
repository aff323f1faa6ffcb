// cli_main.cpp - the `trre` / `trre_dft` command lines, unchanged contract (trre.1; trre_nft.c:713-803,
// trre_dft.c:1199-1302): `trre [-d] [-m] [-a] EXPR [FILE]`, stdin when no FILE, same stderr texts and
// exit codes.  Scan mode hands the WHOLE input to the GPU engine through the C ABI; -m / -a stay on
// the host (trre_match_host).  Built twice: -DTRRE_CLI_FLAVOR=0 -> bin/trre, =1 -> bin/trre_dft.
// GPU knobs come from the environment so the argv surface stays the reference's: TRRE_DEVICE (default 0),
// TRRE_GPUS (default 1: number of devices, starting at TRRE_DEVICE, the lines are sharded over).
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
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
    std::vector<unsigned char> in;
    {
        unsigned char buf[1 << 16];
        size_t k;
        while ((k = std::fread(buf, 1, sizeof buf, fp)) > 0) in.insert(in.end(), buf, buf + k);
    }
    if (fp != stdin) std::fclose(fp);

    uint8_t *out = nullptr;
    size_t out_n = 0;
    if (!match && !all) {
        int device = 0, gpus = 1;
        if (const char *e = std::getenv("TRRE_DEVICE")) device = std::atoi(e);
        if (const char *e = std::getenv("TRRE_GPUS")) gpus = std::atoi(e);
        if (gpus > 1) {                                           // lines sharded over TRRE_GPUS devices of this box
            std::vector<int> devs;
            for (int k = 0; k < gpus; ++k) devs.push_back(device + k);
            rc = trre_scan_multi(prog, devs.data(), gpus, in.data(), in.size(), &out, &out_n);
        } else {
            rc = trre_scan(prog, device, in.data(), in.size(), &out, &out_n);
        }
        if (rc != TRRE_OK) {
            const char *msg = trre_last_error();
            std::fputs(msg, stderr);
            if (!*msg || msg[std::strlen(msg) - 1] != '\n') std::fputc('\n', stderr);
            return 1;
        }
    } else {
        rc = trre_match_host(prog, match ? 0 : 1, all ? 1 : 0, in.data(), in.size(), &out, &out_n, err, sizeof err);
        if (out_n) std::fwrite(out, 1, out_n, stdout);
        if (rc != TRRE_OK) { std::fflush(stdout); std::fputs(err, stderr); return 1; }
        trre_free(out);
        trre_prog_free(prog);
        return 0;
    }
    if (out_n) std::fwrite(out, 1, out_n, stdout);
    trre_free(out);
    trre_prog_free(prog);
    return 0;
}
