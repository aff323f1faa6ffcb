// pattern.hpp - host side: trre expression -> syntax tree -> flat NFT.
//
// Same language, precedence and quirks as the reference parser/compiler (trre_nft.c:10-288,
// 375-511; trre_dft.c:12-288, 376-523) because the tables - and therefore stdout - depend on them,
// but re-entrant (no globals), index-addressed and exception-reporting instead of exit().
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace trre {

struct PatternError {
    int         code;   // TRRE_E_* from include/trre_b200.h
    std::string text;   // what the reference prints on stderr
};

// ---- syntax tree ----------------------------------------------------------------------------
// kinds follow trre_nft.c:24-29: 'c' literal byte, 'e' epsilon placeholder, '.', '|', ':', '-',
// '*', '+', '?', 'I' (counted repetition; its right child carries lo in `kind`, hi in `val`).
struct SynNode {
    uint8_t kind;
    uint8_t val;
    int32_t lhs = -1;
    int32_t rhs = -1;
};

struct SynTree {
    std::vector<SynNode> nodes;
    int32_t root = -1;
    int32_t add(uint8_t kind, uint8_t val, int32_t l = -1, int32_t r = -1) {
        nodes.push_back(SynNode{kind, val, l, r});
        return (int32_t)nodes.size() - 1;
    }
};

SynTree parse_expression(const char *expr);          // throws PatternError

// ---- flat NFT ---------------------------------------------------------------------------------
enum NftOp : uint8_t { N_PROD = 0, N_CONS = 1, N_SPLIT = 2, N_SPLITNG = 3, N_JOIN = 4, N_FINAL = 5 };

struct NftState {
    uint8_t op;
    uint8_t val;
    int32_t a = -1;      // nexta
    int32_t b = -1;      // nextb (SPLIT: preferred; SPLITNG: fallback)
};

struct Nft {
    std::vector<NftState> st;
    int32_t start = -1;  // -1: the empty range `[z-a]` as whole pattern: nothing ever matches
    bool    has_init_join = false;   // DFT flavour prepends a JOIN (trre_dft.c:516-523)
};

// with_init_join=false: trre_nft.c:506-511; true: trre_dft.c:516-523
Nft build_nft(SynTree &tree, bool with_init_join);  // throws PatternError

}  // namespace trre
