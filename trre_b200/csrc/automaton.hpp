// automaton.hpp - host side: flat NFT -> deterministic scan automaton (DSA) -> device blob.
//
// The DSA is the one table format both flavours are executed from on the GPU:
//
//   state s, byte class k -> entry { next | DEAD, mark, out-string }       (trans[s*ncls + k])
//   state s               -> info  { HAS_FINAL, DECIDED, HAS_DIVERGE, fin-string }
//
// One scan ATTEMPT at input position p (what infer_dft / infer_backtrack compute per call):
//
//     s = start; i = 0; m = HAS_FINAL(start) ? 0 : none; acc = ""
//     if DECIDED(start): emit fin(start); return 0
//     loop: c = in[p+i]; stop at end of payload
//           e = trans[s][class(c)]; stop if DEAD
//           acc += out(e); s = next(e); i += 1; if mark(e): m = i
//           if DECIDED(s): emit acc + fin(s); return m
//     if HAS_FINAL(s): emit acc + fin(s); return m          // winner known once lookahead ended
//     return FAIL
//
// DFT flavour (trre_dft semantics, leftmost-SHORTEST): a state is DECIDED as soon as it is final,
// exactly infer_dft's "stop at the first final state" (trre_dft.c:1120-1125).
// NFT flavour (trre semantics, leftmost-FIRST-by-priority): states are priority-ordered item lists;
// a FINAL item cuts everything of lower priority, survives steps, and wins once every item of higher
// priority has died; `mark` records where it was born (= bytes consumed by the match).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "pattern.hpp"

namespace trre {

enum : uint32_t {
    DSA_DEAD        = 0x00FFFFFFu,   // `next` value of a dead entry
    DSA_E_MARK      = 1u << 24,      // entry: a FINAL item is born on this transition
    DSA_E_DECIDED   = 1u << 25,      // entry: target state is DECIDED (attempt ends there)
    DSA_E_HASFINAL  = 1u << 26,      // entry: target state has a pending FINAL item (wins if the walk stops there)
    DSA_E_DIVERGE   = 1u << 27,      // entry: target state has a pending DIVERGE item
    DSA_S_HAS_FINAL = 1u,            // state: a FINAL item is pending in the list
    DSA_S_DECIDED   = 2u,            // state: nothing of higher priority is alive any more
    DSA_S_DIVERGE   = 4u             // state: an epsilon cycle is pending (reference: stack cap error)
};

// blob header flags
enum : uint32_t {
    PF_LENGTH_PRESERVING = 1u,
    PF_CAN_MATCH_EMPTY   = 2u,
    PF_ACYCLIC           = 4u,
    PF_NUL_IN_OUTPUT     = 8u,       // some PROD byte is 0: NFT flavour truncates at it (fputs)
    PF_HAS_DIVERGE       = 16u,
    PF_MATCHABLE         = 64u,
    PF_BYTEMAP_VAR       = 128u,
    PF_GRAM              = 512u,     // gram tables present (k_gram)
    PF_MATCHABLE_VAR     = 256u,     // k_mv_* apply: like PF_MATCHABLE but matches may print more or less than they consume     // every attempt is decided by one byte; that byte prints 0..255 bytes (k_bmv_*)      // k_match applies: length preserving, acyclic, attempts <= 16 bytes, no empty match
    PF_OUT_LATE          = 32u       // every transition into a non-DECIDED state emits nothing: a match's output
                                     // is out(last transition) + fin(last state)
};

struct BlobHeader {
    uint32_t magic;            // 'TRB2'
    uint32_t abi;
    uint32_t total_bytes;
    uint32_t flavor;
    uint32_t exec_kind;        // TRRE_EXEC_*
    uint32_t flags;            // PF_*
    uint32_t nft_states;
    int32_t  nft_start;
    uint32_t dsa_states;
    uint32_t n_classes;
    uint32_t max_depth;        // 0xffffffff when cyclic
    uint32_t start_info;       // DSA_S_* of state 0
    uint32_t off_classmap;     // u8[256]
    uint32_t off_lut;          // u8[256]   (TRRE_EXEC_BYTEMAP)
    uint32_t off_startrow;     // u8[256]: 1 if an attempt starting with this byte is not immediately dead
    uint32_t off_trans;        // u64[dsa_states * n_classes]
    uint32_t off_sinfo;        // u64[dsa_states]
    uint32_t off_pool;         // bytes
    uint32_t pool_bytes;
    uint32_t off_nft;          // NftFlat[nft_states]
    uint32_t off_pairlut;      // u16[65536] (TRRE_EXEC_BYTEMAP): LUT applied to two bytes at once, 0 if absent
    uint32_t off_row0;         // u32[256]: low word of trans[start][class(byte)] by raw byte (k_match), 0 if absent
    uint32_t off_ctrans;       // u32[dsa_states * n_classes]: low words of trans (k_match), 0 if absent
    uint32_t off_vent;         // (PF_BYTEMAP_VAR) u32[256]: len | first byte << 8 | offset into vpool << 16, by raw byte
    uint32_t off_vpairlen;     // (PF_BYTEMAP_VAR) u16[65536]: len[c0] + len[c1], bank-swizzled like the pair LUT
    uint32_t off_vpool;        // (PF_BYTEMAP_VAR) strings longer than one byte
    uint32_t vpool_bytes;
    uint32_t off_eout;         // u32[dsa_states * n_classes]: (PF_OUT_LATE) offset(24) | length(8) of out(edge)+fin(target) in the
                               // pool for edges into DECIDED states; u32[dsa_states] for pending FINALs follow. 0 if absent
    // (TRRE_EXEC_BYTEMAP) table-free form of the LUT: it is the identity except on swar_n <= 4 byte ranges, each
    // inside one half of the byte space, where it adds a constant mod 256 (cfg2: a-z -32; Caesar: a-y +1, z -25).
    // 0 = no such form (k_bytemap's pair table serves).  swar[k] = lo | hi << 8 | delta << 16
    uint32_t swar_n;
    uint32_t swar[4];
    // (PF_GRAM) every attempt tabulated by the byte classes of its first gram_depth bytes (k_gram, gram_kernel.cuh)
    uint32_t gram_depth;       // D = max_depth, 1..6
    uint32_t gram_entries;     // n_classes ^ D; index = sum cls(byte i) * n_classes^i
    uint32_t gram_len;         // > 0: every match consumes exactly this many bytes
    uint32_t gram_flags;       // GRAM_*
    uint32_t off_gbits;        // u32[(gram_entries + 31) / 32]: the attempt matches
    uint32_t off_gout;         // u32[gram_entries]: GRAM_DIRECT: the printed bytes; else offset into gpool(20) |
                               // printed length(8) << 20 | (consumed - 1) << 28
    uint32_t off_gpool;
    uint32_t gpool_bytes;
    // (lit_words > 0) the matches are lit_words <= 4 WORDS of lit_len = 2..4 bytes; position k of word w accepts one byte
    // range and prints one byte: lit[w * lit_len + k] = lo | hi << 8 | printed << 16; no suffix of a word can be the
    // prefix of a word, so every candidate is a real match and the scan needs no tables and no line structure at all:
    // k_lit compares in registers (lit_kernel.cuh).  cat:dog is one word, (cat:dog|dog:cat) two.
    uint32_t lit_words;
    uint32_t lit_len;
    uint32_t lit[16];
    // (PF_BYTEMAP_VAR) > 0: RUN form - bits 28..31 of a vent entry are the byte's run class (1..15, 0 = none): a match that
    // starts with such a byte swallows every following byte of the same class ([0-9]+:#), i.e. a byte whose predecessor has
    // its run class prints nothing; the pool offset of an entry is then 12 bits
    uint32_t vrun_classes;
    // (lit_words > 0 on a PF_MATCHABLE_VAR program) the words print strings of any length instead: gpool offset(20) |
    // printed length(8) << 20 per word (k_litv_count / k_litv_write: k_lit's comparison, k_bmv's placement)
    uint32_t lit_out[4];
    // (gram_flags & 2) every match prints its own bytes moved by this constant mod 256 (word:WORD dictionaries: 0xE0):
    // k_gram then needs no output table
    uint32_t gram_delta;
    // (vrun_classes > 0) vrun_swar_n = vrun_classes when every run class is ONE byte range (inside one half of the byte
    // space) whose bytes print strings of one length: vrun_swar[k] = lo | hi << 8 | length << 16 for class k + 1 - the count
    // pass then needs no per-byte entries (k_bmv_count<RUNS>); 0 otherwise
    uint32_t vrun_swar_n;
    uint32_t vrun_swar[4];
};

struct NftFlat {               // 12 bytes, device layout of trre_nft.c:334-340
    uint8_t op, val, pad0, pad1;
    int32_t a, b;
};

// entry: lo = next | flags, hi = out_off(24) | out_len(8); out_len == 255 -> u32 length at pool[out_off]
// sinfo: lo = DSA_S_* flags,  hi = fin_off(24) | fin_len(8), same escape
static inline uint64_t pack64(uint32_t lo, uint32_t hi) { return (uint64_t)lo | ((uint64_t)hi << 32); }

struct CompileLimits {
    uint32_t max_states     = 1u << 16;
    uint64_t max_build_bytes = 192ull << 20;   // host memory of the subset construction
    uint32_t max_residual   = 4096;        // longest delayed output; beyond = not subsequential (twins)
    uint64_t max_table_bytes = 128ull << 20;
    uint64_t max_work       = 24ull * 1000 * 1000;   // closure visits of the subset construction
    int      max_recursion  = 20000;
};

// Builds the whole program blob for `expr`.  Throws PatternError.
std::vector<uint8_t> compile_to_blob(const char *expr, int flavor, const CompileLimits &lim = CompileLimits());

// ---- lazy construction (TRRE_EXEC_LAZY) -----------------------------------------------------------------------------
// trre_dft determinizes lazily: a dstate's row is only built for the bytes the input actually shows it
// (trre_dft.c:1135-1175), so it answers transducers whose full subset construction is astronomically large.  When the
// eager construction above runs into its limits, the same expansion step is kept alive on the host and driven by the
// GPU: the per-line kernel walks rows indexed by the raw byte; a state whose row is missing is reported back
// (ST_MISS + request list), the host builds that row (plus a few behind it, breadth first), uploads it and the pass is
// repeated.  The host only ever COMPILES - which rows exist never changes what a row says - the scan stays on the GPU.
constexpr uint32_t LAZY_UNBUILT = 0x00FFFFFEu;     // `next` value of every entry of a row that has not been built
constexpr uint32_t LAZY_POISON  = 0x00FFFFFDu;     // `next` value of an entry whose target's closure does not terminate:
                                                    // trre_dft recurses for ever there (nft_step_, trre_dft.c:874-907)
constexpr uint32_t DSA_S_ROW_BUILT = 8u;            // sinfo flag

class LazyDsa {
  public:
    LazyDsa(const Nft &g, int flavor);
    ~LazyDsa();
    LazyDsa(const LazyDsa &) = delete;
    LazyDsa &operator=(const LazyDsa &) = delete;
    // Build the rows of the requested states and up to `extra` more rows breadth first behind them.
    // Throws PatternError{TRRE_E_UNSUPPORTED} when the limits (states, memory, residual output) are reached.
    void build(const uint32_t *req, size_t n_req, uint32_t extra);
    uint32_t n_states() const;
    const std::vector<uint64_t> &rows() const;     // [n_states * 256] by raw byte: lo = next | DSA_E_*, hi = out ref
    const std::vector<uint64_t> &sinfo() const;    // [n_states]: lo = DSA_S_*, hi = fin ref
    const std::vector<uint8_t> &pool() const;
    const std::vector<uint32_t> &log() const;      // states whose row exists, in build order
    uint32_t start_info() const;

  private:
    struct Impl;
    Impl *im_;
};

// Validates a blob received from elsewhere (e.g. broadcast from rank 0).
bool blob_is_valid(const void *blob, size_t nbytes, std::string *why);

}  // namespace trre
