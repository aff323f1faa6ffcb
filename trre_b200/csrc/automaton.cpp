// automaton.cpp - flat NFT -> deterministic scan automaton -> blob (host, cold path).
//
// Two subset constructions over the same flat NFT, one per reference binary:
//
//  * DFT flavour: what trre_dft.c builds lazily (nft_step/_step_ 874-923, truncate_lcp 988-1010,
//    finality probe 1164-1174), built EAGERLY here over every reachable state x byte, including the
//    reference's observable quirks: the output string is shared down the preferred branch and only
//    copied afterwards (874-891), and a CONS(0) state counts as accepting in the probe (893-897).
//  * NFT flavour: trre's backtracking DFS (trre_nft.c:593-657) turned into a priority-ordered
//    subset construction (leftmost-first, like a Pike VM whose thread list order is the DFS order),
//    with delayed output per item and longest-common-prefix extraction.  Not every transducer is
//    subsequential; when the construction does not converge within the limits the program falls
//    back to TRRE_EXEC_BACKTRACK (the DFS itself, on the GPU).
#include "automaton.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <unordered_map>

#include "../../include/trre_b200.h"

namespace trre {
namespace {

constexpr int32_t IT_FINAL = -2;      // pseudo item: an accepting path is pending
constexpr int32_t IT_DIVERGE = -3;
constexpr uint32_t GRAM_DIRECT_FLAG = 1u;   // = GRAM_DIRECT in gram_kernel.cuh    // pseudo item: the reference's DFS never returns from here

struct Item {
    int32_t q;
    std::string w;                    // output produced on this path and not yet emitted
};
using ItemList = std::vector<Item>;

struct RawEdge {
    uint8_t  byte;
    uint32_t next;
    bool     mark;
    uint32_t out_id;
};
struct RawState {
    ItemList items;
    uint32_t flags = 0;
    uint32_t fin_id = 0;
    std::vector<RawEdge> edges;
};

struct NotSubsequential { const char *why; };

class StringPool {
  public:
    StringPool() { intern(std::string()); }
    uint32_t intern(const std::string &s) {
        auto it = ids_.find(s);
        if (it != ids_.end()) return it->second;
        uint32_t id = (uint32_t)strs_.size();
        strs_.push_back(s);
        ids_.emplace(s, id);
        return id;
    }
    const std::string &get(uint32_t id) const { return strs_[id]; }
    size_t count() const { return strs_.size(); }

  private:
    std::vector<std::string> strs_;
    std::unordered_map<std::string, uint32_t> ids_;
};

std::string key_of(const ItemList &l) {
    std::string k;
    k.reserve(l.size() * 8);
    for (const Item &it : l) {
        uint32_t q = (uint32_t)it.q, n = (uint32_t)it.w.size();
        k.append(reinterpret_cast<const char *>(&q), 4);
        k.append(reinterpret_cast<const char *>(&n), 4);
        k.append(it.w);
    }
    return k;
}

// moves the longest common prefix of all residuals into `prefix` (trre_dft.c:988-1010)
void strip_lcp(ItemList &l, std::string &prefix) {
    size_t first = 0;
    while (first < l.size() && l[first].q == IT_DIVERGE) ++first;   // its output is never printed
    if (first == l.size()) return;
    size_t n = l[first].w.size();
    for (const Item &it : l) {
        if (it.q == IT_DIVERGE) continue;
        size_t k = 0, lim = std::min(n, it.w.size());
        while (k < lim && it.w[k] == l[first].w[k]) ++k;
        n = k;
        if (!n) return;
    }
    prefix.assign(l[first].w, 0, n);
    for (Item &it : l)
        if (it.q != IT_DIVERGE) it.w.erase(0, n);
}

class Determinizer {
  public:
    Determinizer(const Nft &g, int flavor, const CompileLimits &lim, StringPool &pool)
        : g_(g), flavor_(flavor), lim_(lim), pool_(pool), stamp_(g.st.size(), 0), gray_(g.st.size(), 0) {}

    std::vector<RawState> states;

    void run() {
        init();
        for (uint32_t s = 0; s < states.size(); ++s) expand(s);
    }

    // one state's row: its transitions on every byte, creating the target states (the step trre_dft takes lazily,
    // trre_dft.c:1135-1175; the eager construction takes it for every reachable state)
    void expand(uint32_t s) {
        if (states[s].flags & DSA_S_DECIDED) return;              // the attempt ends here
        if (flavor_ == TRRE_FLAVOR_DFT) expand_dft(s); else expand_nft(s);
    }

    // DFT flavour, lazy use: whether a freshly created state is final is needed by the transitions INTO it, before
    // its own row exists - the reference computes it at creation (trre_dft.c:1164-1174)
    void finality_only(uint32_t sid) {
        if (flavor_ != TRRE_FLAVOR_DFT || sid == 0) return;
        ++gen_;
        arrivals_.clear();
        const ItemList items = states[sid].items;
        for (const Item &it : items) {
            std::string o = it.w;
            close_dft(g_.st[it.q].a, o);
        }
        for (const Arrival &a : arrivals_) {
            const NftState &n = g_.st[a.q];
            if (n.op == N_FINAL || (n.op == N_CONS && n.val == 0)) {
                states[sid].flags = DSA_S_HAS_FINAL | DSA_S_DECIDED;
                states[sid].fin_id = pool_.intern(a.w);
                break;
            }
        }
        arrivals_.clear();
    }

    void init() {
        ItemList init;
        uint32_t flags0 = 0;
        if (flavor_ == TRRE_FLAVOR_DFT) {
            // dstart = {(init JOIN, "")}; its `final` stays -1 for ever, so the start
            // state is never accepting and an empty match cannot happen.
            init.push_back(Item{g_.start, std::string()});
        } else {
            begin_step();
            close_nft(g_.start);
            init = std::move(cur_);
            flags0 = flags_of(init);
        }
        add_state(std::move(init), flags0);
    }

  private:
    const Nft &g_;
    int flavor_;
    const CompileLimits &lim_;
    StringPool &pool_;
    std::unordered_map<std::string, uint32_t> index_;

    // scratch of one closure step
    std::vector<uint32_t> stamp_;
    std::vector<uint8_t> gray_;
    uint32_t gen_ = 0;
    ItemList cur_;
    std::string wbuf_;
    bool cut_ = false, born_final_ = false;
    int depth_ = 0;

    // host memory the construction may take before it is declared non-convergent
    uint64_t spent_ = 0;
    uint64_t work_ = 0;
    void tick() { if (++work_ > lim_.max_work) throw NotSubsequential{"work limit"}; }
    void charge(uint64_t bytes) {
        spent_ += bytes;
        if (spent_ > lim_.max_build_bytes) throw NotSubsequential{"memory limit"};
    }

    void begin_step() {
        ++gen_;
        cur_.clear();
        wbuf_.clear();
        cut_ = born_final_ = false;
    }

    static uint32_t flags_of(const ItemList &l) {
        uint32_t f = 0;
        if (!l.empty() && l.back().q == IT_FINAL) f |= DSA_S_HAS_FINAL;
        if (!l.empty() && l.back().q == IT_DIVERGE) f |= DSA_S_DIVERGE;
        if (f && l.size() == 1) f |= DSA_S_DECIDED;
        return f;
    }

    uint32_t add_state(ItemList &&l, uint32_t flags) {
        std::string key = key_of(l);
        auto it = index_.find(key);
        if (it != index_.end()) return it->second;
        if (states.size() >= lim_.max_states) throw NotSubsequential{"state limit"};
        charge(key.size() + 64 + l.size() * sizeof(Item));
        uint32_t id = (uint32_t)states.size();
        RawState rs;
        rs.flags = flags;
        if (flags & DSA_S_HAS_FINAL) rs.fin_id = pool_.intern(l.back().w);
        rs.items = std::move(l);
        states.push_back(std::move(rs));
        index_.emplace(std::move(key), id);
        return id;
    }

    struct DepthGuard {
        Determinizer &d;
        explicit DepthGuard(Determinizer &dd) : d(dd) {
            if (++d.depth_ > d.lim_.max_recursion)
                throw PatternError{TRRE_E_UNSUPPORTED, "error: expression nests too deeply for the table builder\n"};
        }
        ~DepthGuard() { --d.depth_; }
    };

    // ---- NFT flavour ---------------------------------------------------------------------

    // epsilon closure in DFS priority order (the order infer_backtrack explores, trre_nft.c:610-650)
    void close_nft(int32_t s) {
        if (s < 0 || cut_) return;
        if (gray_[s]) {                                   // epsilon cycle on the current DFS path: the
            cur_.push_back(Item{IT_DIVERGE, std::string()});   // reference pushes for ever (stack cap)
            cut_ = true;
            return;
        }
        if (stamp_[s] == gen_) return;                    // reached before with higher priority
        stamp_[s] = gen_;
        tick();
        DepthGuard guard(*this);
        const NftState &n = g_.st[s];
        switch (n.op) {
            case N_CONS:
                cur_.push_back(Item{s, wbuf_});
                return;
            case N_FINAL:
                cur_.push_back(Item{IT_FINAL, wbuf_});
                cut_ = born_final_ = true;                // everything after it has lower priority
                return;
            default: break;
        }
        gray_[s] = 1;
        switch (n.op) {
            case N_PROD: {
                size_t keep = wbuf_.size();
                if (keep >= lim_.max_residual) throw NotSubsequential{"residual limit"};
                wbuf_.push_back((char)n.val);
                close_nft(n.a);
                wbuf_.resize(keep);
                break;
            }
            case N_JOIN:    close_nft(n.a); break;
            case N_SPLIT:   close_nft(n.b); close_nft(n.a); break;
            case N_SPLITNG: close_nft(n.a); close_nft(n.b); break;
            default: break;
        }
        gray_[s] = 0;
    }

    void expand_nft(uint32_t sid) {
        const ItemList items = states[sid].items;         // copy: `states` reallocates while we add
        std::vector<std::vector<uint32_t>> by_byte(256);
        bool pending = false;
        for (uint32_t k = 0; k < items.size(); ++k) {
            if (items[k].q >= 0) by_byte[g_.st[items[k].q].val].push_back(k); else pending = true;
        }
        uint32_t idle_next = 0, idle_out = 0;
        bool idle_done = false;
        for (int c = 1; c < 256; ++c) {
            if (c == '\n') continue;                      // neither byte can occur inside a payload
            const bool live = !by_byte[c].empty();
            if (!live && !pending) continue;              // dead
            if (!live && idle_done) {                     // only the pending item survives: same result
                states[sid].edges.push_back(RawEdge{(uint8_t)c, idle_next, false, idle_out});
                charge(sizeof(RawEdge));
                continue;
            }
            begin_step();
            for (uint32_t k : by_byte[c]) {               // in list (= priority) order
                if (cut_) break;
                wbuf_ = items[k].w;
                close_nft(g_.st[items[k].q].a);
            }
            if (pending && !cut_) cur_.push_back(items.back());   // FINAL / DIVERGE persist and stay last
            if (cur_.empty()) continue;
            bool mark = born_final_;
            std::string out;
            strip_lcp(cur_, out);
            for (const Item &it : cur_)
                if (it.w.size() > lim_.max_residual) throw NotSubsequential{"residual limit"};
            uint32_t flags = flags_of(cur_);
            ItemList next = std::move(cur_);
            cur_ = ItemList();
            uint32_t nid = add_state(std::move(next), flags);
            uint32_t oid = pool_.intern(out);
            states[sid].edges.push_back(RawEdge{(uint8_t)c, nid, mark, oid});
            charge(sizeof(RawEdge));
            if (!live) { idle_done = true; idle_next = nid; idle_out = oid; }
        }
    }

    // ---- DFT flavour -----------------------------------------------------------------------

    struct Arrival { int32_t q; std::string w; };
    std::vector<Arrival> arrivals_;                       // first arrival per CONS / FINAL state, in order

    // nft_step_() for all bytes at once: `visited` is only ever set on the state that matches the
    // step's byte, and a CONS state has exactly one byte, so "first arrival per state" is the same
    // for every byte.  `o` is passed BY REFERENCE down the preferred branch and copied afterwards -
    // the reference's aliasing (trre_dft.c:879-891), which changes the emitted bytes when a PROD
    // sits on a dead preferred path.
    void close_dft(int32_t s, std::string &o) {
        if (s < 0) return;
        tick();
        DepthGuard guard(*this);
        const NftState &n = g_.st[s];
        switch (n.op) {
            case N_SPLIT:   { close_dft(n.b, o); std::string o2 = o; close_dft(n.a, o2); break; }
            case N_SPLITNG: { close_dft(n.a, o); std::string o2 = o; close_dft(n.b, o2); break; }
            case N_JOIN:    close_dft(n.a, o); break;
            case N_PROD:
                if (o.size() >= lim_.max_residual) throw NotSubsequential{"residual limit"};
                o.push_back((char)n.val);
                close_dft(n.a, o);
                break;
            case N_CONS:
            case N_FINAL:
                if (stamp_[s] != gen_) { stamp_[s] = gen_; arrivals_.push_back(Arrival{s, o}); }
                break;
        }
    }

    void expand_dft(uint32_t sid) {
        ++gen_;
        arrivals_.clear();
        {
            const ItemList items = states[sid].items;
            for (const Item &it : items) {
                std::string o = it.w;
                close_dft(g_.st[it.q].a, o);
            }
        }
        // finality of THIS state = probe with c == 0: first arrival that is FINAL or CONS(0)
        // (trre_dft.c:1164-1174).  The start state is never probed.
        if (sid != 0) {
            for (const Arrival &a : arrivals_) {
                const NftState &n = g_.st[a.q];
                if (n.op == N_FINAL || (n.op == N_CONS && n.val == 0)) {
                    states[sid].flags = DSA_S_HAS_FINAL | DSA_S_DECIDED;
                    states[sid].fin_id = pool_.intern(a.w);
                    break;
                }
            }
            if (states[sid].flags & DSA_S_DECIDED) return;   // infer_dft returns at the first final state
        }
        std::vector<std::vector<uint32_t>> by_byte(256);
        for (uint32_t k = 0; k < arrivals_.size(); ++k) {
            const NftState &n = g_.st[arrivals_[k].q];
            if (n.op == N_CONS) by_byte[n.val].push_back(k);
        }
        std::vector<Arrival> arr = std::move(arrivals_);
        arrivals_ = std::vector<Arrival>();
        for (int c = 1; c < 256; ++c) {
            if (c == '\n' || by_byte[c].empty()) continue;
            ItemList next;
            for (uint32_t k : by_byte[c]) next.push_back(Item{arr[k].q, arr[k].w});
            std::string out;
            strip_lcp(next, out);
            uint32_t nid = add_state(std::move(next), 0);
            states[sid].edges.push_back(RawEdge{(uint8_t)c, nid, false, pool_.intern(out)});
            charge(sizeof(RawEdge));
        }
    }
};

// ------------------------------------------------------------------------------------------------

struct Program {
    int flavor = 0;
    uint32_t exec_kind = TRRE_EXEC_TABLE;
    uint32_t flags = 0;
    uint32_t max_depth = 0;
    uint32_t n_classes = 1;
    uint8_t classmap[256];
    uint8_t lut[256];
    uint8_t startrow[256];
    std::vector<uint64_t> trans;
    std::vector<uint64_t> sinfo;
    std::vector<uint8_t> pool;
    std::vector<uint32_t> eout;      // see BlobHeader::off_eout
    bool vmap = false;               // PF_BYTEMAP_VAR
    std::string vstr[256];           // what an attempt starting with this byte prints (raw copy: the byte itself)
    uint8_t vrun[256] = {};          // > 0: the match then swallows every following byte of the same run class (X+:s)
    uint32_t vrun_classes = 0;
};

class PoolWriter {
  public:
    explicit PoolWriter(const StringPool &sp) : sp_(sp), refs_(sp.count(), 0xFFFFFFFFu) {}
    // -> hi word: off(24) | len(8)
    uint32_t ref(uint32_t id) {
        if (id >= refs_.size()) refs_.resize(sp_.count(), 0xFFFFFFFFu);     // strings interned since (lazy construction)
        if (refs_[id] != 0xFFFFFFFFu) return refs_[id];
        const std::string &s = sp_.get(id);
        uint32_t r;
        if (s.empty()) {
            r = 0;
        } else if (s.size() < 255) {
            uint32_t off = place(s.data(), s.size());
            r = off | ((uint32_t)s.size() << 24);
        } else {
            while (bytes.size() & 3) bytes.push_back(0);
            uint32_t len = (uint32_t)s.size();
            uint32_t off = (uint32_t)bytes.size();
            bytes.insert(bytes.end(), reinterpret_cast<uint8_t *>(&len), reinterpret_cast<uint8_t *>(&len) + 4);
            bytes.insert(bytes.end(), s.begin(), s.end());
            r = off | (255u << 24);
        }
        if (bytes.size() >= (1u << 24)) throw NotSubsequential{"output pool limit"};
        refs_[id] = r;
        return r;
    }
    std::vector<uint8_t> bytes;

  private:
    const StringPool &sp_;
    std::vector<uint32_t> refs_;
    uint32_t place(const char *p, size_t n) {
        uint32_t off = (uint32_t)bytes.size();
        bytes.insert(bytes.end(), p, p + n);
        return off;
    }
};

void analyse_and_pack(const std::vector<RawState> &st, const StringPool &sp, int flavor,
                      const CompileLimits &lim, Program &P) {
    const uint32_t S = (uint32_t)st.size();
    // ---- byte classes: bytes whose whole column (over all states) is identical ---------------
    std::vector<std::vector<uint64_t>> col(256);
    for (uint32_t s = 0; s < S; ++s)
        for (const RawEdge &e : st[s].edges) {
            col[e.byte].push_back(((uint64_t)s << 32) | e.next | (e.mark ? (uint32_t)DSA_E_MARK : 0u));
            col[e.byte].push_back(e.out_id);
        }
    std::map<std::vector<uint64_t>, uint32_t> cls_of;
    cls_of.emplace(std::vector<uint64_t>(), 0);            // class 0 = dead everywhere (incl. 0 and '\n')
    for (int c = 0; c < 256; ++c) {
        auto it = cls_of.find(col[c]);
        if (it == cls_of.end()) it = cls_of.emplace(col[c], (uint32_t)cls_of.size()).first;
        P.classmap[c] = (uint8_t)it->second;
    }
    if (cls_of.size() > 256) throw NotSubsequential{"class limit"};
    P.n_classes = (uint32_t)cls_of.size();
    if ((uint64_t)S * P.n_classes * 8 > lim.max_table_bytes) throw NotSubsequential{"table size limit"};

    // ---- pack --------------------------------------------------------------------------------
    PoolWriter pw(sp);
    P.trans.assign((size_t)S * P.n_classes, pack64(DSA_DEAD, 0));
    P.sinfo.resize(S);
    for (uint32_t s = 0; s < S; ++s) {
        P.sinfo[s] = pack64(st[s].flags, (st[s].flags & DSA_S_HAS_FINAL) ? pw.ref(st[s].fin_id) : 0);
        for (const RawEdge &e : st[s].edges) {
            // DFT flavour: the match ends where the walk reaches its first final state
            const bool decided = (st[e.next].flags & DSA_S_DECIDED) != 0;
            const bool mark = e.mark || (flavor == TRRE_FLAVOR_DFT && decided);
            uint32_t lo = e.next | (mark ? (uint32_t)DSA_E_MARK : 0u) | (decided ? (uint32_t)DSA_E_DECIDED : 0u) |
                          ((st[e.next].flags & DSA_S_HAS_FINAL) ? (uint32_t)DSA_E_HASFINAL : 0u) |
                          ((st[e.next].flags & DSA_S_DIVERGE) ? (uint32_t)DSA_E_DIVERGE : 0u);
            P.trans[(size_t)s * P.n_classes + P.classmap[e.byte]] = pack64(lo, pw.ref(e.out_id));
        }
    }
    // combined "what a match prints" strings for the position-parallel kernel: one copy per match
    {
        bool late = true;
        for (uint32_t s = 0; s < S && late; ++s)
            for (const RawEdge &e : st[s].edges)
                if (!(st[e.next].flags & DSA_S_DECIDED) && !sp.get(e.out_id).empty()) { late = false; break; }
        if (late) {
            std::unordered_map<std::string, uint32_t> placed;
            bool fits = true;
            auto place = [&](const std::string &str) -> uint32_t {             // -> offset(24) | length(8)
                if (str.empty()) return 0;
                if (str.size() > 255) { fits = false; return 0; }
                auto it = placed.find(str);
                if (it != placed.end()) return it->second;
                uint32_t ref = (uint32_t)pw.bytes.size() | ((uint32_t)str.size() << 24);
                pw.bytes.insert(pw.bytes.end(), str.begin(), str.end());
                placed.emplace(str, ref);
                return ref;
            };
            P.eout.assign((size_t)S * P.n_classes + S, 0);
            for (uint32_t s = 0; s < S; ++s) {
                if ((st[s].flags & DSA_S_HAS_FINAL) && !(st[s].flags & DSA_S_DECIDED))
                    P.eout[(size_t)S * P.n_classes + s] = place(sp.get(st[s].fin_id));
                for (const RawEdge &e : st[s].edges)
                    if (st[e.next].flags & DSA_S_DECIDED)
                        P.eout[(size_t)s * P.n_classes + P.classmap[e.byte]] = place(sp.get(e.out_id) + sp.get(st[e.next].fin_id));
            }
            if (pw.bytes.size() >= (1u << 24) || !fits) P.eout.clear();
        }
    }
    P.pool = std::move(pw.bytes);
    while (P.pool.size() & 15) P.pool.push_back(0);
    for (int c = 0; c < 256; ++c) P.startrow[c] = 0;
    for (const RawEdge &e : st[0].edges) P.startrow[e.byte] = 1;

    // ---- flags ---------------------------------------------------------------------------------
    if (st[0].flags & DSA_S_HAS_FINAL) P.flags |= PF_CAN_MATCH_EMPTY;
    {
        bool late = true;
        for (uint32_t s = 0; s < S && late; ++s)
            for (const RawEdge &e : st[s].edges)
                if (!(st[e.next].flags & DSA_S_DECIDED) && !sp.get(e.out_id).empty()) { late = false; break; }
        if (late) P.flags |= PF_OUT_LATE;
    }
    for (uint32_t s = 0; s < S; ++s)
        if (st[s].flags & DSA_S_DIVERGE) P.flags |= PF_HAS_DIVERGE;

    // longest attempt / acyclicity: iterative DFS with colours over non-decided states
    {
        std::vector<uint8_t> colour(S, 0);
        std::vector<uint32_t> depth(S, 0);
        bool cyclic = false;
        std::vector<std::pair<uint32_t, size_t>> stack;
        stack.emplace_back(0u, 0u);
        colour[0] = 1;
        while (!stack.empty() && !cyclic) {
            auto &top = stack.back();
            uint32_t s = top.first;
            if (top.second < st[s].edges.size()) {
                uint32_t t = st[s].edges[top.second++].next;
                if (colour[t] == 1) cyclic = true;
                else if (colour[t] == 0) { colour[t] = 1; stack.emplace_back(t, 0u); }
            } else {
                uint32_t d = 0;
                for (const RawEdge &e : st[s].edges) d = std::max(d, depth[e.next] + 1);
                depth[s] = d;
                colour[s] = 2;
                stack.pop_back();
            }
        }
        if (cyclic) P.max_depth = 0xFFFFFFFFu;
        else { P.max_depth = depth[0]; P.flags |= PF_ACYCLIC; }
    }

    // length preservation: along every path |emitted| - consumed (bal) and consumed - match_end (lag)
    // must be path independent, and every accepting outcome must emit exactly match_end bytes
    {
        const int64_t UNSET = INT64_MIN;
        std::vector<int64_t> bal(S, UNSET), lag(S, UNSET);
        bool lp = true;
        std::vector<uint32_t> work;
        bal[0] = 0;
        lag[0] = 0;
        work.push_back(0);
        auto accepts_ok = [&](uint32_t t, int64_t b, int64_t l) {
            // an attempt that ends in t prints acc + fin(t) and consumes i - lag bytes
            return b + (int64_t)sp.get(st[t].fin_id).size() == -l;
        };
        if ((st[0].flags & DSA_S_HAS_FINAL) && !accepts_ok(0, 0, 0)) lp = false;
        while (!work.empty() && lp) {
            uint32_t s = work.back();
            work.pop_back();
            for (const RawEdge &e : st[s].edges) {
                const uint32_t t = e.next;
                int64_t b = bal[s] + (int64_t)sp.get(e.out_id).size() - 1;
                int64_t l = e.mark ? 0 : lag[s] + 1;
                if (flavor == TRRE_FLAVOR_DFT) l = 0;          // the match ends where the walk ends
                if ((st[t].flags & DSA_S_HAS_FINAL) && !accepts_ok(t, b, l)) { lp = false; break; }
                if (st[t].flags & DSA_S_DECIDED) continue;     // terminal: nothing to propagate, any path may end here
                if (bal[t] == UNSET) { bal[t] = b; lag[t] = l; work.push_back(t); }
                else if (bal[t] != b || ((st[t].flags & DSA_S_HAS_FINAL) && lag[t] != l)) { lp = false; break; }
            }
        }
        if (lp) P.flags |= PF_LENGTH_PRESERVING;
    }

    // byte map: every attempt is decided by its first byte and emits exactly one byte
    {
        bool bm = !(st[0].flags & (DSA_S_HAS_FINAL | DSA_S_DIVERGE));
        for (int c = 0; c < 256; ++c) P.lut[c] = (uint8_t)c;
        for (const RawEdge &e : st[0].edges) {
            if (!bm) break;
            const RawState &t = st[e.next];
            bool ok = (t.flags & DSA_S_DECIDED) && (t.flags & DSA_S_HAS_FINAL) &&
                      (flavor == TRRE_FLAVOR_DFT || e.mark);
            if (!ok) { bm = false; break; }
            std::string o = sp.get(e.out_id) + sp.get(t.fin_id);
            if (o.size() != 1) { bm = false; break; }
            P.lut[e.byte] = (uint8_t)o[0];
        }
        if (bm && (P.flags & PF_LENGTH_PRESERVING)) P.exec_kind = TRRE_EXEC_BYTEMAP;
    }
    // variable-length byte map: every attempt is still decided by its first byte, but prints 0..255 bytes
    // (deletions `a:`, expansions `a:xy`, tr-like maps mixed with them).  RUNS (`[0-9]+:#`, `a+:x`, ` +: `): the first byte
    // decides as well, but the match then greedily swallows every following byte of its own set - the first edge leads to
    // a state with a pending FINAL whose only edges are unprinting marked self-loops on exactly the bytes that lead there
    // from the start.  A byte is then covered by an earlier match iff its predecessor is of the same run class: still
    // position parallel, with one byte of left context (k_bmv_*<RUN>).
    if (P.exec_kind != TRRE_EXEC_BYTEMAP) {
        bool vm = !(st[0].flags & (DSA_S_HAS_FINAL | DSA_S_DIVERGE));
        for (int c = 0; c < 256; ++c) P.vstr[c].assign(1, (char)c);
        size_t pool = 0;
        std::vector<uint32_t> run_states;
        for (const RawEdge &e : st[0].edges) {
            if (!vm) break;
            const RawState &t = st[e.next];
            bool ok = (t.flags & DSA_S_DECIDED) && (t.flags & DSA_S_HAS_FINAL) && !(t.flags & DSA_S_DIVERGE) &&
                      (flavor == TRRE_FLAVOR_DFT || e.mark);
            if (!ok && flavor == TRRE_FLAVOR_NFT && e.mark && e.next != 0 && (t.flags & DSA_S_HAS_FINAL) &&
                !(t.flags & (DSA_S_DECIDED | DSA_S_DIVERGE))) {
                size_t k = 0;
                while (k < run_states.size() && run_states[k] != e.next) ++k;
                if (k == run_states.size()) run_states.push_back(e.next);
                if (run_states.size() <= 15) { P.vrun[e.byte] = (uint8_t)(k + 1); ok = true; }
            }
            if (!ok) { vm = false; break; }
            P.vstr[e.byte] = sp.get(e.out_id) + sp.get(t.fin_id);
            if (P.vstr[e.byte].size() > 255) { vm = false; break; }
            if (P.vstr[e.byte].size() > 1) pool += P.vstr[e.byte].size();
        }
        for (size_t k = 0; k < run_states.size() && vm; ++k) {
            // the run state loops silently on exactly its own bytes; any other byte (read ahead, not consumed: no mark) ends
            // the match with what the end of the payload would print
            const RawState &t = st[run_states[k]];
            size_t members = 0, loops = 0;
            for (int c = 0; c < 256; ++c) members += P.vrun[c] == k + 1;
            for (const RawEdge &e2 : t.edges) {
                if (P.vrun[e2.byte] == k + 1) {
                    ++loops;
                    if (e2.next != run_states[k] || !e2.mark || !sp.get(e2.out_id).empty()) vm = false;
                } else {
                    const RawState &t2 = st[e2.next];
                    if (e2.mark || !(t2.flags & DSA_S_DECIDED) || !(t2.flags & DSA_S_HAS_FINAL) || (t2.flags & DSA_S_DIVERGE) ||
                        sp.get(e2.out_id) + sp.get(t2.fin_id) != sp.get(t.fin_id))
                        vm = false;
                }
            }
            if (loops != members) vm = false;
        }
        if (!run_states.empty() && pool >= 4000) vm = false;          // run class bits share the entry with the pool offset
        if (vm && pool < 60000 && !st[0].edges.empty()) { P.vmap = true; P.vrun_classes = (uint32_t)run_states.size(); }
        else { std::memset(P.vrun, 0, sizeof P.vrun); P.vrun_classes = 0; }
    }
}

// Gram tables (k_gram): the outcome of an attempt of an acyclic program is a function of the byte classes of its first
// D = max_depth bytes.  Walk every combination here, once: bit = the attempt matches; out = what it prints, how much
// it consumes.  Mirrors walk_from() in match_kernel.cuh (itself the table form of infer_dft / infer_backtrack).
constexpr uint32_t GRAM_MAPDELTA_FLAG = 2u; // = GRAM_MAPDELTA in gram_kernel.cuh
struct GramTables {
    uint32_t depth = 0, entries = 0, uniform_len = 0, flags = 0, delta = 0;
    std::vector<uint32_t> bits, out;
    std::vector<uint8_t> pool;
};

bool build_grams(const Program &P, GramTables &G) {
    const uint32_t D = P.max_depth, ncls = P.n_classes;
    if (D < 1 || D > 6 || P.trans.empty()) return false;
    uint64_t entries = 1;
    for (uint32_t i = 0; i < D; ++i) { entries *= ncls; if (entries > 640u * 1024u) return false; }
    const uint32_t S = (uint32_t)P.sinfo.size();
    G.depth = D;
    G.entries = (uint32_t)entries;
    G.bits.assign((entries + 31) / 32, 0);
    G.out.assign(entries, 0);
    std::unordered_map<std::string, uint32_t> placed;
    auto pool_str = [&](uint32_t hi, std::string &dst) {
        uint32_t len = hi >> 24, off = hi & 0xFFFFFFu;
        if (len == 255u) { std::memcpy(&len, P.pool.data() + off, 4); off += 4; }
        dst.append(reinterpret_cast<const char *>(P.pool.data() + off), len);
    };
    std::vector<std::string> strs(entries);
    std::vector<uint8_t> consumed(entries, 0);
    bool uniform = true, all_short = true, lp = true;
    uint32_t common = 0;
    std::string str;
    for (uint64_t idx = 0; idx < entries; ++idx) {
        uint32_t s = 0, m = 0, lo_last = 0, steps = 0;
        uint64_t rest = idx;
        str.clear();
        for (uint32_t i = 0; i < D; ++i) {
            const uint32_t k = (uint32_t)(rest % ncls);
            rest /= ncls;
            const uint64_t e = P.trans[(size_t)s * ncls + k];
            const uint32_t lo = (uint32_t)e;
            if ((lo & 0xFFFFFFu) == DSA_DEAD) break;
            pool_str((uint32_t)(e >> 32), str);
            lo_last = lo;
            s = lo & 0xFFFFFFu;
            ++steps;
            if (lo & DSA_E_MARK) m = steps;
            if (lo & DSA_E_DECIDED) break;
        }
        if (!steps || !(lo_last & (DSA_E_DECIDED | DSA_E_HASFINAL)) || !m || s >= S) continue;
        pool_str((uint32_t)(P.sinfo[s] >> 32), str);
        if (str.size() > 255 || m > 16) return false;
        G.bits[idx >> 5] |= 1u << (idx & 31);
        strs[idx] = str;
        consumed[idx] = (uint8_t)m;
        if (!common) common = m; else if (common != m) uniform = false;
        if (str.size() > 4) all_short = false;
        if (str.size() != m) lp = false;
    }
    if (!common) return false;                                    // nothing can match: leave it to the generic kernels
    G.uniform_len = uniform ? common : 0;
    const bool direct = uniform && all_short && lp;
    if (direct) G.flags |= GRAM_DIRECT_FLAG;
    if (direct) {
        // do all matches print their own bytes moved by ONE constant (word:WORD dictionaries: -32)?  Then the device needs
        // no output table at all: matched bytes + delta (k_gram REG = 2)
        int single[256];
        for (uint32_t c = 0; c < ncls; ++c) single[c] = -1;
        {
            int members[256] = {};
            for (int b = 0; b < 256; ++b) { ++members[P.classmap[b]]; single[P.classmap[b]] = b; }
            for (uint32_t c = 0; c < ncls; ++c) if (members[c] != 1) single[c] = -1;
        }
        bool map = true;
        int delta = -1;
        for (uint64_t idx = 0; idx < entries && map; ++idx) {
            if (!consumed[idx]) continue;
            uint64_t rest = idx;
            for (uint32_t k = 0; k < common; ++k) {
                const int b = single[rest % ncls];
                rest /= ncls;
                if (b < 0) { map = false; break; }
                const int d = ((int)(uint8_t)strs[idx][k] - b) & 0xFF;
                if (delta < 0) delta = d; else if (delta != d) { map = false; break; }
            }
        }
        if (map && delta > 0) { G.flags |= GRAM_MAPDELTA_FLAG; G.delta = (uint32_t)delta; }
    }
    for (uint64_t idx = 0; idx < entries; ++idx) {
        if (!consumed[idx]) continue;
        const std::string &o = strs[idx];
        if (direct) {
            uint32_t w = 0;
            for (size_t k = 0; k < o.size(); ++k) w |= (uint32_t)(uint8_t)o[k] << (8 * k);
            G.out[idx] = w;
        } else {
            uint32_t off = 0;
            if (!o.empty()) {
                auto it = placed.find(o);
                if (it == placed.end()) {
                    off = (uint32_t)G.pool.size();
                    G.pool.insert(G.pool.end(), o.begin(), o.end());
                    placed.emplace(o, off);
                } else off = it->second;
            }
            if (G.pool.size() >= (1u << 20)) return false;
            G.out[idx] = off | ((uint32_t)o.size() << 20) | ((uint32_t)(consumed[idx] - 1) << 28);
        }
    }
    while (G.pool.size() & 15) G.pool.push_back(0);
    return true;
}

// The literal form of a gram program (BlobHeader::lit): every match consumes and prints L = 2..4 bytes, whether an
// attempt matches depends on its first L byte classes only, there are at most 4 matching class words, every class in
// them is one byte range inside one half of the byte space, every printed byte is non-zero, and no proper suffix of a
// word is the prefix of a word (then an attempt that starts inside a match can never match, so the set of matches is
// simply the set of candidates - no chain to resolve).
bool build_lit(const Program &P, const GramTables &G, uint32_t *n_words, uint32_t *len, uint32_t (&lit)[16], uint32_t (&lit_out)[4]) {
    const uint32_t L = G.uniform_len, D = G.depth, ncls = P.n_classes;
    const bool direct = (G.flags & GRAM_DIRECT_FLAG) != 0;            // else: G.out = pool offset(20) | printed length(8) << 20 | ..
    if (L < 2 || L > 4 || D < L) return false;
    uint64_t prefixes = 1, completions = 1;
    for (uint32_t i = 0; i < L; ++i) prefixes *= ncls;
    for (uint32_t i = L; i < D; ++i) completions *= ncls;
    struct Word { uint32_t prefix, out; uint64_t seen; };
    std::vector<Word> words;
    for (uint64_t idx = 0; idx < G.entries; ++idx) {
        if (!(G.bits[idx >> 5] >> (idx & 31) & 1u)) continue;
        const uint32_t pre = (uint32_t)(idx % prefixes);
        size_t k = 0;
        while (k < words.size() && words[k].prefix != pre) ++k;
        if (k == words.size()) {
            if (words.size() == 4) return false;
            words.push_back(Word{pre, G.out[idx], 0});
        }
        if (words[k].out != G.out[idx]) return false;                 // what follows the word changes what it prints
        ++words[k].seen;
    }
    if (words.empty()) return false;
    for (const Word &w : words)
        if (w.seen != completions) return false;                      // what follows the word decides whether it matches
    auto cls_at = [&](const Word &w, uint32_t k) { uint32_t v = w.prefix; for (uint32_t i = 0; i < k; ++i) v /= ncls; return v % ncls; };
    for (size_t a = 0; a < words.size(); ++a) {
        if (!direct && (words[a].out >> 28) + 1 != L) return false;   // (consumed - 1) of the variable form
        lit_out[a] = direct ? 0u : (words[a].out & 0x0FFFFFFFu);
        for (uint32_t k = 0; k < L; ++k) {
            const uint32_t c = cls_at(words[a], k);
            int lo = -1, hi = -1, members = 0;
            for (int b = 0; b < 256; ++b)
                if (P.classmap[b] == c) { if (lo < 0) lo = b; hi = b; ++members; }
            if (lo <= 0 || members != hi - lo + 1 || (lo < 0x80) != (hi < 0x80) || (lo <= '\n' && hi >= '\n')) return false;
            const uint32_t printed = direct ? (words[a].out >> (8 * k)) & 0xFFu : 0u;
            if (direct && !printed) return false;
            lit[a * L + k] = (uint32_t)lo | ((uint32_t)hi << 8) | (printed << 16);
        }
        for (size_t b = 0; b < words.size(); ++b)
            for (uint32_t s = 1; s < L; ++s) {                         // words[b] starting s bytes into words[a]
                bool fits = true;
                for (uint32_t k = 0; k + s < L && fits; ++k) fits = cls_at(words[a], s + k) == cls_at(words[b], k);
                if (fits) return false;
            }
    }
    *n_words = (uint32_t)words.size();
    *len = L;
    return true;
}

std::vector<uint8_t> assemble(const Program &P, const Nft &g) {
    BlobHeader h;
    std::memset(&h, 0, sizeof h);
    auto align16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    size_t off = align16(sizeof(BlobHeader));
    h.off_classmap = (uint32_t)off; off += 256;
    h.off_lut = (uint32_t)off; off += 256;
    h.off_startrow = (uint32_t)off; off += 256;
    off = align16(off);
    h.off_trans = (uint32_t)off; off += P.trans.size() * 8; off = align16(off);
    h.off_sinfo = (uint32_t)off; off += P.sinfo.size() * 8; off = align16(off);
    h.off_pool = (uint32_t)off; off += P.pool.size(); off = align16(off);
    h.off_nft = (uint32_t)off; off += g.st.size() * sizeof(NftFlat); off = (off + 127) & ~(size_t)127;
    if (P.exec_kind == TRRE_EXEC_BYTEMAP) { h.off_pairlut = (uint32_t)off; off += 65536 * 2; }
    // compact tables of the position-parallel kernel (k_match): length-preserving, acyclic, short attempts
    // attempts of at most 16 bytes; an automaton with a cycle qualifies as well ([0-9]+:#, (a:b)+): its attempts are short
    // on ordinary text, and the kernel hands the buffer to the per-line path when one is still alive after 16 bytes
    const bool short_attempts = (P.flags & PF_ACYCLIC) ? (P.max_depth >= 1 && P.max_depth <= 16) : true;
    const bool matchable = P.exec_kind == TRRE_EXEC_TABLE && (P.flags & PF_LENGTH_PRESERVING) && short_attempts &&
                           !(P.flags & (PF_CAN_MATCH_EMPTY | PF_HAS_DIVERGE)) && !P.trans.empty();
    std::vector<uint8_t> vpool;
    std::vector<uint32_t> vent(256, 0);
    if (P.vmap) {
        for (int c = 0; c < 256; ++c) {
            const std::string &str = P.vstr[c];
            uint32_t ent = (uint32_t)str.size() | ((str.empty() ? 0u : (uint32_t)(uint8_t)str[0]) << 8);
            if (str.size() > 1) { ent |= (uint32_t)vpool.size() << 16; vpool.insert(vpool.end(), str.begin(), str.end()); }
            vent[c] = ent | ((uint32_t)P.vrun[c] << 28);              // (vrun_classes > 0 implies a pool below 4 KB)
        }
        while (vpool.size() & 15) vpool.push_back(0);
        h.off_vent = (uint32_t)off; off += 1024;
        off = (off + 127) & ~(size_t)127;
        h.off_vpairlen = (uint32_t)off; off += 65536 * 2;
        h.off_vpool = (uint32_t)off; h.vpool_bytes = (uint32_t)vpool.size(); off += vpool.size(); off = align16(off);
    }
    // the same, but a match may print more or fewer bytes than it consumes (s/cat/mouse/): needs the single
    // output string per match (PF_OUT_LATE tables); byte maps have their own kernels
    const bool matchable_var = P.exec_kind == TRRE_EXEC_TABLE && !(P.flags & PF_LENGTH_PRESERVING) && short_attempts &&
                               !(P.flags & (PF_CAN_MATCH_EMPTY | PF_HAS_DIVERGE)) && !P.eout.empty() && !P.vmap;
    if (!P.trans.empty()) {                                   // compact tables: every table program
        h.off_row0 = (uint32_t)off; off += 256 * 4;
        h.off_ctrans = (uint32_t)off; off += P.trans.size() * 4; off = align16(off);
        if ((matchable || matchable_var) && !P.eout.empty()) { h.off_eout = (uint32_t)off; off += P.eout.size() * 4; off = align16(off); }
    }
    GramTables G;
    const bool gram = (matchable || matchable_var) && build_grams(P, G);
    if (gram) {
        off = (off + 127) & ~(size_t)127;
        h.off_gbits = (uint32_t)off; off += G.bits.size() * 4; off = align16(off);
        h.off_gout = (uint32_t)off; off += G.out.size() * 4; off = align16(off);
        h.off_gpool = (uint32_t)off; h.gpool_bytes = (uint32_t)G.pool.size(); off += G.pool.size(); off = align16(off);
        h.gram_depth = G.depth; h.gram_entries = G.entries; h.gram_len = G.uniform_len; h.gram_flags = G.flags;
        h.gram_delta = G.delta;
        // the literal word form: length preserving (k_lit) or with outputs of any length (k_litv_*: same comparison,
        // k_bmv's placement)
        const bool want_lit = (matchable && (G.flags & GRAM_DIRECT_FLAG)) || (matchable_var && !(G.flags & GRAM_DIRECT_FLAG));
        if (!want_lit || !build_lit(P, G, &h.lit_words, &h.lit_len, h.lit, h.lit_out)) {
            h.lit_words = h.lit_len = 0;
            std::memset(h.lit, 0, sizeof h.lit);
            std::memset(h.lit_out, 0, sizeof h.lit_out);
        }
    }
    if (off >= (1ull << 31)) throw PatternError{TRRE_E_UNSUPPORTED, "error: compiled tables exceed 2 GiB\n"};
    h.magic = 0x32425254u;  // "TRB2"
    h.abi = TRRE_B200_ABI_VERSION;
    h.total_bytes = (uint32_t)off;
    h.flavor = (uint32_t)P.flavor;
    h.exec_kind = P.exec_kind;
    h.flags = P.flags | (matchable ? (uint32_t)PF_MATCHABLE : 0u) | (P.vmap ? (uint32_t)PF_BYTEMAP_VAR : 0u) |
              (matchable_var ? (uint32_t)PF_MATCHABLE_VAR : 0u) | (gram ? (uint32_t)PF_GRAM : 0u);
    h.nft_states = (uint32_t)g.st.size();
    h.nft_start = g.start;
    h.dsa_states = (uint32_t)P.sinfo.size();
    h.n_classes = P.n_classes;
    h.max_depth = P.max_depth;
    h.start_info = P.sinfo.empty() ? 0 : (uint32_t)(P.sinfo[0] & 0xFFFFFFFFu);
    h.pool_bytes = (uint32_t)P.pool.size();
    h.vrun_classes = P.vrun_classes;
    if (P.vmap && P.vrun_classes >= 1 && P.vrun_classes <= 4) {
        bool ok = true;
        for (uint32_t k = 1; k <= P.vrun_classes && ok; ++k) {
            int lo = -1, hi = -1, members = 0;
            size_t len = 0;
            for (int c = 0; c < 256; ++c)
                if (P.vrun[c] == k) {
                    if (lo < 0) { lo = c; len = P.vstr[c].size(); }
                    hi = c; ++members;
                    if (P.vstr[c].size() != len) ok = false;
                }
            if (lo < 0 || members != hi - lo + 1 || (lo < 0x80) != (hi < 0x80) || len > 255) ok = false;
            if (ok) h.vrun_swar[k - 1] = (uint32_t)lo | ((uint32_t)hi << 8) | ((uint32_t)len << 16);
        }
        if (ok) h.vrun_swar_n = P.vrun_classes;
        else std::memset(h.vrun_swar, 0, sizeof h.vrun_swar);
    }
    if (P.exec_kind == TRRE_EXEC_BYTEMAP) {
        // maximal runs of bytes that move by the same non-zero delta, cut at 0x80 (k_bytemap_swar tests 7-bit ranges)
        std::vector<uint32_t> runs;
        for (int c = 0; c < 256;) {
            const uint32_t d = (uint32_t)(uint8_t)(P.lut[c] - c);
            int e = c;
            while (e + 1 < 256 && ((e + 1) & 0x7F) != 0 && (uint8_t)(P.lut[e + 1] - (e + 1)) == d) ++e;
            if (d) runs.push_back((uint32_t)c | ((uint32_t)e << 8) | (d << 16));
            c = e + 1;
        }
        if (!runs.empty() && runs.size() <= 4) {
            h.swar_n = (uint32_t)runs.size();
            for (size_t k = 0; k < runs.size(); ++k) h.swar[k] = runs[k];
        }
    }
    std::vector<uint8_t> blob(off, 0);
    std::memcpy(blob.data(), &h, sizeof h);
    std::memcpy(blob.data() + h.off_classmap, P.classmap, 256);
    std::memcpy(blob.data() + h.off_lut, P.lut, 256);
    std::memcpy(blob.data() + h.off_startrow, P.startrow, 256);
    if (!P.trans.empty()) std::memcpy(blob.data() + h.off_trans, P.trans.data(), P.trans.size() * 8);
    if (!P.sinfo.empty()) std::memcpy(blob.data() + h.off_sinfo, P.sinfo.data(), P.sinfo.size() * 8);
    if (!P.pool.empty()) std::memcpy(blob.data() + h.off_pool, P.pool.data(), P.pool.size());
    NftFlat *nf = reinterpret_cast<NftFlat *>(blob.data() + h.off_nft);
    for (size_t k = 0; k < g.st.size(); ++k) {
        nf[k].op = g.st[k].op; nf[k].val = g.st[k].val; nf[k].pad0 = nf[k].pad1 = 0;
        nf[k].a = g.st[k].a; nf[k].b = g.st[k].b;
    }
    if (h.off_row0) {
        uint32_t *row0 = reinterpret_cast<uint32_t *>(blob.data() + h.off_row0);
        uint32_t *ct = reinterpret_cast<uint32_t *>(blob.data() + h.off_ctrans);
        for (int c = 0; c < 256; ++c) row0[c] = (uint32_t)(P.trans[P.classmap[c]] & 0xFFFFFFFFu);
        for (size_t k = 0; k < P.trans.size(); ++k) ct[k] = (uint32_t)(P.trans[k] & 0xFFFFFFFFu);
    }
    if (h.off_eout) std::memcpy(blob.data() + h.off_eout, P.eout.data(), P.eout.size() * 4);
    if (gram) {
        std::memcpy(blob.data() + h.off_gbits, G.bits.data(), G.bits.size() * 4);
        std::memcpy(blob.data() + h.off_gout, G.out.data(), G.out.size() * 4);
        if (!G.pool.empty()) std::memcpy(blob.data() + h.off_gpool, G.pool.data(), G.pool.size());
    }
    if (h.off_vent) {
        std::memcpy(blob.data() + h.off_vent, vent.data(), 1024);
        if (!vpool.empty()) std::memcpy(blob.data() + h.off_vpool, vpool.data(), vpool.size());
        uint16_t *pl = reinterpret_cast<uint16_t *>(blob.data() + h.off_vpairlen);
        for (uint32_t v = 0; v < 65536; ++v)
            pl[v ^ (((v >> 8) & 0x1Fu) << 1)] = (uint16_t)((vent[v & 0xFF] & 0xFFu) + (vent[v >> 8] & 0xFFu));
    }
    if (h.off_pairlut) {
        // pair[c0 | c1 << 8] = LUT[c0] | LUT[c1] << 8 : one shared-memory lookup maps two input bytes
        uint16_t *pair = reinterpret_cast<uint16_t *>(blob.data() + h.off_pairlut);
        // stored swizzled: the shared-memory bank of an entry would otherwise depend on bits 1-5 of the FIRST
        // byte only (about 13 distinct banks on lowercase text); XOR-ing the low 5 bits of the second byte
        // into the bank bits spreads a warp's 32 lookups over all banks (k_bytemap::map_word mirrors this)
        for (uint32_t v = 0; v < 65536; ++v)
            pair[v ^ (((v >> 8) & 0x1Fu) << 1)] = (uint16_t)(P.lut[v & 0xFF] | ((uint32_t)P.lut[v >> 8] << 8));
    }
    return blob;
}

}  // namespace

std::vector<uint8_t> compile_to_blob(const char *expr, int flavor, const CompileLimits &lim) {
    if (flavor != TRRE_FLAVOR_NFT && flavor != TRRE_FLAVOR_DFT)
        throw PatternError{TRRE_E_ARG, "error: unknown flavor\n"};
    SynTree tree = parse_expression(expr);
    Nft g = build_nft(tree, flavor == TRRE_FLAVOR_DFT);

    Program P;
    P.flavor = flavor;
    // can a produced NUL ever be printed?  CONS(0) and CONS('\n') never fire (the payload is a C string
    // inside one getline record), so what lies behind them - e.g. the PROD(0) of `.` - is unreachable
    bool nul_out = false;
    {
        std::vector<uint8_t> seen(g.st.size(), 0);
        std::vector<int32_t> todo;
        if (g.start >= 0) { todo.push_back(g.start); seen[g.start] = 1; }
        while (!todo.empty()) {
            int32_t s = todo.back();
            todo.pop_back();
            const NftState &n = g.st[s];
            if (n.op == N_PROD && n.val == 0) nul_out = true;
            if (n.op == N_CONS && (n.val == 0 || n.val == '\n')) continue;
            for (int32_t t : {n.a, n.b})
                if (t >= 0 && !seen[t]) { seen[t] = 1; todo.push_back(t); }
        }
    }
    if (nul_out) P.flags |= PF_NUL_IN_OUTPUT;

    bool tables = true;
    // fputs() stops at a produced NUL (trre_nft.c:644-645): only the DFS executor models that
    if (flavor == TRRE_FLAVOR_NFT && nul_out) tables = false;
    if (tables) {
        try {
            StringPool sp;
            Determinizer det(g, flavor, lim, sp);
            det.run();
            analyse_and_pack(det.states, sp, flavor, lim, P);
        } catch (const PatternError &pe) {
            if (pe.code != TRRE_E_UNSUPPORTED) throw;
            tables = false;
        } catch (const NotSubsequential &ns) {
            if (std::getenv("TRRE_DEBUG")) std::fprintf(stderr, "trre_b200: table construction abandoned: %s\n", ns.why);
            tables = false;
        }
    }
    if (!tables) {
        uint32_t keep = P.flags & PF_NUL_IN_OUTPUT;
        P = Program();
        P.flavor = flavor;
        P.flags = keep;
        // no table ahead of time.  DFT flavour: build rows on demand, as trre_dft itself does (LazyDsa); NFT flavour:
        // the reference's own DFS on the GPU
        P.exec_kind = flavor == TRRE_FLAVOR_DFT ? TRRE_EXEC_LAZY : TRRE_EXEC_BACKTRACK;
        P.n_classes = 1;
        for (int c = 0; c < 256; ++c) { P.classmap[c] = 0; P.lut[c] = (uint8_t)c; P.startrow[c] = 1; }
        P.max_depth = 0xFFFFFFFFu;
    }
    return assemble(P, g);
}

// ---- LazyDsa ---------------------------------------------------------------------------------------------------------

struct LazyDsa::Impl {
    Nft g;
    int flavor;
    CompileLimits lim;
    StringPool sp;
    Determinizer det;
    PoolWriter pw;
    std::vector<uint64_t> rows, sinfo;
    std::vector<uint32_t> log;
    std::vector<uint8_t> built, poisoned;

    // a closure that does not terminate (epsilon cycle, with or without output) is not a resource limit: the reference
    // would recurse for ever if the input ever led there.  Such states are poisoned, not fatal - the input may never
    // reach them (a row is built for all 256 bytes at once, the reference only follows the byte it sees).
    static bool is_runaway(const char *why) { return std::strcmp(why, "residual limit") == 0; }

    static CompileLimits lazy_limits() {
        CompileLimits l;
        l.max_states = 1u << 18;                   // 2 KB of row per state on the device
        l.max_build_bytes = 2048ull << 20;
        l.max_residual = 1u << 16;
        l.max_work = ~0ull >> 1;
        return l;
    }
    Impl(const Nft &gg, int fl) : g(gg), flavor(fl), lim(lazy_limits()), det(g, fl, lim, sp), pw(sp) {
        det.init();
        sync_states();
    }
    // rows / sinfo for states created since the last call
    void sync_states() {
        const size_t n = det.states.size();
        const size_t old = sinfo.size();
        if (n == old) return;
        rows.resize(n * 256, pack64(LAZY_UNBUILT, 0));
        sinfo.resize(n);
        built.resize(n, 0);
        poisoned.resize(n, 0);
        for (size_t s = old; s < n; ++s) {
            try {
                det.finality_only((uint32_t)s);
            } catch (const NotSubsequential &ns) {
                if (!is_runaway(ns.why)) throw;
                poisoned[s] = 1;
            } catch (const PatternError &) {                   // recursion depth
                poisoned[s] = 1;
            }
            const RawState &st = det.states[s];
            sinfo[s] = pack64(st.flags, (st.flags & DSA_S_HAS_FINAL) ? pw.ref(st.fin_id) : 0);
        }
    }
    void build_row(uint32_t s) {
        if (s >= built.size() || built[s]) return;
        bool runaway = poisoned[s] != 0;
        if (!runaway) {
            try {
                det.expand(s);
            } catch (const NotSubsequential &ns) {
                if (!is_runaway(ns.why)) throw;
                runaway = true;
            } catch (const PatternError &) {
                runaway = true;
            }
        }
        sync_states();                                    // the targets' flags are needed by the entries below
        RawState &st = det.states[s];
        uint64_t *row = rows.data() + (size_t)s * 256;
        for (int c = 0; c < 256; ++c) row[c] = pack64(runaway ? LAZY_POISON : DSA_DEAD, 0);
        if (runaway) st.edges.clear();
        for (const RawEdge &e : st.edges) {
            if (poisoned[e.next]) { row[e.byte] = pack64(LAZY_POISON, 0); continue; }
            const uint32_t tf = det.states[e.next].flags;
            const bool decided = (tf & DSA_S_DECIDED) != 0;
            const bool mark = e.mark || (flavor == TRRE_FLAVOR_DFT && decided);
            const uint32_t lo = e.next | (mark ? (uint32_t)DSA_E_MARK : 0u) | (decided ? (uint32_t)DSA_E_DECIDED : 0u) |
                                ((tf & DSA_S_HAS_FINAL) ? (uint32_t)DSA_E_HASFINAL : 0u) |
                                ((tf & DSA_S_DIVERGE) ? (uint32_t)DSA_E_DIVERGE : 0u);
            row[e.byte] = pack64(lo, pw.ref(e.out_id));
        }
        std::vector<RawEdge>().swap(st.edges);            // the packed row is what remains
        built[s] = 1;
        sinfo[s] |= DSA_S_ROW_BUILT;
        log.push_back(s);
    }
};

LazyDsa::LazyDsa(const Nft &g, int flavor) : im_(new Impl(g, flavor)) {}
LazyDsa::~LazyDsa() { delete im_; }
uint32_t LazyDsa::n_states() const { return (uint32_t)im_->sinfo.size(); }
const std::vector<uint64_t> &LazyDsa::rows() const { return im_->rows; }
const std::vector<uint64_t> &LazyDsa::sinfo() const { return im_->sinfo; }
const std::vector<uint8_t> &LazyDsa::pool() const { return im_->pw.bytes; }
const std::vector<uint32_t> &LazyDsa::log() const { return im_->log; }
uint32_t LazyDsa::start_info() const { return (uint32_t)(im_->sinfo[0] & 0xFFFFFFFFu); }

void LazyDsa::build(const uint32_t *req, size_t n_req, uint32_t extra) {
    try {
        std::vector<uint32_t> frontier;
        size_t first_new = im_->sinfo.size();
        for (size_t k = 0; k < n_req; ++k) im_->build_row(req[k]);
        // breadth first behind the requests: the states they created, then the states those create ...
        for (size_t s = first_new; s < im_->sinfo.size() && extra; ++s) {
            if (im_->built[s] || (im_->det.states[s].flags & DSA_S_DECIDED)) continue;
            im_->build_row((uint32_t)s);
            --extra;
        }
    } catch (const NotSubsequential &ns) {
        throw PatternError{TRRE_E_UNSUPPORTED,
                           std::string("error: the scan automaton outgrew the on-demand table limits (") + ns.why + ")\n"};
    }
}

bool blob_is_valid(const void *blob, size_t nbytes, std::string *why) {
    auto bad = [&](const char *w) { if (why) *why = w; return false; };
    if (!blob || nbytes < sizeof(BlobHeader)) return bad("blob too small");
    BlobHeader h;
    std::memcpy(&h, blob, sizeof h);
    if (h.magic != 0x32425254u) return bad("bad magic");
    if (h.abi != TRRE_B200_ABI_VERSION) return bad("ABI version mismatch");
    if (h.total_bytes != nbytes) return bad("size mismatch");
    if (h.flavor > 1 || h.exec_kind > TRRE_EXEC_LAZY) return bad("unknown flavor / executor");
    if (h.n_classes < 1 || h.n_classes > 256) return bad("class count out of range");
    // every section the header names must lie inside the blob (0 = absent for the optional ones)
    auto inside = [&](uint64_t off, uint64_t bytes, bool optional) {
        if (off == 0 && optional) return true;
        return off >= sizeof(BlobHeader) && off + bytes <= nbytes;
    };
    const uint64_t cells = (uint64_t)h.dsa_states * h.n_classes;
    if (!inside(h.off_classmap, 256, false) || !inside(h.off_lut, 256, false) || !inside(h.off_startrow, 256, false) ||
        !inside(h.off_trans, cells * 8, cells == 0) || !inside(h.off_sinfo, (uint64_t)h.dsa_states * 8, h.dsa_states == 0) ||
        !inside(h.off_pool, h.pool_bytes, h.pool_bytes == 0) || !inside(h.off_nft, (uint64_t)h.nft_states * sizeof(NftFlat), h.nft_states == 0) ||
        !inside(h.off_pairlut, 65536 * 2, true) || !inside(h.off_row0, 1024, true) || !inside(h.off_ctrans, cells * 4, true) ||
        !inside(h.off_vent, 1024, true) || !inside(h.off_vpairlen, 65536 * 2, true) || !inside(h.off_vpool, h.vpool_bytes, true) ||
        !inside(h.off_eout, (cells + h.dsa_states) * 4, true))
        return bad("section out of range");
    const uint8_t *b = static_cast<const uint8_t *>(blob);
    // flags must agree with the sections they promise
    if (h.exec_kind == TRRE_EXEC_BYTEMAP && !h.off_pairlut) return bad("byte map without its pair table");
    if ((h.flags & PF_BYTEMAP_VAR) && !(h.off_vent && h.off_vpairlen)) return bad("variable byte map without its tables");
    if ((h.flags & (PF_MATCHABLE | PF_MATCHABLE_VAR)) && !(h.off_row0 && h.off_ctrans)) return bad("match program without compact tables");
    if ((h.exec_kind == TRRE_EXEC_TABLE) && h.dsa_states == 0) return bad("table program without states");
    if (h.swar_n > 4) return bad("byte-map range count");
    for (uint32_t k = 0; k < h.swar_n; ++k) {
        const uint32_t lo = h.swar[k] & 0xFFu, hi = (h.swar[k] >> 8) & 0xFFu;
        if (lo > hi || (lo < 0x80u) != (hi < 0x80u)) return bad("byte-map range");
    }
    if (h.flags & PF_GRAM) {
        uint64_t e = 1;
        if (h.gram_depth < 1 || h.gram_depth > 6) return bad("gram depth");
        for (uint32_t i = 0; i < h.gram_depth; ++i) e *= h.n_classes;
        if (e != h.gram_entries || e > 640u * 1024u) return bad("gram table size");
        if (!inside(h.off_gbits, (e + 31) / 32 * 4, false) || !inside(h.off_gout, e * 4, false) ||
            !inside(h.off_gpool, h.gpool_bytes, h.gpool_bytes == 0) || h.gram_len > 16)
            return bad("gram section out of range");
        if (!(h.gram_flags & GRAM_DIRECT_FLAG)) {
            const uint32_t *go = reinterpret_cast<const uint32_t *>(b + h.off_gout);
            for (uint64_t k = 0; k < e; ++k)
                if ((uint64_t)(go[k] & 0xFFFFFu) + ((go[k] >> 20) & 0xFFu) > h.gpool_bytes && go[k]) return bad("gram output reference");
        }
    }
    if (h.off_vent) {
        const uint32_t *ve = reinterpret_cast<const uint32_t *>(b + h.off_vent);
        for (int c = 0; c < 256; ++c) {
            const uint32_t len = ve[c] & 0xFFu, off = h.vrun_classes ? ((ve[c] >> 16) & 0xFFFu) : (ve[c] >> 16);
            if (len > 1 && (uint64_t)off + len > h.vpool_bytes) return bad("byte-map string out of range");
            if (!h.vrun_classes ? false : (ve[c] >> 28) > h.vrun_classes) return bad("run class out of range");
        }
    }
    if (h.vrun_swar_n && (h.vrun_swar_n != h.vrun_classes || h.vrun_swar_n > 4)) return bad("run ranges");
    for (uint32_t k = 0; k < h.vrun_swar_n; ++k) {
        const uint32_t lo = h.vrun_swar[k] & 0xFFu, hi = (h.vrun_swar[k] >> 8) & 0xFFu;
        if (!lo || lo > hi || (lo < 0x80u) != (hi < 0x80u)) return bad("run range");
    }
    if (h.vrun_classes > 15 || (h.vrun_classes && (!(h.flags & PF_BYTEMAP_VAR) || h.vpool_bytes > 4096))) return bad("run form");
    if (h.lit_words) {
        const bool var = (h.flags & PF_MATCHABLE_VAR) != 0;
        if (h.lit_words > 4 || h.lit_len < 2 || h.lit_len > 4 || !(h.flags & PF_GRAM) || !(h.flags & (PF_MATCHABLE | PF_MATCHABLE_VAR)))
            return bad("literal word form");
        for (uint32_t k = 0; k < h.lit_words * h.lit_len; ++k) {
            const uint32_t lo = h.lit[k] & 0xFFu, hi = (h.lit[k] >> 8) & 0xFFu;
            if (!lo || lo > hi || (lo < 0x80u) != (hi < 0x80u) || (!var && !((h.lit[k] >> 16) & 0xFFu))) return bad("literal word range");
        }
        for (uint32_t w = 0; var && w < h.lit_words; ++w)
            if ((uint64_t)(h.lit_out[w] & 0xFFFFFu) + ((h.lit_out[w] >> 20) & 0xFFu) > h.gpool_bytes) return bad("literal word output");
    }
    // table contents: every class < n_classes, every next state < dsa_states (or DEAD), every string inside the pool
    const uint8_t *cm = b + h.off_classmap;
    for (int c = 0; c < 256; ++c)
        if (cm[c] >= h.n_classes) return bad("class map entry out of range");
    auto ref_ok = [&](uint32_t hi) {
        const uint32_t len = hi >> 24, off = hi & 0xFFFFFFu;
        if (len == 255u) {
            if ((uint64_t)off + 4 > h.pool_bytes) return false;
            uint32_t l2;
            std::memcpy(&l2, b + h.off_pool + off, 4);
            return (uint64_t)off + 4 + l2 <= h.pool_bytes;
        }
        return (uint64_t)off + len <= h.pool_bytes;
    };
    if (cells) {
        const uint64_t *tr = reinterpret_cast<const uint64_t *>(b + h.off_trans);
        for (uint64_t k = 0; k < cells; ++k) {
            const uint32_t nx = (uint32_t)tr[k] & 0xFFFFFFu;
            if (nx != DSA_DEAD && nx >= h.dsa_states) return bad("transition target out of range");
            if (!ref_ok((uint32_t)(tr[k] >> 32))) return bad("transition output out of range");
        }
        const uint64_t *si = reinterpret_cast<const uint64_t *>(b + h.off_sinfo);
        for (uint32_t k = 0; k < h.dsa_states; ++k)
            if (!ref_ok((uint32_t)(si[k] >> 32))) return bad("final output out of range");
        if (h.off_ctrans) {
            const uint32_t *ct = reinterpret_cast<const uint32_t *>(b + h.off_ctrans);
            for (uint64_t k = 0; k < cells; ++k)
                if ((ct[k] & 0xFFFFFFu) != DSA_DEAD && (ct[k] & 0xFFFFFFu) >= h.dsa_states) return bad("compact transition out of range");
        }
        if (h.off_eout) {
            const uint32_t *eo = reinterpret_cast<const uint32_t *>(b + h.off_eout);
            for (uint64_t k = 0; k < cells + h.dsa_states; ++k)
                if ((uint64_t)(eo[k] & 0xFFFFFFu) + (eo[k] >> 24) > h.pool_bytes) return bad("match output out of range");
        }
    }
    if (h.nft_states) {
        if (h.nft_start < -1 || h.nft_start >= (int32_t)h.nft_states) return bad("NFT start out of range");
        const NftFlat *nf = reinterpret_cast<const NftFlat *>(b + h.off_nft);
        for (uint32_t k = 0; k < h.nft_states; ++k)
            if (nf[k].op > N_FINAL || nf[k].a < -1 || nf[k].a >= (int32_t)h.nft_states || nf[k].b < -1 || nf[k].b >= (int32_t)h.nft_states)
                return bad("NFT link out of range");
    } else if (h.exec_kind == TRRE_EXEC_BACKTRACK || h.exec_kind == TRRE_EXEC_LAZY) {
        return bad("executor needs the NFT");
    }
    return true;
}

}  // namespace trre
