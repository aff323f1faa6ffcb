// pattern.cpp - trre expression -> syntax tree -> flat NFT (host, cold path).
//
// Behavioural contract = the reference's parser and NFT builder, quirks included
// (SURVEY.md s4.5 lists them; each is cited below where it is reproduced).
#include "pattern.hpp"

#include "../../include/trre_b200.h"

namespace trre {

namespace {

[[noreturn]] void fail(int code, std::string text) { throw PatternError{code, std::move(text)}; }

// binding strength of stacked operators (trre_nft.c:10-22); brackets and parens bind loosest
int binding(uint8_t op) {
    switch (op) {
        case '|': return 1;
        case '-': return 2;
        case ':': return 3;
        case '.': return 4;
        case '?': case '*': case '+': case 'I': return 5;
        case '\\': return 6;
        default: return -1;
    }
}

class ExprParser {
  public:
    explicit ExprParser(const char *expr) : src_(reinterpret_cast<const uint8_t *>(expr)) {}

    SynTree run() {
        size_t i = 0;
        bool after_operand = false;                       // "state" of trre_nft.c:198
        while (src_[i]) {
            uint8_t c = src_[i];
            bool reread = after_operand ? on_operator_slot(c, i, after_operand)
                                        : on_operand_slot(c, i, after_operand);
            if (stop_) break;
            if (!reread) ++i;
        }
        while (!ops_.empty()) fold();
        if (vals_.empty())                                // assert(operands != opd), trre_nft.c:285
            fail(TRRE_E_ASSERT, "parse: Assertion `operands != opd' failed.\n");
        tree_.root = pop_val();
        return std::move(tree_);
    }

  private:
    const uint8_t *src_;
    SynTree tree_;
    std::vector<uint8_t> ops_;
    std::vector<int32_t> vals_;
    bool stop_ = false;

    int32_t pop_val() {
        // the reference reads below its operand array here; there is nothing to be compatible with
        if (vals_.empty()) fail(TRRE_E_UNSUPPORTED, "error: malformed expression (operand stack underflow)\n");
        int32_t v = vals_.back();
        vals_.pop_back();
        return v;
    }

    // reduce(), trre_nft.c:93-108
    void fold() {
        uint8_t op = ops_.back();
        ops_.pop_back();
        switch (op) {
            case '|': case '.': case ':': case '-': {
                int32_t r = pop_val();
                int32_t l = pop_val();
                vals_.push_back(tree_.add(op, 0, l, r));
                break;
            }
            case '(':
                fail(TRRE_E_PATTERN, "error: unmached parenthesis\n");
            default:
                break;                                    // e.g. '[' is dropped silently
        }
    }

    // reduce_op(), trre_nft.c:111-115
    void push_binary(uint8_t op) {
        while (!ops_.empty() && binding(ops_.back()) >= binding(op)) fold();
        ops_.push_back(op);
    }

    void apply_postfix(uint8_t op, bool lazy) {           // reduce_postfix(), trre_nft.c:76-90
        int32_t body = pop_val();
        vals_.push_back(tree_.add(op, lazy ? 1 : 0, body, -1));
    }

    // expecting an operand (trre_nft.c:201-239); returns true when `c` must be looked at again
    bool on_operand_slot(uint8_t c, size_t &i, bool &after_operand) {
        switch (c) {
            case '(':
                ops_.push_back(c);
                return false;
            case '[':
                ops_.push_back(c);
                i = bracket(i + 1);
                after_operand = true;
                return false;
            case '\\':
                ++i;
                vals_.push_back(tree_.add('c', src_[i]));
                after_operand = true;
                if (!src_[i]) stop_ = true;               // reference steps over the terminator (UB)
                return false;
            case '.': {                                   // any byte == [\x00-\xff], trre_nft.c:215-221
                int32_t lo = tree_.add('c', 0), hi = tree_.add('c', 255);
                vals_.push_back(tree_.add('-', 0, lo, hi));
                after_operand = true;
                return false;
            }
            case ':':                                     // epsilon as implicit left operand (:222-225)
                vals_.push_back(tree_.add('e', c));
                after_operand = true;
                return true;
            case '|': case '*': case '+': case '?': case ')': case '{': case '}':
                if (!ops_.empty() && ops_.back() == ':') {   // epsilon as implicit right operand (:228-231)
                    vals_.push_back(tree_.add('e', c));
                    after_operand = true;
                    return true;
                }
                fail(TRRE_E_PATTERN, std::string("error: unexpected symbol ") + (char)c + "\n");
            default:
                vals_.push_back(tree_.add('c', c));
                after_operand = true;
                return false;
        }
    }

    // expecting a postfix or binary operator (trre_nft.c:240-278)
    bool on_operator_slot(uint8_t c, size_t &i, bool &after_operand) {
        switch (c) {
            case '*': case '+': case '?':
                if (src_[i + 1] == '?') { apply_postfix(c, true); ++i; }
                else apply_postfix(c, false);
                return false;
            case '|':
                push_binary(c);
                after_operand = false;
                return false;
            case ':':
                // a ':' that ends the expression gets its epsilon BEFORE the fold (:255-258), so
                // "xo:" means x:(o then PROD ':') - kept, it changes the tables
                if (src_[i + 1] == 0) vals_.push_back(tree_.add('e', c));
                push_binary(c);
                after_operand = false;
                return false;
            case '{':
                i = braces(i + 1);
                return false;
            case ')':
                while (!ops_.empty() && ops_.back() != '(') fold();
                if (ops_.empty()) fail(TRRE_E_PATTERN, "error: unmached parenthesis");
                ops_.pop_back();
                return false;
            default:                                      // juxtaposition = concatenation
                push_binary('.');
                after_operand = false;
                return true;
        }
    }

    // [...] trre_nft.c:158-193: no escapes, '.' literal, juxtaposition = alternation
    size_t bracket(size_t i) {
        bool after_operand = false;
        while (src_[i]) {
            uint8_t c = src_[i];
            if (!after_operand) {
                if (c == ':' || c == '-' || c == '[' || c == ']')
                    fail(TRRE_E_PATTERN, std::string("error: unexpected symbol in square brackets: ") + (char)c);
                vals_.push_back(tree_.add('c', c));
                after_operand = true;
                ++i;
            } else if (c == ':' || c == '-') {
                push_binary(c);
                after_operand = false;
                ++i;
            } else if (c == ']') {
                while (!ops_.empty() && ops_.back() != '[') fold();
                ops_.pop_back();
                return i;
            } else {
                push_binary('|');
                after_operand = false;                    // and look at c again as an operand
            }
        }
        fail(TRRE_E_PATTERN, "error: unmached square brackets");
    }

    // {m} {m,} {,n} {m,n} and a trailing '?', trre_nft.c:117-156; bounds are 8-bit (:142)
    size_t braces(size_t i) {
        int commas = 0, number = 0, low = 0;
        for (; src_[i]; ++i) {
            uint8_t c = src_[i];
            if (c >= '0' && c <= '9') {
                number = number * 10 + (c - '0');
            } else if (c == ',') {
                low = number;
                number = 0;
                ++commas;
            } else if (c == '}') {
                bool lazy = src_[i + 1] == '?';
                if (lazy) ++i;
                if (commas == 0) low = number;
                else if (commas > 1) fail(TRRE_E_PATTERN, "error: more then one comma in curly brackets\n");
                int32_t bounds = tree_.add((uint8_t)low, (uint8_t)number);
                int32_t body = pop_val();
                vals_.push_back(tree_.add('I', lazy ? 1 : 0, body, bounds));
                return i;
            } else {
                fail(TRRE_E_PATTERN, std::string("error: unexpected symbol in curly brackets: ") + (char)c + "\n");
            }
        }
        fail(TRRE_E_PATTERN, "error: unmached curly brackets");
    }
};

// ---- NFT construction (trre_nft.c:375-504) ---------------------------------------------------

struct Piece { int32_t head, tail; };

class NftBuilder {
  public:
    NftBuilder(SynTree &t, Nft &out) : t_(t), g_(out) {}

    // side: 0 = both tapes (literal = CONS then PROD), 1 = input tape only, 2 = output tape only
    Piece emit(int32_t node, int side) {
        if (node < 0) return {-1, -1};
        const SynNode n = t_.nodes[node];                 // by value: t_.nodes may grow below
        switch (n.kind) {
            case '.': {
                Piece l = emit(n.lhs, side), r = emit(n.rhs, side);
                link(l.tail, r.head);
                return {l.head, r.tail};
            }
            case '|': {                                   // first alternative preferred (:394)
                Piece l = emit(n.lhs, side), r = emit(n.rhs, side);
                int32_t fork = add(N_SPLITNG, l.head, r.head);
                int32_t join = add(N_JOIN);
                link(l.tail, join);
                link(r.tail, join);
                return {fork, join};
            }
            case '*': {
                Piece l = emit(n.lhs, side);
                int32_t fork = add(n.val ? N_SPLITNG : N_SPLIT, -1, l.head);
                link(l.tail, fork);
                return {fork, fork};
            }
            case '?': {
                Piece l = emit(n.lhs, side);
                int32_t join = add(N_JOIN);
                int32_t fork = add(n.val ? N_SPLITNG : N_SPLIT, join, l.head);
                link(l.tail, join);
                return {fork, join};
            }
            case '+': {
                Piece l = emit(n.lhs, side);
                int32_t fork = add(n.val ? N_SPLITNG : N_SPLIT, -1, l.head);
                link(l.tail, fork);
                return {l.head, fork};
            }
            case ':': {
                if (t_.nodes[n.lhs].kind == 'e') return emit(n.rhs, 2);
                if (t_.nodes[n.rhs].kind == 'e') return emit(n.lhs, 1);
                Piece l = emit(n.lhs, 1), r = emit(n.rhs, 2);
                link(l.tail, r.head);
                return {l.head, r.tail};
            }
            case '-':
                return range(n, side);
            case 'I':
                return counted(n, side);
            default:                                      // literal byte, or a stray epsilon node whose
                return literal(n.val, side);              // val is the char that created it (:486-502)
        }
    }

  private:
    SynTree &t_;
    Nft &g_;

    int32_t add(uint8_t op, int32_t a = -1, int32_t b = -1, uint8_t val = 0) {
        g_.st.push_back(NftState{op, val, a, b});
        return (int32_t)g_.st.size() - 1;
    }
    void link(int32_t from, int32_t to) { g_.st[from].a = to; }

    Piece literal(uint8_t byte, int side) {
        if (side == 0) {
            int32_t c = add(N_CONS, -1, -1, byte), p = add(N_PROD, -1, -1, byte);
            link(c, p);
            return {c, p};
        }
        int32_t s = add(side == 1 ? N_CONS : N_PROD, -1, -1, byte);
        return {s, s};
    }

    // [x-y] and [a:x-c:z]: a chain of SPLITNG, lowest byte first (trre_nft.c:425-457)
    Piece range(const SynNode &n, int side) {
        const SynNode lo = t_.nodes[n.lhs], hi = t_.nodes[n.rhs];
        if (lo.kind == 'c' && hi.kind == 'c') {
            int32_t join = add(N_JOIN), chain = -1;
            for (int c = hi.val; c >= (int)lo.val; --c) {
                Piece p = literal((uint8_t)c, side);
                int32_t fork = add(N_SPLITNG, p.head, chain);
                link(p.tail, join);
                chain = fork;
            }
            return {chain, join};                         // chain == -1 for an empty range like [z-a]
        }
        if (lo.kind == ':' && hi.kind == ':') {           // the 4th byte (hi's output) is ignored (:440-442)
            int from_in = t_.nodes[lo.lhs].val, from_out = t_.nodes[lo.rhs].val, to_in = t_.nodes[hi.lhs].val;
            int32_t join = add(N_JOIN), chain = -1;
            for (int k = to_in - from_in; k >= 0; --k) {
                // == nft(':'('c' from_in+k, 'c' from_out+k)): CONS then PROD, bytes wrap mod 256
                int32_t c = add(N_CONS, -1, -1, (uint8_t)(from_in + k));
                int32_t p = add(N_PROD, -1, -1, (uint8_t)(from_out + k));
                link(c, p);
                int32_t fork = add(N_SPLITNG, c, chain);
                link(p, join);
                chain = fork;
            }
            return {chain, join};
        }
        fail(TRRE_E_PATTERN, "error: unexpected range syntax\n");
    }

    // x{lo,hi}: lo copies, then either a greedy star (hi==0, even for the lazy form) or hi-lo
    // optional copies sharing one exit (trre_nft.c:458-485)
    Piece counted(const SynNode &n, int side) {
        const int lo = t_.nodes[n.rhs].kind, hi = t_.nodes[n.rhs].val;
        int32_t head = add(N_JOIN), tail = head;
        for (int k = 0; k < lo; ++k) {
            Piece p = emit(n.lhs, side);
            link(tail, p.head);
            tail = p.tail;
        }
        if (hi == 0) {
            int32_t star = t_.add('*', 0, n.lhs, -1);
            Piece p = emit(star, side);
            link(tail, p.head);
            tail = p.tail;
        } else {
            int32_t exit = add(N_JOIN);
            for (int k = lo; k < hi; ++k) {
                Piece p = emit(n.lhs, side);
                link(tail, add(n.val ? N_SPLITNG : N_SPLIT, exit, p.head));
                tail = p.tail;
            }
            link(tail, exit);
            tail = exit;
        }
        return {head, tail};
    }
};

}  // namespace

SynTree parse_expression(const char *expr) { return ExprParser(expr).run(); }

Nft build_nft(SynTree &tree, bool with_init_join) {
    Nft g;
    g.has_init_join = with_init_join;
    g.st.push_back(NftState{N_FINAL, 0, -1, -1});         // state 0 is the accepting state
    int32_t init = -1;
    if (with_init_join) {
        g.st.push_back(NftState{N_JOIN, 0, -1, -1});
        init = 1;
    }
    NftBuilder b(tree, g);
    Piece p = b.emit(tree.root, 0);
    g.st[p.tail].a = 0;
    if (with_init_join) {
        g.st[init].a = p.head;
        g.start = init;
    } else {
        g.start = p.head;
    }
    // Path compression of JOIN chains: a JOIN is a pure epsilon move with one successor, so every edge
    // may point straight at the first non-JOIN state behind it.  Semantics (both flavours, including the
    // DFS order and the shared-output quirk of trre_dft) are untouched; a K-way alternation otherwise
    // costs K JOIN hops per accepted word (left-nested '|' nodes, one JOIN each).
    {
        const int32_t n = (int32_t)g.st.size();
        std::vector<int32_t> target(n, -2);
        auto resolve = [&](int32_t s) {
            std::vector<int32_t> path;
            int32_t t = s;
            while (t >= 0 && g.st[t].op == N_JOIN && g.st[t].a >= 0 && target[t] == -2 && t != init) {
                path.push_back(t);
                t = g.st[t].a;
                if (path.size() > (size_t)n) break;       // cannot happen: cycles always contain a SPLIT
            }
            int32_t dst = (t >= 0 && target[t] >= -1 && g.st[t].op == N_JOIN && t != init) ? target[t] : t;
            for (int32_t q : path) target[q] = dst;
            return dst;
        };
        for (int32_t s = 0; s < n; ++s) {
            if (g.st[s].a >= 0) g.st[s].a = resolve(g.st[s].a);
            if (g.st[s].b >= 0) g.st[s].b = resolve(g.st[s].b);
        }
        if (!with_init_join && g.start >= 0) g.start = resolve(g.start);
    }
    return g;
}

}  // namespace trre
