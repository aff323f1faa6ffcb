/*
 * trre_oracle.c - CPU ORACLE for the trre scan path.  TEST INFRASTRUCTURE ONLY - see trre_oracle.h.
 *
 * Restates, in plain C over index-addressed arrays and a memory buffer instead of stdout, what
 * the reference (c0stya/trre @ c474a68) computes.  Citations are file:line into the reference.
 * Not shipped, not linked into the product; the product fails loudly without its CUDA library.
 *
 * What is restated (and why each quirk is kept - the tables and therefore stdout depend on it):
 *   precedence table                     trre_nft.c:10-22
 *   two-stack operator-precedence parser trre_nft.c:76-288   (shared verbatim by trre_dft.c:77-288)
 *   Thompson-style NFT construction      trre_nft.c:375-511 / trre_dft.c:376-523 (extra init JOIN)
 *   backtracking matcher                 trre_nft.c:593-657  (stack cap trre_nft.c:36,548-556)
 *   scan / match drivers (getline model) trre_nft.c:775-797 / trre_dft.c:1272-1293
 *   lazy determinizer + DFT walk         trre_dft.c:874-923, 988-1010, 1028-1049, 1110-1196
 */
#include "trre_oracle.h"

#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* run context: arena, error exit, output sink                                                 */
/* ------------------------------------------------------------------------------------------ */

typedef struct Block { struct Block *prev; } Block;

typedef struct {
    jmp_buf   bail;
    int       rc;
    char     *err;
    size_t    errcap;
    Block    *blocks;           /* every allocation, freed together */
    uint8_t  *out;              /* stands in for stdout */
    size_t    out_n, out_cap;
} Run;

static size_t g_last_nstates, g_last_dstates;

static void die(Run *R, int rc, const char *fmt, ...) {
    va_list ap;
    if (R->err && R->errcap) {
        va_start(ap, fmt);
        vsnprintf(R->err, R->errcap, fmt, ap);
        va_end(ap);
    }
    R->rc = rc;
    longjmp(R->bail, 1);
}

static void *take(Run *R, size_t n) {
    Block *b = (Block *)malloc(sizeof(Block) + (n ? n : 1));
    if (!b) die(R, 1, "error: node memory allocation failed\n");
    b->prev = R->blocks;
    R->blocks = b;
    return (void *)(b + 1);
}

static void release_to(Run *R, Block *mark) {
    while (R->blocks != mark) { Block *p = R->blocks->prev; free(R->blocks); R->blocks = p; }
}
static void release_all(Run *R) { release_to(R, NULL); }

static void emit(Run *R, const uint8_t *p, size_t n) {
    if (!n) return;
    if (R->out_n + n > R->out_cap) {
        size_t cap = R->out_cap ? R->out_cap : 4096;
        while (cap < R->out_n + n) cap *= 2;
        uint8_t *q = (uint8_t *)realloc(R->out, cap);
        if (!q) die(R, 1, "error: memory reallocation failed for output\n");
        R->out = q; R->out_cap = cap;
    }
    memcpy(R->out + R->out_n, p, n);
    R->out_n += n;
}
static void emit1(Run *R, uint8_t c) { emit(R, &c, 1); }

/* ------------------------------------------------------------------------------------------ */
/* AST + parser  (trre_nft.c:10-288)                                                           */
/* ------------------------------------------------------------------------------------------ */

typedef struct { uint8_t type, val; int l, r; } Ast;       /* trre_nft.c:24-29; children by index */

typedef struct {
    Run  *R;
    Ast  *nodes; int n_nodes, cap_nodes;
    /* the reference keeps two fixed 1024-entry stacks (trre_nft.c:38-39); we grow instead */
    uint8_t *ops; int n_ops, cap_ops;
    int     *opd; int n_opd, cap_opd;
} Parser;

static int precedence(int c) {                                /* trre_nft.c:10-22 */
    if (c == '|') return 1;
    if (c == '-') return 2;
    if (c == ':') return 3;
    if (c == '.') return 4;
    if (c == '?' || c == '*' || c == '+' || c == 'I') return 5;
    if (c == '\\') return 6;
    return -1;
}

static int ast_new(Parser *P, uint8_t type, uint8_t val, int l, int r) {   /* trre_nft.c:48-66 */
    if (P->n_nodes == P->cap_nodes) {
        int cap = P->cap_nodes ? P->cap_nodes * 2 : 64;
        Ast *q = (Ast *)take(P->R, (size_t)cap * sizeof(Ast));
        if (P->n_nodes) memcpy(q, P->nodes, (size_t)P->n_nodes * sizeof(Ast));
        P->nodes = q; P->cap_nodes = cap;
    }
    Ast *a = &P->nodes[P->n_nodes];
    a->type = type; a->val = val; a->l = l; a->r = r;
    return P->n_nodes++;
}

static void op_push(Parser *P, uint8_t c) {
    if (P->n_ops == P->cap_ops) {
        int cap = P->cap_ops ? P->cap_ops * 2 : 64;
        uint8_t *q = (uint8_t *)take(P->R, (size_t)cap);
        if (P->n_ops) memcpy(q, P->ops, (size_t)P->n_ops);
        P->ops = q; P->cap_ops = cap;
    }
    P->ops[P->n_ops++] = c;
}
static void opd_push(Parser *P, int node) {
    if (P->n_opd == P->cap_opd) {
        int cap = P->cap_opd ? P->cap_opd * 2 : 64;
        int *q = (int *)take(P->R, (size_t)cap * sizeof(int));
        if (P->n_opd) memcpy(q, P->opd, (size_t)P->n_opd * sizeof(int));
        P->opd = q; P->cap_opd = cap;
    }
    P->opd[P->n_opd++] = node;
}
static int opd_pop(Parser *P) {
    /* the reference would read below its array here (undefined behaviour); we refuse instead */
    if (P->n_opd == 0) die(P->R, 2, "error: malformed expression (operand stack underflow)\n");
    return P->opd[--P->n_opd];
}

static void fold_top(Parser *P) {                             /* reduce(), trre_nft.c:93-108 */
    uint8_t op = P->ops[--P->n_ops];
    if (op == '|' || op == '.' || op == ':' || op == '-') {
        int r = opd_pop(P), l = opd_pop(P);
        opd_push(P, ast_new(P, op, 0, l, r));
    } else if (op == '(') {
        die(P->R, 1, "error: unmached parenthesis\n");
    }
    /* anything else (e.g. a stray '[') is silently dropped, as in the reference */
}

static void shift_op(Parser *P, uint8_t op) {                 /* reduce_op(), trre_nft.c:111-115 */
    while (P->n_ops && precedence(P->ops[P->n_ops - 1]) >= precedence(op)) fold_top(P);
    op_push(P, op);
}

static void postfix(Parser *P, uint8_t op, int ng) {          /* reduce_postfix(), trre_nft.c:76-90 */
    int l = opd_pop(P);
    opd_push(P, ast_new(P, op, (uint8_t)ng, l, -1));
}

/* {m}, {m,}, {,n}, {m,n} with optional '?' - trre_nft.c:117-156.  Returns index of the '}'
 * (or of the '?' after it).  Counts are stored in the 8-bit type/val of the right child (:142). */
static size_t parse_counts(Parser *P, const uint8_t *e, size_t i) {
    int commas = 0, ng = 0, count = 0, lo = 0;
    for (; e[i]; i++) {
        uint8_t c = e[i];
        if (c >= '0' && c <= '9') {
            count = count * 10 + (c - '0');
        } else if (c == ',') {
            lo = count; count = 0; commas++;
        } else if (c == '}') {
            if (e[i + 1] == '?') { ng = 1; i++; }
            if (commas == 0) lo = count;
            else if (commas > 1) die(P->R, 1, "error: more then one comma in curly brackets\n");
            int bounds = ast_new(P, (uint8_t)lo, (uint8_t)count, -1, -1);
            int it = ast_new(P, 'I', (uint8_t)ng, opd_pop(P), bounds);
            opd_push(P, it);
            return i;
        } else {
            die(P->R, 1, "error: unexpected symbol in curly brackets: %c\n", c);
        }
    }
    die(P->R, 1, "error: unmached curly brackets");
    return i;
}

/* [...] - trre_nft.c:158-193.  No escapes inside; juxtaposition is alternation. */
static size_t parse_class(Parser *P, const uint8_t *e, size_t i) {
    int want_operator = 0;
    while (e[i]) {
        uint8_t c = e[i];
        if (!want_operator) {
            if (c == ':' || c == '-' || c == '[' || c == ']')
                die(P->R, 1, "error: unexpected symbol in square brackets: %c", c);
            opd_push(P, ast_new(P, 'c', c, -1, -1));
            want_operator = 1;
        } else if (c == ':' || c == '-') {
            shift_op(P, c);
            want_operator = 0;
        } else if (c == ']') {
            while (P->n_ops && P->ops[P->n_ops - 1] != '[') fold_top(P);
            P->n_ops--;                                       /* drop the '[' */
            return i;
        } else {
            shift_op(P, '|');
            want_operator = 0;
            continue;                                         /* re-read c as an operand */
        }
        i++;
    }
    die(P->R, 1, "error: unmached square brackets");
    return i;
}

static int parse_expr(Parser *P, const char *expr) {          /* parse(), trre_nft.c:196-288 */
    const uint8_t *e = (const uint8_t *)expr;
    size_t i = 0;
    int want_operator = 0;
    while (e[i]) {
        uint8_t c = e[i];
        if (!want_operator) {
            if (c == '(') {
                op_push(P, c);
            } else if (c == '[') {
                op_push(P, c);
                i = parse_class(P, e, i + 1);
                want_operator = 1;
            } else if (c == '\\') {
                i++;
                opd_push(P, ast_new(P, 'c', e[i], -1, -1));
                want_operator = 1;
                if (!e[i]) break;          /* reference walks past the terminator here (UB) */
            } else if (c == '.') {                            /* any byte = range 0..255 (:215-221) */
                int lo = ast_new(P, 'c', 0, -1, -1), hi = ast_new(P, 'c', 255, -1, -1);
                opd_push(P, ast_new(P, '-', 0, lo, hi));
                want_operator = 1;
            } else if (c == ':') {                            /* implicit epsilon on the left (:222-225) */
                opd_push(P, ast_new(P, 'e', c, -1, -1));
                want_operator = 1;
                continue;
            } else if (c == '|' || c == '*' || c == '+' || c == '?' ||
                       c == ')' || c == '{' || c == '}') {
                if (P->n_ops && P->ops[P->n_ops - 1] == ':') {   /* implicit epsilon on the right */
                    opd_push(P, ast_new(P, 'e', c, -1, -1));
                    want_operator = 1;
                    continue;
                }
                die(P->R, 1, "error: unexpected symbol %c\n", c);
            } else {
                opd_push(P, ast_new(P, 'c', c, -1, -1));
                want_operator = 1;
            }
        } else {
            if (c == '*' || c == '+' || c == '?') {
                if (e[i + 1] == '?') { postfix(P, c, 1); i++; }
                else postfix(P, c, 0);
            } else if (c == '|') {
                shift_op(P, c);
                want_operator = 0;
            } else if (c == ':') {
                /* trailing ':' pushes its epsilon BEFORE folding (:255-258): "xo:" == x:(o + PROD ':') */
                if (e[i + 1] == 0) opd_push(P, ast_new(P, 'e', c, -1, -1));
                shift_op(P, c);
                want_operator = 0;
            } else if (c == '{') {
                i = parse_counts(P, e, i + 1);
            } else if (c == ')') {
                while (P->n_ops && P->ops[P->n_ops - 1] != '(') fold_top(P);
                if (!P->n_ops) die(P->R, 1, "error: unmached parenthesis");
                P->n_ops--;
            } else {
                shift_op(P, '.');                             /* implicit concatenation */
                want_operator = 0;
                continue;
            }
        }
        i++;
    }
    while (P->n_ops) fold_top(P);
    if (P->n_opd == 0) die(P->R, 134, "Assertion `operands != opd' failed.\n");   /* :285 */
    return opd_pop(P);
}

/* ------------------------------------------------------------------------------------------ */
/* NFT  (trre_nft.c:323-511, trre_dft.c:322-523)                                               */
/* ------------------------------------------------------------------------------------------ */

enum { PROD, CONS, SPLIT, SPLITNG, JOIN, FINAL };            /* trre_nft.c:323-330 */

typedef struct { uint8_t type, val, visited; int a, b; } Nst;
typedef struct { int head, tail; } Frag;

typedef struct {
    Run *R; Parser *P;
    Nst *st; int n, cap;
} Nft;

static int nst_new(Nft *N, int type, int a, int b) {          /* create_nstate, trre_nft.c:343-357 */
    if (N->n == N->cap) {
        int cap = N->cap ? N->cap * 2 : 256;
        Nst *q = (Nst *)take(N->R, (size_t)cap * sizeof(Nst));
        if (N->n) memcpy(q, N->st, (size_t)N->n * sizeof(Nst));
        N->st = q; N->cap = cap;
    }
    Nst *s = &N->st[N->n];
    s->type = (uint8_t)type; s->val = 0; s->visited = 0; s->a = a; s->b = b;
    return N->n++;
}

static Frag frag(int h, int t) { Frag f; f.head = h; f.tail = t; return f; }

/* nft(), trre_nft.c:375-504.  mode 0: literal = CONS then PROD; 1: CONS only; 2: PROD only. */
static Frag build(Nft *N, int node, int mode) {
    Ast *A = N->P->nodes;
    if (node < 0) return frag(-1, -1);
    Ast n = A[node];
    Frag l, r;
    int split, join, prev;
    switch (n.type) {
    case '.':
        l = build(N, n.l, mode); r = build(N, n.r, mode);
        N->st[l.tail].a = r.head;
        return frag(l.head, r.tail);
    case '|':                                                 /* first alternative preferred (:394) */
        l = build(N, n.l, mode); r = build(N, n.r, mode);
        split = nst_new(N, SPLITNG, l.head, r.head);
        join = nst_new(N, JOIN, -1, -1);
        N->st[l.tail].a = join; N->st[r.tail].a = join;
        return frag(split, join);
    case '*':
        l = build(N, n.l, mode);
        split = nst_new(N, n.val ? SPLITNG : SPLIT, -1, l.head);
        N->st[l.tail].a = split;
        return frag(split, split);
    case '?':
        l = build(N, n.l, mode);
        join = nst_new(N, JOIN, -1, -1);
        split = nst_new(N, n.val ? SPLITNG : SPLIT, join, l.head);
        N->st[l.tail].a = join;
        return frag(split, join);
    case '+':
        l = build(N, n.l, mode);
        split = nst_new(N, n.val ? SPLITNG : SPLIT, -1, l.head);
        N->st[l.tail].a = split;
        return frag(l.head, split);
    case ':':
        if (A[n.l].type == 'e') return build(N, n.r, 2);
        if (A[n.r].type == 'e') return build(N, n.l, 1);
        l = build(N, n.l, 1); r = build(N, n.r, 2);
        N->st[l.tail].a = r.head;
        return frag(l.head, r.tail);
    case '-':
        if (A[n.l].type == 'c' && A[n.r].type == 'c') {       /* [x-y]: SPLITNG chain, low byte first (:426-435) */
            join = nst_new(N, JOIN, -1, -1);
            prev = -1;
            for (int c = A[n.r].val; c >= (int)A[n.l].val; c--) {
                int lit = ast_new(N->P, 'c', (uint8_t)c, -1, -1);
                A = N->P->nodes;
                l = build(N, lit, mode);
                split = nst_new(N, SPLITNG, l.head, prev);
                N->st[l.tail].a = join;
                prev = split;
            }
            return frag(prev, join);
        } else if (A[n.l].type == ':' && A[n.r].type == ':') { /* [a:x-c:z]; 4th byte ignored (:436-452) */
            int llv = A[A[n.l].l].val, lrv = A[A[n.l].r].val, rlv = A[A[n.r].l].val;
            join = nst_new(N, JOIN, -1, -1);
            prev = -1;
            for (int c = rlv - llv; c >= 0; c--) {
                int x = ast_new(N->P, 'c', (uint8_t)(llv + c), -1, -1);
                int y = ast_new(N->P, 'c', (uint8_t)(lrv + c), -1, -1);
                int pr = ast_new(N->P, ':', 0, x, y);
                A = N->P->nodes;
                l = build(N, pr, mode);
                split = nst_new(N, SPLITNG, l.head, prev);
                N->st[l.tail].a = join;
                prev = split;
            }
            return frag(prev, join);
        }
        die(N->R, 1, "error: unexpected range syntax\n");
        return frag(-1, -1);
    case 'I': {                                               /* {lo,hi}: :458-485 */
        int lo = A[n.r].type, hi = A[n.r].val;
        int head = nst_new(N, JOIN, -1, -1), tail = head;
        for (int k = 0; k < lo; k++) {
            l = build(N, n.l, mode);
            N->st[tail].a = l.head; tail = l.tail;
        }
        if (hi == 0) {                                        /* open upper bound: always a GREEDY star (:469-472) */
            int star = ast_new(N->P, '*', 0, n.l, -1);
            l = build(N, star, mode);
            N->st[tail].a = l.head; tail = l.tail;
        } else {
            int fin = nst_new(N, JOIN, -1, -1);
            for (int k = lo; k < hi; k++) {
                l = build(N, n.l, mode);
                N->st[tail].a = nst_new(N, n.val ? SPLITNG : SPLIT, fin, l.head);
                tail = l.tail;
            }
            N->st[tail].a = fin; tail = fin;
        }
        return frag(head, tail);
    }
    default:                                                  /* literal, or a stray epsilon node (:486-502) */
        if (mode == 0) {
            int cs = nst_new(N, CONS, -1, -1), ps = nst_new(N, PROD, -1, -1);
            N->st[cs].val = n.val; N->st[ps].val = n.val;
            N->st[cs].a = ps;
            return frag(cs, ps);
        } else {
            int s = nst_new(N, mode == 1 ? CONS : PROD, -1, -1);
            N->st[s].val = n.val;
            return frag(s, s);
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* getline record model shared by both drivers (trre_nft.c:776-777 / trre_dft.c:1273-1274)      */
/* ------------------------------------------------------------------------------------------ */

/* Next record starting at *pos: returns 0 at EOF; otherwise sets [*beg,*end) to the C-string
 * payload: the record minus its LAST byte (whatever it is), cut at the first NUL. */
static int next_record(const uint8_t *in, size_t n, size_t *pos, size_t *beg, size_t *end) {
    if (*pos >= n) return 0;
    size_t b = *pos, e = b;
    while (e < n && in[e] != '\n') e++;
    size_t reclen = (e < n) ? e - b + 1 : e - b;              /* includes the '\n' if present */
    *pos = b + reclen;
    size_t payload = reclen - 1;                              /* line[read-1] = '\0' */
    const uint8_t *z = (const uint8_t *)memchr(in + b, 0, payload);
    *beg = b;
    *end = z ? (size_t)(z - in) : b + payload;
    return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* NFT flavour: backtracking matcher (trre_nft.c:593-657)                                      */
/* ------------------------------------------------------------------------------------------ */

typedef struct { int s; size_t i, o; } Saved;

typedef struct {
    Run *R; Nft *N;
    Saved *stk; size_t n_stk, cap_stk;
    uint8_t *obuf; size_t ocap;
} Backtracker;

/* capacity doubles from 32 and may not exceed 100000 (trre_nft.c:35-36, 548-556): the last
 * capacity reached is 65536, the push that needs more fails. */
static void save(Backtracker *B, int s, size_t i, size_t o) {
    if (B->n_stk == B->cap_stk) {
        if (B->cap_stk * 2 > 100000) die(B->R, 1, "error: stack max capacity reached\n");
        size_t cap = B->cap_stk * 2;
        Saved *q = (Saved *)take(B->R, cap * sizeof(Saved));
        memcpy(q, B->stk, B->n_stk * sizeof(Saved));
        B->stk = q; B->cap_stk = cap;
    }
    B->stk[B->n_stk].s = s; B->stk[B->n_stk].i = i; B->stk[B->n_stk].o = o;
    B->n_stk++;
}

/* fputs(output): bytes up to the first NUL (trre_nft.c:637-638, 644-645) */
static void put_cstr(Run *R, const uint8_t *p, size_t n) {
    const uint8_t *z = (const uint8_t *)memchr(p, 0, n);
    emit(R, p, z ? (size_t)(z - p) : n);
}

/* input = [p, e) stands for the NUL-terminated tail of the line */
static long backtrack(Backtracker *B, int start, const uint8_t *p, const uint8_t *e, int mode, int all) {
    Nst *st = B->N->st;
    size_t i = 0, o = 0, len = (size_t)(e - p);
    int s = start;
    unsigned long long steps = 0;
    B->n_stk = 0;
    while (B->n_stk || s >= 0) {
        /* a producing epsilon cycle never ends in the reference (README.md:157-166); refuse instead */
        if (++steps > 200000000ull) die(B->R, 2, "error: attempt does not terminate in the reference\n");
        if (s < 0) {
            B->n_stk--;
            s = B->stk[B->n_stk].s; i = B->stk[B->n_stk].i; o = B->stk[B->n_stk].o;
            if (s < 0) continue;
        }
        if (o + 2 > B->ocap) {
            size_t cap = B->ocap * 2;
            uint8_t *q = (uint8_t *)take(B->R, cap);
            memcpy(q, B->obuf, o);
            B->obuf = q; B->ocap = cap;
        }
        switch (st[s].type) {
        case CONS:
            if (i < len && st[s].val == p[i]) { i++; s = st[s].a; } else s = -1;
            break;
        case PROD:    B->obuf[o++] = st[s].val; s = st[s].a; break;
        case SPLIT:   save(B, st[s].a, i, o); s = st[s].b; break;      /* greedy: body first */
        case SPLITNG: save(B, st[s].b, i, o); s = st[s].a; break;
        case JOIN:    s = st[s].a; break;
        case FINAL:
            if (mode == TRRE_ORACLE_MATCH) {
                if (i == len) {
                    put_cstr(B->R, B->obuf, o);
                    emit1(B->R, '\n');
                    if (!all) return (long)i;
                }
            } else {
                put_cstr(B->R, B->obuf, o);
                if (!all) return (long)i;
            }
            s = -1;
            break;
        }
    }
    return -1;
}

static void run_nft(Run *R, Nft *N, int start, int mode, int all, const uint8_t *in, size_t n) {
    Backtracker B;
    B.R = R; B.N = N;
    B.cap_stk = 32; B.n_stk = 0; B.stk = (Saved *)take(R, B.cap_stk * sizeof(Saved));
    B.ocap = 32; B.obuf = (uint8_t *)take(R, B.ocap);
    size_t pos = 0, b, e;
    while (next_record(in, n, &pos, &b, &e)) {
        if (mode == TRRE_ORACLE_SCAN) {                       /* trre_nft.c:775-790 */
            size_t ch = b;
            while (ch < e) {
                long k = backtrack(&B, start, in + ch, in + e, mode, all);
                if (k > 0) ch += (size_t)k; else emit1(R, in[ch++]);
            }
            backtrack(&B, start, in + e, in + e, mode, all);  /* one more attempt on the empty tail (:788) */
            emit1(R, '\n');
        } else {                                              /* trre_nft.c:791-797 */
            backtrack(&B, start, in + b, in + e, mode, all);
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* DFT flavour: lazy determinizer + walk (trre_dft.c:720-1196)                                 */
/* ------------------------------------------------------------------------------------------ */

typedef struct { uint8_t *p; size_t n, cap; } Str;            /* stands in for the linked-list str (:720-817) */
typedef struct { int s; Str *w; } Item;                       /* slitem (:827-831) */
typedef struct { Item *it; size_t n, cap; } List;             /* slist  (:821-825) */

#define DEAD_MARK ((Str *)1)                                  /* "(struct str*)1" sentinel (:1140) */

typedef struct {
    List *states; Str *final_out; int final;                  /* final: -1 unexplored, 0, 1 (:926-932) */
    Str *out[256]; int next[256];
} Dst;

typedef struct {
    Run *R; Nft *N;
    Dst **ds; size_t n_ds, cap_ds;
    int *bucket; size_t n_bucket;                             /* content hash instead of the pointer-keyed BST (:1052-1100) */
    int *chain; size_t cap_chain;
    Str *acc;                                                 /* per-attempt output accumulator, reused */
    int depth;
    unsigned long long work;
} Dft;

static Str *str_new(Run *R) { Str *s = (Str *)take(R, sizeof(Str)); s->p = NULL; s->n = s->cap = 0; return s; }
static void str_push(Run *R, Str *s, uint8_t c) {
    if (s->n == s->cap) {
        size_t cap = s->cap ? s->cap * 2 : 8;
        uint8_t *q = (uint8_t *)take(R, cap);
        if (s->n) memcpy(q, s->p, s->n);
        s->p = q; s->cap = cap;
    }
    s->p[s->n++] = c;
}
static Str *str_dup(Run *R, const Str *a) {
    Str *s = str_new(R);
    for (size_t k = 0; k < a->n; k++) str_push(R, s, a->p[k]);
    return s;
}
static void str_cat(Run *R, Str *d, const Str *a) { for (size_t k = 0; k < a->n; k++) str_push(R, d, a->p[k]); }

static List *list_new(Run *R) { List *l = (List *)take(R, sizeof(List)); l->it = NULL; l->n = l->cap = 0; return l; }
static void list_push(Run *R, List *l, int s, Str *w) {
    if (l->n == l->cap) {
        size_t cap = l->cap ? l->cap * 2 : 8;
        Item *q = (Item *)take(R, cap * sizeof(Item));
        if (l->n) memcpy(q, l->it, l->n * sizeof(Item));
        l->it = q; l->cap = cap;
    }
    l->it[l->n].s = s; l->it[l->n].w = w; l->n++;
}

/* nft_step_(), trre_dft.c:874-907.  NB the aliasing the reference has: `o` is ONE mutable string
 * handed down the preferred branch (PROD appends in place, :891) and only afterwards copied for
 * the other branch (:881, :885), so bytes produced on a dead preferred path leak into the copy. */
static void closure_(Dft *D, int s, Str *o, int c, List *sl);
static void closure(Dft *D, int s, Str *o, int c, List *sl) {
    /* an epsilon cycle recurses until the reference's C stack overflows (SIGSEGV); refuse instead */
    /* generators such as a:(....) enumerate every output string; the reference runs out of memory */
    if (++D->work > 4000000ull) die(D->R, 2, "error: closure does not finish in the reference (output enumeration)\n");
    if (++D->depth > 30000) die(D->R, 2, "error: epsilon cycle (the reference overflows its C stack here)\n");
    closure_(D, s, o, c, sl);
    D->depth--;
}
static void closure_(Dft *D, int s, Str *o, int c, List *sl) {
    Nst *st = D->N->st;
    if (s < 0) return;
    switch (st[s].type) {
    case SPLIT:   closure(D, st[s].b, o, c, sl); closure(D, st[s].a, str_dup(D->R, o), c, sl); break;
    case SPLITNG: closure(D, st[s].a, o, c, sl); closure(D, st[s].b, str_dup(D->R, o), c, sl); break;
    case JOIN:    closure(D, st[s].a, o, c, sl); break;
    case PROD:    str_push(D->R, o, st[s].val); closure(D, st[s].a, o, c, sl); break;
    case CONS:
        if (c == st[s].val && !st[s].visited) { st[s].visited = 1; list_push(D->R, sl, s, o); }
        break;
    case FINAL:
        if (c == 0 && !st[s].visited) { list_push(D->R, sl, s, o); st[s].visited = 1; }
        break;
    }
}

static List *step(Dft *D, const List *from, int c) {          /* nft_step(), trre_dft.c:910-923 */
    List *sl = list_new(D->R);
    D->work = 0;                                              /* budget is per step */
    for (size_t k = 0; k < from->n; k++)
        closure(D, D->N->st[from->it[k].s].a, str_dup(D->R, from->it[k].w), c, sl);
    for (size_t k = 0; k < sl->n; k++) D->N->st[sl->it[k].s].visited = 0;
    return sl;
}

static void strip_common_prefix(Dft *D, List *sl, Str *prefix) {   /* truncate_lcp(), trre_dft.c:988-1010 */
    for (;;) {
        if (sl->it[0].w->n == 0) return;
        uint8_t ch = sl->it[0].w->p[0];
        for (size_t k = 0; k < sl->n; k++)
            if (sl->it[k].w->n == 0 || sl->it[k].w->p[0] != ch) return;
        str_push(D->R, prefix, ch);
        for (size_t k = 0; k < sl->n; k++) { sl->it[k].w->p++; sl->it[k].w->n--; sl->it[k].w->cap--; }
    }
}

static int same_list(const List *a, const List *b) {          /* list_cmp()==0, trre_dft.c:1028-1049 */
    if (a->n != b->n) return 0;
    for (size_t k = 0; k < a->n; k++) {
        if (a->it[k].s != b->it[k].s) return 0;
        if (a->it[k].w->n != b->it[k].w->n) return 0;
        if (a->it[k].w->n && memcmp(a->it[k].w->p, b->it[k].w->p, a->it[k].w->n)) return 0;
    }
    return 1;
}
static uint64_t list_hash(const List *a) {
    uint64_t h = 1469598103934665603ull;
    for (size_t k = 0; k < a->n; k++) {
        h = (h ^ (uint64_t)a->it[k].s) * 1099511628211ull;
        for (size_t j = 0; j < a->it[k].w->n; j++) h = (h ^ a->it[k].w->p[j]) * 1099511628211ull;
        h = (h ^ 0xffu) * 1099511628211ull;
    }
    return h;
}

static int dst_new(Dft *D, List *states) {                    /* dstate_create(), trre_dft.c:934-942 */
    if (D->n_ds == D->cap_ds) {
        size_t cap = D->cap_ds ? D->cap_ds * 2 : 64;
        Dst **q = (Dst **)take(D->R, cap * sizeof(Dst *));
        int *c2 = (int *)take(D->R, cap * sizeof(int));
        if (D->n_ds) { memcpy(q, D->ds, D->n_ds * sizeof(Dst *)); memcpy(c2, D->chain, D->n_ds * sizeof(int)); }
        D->ds = q; D->chain = c2; D->cap_ds = cap;
    }
    Dst *d = (Dst *)take(D->R, sizeof(Dst));
    d->states = states; d->final_out = NULL; d->final = -1;
    for (int c = 0; c < 256; c++) { d->out[c] = NULL; d->next[c] = -1; }
    D->ds[D->n_ds] = d;
    size_t h = (size_t)(list_hash(states) % D->n_bucket);
    D->chain[D->n_ds] = D->bucket[h];
    D->bucket[h] = (int)D->n_ds;
    return (int)D->n_ds++;
}
static int dst_find(Dft *D, const List *sl) {                 /* bt_lookup(), trre_dft.c:1074-1087 */
    size_t h = (size_t)(list_hash(sl) % D->n_bucket);
    for (int k = D->bucket[h]; k >= 0; k = D->chain[k])
        if (same_list(sl, D->ds[k]->states)) return k;
    return -1;
}

/* infer_dft() in SCAN mode, trre_dft.c:1110-1196.  Input = [p,e). */
static long walk_dft(Dft *D, int dstart, const uint8_t *p, const uint8_t *e) {
    Run *R = D->R;
    int ds = dstart;
    Str *acc = D->acc;                                        /* `out = str_create()` per call in the reference (:1112) */
    long i = 0;
    acc->n = 0;
    for (const uint8_t *c = p; c < e; c++, i++) {
        Dst *d = D->ds[ds];
        if (d->final == 1) {                                  /* stop at the FIRST final state (:1120-1125) */
            emit(R, acc->p, acc->n); emit(R, d->final_out->p, d->final_out->n);
            return i;
        }
        if (d->next[*c] >= 0) {
            str_cat(R, acc, d->out[*c]);
            ds = d->next[*c];
        } else if (d->out[*c] != NULL) {                      /* explored, dead (:1132-1134) */
            break;
        } else {                                              /* explore (:1135-1175) */
            /* build into a scratch region, keep copies of what survives, drop the rest */
                    Block *mark;
            {   /* make sure acc cannot grow inside the scratch region */
                size_t need = acc->n + 4096;
                if (acc->cap < need) { uint8_t *q = (uint8_t *)take(R, need); if (acc->n) memcpy(q, acc->p, acc->n); acc->p = q; acc->cap = need; }
            }
            mark = R->blocks;
            List *sl = step(D, d->states, *c);
            if (sl->n == 0) { release_to(R, mark); d->out[*c] = DEAD_MARK; break; }
            Str *prefix = str_new(R);
            strip_common_prefix(D, sl, prefix);
            int nx = dst_find(D, sl);
            /* copy survivors out of the scratch region */
            size_t plen = prefix->n;
            uint8_t *ptmp = (uint8_t *)malloc(plen ? plen : 1);
            if (plen) memcpy(ptmp, prefix->p, plen);
            List *keep = NULL;
            if (nx < 0) {
                /* serialise the list outside the arena */
                size_t tot = 0;
                for (size_t k = 0; k < sl->n; k++) tot += sl->it[k].w->n;
                uint8_t *bytes = (uint8_t *)malloc(tot ? tot : 1);
                size_t *lens = (size_t *)malloc(sl->n * sizeof(size_t));
                int *sts = (int *)malloc(sl->n * sizeof(int));
                size_t off = 0, cnt = sl->n;
                for (size_t k = 0; k < cnt; k++) {
                    lens[k] = sl->it[k].w->n; sts[k] = sl->it[k].s;
                    if (lens[k]) memcpy(bytes + off, sl->it[k].w->p, lens[k]);
                    off += lens[k];
                }
                release_to(R, mark);
                keep = list_new(R);
                off = 0;
                for (size_t k = 0; k < cnt; k++) {
                    Str *w = str_new(R);
                    for (size_t j = 0; j < lens[k]; j++) str_push(R, w, bytes[off + j]);
                    off += lens[k];
                    list_push(R, keep, sts[k], w);
                }
                free(bytes); free(lens); free(sts);
                nx = dst_new(D, keep);
            } else {
                release_to(R, mark);
            }
            prefix = str_new(R);
            for (size_t j = 0; j < plen; j++) str_push(R, prefix, ptmp[j]);
            free(ptmp);
            d = D->ds[ds];
            d->next[*c] = nx; d->out[*c] = prefix;
            if (acc->n + prefix->n > acc->cap) str_cat(R, acc, prefix);      /* rare: very long prefix */
            else { if (prefix->n) memcpy(acc->p + acc->n, prefix->p, prefix->n); acc->n += prefix->n; }
            ds = nx;
            Dst *t = D->ds[ds];
            if (t->final < 0) {                               /* finality probe with c = '\0' (:1164-1174); a CONS(0) counts */
                Block *m2 = R->blocks;
                List *fl = step(D, t->states, 0);
                if (fl->n) {
                    size_t fn = fl->it[0].w->n;
                    uint8_t *ft = (uint8_t *)malloc(fn ? fn : 1);
                    if (fn) memcpy(ft, fl->it[0].w->p, fn);
                    release_to(R, m2);
                    t->final = 1; t->final_out = str_new(R);
                    for (size_t j = 0; j < fn; j++) str_push(R, t->final_out, ft[j]);
                    free(ft);
                } else { release_to(R, m2); t->final = 0; }
            }
        }
    }
    if (D->ds[ds]->final == 1) {                              /* :1178-1182 */
        emit(R, acc->p, acc->n); emit(R, D->ds[ds]->final_out->p, D->ds[ds]->final_out->n);
        return i;
    }
    return -1;
}

static void run_dft(Run *R, Nft *N, int init, int mode, const uint8_t *in, size_t n) {
    Dft D;
    D.R = R; D.N = N; D.ds = NULL; D.n_ds = D.cap_ds = 0; D.chain = NULL;
    D.acc = str_new(R); D.depth = 0; D.work = 0;
    D.n_bucket = 1u << 16;
    D.bucket = (int *)take(R, D.n_bucket * sizeof(int));
    for (size_t k = 0; k < D.n_bucket; k++) D.bucket[k] = -1;
    List *l0 = list_new(R);                                   /* {(init, "")} (:1255-1259) */
    list_push(R, l0, init, str_new(R));
    int dstart = dst_new(&D, l0);
    size_t pos = 0, b, e;
    while (next_record(in, n, &pos, &b, &e)) {
        if (mode == TRRE_ORACLE_SCAN) {                       /* trre_dft.c:1272-1286 */
            size_t ch = b;
            while (ch < e) {
                long k = walk_dft(&D, dstart, in + ch, in + e);
                if (k > 0) ch += (size_t)k; else emit1(R, in[ch++]);
            }
            walk_dft(&D, dstart, in + e, in + e);             /* no-op: dstart->final stays -1 */
            emit1(R, '\n');
        } else {                                              /* match mode prints a bare newline (:1287-1293) */
            emit1(R, '\n');
        }
    }
    g_last_dstates = D.n_ds;
}

/* ------------------------------------------------------------------------------------------ */
/* entry                                                                                       */
/* ------------------------------------------------------------------------------------------ */

int trre_oracle_run(const char *expr, int flavor, int mode, int all,
                    const uint8_t *in, size_t n, uint8_t **out, size_t *out_n,
                    char *err, size_t errcap) {
    Run R;
    memset(&R, 0, sizeof R);
    R.err = err; R.errcap = errcap;
    if (err && errcap) err[0] = 0;
    if (out) *out = NULL;
    if (out_n) *out_n = 0;
    if (setjmp(R.bail) == 0) {
        if (flavor == TRRE_ORACLE_DFT && all) die(&R, 1, "Not supported yet\n");   /* trre_dft.c:1225-1227 */
        Parser P; memset(&P, 0, sizeof P); P.R = &R;
        int root = parse_expr(&P, expr);
        Nft N; memset(&N, 0, sizeof N); N.R = &R; N.P = &P;
        int fin = nst_new(&N, FINAL, -1, -1);
        int start;
        if (flavor == TRRE_ORACLE_NFT) {                      /* create_nft, trre_nft.c:506-511 */
            Frag f = build(&N, root, 0);
            N.st[f.tail].a = fin;
            start = f.head;
        } else {                                              /* trre_dft.c:516-523 */
            int init = nst_new(&N, JOIN, -1, -1);
            Frag f = build(&N, root, 0);
            N.st[f.tail].a = fin;
            N.st[init].a = f.head;
            start = init;
        }
        g_last_nstates = (size_t)N.n;
        if (flavor == TRRE_ORACLE_NFT) run_nft(&R, &N, start, mode, all, in, n);
        else run_dft(&R, &N, start, mode, in, n);
    }
    release_all(&R);
    if (out) *out = R.out; else free(R.out);
    if (out_n) *out_n = R.out_n;
    return R.rc;
}

void trre_oracle_free(void *p) { free(p); }
size_t trre_oracle_last_nstates(void) { return g_last_nstates; }
size_t trre_oracle_last_dstates(void) { return g_last_dstates; }

/* ---- synthetic corpus (BASELINE.md s3 / SURVEY.md s8d) ----------------------------------------------------------
 * Second, independent implementation of the generator the product exports as trre_synth_text(): bench.py's
 * --impl reference arm builds its input with this one so that it never maps the product's library, and
 * tests/test_bench_contract.py checks the two byte for byte.  splitmix64 stream: line target 40 + r % 81; per token
 * one draw r (low 32 bits < p_hit * 2^32: hit word number (r >> 32) % n_words; else 1 + (r >> 32) % maxtok random
 * characters, 12 bits each from two further draws); single spaces; the buffer's last byte is '\n'. */
static uint64_t corpus_next(uint64_t *x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int trre_oracle_synth_text(uint64_t seed, uint8_t *buf, size_t n, const char *words, double p_hit, int lowercase_only) {
    static const char mixed[] = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789.,;";
    const char *alpha = mixed;                                /* its first 26 characters are the lowercase set */
    const uint64_t nalpha = lowercase_only ? 26 : 65, maxtok = lowercase_only ? 8 : 10;
    const char **hit = NULL;
    size_t *hit_len = NULL, n_hits = 0, cap = 0;
    if (!buf && n) return 1;
    for (const char *w = words; w && *w;) {
        const char *e = strchr(w, '\n');
        size_t len = e ? (size_t)(e - w) : strlen(w);
        if (len) {
            if (n_hits == cap) {
                cap = cap ? 2 * cap : 64;
                hit = (const char **)realloc((void *)hit, cap * sizeof *hit);
                hit_len = (size_t *)realloc(hit_len, cap * sizeof *hit_len);
                if (!hit || !hit_len) return 1;
            }
            hit[n_hits] = w; hit_len[n_hits++] = len;
        }
        w += len + (e ? 1 : 0);
    }
    const uint64_t thresh = (p_hit <= 0 || !n_hits) ? 0 : (uint64_t)(p_hit * 4294967296.0);
    uint64_t st = seed;
    size_t pos = 0, line_len = 0, target = 40 + corpus_next(&st) % 81;
    while (pos < n) {
        const uint64_t r = corpus_next(&st);
        if ((r & 0xFFFFFFFFull) < thresh) {
            const size_t k = (size_t)((r >> 32) % n_hits);
            for (size_t c = 0; c < hit_len[k] && pos < n; ++c, ++line_len) buf[pos++] = (uint8_t)hit[k][c];
        } else {
            const uint64_t len = 1 + (r >> 32) % maxtok, b0 = corpus_next(&st), b1 = corpus_next(&st);
            for (uint64_t c = 0; c < len && pos < n; ++c, ++line_len)
                buf[pos++] = (uint8_t)alpha[((c < 5 ? b0 >> (c * 12) : b1 >> ((c - 5) * 12)) & 0xFFF) % nalpha];
        }
        if (pos >= n) break;
        if (line_len >= target) { buf[pos++] = '\n'; line_len = 0; target = 40 + corpus_next(&st) % 81; }
        else { buf[pos++] = ' '; ++line_len; }
    }
    if (n) buf[n - 1] = '\n';
    free((void *)hit);
    free(hit_len);
    return 0;
}
