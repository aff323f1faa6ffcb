/*
 * trre_oracle.h - CPU ORACLE for the trre scan path.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's algorithm (c0stya/trre @ c474a68) for the path
 * BASELINE.json names: parser -> NFT -> {backtracking scan | lazily determinized DFT scan} over
 * getline records.  It exists to CHECK the CUDA engine; it is never linked into, imported by or
 * executed from the product (trre_b200/, include/, bin/).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Parity status: PINNED.  tests/test_oracle_pinned.py checks this restatement against
 *   (a) all 69 known-answer vectors of the reference's test.sh (18 scan + 51 match),
 *   (b) the probe log of SURVEY.md s4.4,
 *   (c) outputs of the real reference binaries (oracle/_ref/trre, oracle/_ref/trre_dft, compiled
 *       from /root/reference by oracle/Makefile) on seeded random patterns x inputs, and the
 *       committed fixtures under tests/golden/ generated from those binaries.
 */
#ifndef TRRE_ORACLE_H
#define TRRE_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { TRRE_ORACLE_NFT = 0, TRRE_ORACLE_DFT = 1 };           /* which reference binary to mimic */
enum { TRRE_ORACLE_SCAN = 0, TRRE_ORACLE_MATCH = 1 };        /* default mode / -m               */

/* Return codes: 0 ok; 1 = the reference would have printed `err` to stderr and exit(1)
 * (whatever it had written to stdout before that is still returned in *out);
 * 134 = the reference would have hit assert() in parse() (empty pattern);
 * 2 = pattern is outside what the reference can process without undefined behaviour. */
int trre_oracle_run(const char *expr, int flavor, int mode, int all,
                    const uint8_t *in, size_t n,
                    uint8_t **out, size_t *out_n,
                    char *err, size_t errcap);

void trre_oracle_free(void *p);

/* Introspection used by tests: number of NFT states / lazily created DFT states of the last run
 * in this thread (not thread safe; test use only). */
size_t trre_oracle_last_nstates(void);
size_t trre_oracle_last_dstates(void);

/* The synthetic corpus of BASELINE.md s3, implemented a second time (the product exports trre_synth_text): what
 * bench.py --impl reference feeds the reference binary, without loading the product's library.  `words`: hit list,
 * '\n'-separated, or NULL. */
int trre_oracle_synth_text(uint64_t seed, uint8_t *buf, size_t n, const char *words, double p_hit, int lowercase_only);

#ifdef __cplusplus
}
#endif
#endif
