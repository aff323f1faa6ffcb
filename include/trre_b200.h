/*
 * trre_b200.h - C ABI of the B200-native scan engine for trre (transductive regular expressions).
 *
 * This is the drop-in boundary for ONE path of the reference (c0stya/trre @ c474a68): the scan-mode
 * loop that walks a compiled transducer over the bytes of every input line and writes the
 * transformed bytes.  The reference has no library API; the functions below replace, as whole-buffer
 * calls, what its two `main()`s do per line:
 *
 *   reference                                               replaced by
 *   ------------------------------------------------------  -----------------------------------------
 *   parse() + create_nft()        trre_nft.c:196-288,506-511  trre_compile(expr, TRRE_FLAVOR_NFT, ..)
 *   parse() + create_nft() + the lazily filled dstate graph
 *                                 trre_dft.c:196-288,516-523,
 *                                 926-942, 1135-1175           trre_compile(expr, TRRE_FLAVOR_DFT, ..)
 *   scan driver + infer_backtrack trre_nft.c:775-790, 593-657  trre_scan / trre_scan_device (NFT flavour)
 *   scan driver + infer_dft       trre_dft.c:1272-1286,
 *                                 1110-1196                    trre_scan / trre_scan_device (DFT flavour)
 *   -a in scan mode               trre_nft.c:635-648, 782-788  trre_prog_set_mode(p, 0, 1) + trre_scan*  (GPU, DFS executor)
 *   match driver (-m, -ma)        trre_nft.c:635-642, 791-797  trre_prog_set_mode(p, 1, all) + trre_scan* (GPU, DFS executor)
 *                                                              trre_match_host: the same on the host, kept as a cross-check
 *
 * Plain pointers and sizes only; no CUDA or torch types in any signature (a stream is passed as
 * `void *`, i.e. a cudaStream_t, 0 = default stream).  All scan entry points run on the GPU: there
 * is no CPU fallback - without a usable CUDA device they return TRRE_E_CUDA.
 *
 * Output contract: byte-for-byte what the corresponding reference binary writes to stdout for the
 * same pattern and the same stdin bytes (getline records, last byte of each record dropped, C-string
 * payload, one '\n' per record - trre_nft.c:776-789).
 */
#ifndef TRRE_B200_H
#define TRRE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRRE_B200_ABI_VERSION 3

/* which reference binary's semantics: `trre` (leftmost, first-by-priority, empty matches allowed)
 * or `trre_dft` (leftmost, shortest non-empty match). */
enum { TRRE_FLAVOR_NFT = 0, TRRE_FLAVOR_DFT = 1 };

/* return codes */
enum {
    TRRE_OK            = 0,
    TRRE_E_PATTERN     = 1,    /* the reference prints `err` on stderr and exits 1 (parser / range errors)  */
    TRRE_E_UNSUPPORTED = 2,    /* pattern outside what the engine (or the reference without UB) can run     */
    TRRE_E_CUDA        = 3,    /* no device / CUDA runtime error; text in `err` or trre_last_error()        */
    TRRE_E_NOSPACE     = 4,    /* out_cap too small; *out_n holds the size needed                           */
    TRRE_E_STACK       = 5,    /* "error: stack max capacity reached" (trre_nft.c:551-554)                  */
    TRRE_E_ARG         = 7,
    TRRE_E_ASSERT      = 134   /* empty pattern: the reference dies in assert() (trre_nft.c:285)            */
};

/* how the compiled program will be executed on the device */
enum {
    TRRE_EXEC_BYTEMAP   = 0,   /* every attempt is decided by one byte and emits one byte: 256-entry LUT     */
    TRRE_EXEC_TABLE     = 1,   /* deterministic scan automaton walked per line                               */
    TRRE_EXEC_BACKTRACK = 2,   /* NFT flavour only: per-line priority DFS over the flat NFT                  */
    TRRE_EXEC_LAZY      = 3    /* scan automaton too large to build ahead of time: rows are built on demand,  */
                               /* as trre_dft does (trre_dft.c:1135-1175); the scan itself stays on the GPU   */
};

/* the kernel family that serves a regular buffer (no NUL byte, final newline); irregular buffers and
 * everything else run the per-line kernels */
enum {
    TRRE_FAST_NONE      = 0,   /* per-line kernels only (cycles, empty matches, DFS executor)                */
    TRRE_FAST_BYTEMAP   = 1,   /* k_bytemap: one byte in, one byte out                                       */
    TRRE_FAST_MATCH     = 2,   /* k_lit (literal_words > 0) / k_gram (gram_tables) / k_match: length-preserving  */
                               /* program with short attempts                                                */
    TRRE_FAST_BYTEMAP_VAR = 3, /* k_bmv_*: one byte in, 0..255 bytes out; also runs (`[0-9]+:#`: RUN form)       */
    TRRE_FAST_MATCH_VAR = 4,   /* k_litv / k_gram_var / k_match count + write: matches of any output length     */
    TRRE_FAST_INPLACE   = 5    /* k_lines in place: other length-preserving programs                         */
};

typedef struct trre_prog trre_prog;

typedef struct trre_prog_info {
    uint32_t abi_version;
    uint32_t flavor;
    uint32_t exec_kind;          /* TRRE_EXEC_*                                                             */
    uint32_t nft_states;         /* flat NFT (trre_nft.c:334-340 as an index table)                          */
    uint32_t dsa_states;         /* deterministic scan-automaton states (0 for TRRE_EXEC_BACKTRACK)          */
    uint32_t byte_classes;       /* columns of the transition table after byte-class compression            */
    uint32_t max_depth;          /* longest walk of one attempt; 0xffffffff if the automaton has a cycle     */
    uint32_t length_preserving;  /* 1: every match emits exactly as many bytes as it consumes                */
    uint32_t can_match_empty;    /* 1: an attempt can succeed without consuming                              */
    uint32_t table_bytes;        /* size of the device-resident tables                                       */
    uint32_t fast_path;          /* TRRE_FAST_*: which position-parallel kernel serves a regular buffer      */
    uint32_t gram_tables;        /* 1: every attempt is tabulated by byte classes (k_gram / k_gram_var serve   */
                                 /*    TRRE_FAST_MATCH / TRRE_FAST_MATCH_VAR instead of the walking k_match)   */
    uint32_t literal_words;      /* > 0: the matches are that many fixed words of 2..4 bytes that cannot overlap */
                                 /*    (cat:dog, (cat:dog|dog:cat)): k_lit compares in registers, no tables     */
} trre_prog_info;

/* ---- host side: pattern -> program (no GPU needed) ------------------------------------------- */

/* Parse `expr` and build the tables.  On TRRE_E_PATTERN, `err` holds the reference's own message. */
int  trre_compile(const char *expr, int flavor, trre_prog **out, char *err, size_t errcap);
void trre_prog_free(trre_prog *p);
int  trre_prog_get_info(const trre_prog *p, trre_prog_info *info);

/* The flat tables as one relocatable blob: this is what a multi-GPU launcher broadcasts from
 * rank 0 (one ncclBroadcast / torch.distributed.broadcast) instead of recompiling per rank. */
const void *trre_prog_blob(const trre_prog *p, size_t *nbytes);
int  trre_prog_from_blob(const void *blob, size_t nbytes, trre_prog **out, char *err, size_t errcap);

/* TRRE_EXEC_LAZY programs (trre_dft.c:1135-1175: rows of the deterministic transducer exist only for the states the
 * input reached).  During a scan the GPU reports the rows it misses and the library builds them; these two calls
 * expose the same builder to tooling: build the rows of `states` (+ up to `extra` more, breadth first), and look at
 * what exists.  rows: u64[n_states * 256] by raw byte, lo = next | flags (0x00FFFFFE: row not built, 0x00FFFFFF:
 * dead), hi = output ref into pool (offset(24) | length(8), length 255: u32 length at the offset); sinfo: u64 per
 * state, lo = flags (1 final pending, 2 decided, 8 row built), hi = ref of the final output.  The pointers stay
 * valid until the next build. */
int  trre_lazy_build(trre_prog *p, const uint32_t *states, size_t n, uint32_t extra);
int  trre_lazy_view(const trre_prog *p, uint32_t *n_states, const uint64_t **rows, const uint64_t **sinfo,
                    const uint8_t **pool, size_t *pool_bytes);

/* The CLI's -m / -a flags (trre_nft.c:728-741): which output every trre_scan* call of this program produces from now on.
 *   match=0, all=0  scan mode (the default)
 *   match=0, all=1  `trre -a`: at every position ALL accepting paths print, in DFS priority order, then the raw byte is
 *                   copied (an attempt never "wins" with -a: trre_nft.c:646-648 falls through and returns -1)
 *   match=1, all=0  `trre -m`: a record prints the output of its first accepting path over the WHOLE record + '\n',
 *                   records that do not match print nothing (trre_nft.c:635-642, 792-796)
 *   match=1, all=1  `trre -ma`: every accepting path
 * These modes run the reference's priority DFS over the flat NFT on the GPU, one record per thread (k_lines with the DFS
 * executor), whatever tables the program has.  DFT flavour: -m prints one bare newline per record (trre_dft.c:1287-1292);
 * -a returns TRRE_E_UNSUPPORTED with the reference's text "Not supported yet" (trre_dft.c:1226).  The mode is not part
 * of the blob: set it again on a program made by trre_prog_from_blob.  It is plain state of the program object: do not
 * change it while a scan of the same program is running in another thread.  One record may take at most 2^26 DFS steps
 * and, with all != 0, print at most 16 MiB (TRRE_E_UNSUPPORTED beyond: the reference would run for hours there). */
int  trre_prog_set_mode(trre_prog *p, int match, int all);

/* 256-entry map for TRRE_EXEC_BYTEMAP programs (returns TRRE_E_ARG otherwise). */
int  trre_prog_bytemap(const trre_prog *p, uint8_t lut[256]);

/* ---- device side: the scan path --------------------------------------------------------------- */

/* Whole-buffer scan with HOST pointers: H2D, scan on `device`, D2H.  *out is malloc'd, release with
 * trre_free().  Equivalent to `trre EXPR < in > out` (or trre_dft). */
int  trre_scan(const trre_prog *p, int device,
               const uint8_t *in, size_t n, uint8_t **out, size_t *out_n);

/* Same, into a caller-provided host buffer (pinned memory makes the copies asynchronous). */
int  trre_scan_into(const trre_prog *p, int device,
                    const uint8_t *in, size_t n, uint8_t *out, size_t out_cap, size_t *out_n);

/* trre_scan over `ndev` GPUs of one box: the buffer is cut at line boundaries (trre_shard_lines), shard k is
 * scanned on devices[k] (one host thread per device, each with its own H2D / kernels / D2H pipeline), the
 * outputs are concatenated in shard order - byte-identical to the single-GPU result because lines are
 * independent (trre_nft.c:778-781).  The CLI uses it when TRRE_GPUS=N is set. */
int  trre_scan_multi(const trre_prog *p, const int *devices, int ndev,
                     const uint8_t *in, size_t n, uint8_t **out, size_t *out_n);

/* Device-resident scan: d_in/d_out are device pointers on `device`; work is enqueued on `stream`
 * (a cudaStream_t).  If out_n != NULL the call synchronises the stream and returns the output size
 * and status; if out_n == NULL it only enqueues: the result is complete, and its size/status known,
 * once trre_scan_status returns, so d_in/d_out must stay valid until then (a fast kernel that meets an
 * irregular buffer - NUL byte, no final newline - is rerun on the general path from there).
 * Tables are uploaded to the device on first use and stay resident.  One call at a time per
 * (program, device): the per-device workspace is not shared between host threads. */
int  trre_scan_device(const trre_prog *p, int device,
                      const uint8_t *d_in, size_t n, uint8_t *d_out, size_t out_cap,
                      size_t *out_n, void *stream);

/* Synchronise `stream` and report the outcome of the last trre_scan_device enqueued for (p, device). */
int  trre_scan_status(const trre_prog *p, int device, size_t *out_n, void *stream);

/* Number of kernels the last trre_scan_device call launched for (p, device). */
int  trre_scan_launches(const trre_prog *p, int device);

/* ---- line-aligned sharding (multi-GPU: lines are the independent unit, trre_nft.c:778-781) ----- */

/* Cut [0,n) into `parts` contiguous ranges whose boundaries sit just after a '\n'.  cuts[] has
 * parts+1 entries, cuts[0]=0, cuts[parts]=n. */
int  trre_shard_lines(const uint8_t *in, size_t n, int parts, size_t *cuts);

/* ---- host-only helpers ------------------------------------------------------------------------ */

/* Match mode (`-m`, with all!=0 `-ma`; with scan!=0 it is `-a` in scan mode) on the HOST: C++ over the flat NFT.
 * The CLI runs these flags on the GPU (trre_prog_set_mode); this stays as the host-side cross-check of that path. */
int  trre_match_host(const trre_prog *p, int scan, int all,
                     const uint8_t *in, size_t n, uint8_t **out, size_t *out_n,
                     char *err, size_t errcap);

/* Deterministic synthetic text of BASELINE.md s3 / SURVEY s8d (printable ASCII, lines U{40..120},
 * ends with '\n').  `words` = optional hit list ('\n'-separated), used with probability p_hit. */
int  trre_synth_text(uint64_t seed, uint8_t *buf, size_t n,
                     const char *words, double p_hit, int lowercase_only);

/* Page-locked host memory for trre_scan_into: with it the H2D / D2H copies of the chunk pipeline are truly asynchronous
 * (pageable memory is staged by the driver).  NULL when there is no usable CUDA device. */
void       *trre_host_alloc(size_t nbytes);
void        trre_host_free(void *p);

void        trre_free(void *p);
const char *trre_last_error(void);
const char *trre_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TRRE_B200_H */
