/*
 * ref_shim.c -- runs the reference's OWN AMD64 assembly (asm_amd64.s) from C.
 *
 * TEST INFRASTRUCTURE ONLY (same rule as minlz_oracle.h): only tests/,
 * __graft_entry__.smoke() and bench.py's CPU legs may load the library this
 * builds (oracle/_ref/libminlz_ref.so).  The product never links or loads it.
 *
 * The reference's hot loops are Go-assembler (Plan 9 syntax) functions with the
 * Go ABI0 calling convention: every argument and result lives in a stack block
 * directly above the return address (`dst_base+0(FP)` ... `ret+56(FP)`).
 * oracle/p9_to_gas.py transliterates the instructions into GNU as syntax
 * (symbols `p9_<name>`); this file supplies what surrounds them in Go:
 *
 *   p9_call         the ABI0 call frame (copy the argument block onto the
 *                   stack, call, copy it back so results can be read), saving
 *                   the SysV callee-saved registers the Go code is free to use;
 *   mzr_*           the per-size dispatch of encode_amd64.go:37-271 and
 *                   decode_amd64.go:21-30 (which variant runs for which block
 *                   length, and the scratch table each one is handed), restated;
 *   mzr_*_batch_mt  one block per task over a pthread pool, for bench.py.
 *
 * Nothing here is a codec: all token work happens inside the reference's
 * assembly.  The scratch tables are caller-zeroed-or-not exactly as in Go: the
 * asm zeroes `tmp` itself (asm_amd64.s, zero_loop_*), sync.Pool hands it dirty.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* int64 p9_call(void *fn, void *frame, size_t nbytes): nbytes <= 128. */
__asm__(
    ".text\n"
    ".globl p9_call\n"
    ".type p9_call,@function\n"
    "p9_call:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  subq $152, %rsp\n"          /* 128 B argument block + frame ptr + nbytes; keeps rsp 16-aligned */
    "  movq %rsi, 128(%rsp)\n"
    "  movq %rdx, 136(%rsp)\n"
    "  movq %rdi, %rax\n"
    "  movq %rsp, %rdi\n"
    "  movq %rdx, %rcx\n"
    "  rep movsb\n"                /* frame -> stack */
    "  call *%rax\n"
    "  movq 128(%rsp), %rdi\n"
    "  movq 136(%rsp), %rcx\n"
    "  movq %rsp, %rsi\n"
    "  rep movsb\n"                /* stack -> frame (results) */
    "  addq $152, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n"
    ".size p9_call, .-p9_call\n");
extern void p9_call(void *fn, void *frame, size_t nbytes);

typedef struct { const uint8_t *p; int64_t len, cap; } goslice;

#define DECL(n) extern void p9_##n(void)
DECL(encodeBlockAsm); DECL(encodeBlockAsm2MB); DECL(encodeBlockAsm512K); DECL(encodeBlockAsm64K);
DECL(encodeBlockAsm16K); DECL(encodeBlockAsm4K); DECL(encodeBlockAsm1K);
DECL(encodeFastBlockAsm); DECL(encodeFastBlockAsm2MB); DECL(encodeFastBlockAsm512K); DECL(encodeFastBlockAsm64K);
DECL(encodeFastBlockAsm16K); DECL(encodeFastBlockAsm4K); DECL(encodeFastBlockAsm1K);
DECL(encodeBetterBlockAsm); DECL(encodeBetterBlockAsm2MB); DECL(encodeBetterBlockAsm512K);
DECL(encodeBetterBlockAsm64K); DECL(encodeBetterBlockAsm16K); DECL(encodeBetterBlockAsm4K);
DECL(encodeBetterBlockAsm1K);
DECL(emitLiteral); DECL(emitRepeat); DECL(emitCopy); DECL(emitCopyLits2); DECL(emitCopyLits3);
DECL(matchLen); DECL(decodeBlockAsm);

/* func encode*Asm*(dst []byte, src []byte, tmp *[N]byte) int      -- frame $24-64 */
static int64_t call_enc(void (*fn)(void), uint8_t *dst, size_t dst_len, const uint8_t *src, size_t n, void *tmp) {
    struct { goslice dst, src; void *tmp; int64_t ret; } f = {
        {dst, (int64_t)dst_len, (int64_t)dst_len}, {src, (int64_t)n, (int64_t)n}, tmp, -12345};
    p9_call((void *)fn, &f, sizeof f);
    return f.ret;
}

/* Largest scratch any variant takes (encodeBetterBlockAsm: 589824 B).  One per
 * thread, never cleared between calls -- like a sync.Pool entry. */
static __thread uint8_t *tl_tmp;
static void *scratch(void) {
    if (!tl_tmp && posix_memalign((void **)&tl_tmp, 64, 589824)) abort();
    return tl_tmp;
}
static void scratch_release(void) { free(tl_tmp); tl_tmp = NULL; }

/* encode_amd64.go:119-189 encodeBlock (LevelFastest). dst_len = len(dst). */
int64_t mzr_encode_block_l1(uint8_t *dst, size_t dst_len, const uint8_t *src, size_t n) {
    void (*fn)(void);
    if (n > (2u << 20)) fn = p9_encodeBlockAsm;
    else if (n > (512u << 10)) fn = p9_encodeBlockAsm2MB;
    else if (n > (64u << 10)) fn = p9_encodeBlockAsm512K;
    else if (n > (16u << 10)) fn = p9_encodeBlockAsm64K;
    else if (n > (4u << 10)) fn = p9_encodeBlockAsm16K;
    else if (n > (1u << 10)) fn = p9_encodeBlockAsm4K;
    else if (n > 16) fn = p9_encodeBlockAsm1K;
    else return 0;
    return call_enc(fn, dst, dst_len, src, n, scratch());
}

/* encode_amd64.go:37-107 encodeBlockFast (LevelSuperFast) */
int64_t mzr_encode_block_l0(uint8_t *dst, size_t dst_len, const uint8_t *src, size_t n) {
    void (*fn)(void);
    if (n > (2u << 20)) fn = p9_encodeFastBlockAsm;
    else if (n > (512u << 10)) fn = p9_encodeFastBlockAsm2MB;
    else if (n > (64u << 10)) fn = p9_encodeFastBlockAsm512K;
    else if (n > (16u << 10)) fn = p9_encodeFastBlockAsm64K;
    else if (n > (4u << 10)) fn = p9_encodeFastBlockAsm16K;
    else if (n > (1u << 10)) fn = p9_encodeFastBlockAsm4K;
    else if (n > 32) fn = p9_encodeFastBlockAsm1K;
    else return 0;
    return call_enc(fn, dst, dst_len, src, n, scratch());
}

/* encode_amd64.go:201-271 encodeBlockBetter (LevelBalanced) */
int64_t mzr_encode_block_l2(uint8_t *dst, size_t dst_len, const uint8_t *src, size_t n) {
    void (*fn)(void);
    if (n > (2u << 20)) fn = p9_encodeBetterBlockAsm;
    else if (n > (512u << 10)) fn = p9_encodeBetterBlockAsm2MB;
    else if (n > (64u << 10)) fn = p9_encodeBetterBlockAsm512K;
    else if (n > (16u << 10)) fn = p9_encodeBetterBlockAsm64K;
    else if (n > (4u << 10)) fn = p9_encodeBetterBlockAsm16K;
    else if (n > (1u << 10)) fn = p9_encodeBetterBlockAsm4K;
    else if (n > 16) fn = p9_encodeBetterBlockAsm1K;
    else return 0;
    return call_enc(fn, dst, dst_len, src, n, scratch());
}

/* decode_amd64.go:21-30 minLZDecode -> func decodeBlockAsm(dst []byte, src []byte) int   -- frame $8-56 */
int mzr_decode_block(uint8_t *dst, size_t dst_len, const uint8_t *src, size_t src_len) {
    struct { goslice dst, src; int64_t ret; } f = {
        {dst, (int64_t)dst_len, (int64_t)dst_len}, {src, (int64_t)src_len, (int64_t)src_len}, -12345};
    p9_call((void *)p9_decodeBlockAsm, &f, sizeof f);
    return (int)f.ret;
}

/* asm_amd64.go:162-210 emitters and matchLen (used to pin the oracle's restated emitters) */
int64_t mzr_emit_literal(uint8_t *dst, size_t dst_len, const uint8_t *lit, size_t n) {
    struct { goslice dst, lit; int64_t ret; } f = {{dst, (int64_t)dst_len, (int64_t)dst_len}, {lit, (int64_t)n, (int64_t)n}, -1};
    p9_call((void *)p9_emitLiteral, &f, sizeof f);
    return f.ret;
}
int64_t mzr_emit_repeat(uint8_t *dst, size_t dst_len, int64_t length) {
    struct { goslice dst; int64_t length, ret; } f = {{dst, (int64_t)dst_len, (int64_t)dst_len}, length, -1};
    p9_call((void *)p9_emitRepeat, &f, sizeof f);
    return f.ret;
}
int64_t mzr_emit_copy(uint8_t *dst, size_t dst_len, int64_t offset, int64_t length) {
    struct { goslice dst; int64_t offset, length, ret; } f = {{dst, (int64_t)dst_len, (int64_t)dst_len}, offset, length, -1};
    p9_call((void *)p9_emitCopy, &f, sizeof f);
    return f.ret;
}
int64_t mzr_emit_copy_lits2(uint8_t *dst, size_t dst_len, const uint8_t *lits, size_t nl, int64_t offset, int64_t length) {
    struct { goslice dst, lits; int64_t offset, length, ret; } f = {
        {dst, (int64_t)dst_len, (int64_t)dst_len}, {lits, (int64_t)nl, (int64_t)nl}, offset, length, -1};
    p9_call((void *)p9_emitCopyLits2, &f, sizeof f);
    return f.ret;
}
int64_t mzr_emit_copy_lits3(uint8_t *dst, size_t dst_len, const uint8_t *lits, size_t nl, int64_t offset, int64_t length) {
    struct { goslice dst, lits; int64_t offset, length, ret; } f = {
        {dst, (int64_t)dst_len, (int64_t)dst_len}, {lits, (int64_t)nl, (int64_t)nl}, offset, length, -1};
    p9_call((void *)p9_emitCopyLits3, &f, sizeof f);
    return f.ret;
}
int64_t mzr_match_len(const uint8_t *a, size_t na, const uint8_t *b, size_t nb) {
    struct { goslice a, b; int64_t ret; } f = {{a, (int64_t)na, (int64_t)na}, {b, (int64_t)nb, (int64_t)nb}, -1};
    p9_call((void *)p9_matchLen, &f, sizeof f);
    return f.ret;
}

/* ---- batch drivers (bench.py CPU legs): one block per task, atomic work counter ---- */
typedef struct {
    int level, nblk;
    const uint8_t *src; const uint64_t *src_off;
    uint8_t *dst; const uint64_t *dst_off;
    uint32_t *out_len; int32_t *status;
    int next;
} batch_t;

static void *enc_worker(void *arg) {
    batch_t *b = (batch_t *)arg;
    for (;;) {
        int i = __atomic_fetch_add(&b->next, 1, __ATOMIC_RELAXED);
        if (i >= b->nblk) break;
        const uint8_t *s = b->src + b->src_off[i];
        size_t n = (size_t)(b->src_off[i + 1] - b->src_off[i]);
        uint8_t *d = b->dst + b->dst_off[i];
        size_t cap = (size_t)(b->dst_off[i + 1] - b->dst_off[i]);
        int64_t r = b->level == 2 ? mzr_encode_block_l2(d, cap, s, n)
                  : b->level == 1 ? mzr_encode_block_l1(d, cap, s, n)
                                  : mzr_encode_block_l0(d, cap, s, n);
        b->out_len[i] = (uint32_t)r;
    }
    scratch_release();
    return NULL;
}
static void *dec_worker(void *arg) {
    batch_t *b = (batch_t *)arg;
    for (;;) {
        int i = __atomic_fetch_add(&b->next, 1, __ATOMIC_RELAXED);
        if (i >= b->nblk) break;
        b->status[i] = mzr_decode_block(b->dst + b->dst_off[i], (size_t)(b->dst_off[i + 1] - b->dst_off[i]),
                                        b->src + b->src_off[i], (size_t)(b->src_off[i + 1] - b->src_off[i]));
    }
    return NULL;
}
static int run_pool(void *(*fn)(void *), batch_t *b, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    for (int t = 0; t < nthreads; t++)
        if (pthread_create(&th[t], NULL, fn, b)) return -1;
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    return 0;
}
int mzr_encode_batch_mt(int level, int nblk, const uint8_t *src, const uint64_t *src_off, uint8_t *dst,
                        const uint64_t *dst_off, uint32_t *out_len, int nthreads) {
    batch_t b = {level, nblk, src, src_off, dst, dst_off, out_len, NULL, 0};
    return run_pool(enc_worker, &b, nthreads);
}
int mzr_decode_batch_mt(int nblk, const uint8_t *src, const uint64_t *src_off, uint8_t *dst,
                        const uint64_t *dst_off, int32_t *status, int nthreads) {
    batch_t b = {0, nblk, src, src_off, dst, dst_off, NULL, status, 0};
    return run_pool(dec_worker, &b, nthreads);
}
