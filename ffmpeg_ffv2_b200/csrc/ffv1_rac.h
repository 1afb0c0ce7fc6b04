/*
 * ffv1_rac.h -- FFV1's binary adaptive range coder and its symbol binarisation,
 * as plain inline functions usable from host C, from CUDA device code and from
 * the CPU emulation harness in tests/emul (same source, three compilers).
 *
 * Bit-exact counterpart of libavcodec/rangecoder.h:71-152, rangecoder.c:42-123
 * (renorm with carry / outstanding bytes, put_rac, get_rac, termination) and of
 * put_symbol_inline ffv1enc.c:185-231 / get_symbol_inline ffv1dec.c:42-64.
 */
#ifndef FFGPU_FFV1_RAC_H
#define FFGPU_FFV1_RAC_H

#include <stdint.h>

#if defined(__CUDACC__)
#define FFGPU_HD __host__ __device__ __forceinline__
#define FFGPU_HD_COLD static __host__ __device__ __noinline__   /* once per slice: keep it out of line */
#else
#define FFGPU_HD static inline
#define FFGPU_HD_COLD static inline
#endif

/* adaptive-probability transition tables: one_state / zero_state (256 B each) */
typedef struct FFRacTables {
    uint8_t one[256];
    uint8_t zero[256];
} FFRacTables;

/* ---------------- encoder ---------------- */
typedef struct FFRacEnc {
    int32_t  low;
    int32_t  range;
    int32_t  pending;      /* outstanding_byte, -1 until the first renorm      */
    int32_t  run;          /* outstanding_count: carry-transparent bytes        */
    uint8_t *buf;
    uint32_t pos;          /* bytes emitted                                    */
    uint32_t cap;          /* capacity; writes beyond it are dropped + flagged */
    uint32_t overflow;
} FFRacEnc;

FFGPU_HD void ffrac_enc_init(FFRacEnc *c, uint8_t *buf, uint32_t cap)
{
    c->low = 0;
    c->range = 0xFF00;
    c->pending = -1;
    c->run = 0;
    c->buf = buf;
    c->pos = 0;
    c->cap = cap;
    c->overflow = 0;
}

FFGPU_HD void ffrac_emit(FFRacEnc *c, int b)
{
    if (c->pos < c->cap)
        c->buf[c->pos] = (uint8_t)b;
    else
        c->overflow = 1;
    c->pos++;
}

/* one byte leaves the low register; resolve the carry into the held-back bytes */
FFGPU_HD void ffrac_enc_renorm(FFRacEnc *c)
{
    while (c->range < 0x100) {
        if (c->pending < 0) {
            c->pending = c->low >> 8;
        } else if (c->low <= 0xFF00) {
            ffrac_emit(c, c->pending);
            for (; c->run; c->run--)
                ffrac_emit(c, 0xFF);
            c->pending = c->low >> 8;
        } else if (c->low >= 0x10000) {
            ffrac_emit(c, c->pending + 1);
            for (; c->run; c->run--)
                ffrac_emit(c, 0x00);
            c->pending = (c->low >> 8) & 0xFF;
        } else {
            c->run++;
        }
        c->low = (c->low & 0xFF) << 8;
        c->range <<= 8;
    }
}

FFGPU_HD void ffrac_put(FFRacEnc *c, const FFRacTables *t, uint8_t *state, int bit)
{
    const int s = *state;
    const int r1 = (c->range * s) >> 8;
    if (!bit) {
        c->range -= r1;
        *state = t->zero[s];
    } else {
        c->low += c->range - r1;
        c->range = r1;
        *state = t->one[s];
    }
    ffrac_enc_renorm(c);
}

/* ff_rac_terminate(c, version): returns the number of bytes written */
FFGPU_HD uint32_t ffrac_enc_finish(FFRacEnc *c, const FFRacTables *t, int version)
{
    if (version == 1) {
        uint8_t s = 129;
        ffrac_put(c, t, &s, 0);
    }
    c->range = 0xFF;
    c->low += 0xFF;
    ffrac_enc_renorm(c);
    c->range = 0xFF;
    ffrac_enc_renorm(c);
    return c->pos;
}

FFGPU_HD int ffrac_ilog2(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)(v | 1));
#else
    return 31 - __builtin_clz(v | 1);
#endif
}

/* put_symbol_inline: zero flag, unary exponent, mantissa MSB-first, sign */
FFGPU_HD void ffrac_put_symbol(FFRacEnc *c, const FFRacTables *t, uint8_t *st, int v, int is_signed)
{
    if (!v) {
        ffrac_put(c, t, st, 1);
        return;
    }
    {
        const int a = v < 0 ? -v : v;
        const int e = ffrac_ilog2((uint32_t)a);
        int i;
        ffrac_put(c, t, st, 0);
        for (i = 0; i < e; i++)
            ffrac_put(c, t, st + 1 + (i < 9 ? i : 9), 1);
        ffrac_put(c, t, st + 1 + (e < 9 ? e : 9), 0);
        for (i = e - 1; i >= 0; i--)
            ffrac_put(c, t, st + 22 + (i < 9 ? i : 9), (a >> i) & 1);
        if (is_signed)
            ffrac_put(c, t, st + 11 + (e < 10 ? e : 10), v < 0);
    }
}

/* ---------------- decoder ---------------- */
typedef struct FFRacDec {
    int32_t  low;
    int32_t  range;
    const uint8_t *buf;
    uint32_t pos;          /* next byte to read                 */
    uint32_t end;          /* bytestream_end - bytestream_start */
    int32_t  overread;
} FFRacDec;

/* ff_init_range_decoder: needs size >= 2 */
FFGPU_HD void ffrac_dec_init(FFRacDec *c, const uint8_t *buf, uint32_t size)
{
    c->buf = buf;
    c->end = size;
    c->range = 0xFF00;
    c->low = (buf[0] << 8) | buf[1];
    c->pos = 2;
    c->overread = 0;
    if (c->low >= 0xFF00) {
        c->low = 0xFF00;
        c->end = c->pos;
    }
}

FFGPU_HD int ffrac_get(FFRacDec *c, const FFRacTables *t, uint8_t *state)
{
    const int s = *state;
    const int r1 = (c->range * s) >> 8;
    int bit;
    c->range -= r1;
    if (c->low < c->range) {
        *state = t->zero[s];
        bit = 0;
    } else {
        c->low -= c->range;
        c->range = r1;
        *state = t->one[s];
        bit = 1;
    }
    if (c->range < 0x100) {
        c->range <<= 8;
        c->low <<= 8;
        if (c->pos < c->end)
            c->low += c->buf[c->pos++];
        else
            c->overread++;
    }
    return bit;
}

/* get_symbol_inline.  On the e > 31 error the reference returns AVERROR_INVALIDDATA *as the
 * symbol value* (ffv1dec.c:54-55); callers see the same number here. */
#define FFRAC_SYMBOL_ERROR (-1094995529)
/* -v modulo 2^32: a damaged stream can produce 0x80000000 (32 mantissa bits), whose signed
 * negation is undefined; the reference wraps (ffv1dec.c:61-63, :208-209) */
#define FF_NEG32(v) ((int)(0u - (uint32_t)(v)))
FFGPU_HD int ffrac_get_symbol(FFRacDec *c, const FFRacTables *t, uint8_t *st, int is_signed)
{
    int e = 0, i;
    uint32_t a = 1;
    if (ffrac_get(c, t, st))
        return 0;
    while (ffrac_get(c, t, st + 1 + (e < 9 ? e : 9))) {
        e++;
        if (e > 31)
            return FFRAC_SYMBOL_ERROR;
    }
    for (i = e - 1; i >= 0; i--)
        a += a + (uint32_t)ffrac_get(c, t, st + 22 + (i < 9 ? i : 9));
    if (is_signed && ffrac_get(c, t, st + 11 + (e < 10 ? e : 10)))
        return (int)(0u - a);
    return (int)a;
}

#endif /* FFGPU_FFV1_RAC_H */
