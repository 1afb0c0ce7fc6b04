/*
 * oracle/ffv1_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * A from-scratch CPU restatement of the FFV1 (v0/v1/v3) slice pixel path of the
 * reference FFmpeg fork, written as the parity checker for the CUDA product.
 * It is organised differently from the reference (whole-slice sample arrays
 * with explicit neighbour rules instead of rotating line buffers; a resolved
 * stream description instead of a cloned codec context per slice) but must
 * produce byte-identical packets and identical decoded frames.
 *
 * PARITY PIN: tests/test_oracle_vs_ref.py checks this file byte-for-byte
 * against the UNMODIFIED reference built by oracle/Makefile (oracle/_ref) over
 * every pixel format / coder / context / slice option below, and
 * tests/golden/ holds packet digests generated from that reference (and
 * chained to the reference's FATE refs tests/ref/vsynth/vsynth*-ffv1*).
 *
 * Each function cites the reference code it restates (paths relative to
 * /root/reference/).
 */
#include "ffv1_oracle.h"

#include <limits.h>
#include <stddef.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define MAX_SLICES        1024   /* libavcodec/ffv1.h:77 */
#define CONTEXT_SIZE      32     /* libavcodec/ffv1.h:52 */
#define MAX_QUANT_TABLES  8
#define MAX_CTX_INPUTS    5
#define AC_GOLOMB   0
#define AC_DEFAULT  1
#define AC_CUSTOM   2

#define MINV(a, b) ((a) < (b) ? (a) : (b))
#define MAXV(a, b) ((a) > (b) ? (a) : (b))
#define CEIL_RSHIFT(a, b) (-((-(a)) >> (b)))

/* ------------------------------------------------------------------ */
/* pixel formats                                                       */
/* ------------------------------------------------------------------ */

enum Layout {
    LAY_PLANAR,    /* gray / yuv / yuva planar, 1 or 2 bytes per sample        */
    LAY_YA8,       /* gray+alpha interleaved, 2 bytes per pixel                */
    LAY_BGR32,     /* bgr0 / bgra: one little-endian u32 per pixel             */
    LAY_GBRP,      /* planar G,B,R[,A], 16-bit containers                      */
    LAY_RGB48,     /* packed R,G,B[,A] 16-bit little-endian                    */
};

typedef struct PixFmt {
    const char *name;
    int layout;
    int depth;         /* native sample depth                                   */
    int hs, vs;        /* log2 chroma subsampling                               */
    int chroma;        /* has chroma planes                                     */
    int alpha;
    int nplanes;       /* memory planes                                         */
} PixFmt;

/* the encoder's pix_fmts list, libavcodec/ffv1enc.c:1333-1355 (native-endian
 * names resolved for a little-endian host) */
static const PixFmt pixfmts[] = {
    { "gray",        LAY_PLANAR, 8, 0, 0, 0, 0, 1 }, { "gray9le",     LAY_PLANAR, 9, 0, 0, 0, 0, 1 },
    { "gray10le",    LAY_PLANAR, 10, 0, 0, 0, 0, 1 }, { "gray12le",    LAY_PLANAR, 12, 0, 0, 0, 0, 1 },
    { "gray16le",    LAY_PLANAR, 16, 0, 0, 0, 0, 1 }, { "ya8",         LAY_YA8,    8, 0, 0, 0, 1, 1 },
    { "yuv444p",     LAY_PLANAR, 8, 0, 0, 1, 0, 3 }, { "yuv440p",     LAY_PLANAR, 8, 0, 1, 1, 0, 3 },
    { "yuv422p",     LAY_PLANAR, 8, 1, 0, 1, 0, 3 }, { "yuv420p",     LAY_PLANAR, 8, 1, 1, 1, 0, 3 },
    { "yuv411p",     LAY_PLANAR, 8, 2, 0, 1, 0, 3 }, { "yuv410p",     LAY_PLANAR, 8, 2, 2, 1, 0, 3 },
    { "yuva444p",    LAY_PLANAR, 8, 0, 0, 1, 1, 4 }, { "yuva422p",    LAY_PLANAR, 8, 1, 0, 1, 1, 4 },
    { "yuva420p",    LAY_PLANAR, 8, 1, 1, 1, 1, 4 },
    { "yuv444p9le",  LAY_PLANAR, 9, 0, 0, 1, 0, 3 }, { "yuv422p9le",  LAY_PLANAR, 9, 1, 0, 1, 0, 3 },
    { "yuv420p9le",  LAY_PLANAR, 9, 1, 1, 1, 0, 3 },
    { "yuv444p10le", LAY_PLANAR, 10, 0, 0, 1, 0, 3 }, { "yuv440p10le", LAY_PLANAR, 10, 0, 1, 1, 0, 3 },
    { "yuv422p10le", LAY_PLANAR, 10, 1, 0, 1, 0, 3 }, { "yuv420p10le", LAY_PLANAR, 10, 1, 1, 1, 0, 3 },
    { "yuv444p12le", LAY_PLANAR, 12, 0, 0, 1, 0, 3 }, { "yuv440p12le", LAY_PLANAR, 12, 0, 1, 1, 0, 3 },
    { "yuv422p12le", LAY_PLANAR, 12, 1, 0, 1, 0, 3 }, { "yuv420p12le", LAY_PLANAR, 12, 1, 1, 1, 0, 3 },
    { "yuv444p14le", LAY_PLANAR, 14, 0, 0, 1, 0, 3 }, { "yuv422p14le", LAY_PLANAR, 14, 1, 0, 1, 0, 3 },
    { "yuv420p14le", LAY_PLANAR, 14, 1, 1, 1, 0, 3 },
    { "yuv444p16le", LAY_PLANAR, 16, 0, 0, 1, 0, 3 }, { "yuv422p16le", LAY_PLANAR, 16, 1, 0, 1, 0, 3 },
    { "yuv420p16le", LAY_PLANAR, 16, 1, 1, 1, 0, 3 },
    { "yuva444p9le", LAY_PLANAR, 9, 0, 0, 1, 1, 4 }, { "yuva422p9le", LAY_PLANAR, 9, 1, 0, 1, 1, 4 },
    { "yuva420p9le", LAY_PLANAR, 9, 1, 1, 1, 1, 4 },
    { "yuva444p10le", LAY_PLANAR, 10, 0, 0, 1, 1, 4 }, { "yuva422p10le", LAY_PLANAR, 10, 1, 0, 1, 1, 4 },
    { "yuva420p10le", LAY_PLANAR, 10, 1, 1, 1, 1, 4 },
    { "yuva444p16le", LAY_PLANAR, 16, 0, 0, 1, 1, 4 }, { "yuva422p16le", LAY_PLANAR, 16, 1, 0, 1, 1, 4 },
    { "yuva420p16le", LAY_PLANAR, 16, 1, 1, 1, 1, 4 },
    { "bgr0",        LAY_BGR32,  8, 0, 0, 1, 0, 1 }, { "bgra",        LAY_BGR32,  8, 0, 0, 1, 1, 1 },
    { "gbrp9le",     LAY_GBRP,   9, 0, 0, 1, 0, 3 }, { "gbrp10le",    LAY_GBRP,  10, 0, 0, 1, 0, 3 },
    { "gbrp12le",    LAY_GBRP,  12, 0, 0, 1, 0, 3 }, { "gbrp14le",    LAY_GBRP,  14, 0, 0, 1, 0, 3 },
    { "gbrp16le",    LAY_GBRP,  16, 0, 0, 1, 0, 3 },
    { "gbrap10le",   LAY_GBRP,  10, 0, 0, 1, 1, 4 }, { "gbrap12le",   LAY_GBRP,  12, 0, 0, 1, 1, 4 },
    { "gbrap16le",   LAY_GBRP,  16, 0, 0, 1, 1, 4 },
    { "rgb48le",     LAY_RGB48, 16, 0, 0, 1, 0, 1 }, { "rgba64le",    LAY_RGB48, 16, 0, 0, 1, 1, 1 },
};

static const PixFmt *find_pixfmt(const char *name)
{
    size_t i;
    if (!name)
        return NULL;
    for (i = 0; i < sizeof(pixfmts) / sizeof(pixfmts[0]); i++)
        if (!strcmp(pixfmts[i].name, name))
            return &pixfmts[i];
    return NULL;
}

static int bytes_per_pixel(const PixFmt *pf, int plane)
{
    (void)plane;
    switch (pf->layout) {
    case LAY_YA8:   return 2;
    case LAY_BGR32: return 4;
    case LAY_RGB48: return pf->alpha ? 8 : 6;
    default:        return pf->depth > 8 ? 2 : 1;
    }
}

int ffv1o_plane_geometry(const char *pix_fmt, int w, int h, int plane, int *bytewidth, int *rows)
{
    const PixFmt *pf = find_pixfmt(pix_fmt);
    int sw = 0, sh = 0;
    if (!pf || plane >= pf->nplanes)
        return -1;
    if (pf->layout == LAY_PLANAR && (plane == 1 || plane == 2)) {
        sw = pf->hs;
        sh = pf->vs;
    }
    *bytewidth = CEIL_RSHIFT(w, sw) * bytes_per_pixel(pf, plane);
    *rows      = CEIL_RSHIFT(h, sh);
    return pf->nplanes;
}

/* ------------------------------------------------------------------ */
/* constant tables                                                     */
/* ------------------------------------------------------------------ */

/* The context quantisation tables are piecewise constant and odd-symmetric
 * (q[256-i] = -q[i]); they are restated as break points.  Reference:
 * quant5_10bit / quant5 / quant9_10bit / quant11, libavcodec/ffv1enc.c:44-118. */
static void fill_quant(int8_t q[256], const int *start, int levels)
{
    /* start[l] = first index (0..127) whose value is l */
    int i, l = 0;
    for (i = 0; i < 128; i++) {
        while (l + 1 < levels && i >= start[l + 1])
            l++;
        q[i] = (int8_t)l;
    }
    q[128] = (int8_t)-(levels - 1);
    for (i = 129; i < 256; i++)
        q[i] = (int8_t)-q[256 - i];
}

static void quant_table_sets(int bits, int16_t qt[2][MAX_CTX_INPUTS][256])
{
    static const int s11[]  = { 0, 1, 2, 5, 12, 35 };        /* quant11       */
    static const int s5[]   = { 0, 1, 4 };                   /* quant5        */
    static const int s9x[]  = { 0, 5, 13, 27, 56 };          /* quant9_10bit  */
    static const int s5x[]  = { 0, 11, 50 };                 /* quant5_10bit  */
    int8_t qa[256], qb[256];
    int i;
    if (bits <= 8) {
        fill_quant(qa, s11, 6);
        fill_quant(qb, s5, 3);
    } else {
        fill_quant(qa, s9x, 5);
        fill_quant(qb, s5x, 3);
    }
    memset(qt, 0, sizeof(int16_t) * 2 * MAX_CTX_INPUTS * 256);
    /* libavcodec/ffv1enc.c:730-752 */
    for (i = 0; i < 256; i++) {
        qt[0][0][i] = qa[i];
        qt[0][1][i] = 11 * qa[i];
        qt[0][2][i] = 11 * 11 * qa[i];
        qt[1][0][i] = qa[i];
        qt[1][1][i] = 11 * qa[i];
        qt[1][2][i] = 11 * 11 * qb[i];
        qt[1][3][i] = 5 * 11 * 11 * qb[i];
        qt[1][4][i] = 5 * 5 * 11 * 11 * qb[i];
    }
}

/* custom state transition table "ver2_state", libavcodec/ffv1enc.c:120-137 */
static const uint8_t custom_transition[256] = {
      0,  10,  10,  10,  10,  16,  16,  16,  28,  16,  16,  29,  42,  49,  20,  49,
     59,  25,  26,  26,  27,  31,  33,  33,  33,  34,  34,  37,  67,  38,  39,  39,
     40,  40,  41,  79,  43,  44,  45,  45,  48,  48,  64,  50,  51,  52,  88,  52,
     53,  74,  55,  57,  58,  58,  74,  60, 101,  61,  62,  84,  66,  66,  68,  69,
     87,  82,  71,  97,  73,  73,  82,  75, 111,  77,  94,  78,  87,  81,  83,  97,
     85,  83,  94,  86,  99,  89,  90,  99, 111,  92,  93, 134,  95,  98, 105,  98,
    105, 110, 102, 108, 102, 118, 103, 106, 106, 113, 109, 112, 114, 112, 116, 125,
    115, 116, 117, 117, 126, 119, 125, 121, 121, 123, 145, 124, 126, 131, 127, 129,
    165, 130, 132, 138, 133, 135, 145, 136, 137, 139, 146, 141, 143, 142, 144, 148,
    147, 155, 151, 149, 151, 150, 152, 157, 153, 154, 156, 168, 158, 162, 161, 160,
    172, 163, 169, 164, 166, 184, 167, 170, 177, 174, 171, 173, 182, 176, 180, 178,
    175, 189, 179, 181, 186, 183, 192, 185, 200, 187, 191, 188, 190, 197, 193, 196,
    197, 194, 195, 196, 198, 202, 199, 201, 210, 203, 207, 204, 205, 206, 208, 214,
    209, 211, 221, 212, 213, 215, 224, 216, 217, 218, 219, 220, 222, 228, 223, 225,
    226, 224, 227, 229, 240, 230, 231, 232, 233, 234, 235, 236, 238, 239, 237, 242,
    241, 243, 242, 244, 245, 246, 247, 248, 249, 250, 251, 252, 252, 253, 254, 255,
};

/* run-length exponents of the Golomb run mode, libavcodec/bitstream.c:39-46 */
static const uint8_t log2_run[41] = {
     0,  0,  0,  0,  1,  1,  1,  1,  2,  2,  2,  2,  3,  3,  3,  3,
     4,  4,  5,  5,  6,  6,  7,  7,  8,  9, 10, 11, 12, 13, 14, 15,
    16, 17, 18, 19, 20, 21, 22, 23, 24,
};

/* ------------------------------------------------------------------ */
/* CRC-32 (IEEE 0x04C11DB7, MSB first, no reflection, no final xor)    */
/* libavutil/crc.c:336 AV_CRC_32_IEEE with av_crc()                    */
/* ------------------------------------------------------------------ */
static uint32_t crc_tab[256];
static pthread_once_t crc_once = PTHREAD_ONCE_INIT;

static void crc_build(void)
{
    int i, j;
    for (i = 0; i < 256; i++) {
        uint32_t c = (uint32_t)i << 24;
        for (j = 0; j < 8; j++)
            c = (c << 1) ^ ((c & 0x80000000u) ? 0x04C11DB7u : 0);
        crc_tab[i] = c;
    }
}

/* av_crc keeps the register byte-swapped on little-endian hosts; the value it
 * returns for this polynomial equals bswap32 of the textbook MSB-first register.
 * ffv1 stores it with AV_WL32, so the bytes on the wire are the textbook
 * register in big-endian order.  ffv1o_crc32 returns the av_crc()-compatible
 * value (what AV_WL32 is applied to). */
static uint32_t bswap32(uint32_t v)
{
    return (v >> 24) | ((v >> 8) & 0xFF00) | ((v << 8) & 0xFF0000) | (v << 24);
}

uint32_t ffv1o_crc32(uint32_t crc, const uint8_t *buf, int len)
{
    uint32_t r;
    int i;
    pthread_once(&crc_once, crc_build);
    r = bswap32(crc);
    for (i = 0; i < len; i++)
        r = (r << 8) ^ crc_tab[(r >> 24) ^ buf[i]];
    return bswap32(r);
}

static void put_le32(uint8_t *p, uint32_t v)
{
    p[0] = v; p[1] = v >> 8; p[2] = v >> 16; p[3] = v >> 24;
}

/* ------------------------------------------------------------------ */
/* binary adaptive range coder                                         */
/* libavcodec/rangecoder.h:35-152, rangecoder.c:42-123                 */
/* ------------------------------------------------------------------ */
typedef struct Rac {
    int low, range;
    int pending;          /* held-back byte (-1 = none yet)             */
    int ff_run;           /* number of 0xFF bytes queued behind pending */
    uint8_t one[256], zero[256];
    uint8_t *start, *p, *end;
    int overread;
} Rac;

/* ff_build_rac_states(c, 0.05*(1LL<<32), 256-8), rangecoder.c:68-106 */
static void rac_default_tables(uint8_t one[256], uint8_t zero[256])
{
    const int64_t unit = (int64_t)1 << 32;
    const int factor = (int)(0.05 * (double)((int64_t)1 << 32));
    const int max_p = 256 - 8;
    int64_t p = unit / 2;
    int prev = 0, i;
    memset(one, 0, 256);
    memset(zero, 0, 256);
    for (i = 0; i < 128; i++) {
        int p8 = (int)((256 * p + unit / 2) >> 32);
        if (p8 <= prev)
            p8 = prev + 1;
        if (prev && prev < 256 && p8 <= max_p)
            one[prev] = (uint8_t)p8;
        p += ((unit - p) * factor + unit / 2) >> 32;
        prev = p8;
    }
    for (i = 256 - max_p; i <= max_p; i++) {
        int p8;
        if (one[i])
            continue;
        p  = (i * unit + 128) >> 8;
        p += ((unit - p) * factor + unit / 2) >> 32;
        p8 = (int)((256 * p + unit / 2) >> 32);
        if (p8 <= i)
            p8 = i + 1;
        if (p8 > max_p)
            p8 = max_p;
        one[i] = (uint8_t)p8;
    }
    for (i = 1; i < 255; i++)
        zero[i] = (uint8_t)(256 - one[256 - i]);
}

void ffv1o_default_state_transition(uint8_t one_state[256])
{
    uint8_t z[256];
    rac_default_tables(one_state, z);
}

/* custom table install: ffv1.c:95-101, ffv1enc.c:1213-1219 */
static void rac_set_custom(Rac *c, const uint8_t trans[256])
{
    int j;
    for (j = 1; j < 256; j++) {
        c->one[j]        = trans[j];
        c->zero[256 - j] = (uint8_t)(256 - c->one[j]);
    }
}

static void rac_enc_init(Rac *c, uint8_t *buf, int size)
{
    c->start = c->p = buf;
    c->end = buf + size;
    c->low = 0;
    c->range = 0xFF00;
    c->ff_run = 0;
    c->pending = -1;
    c->overread = 0;
}

/* renorm_encoder, rangecoder.h:71-94 */
static void rac_enc_shift(Rac *c)
{
    while (c->range < 0x100) {
        if (c->pending < 0) {
            c->pending = c->low >> 8;
        } else if (c->low <= 0xFF00) {
            *c->p++ = (uint8_t)c->pending;
            for (; c->ff_run; c->ff_run--)
                *c->p++ = 0xFF;
            c->pending = c->low >> 8;
        } else if (c->low >= 0x10000) {
            *c->p++ = (uint8_t)(c->pending + 1);
            for (; c->ff_run; c->ff_run--)
                *c->p++ = 0x00;
            c->pending = (c->low >> 8) & 0xFF;
        } else {
            c->ff_run++;
        }
        c->low = (c->low & 0xFF) << 8;
        c->range <<= 8;
    }
}

/* put_rac, rangecoder.h:104-121 */
static void rac_put(Rac *c, uint8_t *state, int bit)
{
    int r1 = (c->range * (*state)) >> 8;
    if (!bit) {
        c->range -= r1;
        *state = c->zero[*state];
    } else {
        c->low += c->range - r1;
        c->range = r1;
        *state = c->one[*state];
    }
    rac_enc_shift(c);
}

/* ff_rac_terminate, rangecoder.c:109-123; returns bytes written */
static int rac_enc_finish(Rac *c, int version)
{
    if (version == 1) {
        uint8_t s = 129;
        rac_put(c, &s, 0);
    }
    c->range = 0xFF;
    c->low += 0xFF;
    rac_enc_shift(c);
    c->range = 0xFF;
    rac_enc_shift(c);
    return (int)(c->p - c->start);
}

/* ff_init_range_decoder, rangecoder.c:53-66 */
static void rac_dec_init(Rac *c, const uint8_t *buf, int size)
{
    rac_enc_init(c, (uint8_t *)buf, size);
    c->low = (c->p[0] << 8) | c->p[1];
    c->p += 2;
    if (c->low >= 0xFF00) {
        c->low = 0xFF00;
        c->end = c->p;
    }
}

/* get_rac + refill, rangecoder.h:123-152 */
static int rac_get(Rac *c, uint8_t *state)
{
    int r1 = (c->range * (*state)) >> 8;
    int bit;
    c->range -= r1;
    if (c->low < c->range) {
        *state = c->zero[*state];
        bit = 0;
    } else {
        c->low -= c->range;
        *state = c->one[*state];
        c->range = r1;
        bit = 1;
    }
    if (c->range < 0x100) {
        c->range <<= 8;
        c->low <<= 8;
        if (c->p < c->end)
            c->low += *c->p++;
        else
            c->overread++;
    }
    return bit;
}

static int ilog2(unsigned v)
{
    int n = 0;
    while (v >>= 1)
        n++;
    return n;
}

/* put_symbol_inline, ffv1enc.c:185-231 */
static void sym_put(Rac *c, uint8_t *st, int v, int is_signed)
{
    int i;
    if (!v) {
        rac_put(c, st + 0, 1);
        return;
    }
    {
        const int a = v < 0 ? -v : v;
        const int e = ilog2((unsigned)a);
        rac_put(c, st + 0, 0);
        for (i = 0; i < e; i++)
            rac_put(c, st + 1 + MINV(i, 9), 1);
        rac_put(c, st + 1 + MINV(e, 9), 0);
        for (i = e - 1; i >= 0; i--)
            rac_put(c, st + 22 + MINV(i, 9), (a >> i) & 1);
        if (is_signed)
            rac_put(c, st + 11 + MINV(e, 10), v < 0);
    }
}

/* get_symbol_inline, ffv1dec.c:42-64 */
static int sym_get(Rac *c, uint8_t *st, int is_signed)
{
    int e = 0, i;
    unsigned a = 1;
    if (rac_get(c, st + 0))
        return 0;
    while (rac_get(c, st + 1 + MINV(e, 9))) {
        e++;
        if (e > 31)
            return FFV1O_INVALIDDATA;
    }
    for (i = e - 1; i >= 0; i--)
        a += a + rac_get(c, st + 22 + MINV(i, 9));
    if (is_signed && rac_get(c, st + 11 + MINV(e, 10)))
        return -(int)a;
    return (int)a;
}

/* ------------------------------------------------------------------ */
/* Golomb-Rice layer                                                    */
/* ------------------------------------------------------------------ */
typedef struct Vlc {          /* VlcState, ffv1.h:61-66 */
    int16_t drift;
    uint16_t error_sum;
    int8_t bias;
    uint8_t count;
} Vlc;

typedef struct BitW {         /* MSB-first writer, put_bits.h */
    uint8_t *start, *end;
    uint64_t nbits;           /* total bits appended */
} BitW;

static void bw_put(BitW *b, int n, unsigned v)
{
    int i;
    for (i = n - 1; i >= 0; i--) {
        uint64_t pos = b->nbits++;
        uint8_t *byte = b->start + (pos >> 3);
        int sh = 7 - (int)(pos & 7);
        if (sh == 7)
            *byte = 0;
        *byte |= (uint8_t)(((v >> i) & 1) << sh);
    }
}

typedef struct BitR {         /* MSB-first reader, get_bits.h (checked reader) */
    const uint8_t *buf;
    int64_t pos, size_bits;
} BitR;

/* 32 bits at the cursor; bytes past the end of the supplied range read as the
 * bytes that follow in memory in the reference (trailer / next slice / padding);
 * callers hand a padded buffer, so plain reads are safe here */
static uint32_t br_peek32(const BitR *r)
{
    const uint8_t *p = r->buf + (r->pos >> 3);
    uint64_t w = ((uint64_t)p[0] << 32) | ((uint64_t)p[1] << 24) | ((uint64_t)p[2] << 16) |
                 ((uint64_t)p[3] << 8) | p[4];
    return (uint32_t)(w >> (8 - (r->pos & 7)));
}

static void br_skip(BitR *r, int n)
{
    r->pos += n;
    if (r->pos > r->size_bits + 8)      /* CONFIG_SAFE_BITSTREAM_READER clamp */
        r->pos = r->size_bits + 8;
}

static unsigned br_get(BitR *r, int n)
{
    unsigned v;
    if (!n)
        return 0;
    v = br_peek32(r) >> (32 - n);
    br_skip(r, n);
    return v;
}

static int fold(int diff, int bits)       /* ffv1.h:151-160 */
{
    const unsigned m = 1u << (bits - 1);
    unsigned u = (unsigned)diff & ((m << 1) - 1);
    return (int)((u ^ m) - m);
}

/* update_vlc_state, ffv1.h:162-188 */
static void vlc_adapt(Vlc *s, int v)
{
    int drift = s->drift, count = s->count;
    s->error_sum += (uint16_t)(v < 0 ? -v : v);
    drift += v;
    if (count == 128) {
        count >>= 1;
        drift >>= 1;
        s->error_sum >>= 1;
    }
    count++;
    if (drift <= -count) {
        s->bias = (int8_t)MAXV(s->bias - 1, -128);
        drift = MAXV(drift + count, -count + 1);
    } else if (drift > 0) {
        s->bias = (int8_t)MINV(s->bias + 1, 127);
        drift = MINV(drift - count, 0);
    }
    s->drift = (int16_t)drift;
    s->count = (uint8_t)count;
}

static int vlc_k(const Vlc *s)
{
    int i = s->count, k = 0;
    while (i < s->error_sum) {
        k++;
        i += i;
    }
    return k;
}

/* put_vlc_symbol (ffv1enc.c:240-262) + set_sr_golomb/set_ur_golomb (golomb.h:676-731) */
static void vlc_put(BitW *b, Vlc *s, int v, int bits)
{
    int k, code, u, e;
    v = fold(v - s->bias, bits);
    k = vlc_k(s);
    code = v ^ ((2 * s->drift + s->count) >> 31);
    u = -2 * code - 1;
    u ^= u >> 31;
    e = u >> k;
    if (e < 12)
        bw_put(b, e + k + 1, (1u << k) + ((unsigned)u & ((1u << k) - 1)));
    else
        bw_put(b, 12 + bits, (unsigned)(u - 12 + 1));
    vlc_adapt(s, v);
}

/* get_vlc_symbol (ffv1dec.c:71-94) + get_sr_golomb/get_ur_golomb (golomb.h:373-413,529-534) */
static int vlc_get(BitR *r, Vlc *s, int bits)
{
    const int k = vlc_k(s);
    uint32_t buf = br_peek32(r);
    int log = ilog2(buf | 1);
    unsigned u;
    int v, ret;
    if (log > 31 - 12) {
        buf >>= log - k;
        buf += (uint32_t)(30 - log) << k;
        br_skip(r, 32 + k - log);
        u = buf;
    } else {
        br_skip(r, 12);
        u = br_get(r, bits) + 12 - 1;
    }
    v = (int)(u >> 1) ^ -(int)(u & 1);
    v ^= (2 * s->drift + s->count) >> 31;
    ret = fold(v + s->bias, bits);
    vlc_adapt(s, v);
    return ret;
}

/* ------------------------------------------------------------------ */
/* resolved stream description                                         */
/* ------------------------------------------------------------------ */
typedef struct Stream {
    int width, height;
    int version, micro_version;
    int ac;
    int colorspace;
    int bits;                    /* bits_per_raw_sample as coded (0 for v0)      */
    int chroma_planes, hs, vs, transparency;
    int nh, nv;
    int ec, intra;
    int qt_count;
    int16_t qt[MAX_QUANT_TABLES][MAX_CTX_INPUTS][256];
    int ctx_count[MAX_QUANT_TABLES];
    uint8_t (*initial[MAX_QUANT_TABLES])[CONTEXT_SIZE];  /* NULL = all 128 */
    uint8_t trans[256];          /* state_transition (one_state)                 */
    int plane_sets;              /* f->plane_count                                */
    int use32, packed_lsb;
    const PixFmt *pf;
    uint8_t def_one[256], def_zero[256];
} Stream;

typedef struct Slice {
    int x, y, w, h;              /* luma rectangle                               */
    int qidx[4];                 /* quant table index per plane set              */
    int nctx[4];
    uint8_t (*rstate[4])[CONTEXT_SIZE];
    Vlc *vstate[4];
    int damaged;
    /* per-frame scratch */
    Rac c;
    uint8_t *buf;                /* encoder: slice bitstream                      */
    int buf_cap, bytes;
    int err;
    int run_index;
    int ps, sar_num, sar_den;    /* decoded slice header                          */
} Slice;

static void slice_rect(const Stream *s, int i, Slice *sl)
{
    /* ff_ffv1_init_slice_contexts, ffv1.c:117-165 */
    int sx = i % s->nh, sy = i / s->nh;
    int x0 = s->width * sx / s->nh, x1 = s->width * (sx + 1) / s->nh;
    int y0 = s->height * sy / s->nv, y1 = s->height * (sy + 1) / s->nv;
    sl->x = x0; sl->w = x1 - x0;
    sl->y = y0; sl->h = y1 - y0;
}

static int slice_alloc_states(const Stream *s, Slice *sl)
{
    int j;
    for (j = 0; j < s->plane_sets; j++) {
        int n = s->ctx_count[sl->qidx[j]];
        if (sl->nctx[j] < n) {
            free(sl->rstate[j]);
            free(sl->vstate[j]);
            sl->rstate[j] = NULL;
            sl->vstate[j] = NULL;
        }
        sl->nctx[j] = n;
        if (s->ac != AC_GOLOMB) {
            if (!sl->rstate[j]) {
                sl->rstate[j] = malloc((size_t)n * CONTEXT_SIZE);
                if (!sl->rstate[j])
                    return FFV1O_ENOMEM;
                memset(sl->rstate[j], 128, (size_t)n * CONTEXT_SIZE);
            }
        } else if (!sl->vstate[j]) {
            int i;
            sl->vstate[j] = calloc(n, sizeof(Vlc));
            if (!sl->vstate[j])
                return FFV1O_ENOMEM;
            for (i = 0; i < n; i++) {
                sl->vstate[j][i].error_sum = 4;
                sl->vstate[j][i].count = 1;
            }
        }
    }
    return 0;
}

/* ff_ffv1_clear_slice_state, ffv1.c:182-207 */
static void slice_reset_states(const Stream *s, Slice *sl)
{
    int j, i;
    for (j = 0; j < s->plane_sets; j++) {
        int n = sl->nctx[j];
        if (s->ac != AC_GOLOMB) {
            if (s->initial[sl->qidx[j]])
                memcpy(sl->rstate[j], s->initial[sl->qidx[j]], (size_t)n * CONTEXT_SIZE);
            else
                memset(sl->rstate[j], 128, (size_t)n * CONTEXT_SIZE);
        } else {
            for (i = 0; i < n; i++) {
                Vlc *v = &sl->vstate[j][i];
                v->drift = 0;
                v->error_sum = 4;
                v->bias = 0;
                v->count = 1;
            }
        }
    }
}

static void slice_free(Slice *sl)
{
    int j;
    for (j = 0; j < 4; j++) {
        free(sl->rstate[j]);
        free(sl->vstate[j]);
    }
    free(sl->buf);
}

/* ------------------------------------------------------------------ */
/* sample geometry: one coded plane of one slice                        */
/* ------------------------------------------------------------------ */
typedef struct Plane {
    int w, h;
    int bits;            /* coding depth                                   */
    int set;             /* plane-context set (state/quant index)          */
    int32_t *s;          /* w*h samples, already wrapped to the coding type */
    /* where the samples live in the picture (YCbCr layouts only) */
    int mem;             /* memory plane                                   */
    int x0, y0;          /* origin in that plane, in samples               */
    int step, off;       /* bytes between samples, byte offset in a pixel  */
} Plane;

/* The reference keeps 2-3 rotating line buffers with 3 samples of padding and
 * patches two border cells per line (ffv1enc.c:284-289, ffv1dec.c:128-139).
 * Unrolled, the neighbourhood of sample (x,y) of a w-wide plane is:
 *   T  = S(x,y-1)                      0 on the first line
 *   L  = S(x-1,y)   ; x==0  -> T(0,y)  (the line above's first sample)
 *   LT = S(x-1,y-1) ; x==0  -> S(0,y-2), 0 on the first two lines
 *   RT = S(x+1,y-1) ; x==w-1-> S(w-1,y-1)
 *   LL = S(x-2,y)   ; x==1  -> L(0,y) ; x==0 -> 0
 *   TT = S(x,y-2)                      0 on the first two lines           */
static inline int at(const Plane *p, int x, int y)
{
    return (y < 0) ? 0 : p->s[(size_t)y * p->w + x];
}
static inline int nb_T(const Plane *p, int x, int y)  { return at(p, x, y - 1); }
static inline int nb_L(const Plane *p, int x, int y)  { return x ? at(p, x - 1, y) : at(p, 0, y - 1); }
static inline int nb_LT(const Plane *p, int x, int y) { return x ? at(p, x - 1, y - 1) : at(p, 0, y - 2); }
static inline int nb_RT(const Plane *p, int x, int y) { return at(p, MINV(x + 1, p->w - 1), y - 1); }
static inline int nb_LL(const Plane *p, int x, int y) { return x >= 2 ? at(p, x - 2, y) : (x == 1 ? nb_L(p, 0, y) : 0); }
static inline int nb_TT(const Plane *p, int x, int y) { return at(p, x, y - 2); }

static inline int median3(int a, int b, int c)   /* mid_pred, mathops.h:98-112 */
{
    int lo = MINV(a, b), hi = MAXV(a, b);
    return MAXV(lo, MINV(hi, c));
}

/* predict, ffv1_template.c:23-30 */
static inline int predict_at(const Plane *p, int x, int y)
{
    int L = nb_L(p, x, y), T = nb_T(p, x, y), LT = nb_LT(p, x, y);
    return median3(L, L + T - LT, T);
}

/* get_context, ffv1_template.c:32-52 */
static inline int context_at(const int16_t q[MAX_CTX_INPUTS][256], const Plane *p, int x, int y)
{
    int L = nb_L(p, x, y), T = nb_T(p, x, y), LT = nb_LT(p, x, y), RT = nb_RT(p, x, y);
    int c = q[0][(L - LT) & 0xFF] + q[1][(LT - T) & 0xFF] + q[2][(T - RT) & 0xFF];
    if (q[3][127] || q[4][127])
        c += q[3][(nb_LL(p, x, y) - L) & 0xFF] + q[4][(nb_TT(p, x, y) - T) & 0xFF];
    return c;
}

static inline int wrap_sample(const Stream *s, int v)
{
    return s->use32 ? v : (int16_t)v;
}

/* ------------------------------------------------------------------ */
/* per-plane coding                                                     */
/* ------------------------------------------------------------------ */

/* encode_line, ffv1enc_template.c:23-123, for every line of a plane (the
 * enclosing loops of encode_plane ffv1enc.c:274-312 / encode_rgb_frame
 * ffv1enc_template.c:125-201 are in encode_slice_pixels below) */
static void enc_line_range(const Stream *s, Slice *sl, const Plane *p, int y)
{
    const int16_t (*q)[256] = s->qt[sl->qidx[p->set]];
    int x;
    if (sl->c.end - sl->c.p < (ptrdiff_t)p->w * 35) {   /* "encoded frame too large" */
        sl->err = FFV1O_INVALIDDATA;
        return;
    }
    for (x = 0; x < p->w; x++) {
        int ctx = context_at(q, p, x, y);
        int diff = at(p, x, y) - predict_at(p, x, y);
        if (ctx < 0) {
            ctx = -ctx;
            diff = -diff;
        }
        diff = fold(diff, p->bits);
        sym_put(&sl->c, sl->rstate[p->set][ctx], diff, 1);
    }
}

static void enc_line_golomb(const Stream *s, Slice *sl, BitW *bw, const Plane *p, int y)
{
    const int16_t (*q)[256] = s->qt[sl->qidx[p->set]];
    int run_index = sl->run_index, run_count = 0, run_mode = 0;
    int x;
    if (bw->end - bw->start - (ptrdiff_t)(bw->nbits >> 3) < (ptrdiff_t)p->w * 4) {
        sl->err = FFV1O_INVALIDDATA;
        return;
    }
    for (x = 0; x < p->w; x++) {
        int ctx = context_at(q, p, x, y);
        int diff = at(p, x, y) - predict_at(p, x, y);
        if (ctx < 0) {
            ctx = -ctx;
            diff = -diff;
        }
        diff = fold(diff, p->bits);
        if (ctx == 0)
            run_mode = 1;
        if (run_mode) {
            if (diff) {
                while (run_count >= 1 << log2_run[run_index]) {
                    run_count -= 1 << log2_run[run_index];
                    run_index++;
                    bw_put(bw, 1, 1);
                }
                bw_put(bw, 1 + log2_run[run_index], (unsigned)run_count);
                if (run_index)
                    run_index--;
                run_count = 0;
                run_mode = 0;
                if (diff > 0)
                    diff--;
            } else {
                run_count++;
            }
        }
        if (!run_mode)
            vlc_put(bw, &sl->vstate[p->set][ctx], diff, p->bits);
    }
    if (run_mode) {
        while (run_count >= 1 << log2_run[run_index]) {
            run_count -= 1 << log2_run[run_index];
            run_index++;
            bw_put(bw, 1, 1);
        }
        if (run_count)
            bw_put(bw, 1, 1);
    }
    sl->run_index = run_index;
}

/* decode_line, ffv1dec_template.c:23-126 */
static int dec_line_range(const Stream *s, Slice *sl, Plane *p, int y)
{
    const int16_t (*q)[256] = s->qt[sl->qidx[p->set]];
    const unsigned mask = (1u << p->bits) - 1;
    int x;
    if (sl->c.overread > 2)
        return FFV1O_INVALIDDATA;
    for (x = 0; x < p->w; x++) {
        int ctx, sign = 0, diff;
        if (!(x & 1023) && sl->c.overread > 2)
            return FFV1O_INVALIDDATA;
        ctx = context_at(q, p, x, y);
        if (ctx < 0) {
            ctx = -ctx;
            sign = 1;
        }
        diff = sym_get(&sl->c, sl->rstate[p->set][ctx], 1);
        if (sign)
            diff = -diff;
        p->s[(size_t)y * p->w + x] =
            wrap_sample(s, (int)(((unsigned)predict_at(p, x, y) + (unsigned)diff) & mask));
    }
    return 0;
}

static int dec_line_golomb(const Stream *s, Slice *sl, BitR *br, Plane *p, int y)
{
    const int16_t (*q)[256] = s->qt[sl->qidx[p->set]];
    const unsigned mask = (1u << p->bits) - 1;
    int run_index = sl->run_index, run_count = 0, run_mode = 0;
    int x;
    if (br->size_bits - br->pos < 1)
        return FFV1O_INVALIDDATA;
    for (x = 0; x < p->w; x++) {
        int ctx, sign = 0, diff;
        if (!(x & 1023) && br->size_bits - br->pos < 1)
            return FFV1O_INVALIDDATA;
        ctx = context_at(q, p, x, y);
        if (ctx < 0) {
            ctx = -ctx;
            sign = 1;
        }
        if (ctx == 0 && run_mode == 0)
            run_mode = 1;
        if (run_mode) {
            if (run_count == 0 && run_mode == 1) {
                if (br_get(br, 1)) {
                    run_count = 1 << log2_run[run_index];
                    if (x + run_count <= p->w)
                        run_index++;
                } else {
                    run_count = log2_run[run_index] ? (int)br_get(br, log2_run[run_index]) : 0;
                    if (run_index)
                        run_index--;
                    run_mode = 2;
                }
            }
            /* a run of zero residuals: every sample equals its prediction (the
             * reference short-cuts the L==LT case to a copy of T, same value) */
            while (run_count > 1 && p->w - x > 1) {
                p->s[(size_t)y * p->w + x] = wrap_sample(s, predict_at(p, x, y));
                x++;
                run_count--;
            }
            run_count--;
            if (run_count < 0) {
                run_mode = 0;
                run_count = 0;
                /* the context of the sample that ends the run: the reference
                 * keeps the one computed before the run was expanded */
                diff = vlc_get(br, &sl->vstate[p->set][ctx], p->bits);
                if (diff >= 0)
                    diff++;
            } else {
                diff = 0;
            }
        } else {
            diff = vlc_get(br, &sl->vstate[p->set][ctx], p->bits);
        }
        if (sign)
            diff = -diff;
        p->s[(size_t)y * p->w + x] =
            wrap_sample(s, (int)(((unsigned)predict_at(p, x, y) + (unsigned)diff) & mask));
    }
    sl->run_index = run_index;
    return 0;
}

/* ------------------------------------------------------------------ */
/* slice <-> picture sample transfer                                    */
/* ------------------------------------------------------------------ */
typedef struct Picture {
    uint8_t *data[4];
    int linesize[4];
} Picture;

static inline unsigned rd16(const uint8_t *p) { return p[0] | (p[1] << 8); }
static inline void wr16(uint8_t *p, unsigned v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }

/* how many coded planes a slice has and their geometry/order:
 * encode_slice ffv1enc.c:1083-1104 / decode_slice ffv1dec.c:322-350 */
static int slice_planes(const Stream *s, const Slice *sl, Plane pl[4])
{
    int n = 0, i;
    memset(pl, 0, 4 * sizeof(Plane));
    if (s->colorspace == 0) {
        const int cw = CEIL_RSHIFT(sl->w, s->hs), ch = CEIL_RSHIFT(sl->h, s->vs);
        const int bits = s->bits <= 8 ? 8 : s->bits;
        const int step = s->bits > 8 ? 2 : 1;
        if (s->pf->layout == LAY_YA8) {
            pl[n++] = (Plane){ sl->w, sl->h, bits, 0, NULL, 0, sl->x, sl->y, 2, 0 };
            pl[n++] = (Plane){ sl->w, sl->h, bits, 1, NULL, 0, sl->x, sl->y, 2, 1 };
        } else {
            pl[n++] = (Plane){ sl->w, sl->h, bits, 0, NULL, 0, sl->x, sl->y, step, 0 };
            if (s->chroma_planes) {
                pl[n++] = (Plane){ cw, ch, bits, 1, NULL, 1, sl->x >> s->hs, sl->y >> s->vs, step, 0 };
                pl[n++] = (Plane){ cw, ch, bits, 1, NULL, 2, sl->x >> s->hs, sl->y >> s->vs, step, 0 };
            }
            if (s->transparency)
                pl[n++] = (Plane){ sl->w, sl->h, bits, 2, NULL, 3, sl->x, sl->y, step, 0 };
        }
    } else {
        /* G, B, R [, A] with sets 0,1,1,2 ; 9 bits for 8-bit input else bits+1 */
        const int bits = s->bits <= 8 ? 9 : s->bits + 1;
        for (i = 0; i < 3 + s->transparency; i++)
            pl[n++] = (Plane){ sl->w, sl->h, bits, (i + 1) / 2, NULL, 0, 0, 0, 0, 0 };
    }
    return n;
}

/* encoder side: picture -> wrapped coding samples.
 * encode_plane ffv1enc.c:291-305, encode_rgb_frame ffv1enc_template.c:150-186 */
static void load_slice(const Stream *s, const Slice *sl, const Picture *pic, Plane pl[4], int n)
{
    int x, y, k;
    if (s->colorspace == 0) {
        for (k = 0; k < n; k++) {
            Plane *p = &pl[k];
            for (y = 0; y < p->h; y++) {
                const uint8_t *row = pic->data[p->mem] + (size_t)(p->y0 + y) * pic->linesize[p->mem] +
                                     (size_t)p->x0 * p->step + p->off;
                for (x = 0; x < p->w; x++) {
                    int v;
                    if (s->bits <= 8)
                        v = row[x * p->step];
                    else if (s->packed_lsb)
                        v = (int)rd16(row + 2 * x);
                    else
                        v = (int)(rd16(row + 2 * x) >> (16 - s->bits));
                    p->s[(size_t)y * p->w + x] = (int16_t)v;
                }
            }
        }
    } else {
        const int bits = s->bits > 0 ? s->bits : 8;
        const int offset = 1 << bits;
        for (y = 0; y < sl->h; y++) {
            for (x = 0; x < sl->w; x++) {
                int r, g, b, a = 0;
                const int X = sl->x + x, Y = sl->y + y;
                if (s->pf->layout == LAY_BGR32) {
                    const uint8_t *q = pic->data[0] + (size_t)Y * pic->linesize[0] + 4 * X;
                    b = q[0]; g = q[1]; r = q[2]; a = q[3];
                } else if (s->pf->layout == LAY_RGB48) {
                    const int ps = s->transparency ? 8 : 6;
                    const uint8_t *q = pic->data[0] + (size_t)Y * pic->linesize[0] + ps * X;
                    r = (int)rd16(q); g = (int)rd16(q + 2); b = (int)rd16(q + 4);
                    if (s->transparency)
                        a = (int)rd16(q + 6);
                } else if (s->use32 || s->transparency) {
                    g = (int)rd16(pic->data[0] + (size_t)Y * pic->linesize[0] + 2 * X);
                    b = (int)rd16(pic->data[1] + (size_t)Y * pic->linesize[1] + 2 * X);
                    r = (int)rd16(pic->data[2] + (size_t)Y * pic->linesize[2] + 2 * X);
                    if (s->transparency)
                        a = (int)rd16(pic->data[3] + (size_t)Y * pic->linesize[3] + 2 * X);
                } else {
                    /* ffv1enc_template.c:169-173: for <16-bit GBRP without alpha the
                     * reference reads plane 0 as "b" and plane 1 as "g" */
                    b = (int)rd16(pic->data[0] + (size_t)Y * pic->linesize[0] + 2 * X);
                    g = (int)rd16(pic->data[1] + (size_t)Y * pic->linesize[1] + 2 * X);
                    r = (int)rd16(pic->data[2] + (size_t)Y * pic->linesize[2] + 2 * X);
                }
                b -= g;
                r -= g;
                g += (b + r) >> 2;          /* slice_rct_by_coef = slice_rct_ry_coef = 1 (v<=3) */
                b += offset;
                r += offset;
                pl[0].s[(size_t)y * sl->w + x] = wrap_sample(s, g);
                pl[1].s[(size_t)y * sl->w + x] = wrap_sample(s, b);
                pl[2].s[(size_t)y * sl->w + x] = wrap_sample(s, r);
                if (n > 3)
                    pl[3].s[(size_t)y * sl->w + x] = wrap_sample(s, a);
            }
        }
    }
}

/* decoder side: one decoded line -> picture.
 * decode_plane ffv1dec.c:142-161, decode_rgb_frame ffv1dec_template.c:160-190 */
static void store_yuv_line(const Stream *s, Picture *pic, const Plane *p, int y)
{
    uint8_t *row = pic->data[p->mem] + (size_t)(p->y0 + y) * pic->linesize[p->mem] +
                   (size_t)p->x0 * p->step + p->off;
    int x;
    for (x = 0; x < p->w; x++) {
        int v = p->s[(size_t)y * p->w + x];
        if (s->bits <= 8)
            row[x * p->step] = (uint8_t)v;
        else if (s->packed_lsb)
            wr16(row + 2 * x, (uint16_t)v);
        else
            wr16(row + 2 * x, (uint16_t)((v << (16 - s->bits)) | ((uint16_t)v >> (2 * s->bits - 16))));
    }
}

static void store_rgb_line(const Stream *s, const Slice *sl, Picture *pic, Plane pl[4], int n, int y)
{
    const int bits = s->bits > 0 ? s->bits : 8;
    const int offset = 1 << bits;
    int x;
    for (x = 0; x < sl->w; x++) {
        int g = pl[0].s[(size_t)y * sl->w + x];
        int b = pl[1].s[(size_t)y * sl->w + x];
        int r = pl[2].s[(size_t)y * sl->w + x];
        int a = n > 3 ? pl[3].s[(size_t)y * sl->w + x] : 0;
        const int X = sl->x + x, Y = sl->y + y;
        b -= offset;
        r -= offset;
        g -= (b + r) >> 2;
        b += g;
        r += g;
        if (s->bits <= 8) {
            uint8_t *q = pic->data[0] + (size_t)Y * pic->linesize[0] + 4 * X;
            unsigned v = (unsigned)b + ((unsigned)g << 8) + ((unsigned)r << 16) + ((unsigned)a << 24);
            put_le32(q, v);
        } else if (s->use32 || s->transparency) {
            wr16(pic->data[0] + (size_t)Y * pic->linesize[0] + 2 * X, (uint16_t)g);
            wr16(pic->data[1] + (size_t)Y * pic->linesize[1] + 2 * X, (uint16_t)b);
            wr16(pic->data[2] + (size_t)Y * pic->linesize[2] + 2 * X, (uint16_t)r);
            if (s->transparency)
                wr16(pic->data[3] + (size_t)Y * pic->linesize[3] + 2 * X, (uint16_t)a);
        } else {
            wr16(pic->data[0] + (size_t)Y * pic->linesize[0] + 2 * X, (uint16_t)b);
            wr16(pic->data[1] + (size_t)Y * pic->linesize[1] + 2 * X, (uint16_t)g);
            wr16(pic->data[2] + (size_t)Y * pic->linesize[2] + 2 * X, (uint16_t)r);
        }
    }
}

static int alloc_planes(Plane pl[4], int n)
{
    int k;
    for (k = 0; k < n; k++) {
        pl[k].s = calloc((size_t)pl[k].w * pl[k].h + 1, sizeof(int32_t));
        if (!pl[k].s)
            return FFV1O_ENOMEM;
    }
    return 0;
}

static void free_planes(Plane pl[4], int n)
{
    int k;
    for (k = 0; k < n; k++)
        free(pl[k].s);
}

/* ------------------------------------------------------------------ */
/* thread fan-out over slices (avctx->execute, pthread_slice.c:95)       */
/* ------------------------------------------------------------------ */
typedef struct Fan {
    void (*fn)(void *ctx, int i);
    void *ctx;
    int count, next;
} Fan;

static void *fan_worker(void *v)
{
    Fan *f = v;
    for (;;) {
        int i = __atomic_fetch_add(&f->next, 1, __ATOMIC_RELAXED);
        if (i >= f->count)
            break;
        f->fn(f->ctx, i);
    }
    return NULL;
}

static void fan_out(int threads, int count, void (*fn)(void *, int), void *ctx)
{
    Fan f = { fn, ctx, count, 0 };
    pthread_t th[64];
    int n = MINV(MINV(threads, count), 64) - 1, i;
    for (i = 0; i < n; i++)
        if (pthread_create(&th[i], NULL, fan_worker, &f)) {
            n = i;
            break;
        }
    fan_worker(&f);
    for (i = 0; i < n; i++)
        pthread_join(th[i], NULL);
}

/* ------------------------------------------------------------------ */
/* encoder                                                              */
/* ------------------------------------------------------------------ */
struct FFV1OEncoder {
    Stream s;
    int gop_size, threads;
    int picture_number;
    int key_frame;
    Slice *sl;
    int nslices;
    uint8_t *extradata;
    int extradata_size;
    const Picture *cur;
};

static void write_quant_table(Rac *c, const int16_t *q)
{
    /* ffv1enc.c:314-327 */
    uint8_t st[CONTEXT_SIZE];
    int last = 0, i;
    memset(st, 128, sizeof(st));
    for (i = 1; i < 128; i++)
        if (q[i] != q[i - 1]) {
            sym_put(c, st, i - last - 1, 0);
            last = i;
        }
    sym_put(c, st, i - last - 1, 0);
}

/* write_extradata, ffv1enc.c:396-467 */
static int make_extradata(FFV1OEncoder *e)
{
    Stream *s = &e->s;
    Rac c;
    uint8_t st[CONTEXT_SIZE];
    int i, j, n;
    uint32_t crc;
    uint8_t *buf = malloc(1 << 16);
    if (!buf)
        return FFV1O_ENOMEM;
    memset(st, 128, sizeof(st));
    rac_enc_init(&c, buf, 1 << 16);
    memcpy(c.one, s->def_one, 256);
    memcpy(c.zero, s->def_zero, 256);
    sym_put(&c, st, s->version, 0);
    if (s->version > 2) {
        s->micro_version = s->version == 3 ? 4 : 2;
        sym_put(&c, st, s->micro_version, 0);
    }
    sym_put(&c, st, s->ac, 0);
    if (s->ac == AC_CUSTOM)
        for (i = 1; i < 256; i++)
            sym_put(&c, st, s->trans[i] - c.one[i], 1);
    sym_put(&c, st, s->colorspace, 0);
    sym_put(&c, st, s->bits, 0);
    rac_put(&c, st, s->chroma_planes);
    sym_put(&c, st, s->hs, 0);
    sym_put(&c, st, s->vs, 0);
    rac_put(&c, st, s->transparency);
    sym_put(&c, st, s->nh - 1, 0);
    sym_put(&c, st, s->nv - 1, 0);
    sym_put(&c, st, s->qt_count, 0);
    for (i = 0; i < s->qt_count; i++)
        for (j = 0; j < 5; j++)
            write_quant_table(&c, s->qt[i][j]);
    for (i = 0; i < s->qt_count; i++)
        rac_put(&c, st, 0);                 /* no 2-pass initial states */
    if (s->version > 2) {
        sym_put(&c, st, s->ec, 0);
        s->intra = e->gop_size < 2;
        sym_put(&c, st, s->intra, 0);
    }
    n = rac_enc_finish(&c, 0);
    crc = ffv1o_crc32(0, buf, n);
    put_le32(buf + n, crc);
    e->extradata = buf;
    e->extradata_size = n + 4;
    return 0;
}

/* encode_init, ffv1enc.c:517-928 (2-pass paths excluded) */
FFV1OEncoder *ffv1o_encoder_open(const FFV1OOptions *o, int *err)
{
    FFV1OEncoder *e = calloc(1, sizeof(*e));
    Stream *s;
    const PixFmt *pf = find_pixfmt(o->pix_fmt);
    int16_t qsets[2][MAX_CTX_INPUTS][256];
    int i, ac = o->coder, ec = o->slicecrc;
    *err = 0;
    if (!e) {
        *err = FFV1O_ENOMEM;
        return NULL;
    }
    s = &e->s;
    if (!o->width || !o->height) {
        *err = FFV1O_INVALIDDATA;
        goto fail;
    }
    s->pf = pf;
    s->width = o->width;
    s->height = o->height;
    s->nh = s->nv = 1;
    e->gop_size = o->gop_size;
    e->threads = o->threads > 0 ? o->threads : 1;
    rac_default_tables(s->def_one, s->def_zero);

    /* version selection, ffv1enc.c:526-558 */
    s->version = 0;
    if (o->slices > 1)
        s->version = MAXV(s->version, 2);
    if (o->slices == 0 && o->level < 0 && o->width * o->height > 720 * 576)
        s->version = MAXV(s->version, 2);
    if (o->level <= 0 && s->version == 2)
        s->version = 3;
    if (o->level >= 0 && o->level <= 4) {
        if (o->level < s->version) {
            *err = FFV1O_EINVAL;
            goto fail;
        }
        s->version = o->level;
    }
    if (ec < 0)
        ec = s->version >= 3;
    if (ec)
        s->version = MAXV(s->version, 3);
    s->ec = ec;
    if ((s->version == 2 || s->version > 3) && o->strict > -2) {
        *err = FFV1O_INVALIDDATA;
        goto fail;
    }
    /* coder normalisation, ffv1enc.c:560-570 */
    if (ac == 1)
        ac = AC_CUSTOM;
    else if (ac == -2)
        ac = AC_DEFAULT;

    /* pixel format, ffv1enc.c:572-699 */
    if (!pf) {
        *err = FFV1O_ENOSYS;                 /* "format not supported" */
        goto fail;
    }
    if (pf->layout == LAY_PLANAR || pf->layout == LAY_YA8) {
        s->colorspace = 0;
        s->chroma_planes = pf->chroma;
        s->transparency = pf->alpha;
        if (pf->depth > 8) {
            s->bits = o->bits_per_raw_sample ? o->bits_per_raw_sample : pf->depth;
            s->packed_lsb = pf->depth < 16;
            if (s->bits <= 8) {
                *err = FFV1O_INVALIDDATA;
                goto fail;
            }
            s->version = MAXV(s->version, 1);
        } else {
            s->bits = 8;
        }
    } else {
        s->colorspace = 1;
        s->chroma_planes = 1;
        s->transparency = pf->alpha;
        if (pf->layout == LAY_BGR32) {
            s->bits = 8;
        } else if (pf->layout == LAY_RGB48) {
            s->bits = 16;
            s->use32 = 1;
            s->version = MAXV(s->version, 1);
        } else {
            s->bits = o->bits_per_raw_sample ? o->bits_per_raw_sample : pf->depth;
            s->use32 = s->bits >= 16;
            s->version = MAXV(s->version, 1);
        }
    }
    if (s->bits > 8 && ac == AC_GOLOMB)
        ac = AC_CUSTOM;                       /* ffv1enc.c:702-708 */
    s->ac = ac;
    if ((unsigned)o->context > 1) {
        *err = FFV1O_EINVAL;
        goto fail;
    }
    if (s->version == 2 || s->version > 3) {  /* not restated: experimental versions */
        *err = FFV1O_ENOSYS;
        goto fail;
    }

    /* state transition table, ffv1enc.c:720-728 */
    if (ac == AC_CUSTOM)
        memcpy(s->trans, custom_transition, 256);
    else
        memcpy(s->trans, s->def_one, 256);

    quant_table_sets(s->bits, qsets);
    s->qt_count = 2;
    memcpy(s->qt[0], qsets[0], sizeof(qsets[0]));
    memcpy(s->qt[1], qsets[1], sizeof(qsets[1]));
    s->ctx_count[0] = (11 * 11 * 11 + 1) / 2;
    s->ctx_count[1] = (11 * 11 * 5 * 5 * 5 + 1) / 2;

    s->plane_sets = s->transparency ? 3 : 2;
    s->hs = pf->layout == LAY_PLANAR ? pf->hs : 0;
    s->vs = pf->layout == LAY_PLANAR ? pf->vs : 0;

    /* slice grid, ffv1enc.c:875-903 */
    if (s->version > 1) {
        const int planes = 1 + 2 * s->chroma_planes + s->transparency;
        const int max_h = CEIL_RSHIFT(o->width, s->hs), max_v = CEIL_RSHIFT(o->height, s->vs);
        int ok = 0;
        s->nv = (o->width > 352 || o->height > 288 || !o->slices) ? 2 : 1;
        s->nv = MINV(s->nv, max_v);
        for (; s->nv < 32 && !ok; s->nv++) {
            for (s->nh = s->nv; s->nh < 2 * s->nv; s->nh++) {
                int maxw = (o->width + s->nh - 1) / s->nh;
                int maxh = (o->height + s->nv - 1) / s->nv;
                if (s->nh > max_h || s->nv > max_v)
                    continue;
                if (maxw * maxh * (int64_t)(s->bits + 1) * planes > 8 << 24)
                    continue;
                if ((o->slices == s->nh * s->nv && o->slices <= MAX_SLICES) || !o->slices) {
                    ok = 1;
                    break;
                }
            }
            if (ok)
                break;
        }
        if (!ok) {
            *err = FFV1O_ENOSYS;
            goto fail;
        }
        if ((*err = make_extradata(e)) < 0)
            goto fail;
    }

    e->nslices = s->nh * s->nv;
    e->sl = calloc(e->nslices, sizeof(Slice));
    if (!e->sl) {
        *err = FFV1O_ENOMEM;
        goto fail;
    }
    for (i = 0; i < e->nslices; i++) {
        Slice *sl = &e->sl[i];
        int j;
        slice_rect(s, i, sl);
        for (j = 0; j < 4; j++)
            sl->qidx[j] = o->context;
        if ((*err = slice_alloc_states(s, sl)) < 0)
            goto fail;
    }
    return e;
fail:
    ffv1o_encoder_close(e);
    return NULL;
}

int ffv1o_encoder_extradata(FFV1OEncoder *e, const uint8_t **data)
{
    *data = e->extradata;
    return e->extradata_size;
}

void ffv1o_encoder_info(FFV1OEncoder *e, int info[8])
{
    info[0] = e->s.version;
    info[1] = e->s.micro_version;
    info[2] = e->s.ac;
    info[3] = e->s.nh;
    info[4] = e->s.nv;
    info[5] = e->s.ec;
    info[6] = e->s.bits;
    info[7] = e->s.colorspace;
}

void ffv1o_encoder_close(FFV1OEncoder *e)
{
    int i;
    if (!e)
        return;
    for (i = 0; i < e->nslices && e->sl; i++)
        slice_free(&e->sl[i]);
    free(e->sl);
    free(e->extradata);
    free(e);
}

/* write_header for version < 2, ffv1enc.c:348-376 */
static void write_frame_header_v01(const Stream *s, Rac *c, int context_model)
{
    uint8_t st[CONTEXT_SIZE];
    int i, j;
    memset(st, 128, sizeof(st));
    sym_put(c, st, s->version, 0);
    sym_put(c, st, s->ac, 0);
    if (s->ac == AC_CUSTOM)
        for (i = 1; i < 256; i++)
            sym_put(c, st, s->trans[i] - c->one[i], 1);
    sym_put(c, st, s->colorspace, 0);
    if (s->version > 0)
        sym_put(c, st, s->bits, 0);
    rac_put(c, st, s->chroma_planes);
    sym_put(c, st, s->hs, 0);
    sym_put(c, st, s->vs, 0);
    rac_put(c, st, s->transparency);
    for (j = 0; j < 5; j++)
        write_quant_table(c, s->qt[context_model][j]);
}

/* encode_slice_header, ffv1enc.c:930-961 (progressive, SAR 0/1 as the harness feeds) */
static void write_slice_header(const Stream *s, Slice *sl)
{
    uint8_t st[CONTEXT_SIZE];
    int j;
    memset(st, 128, sizeof(st));
    sym_put(&sl->c, st, (sl->x + 1) * s->nh / s->width, 0);
    sym_put(&sl->c, st, (sl->y + 1) * s->nv / s->height, 0);
    sym_put(&sl->c, st, (sl->w + 1) * s->nh / s->width - 1, 0);
    sym_put(&sl->c, st, (sl->h + 1) * s->nv / s->height - 1, 0);
    for (j = 0; j < s->plane_sets; j++)
        sym_put(&sl->c, st, sl->qidx[j], 0);
    sym_put(&sl->c, st, 3, 0);      /* progressive */
    sym_put(&sl->c, st, 0, 0);      /* sample_aspect_ratio.num */
    sym_put(&sl->c, st, 1, 0);      /* sample_aspect_ratio.den */
}

/* encode_slice, ffv1enc.c:1045-1120 (coder already initialised by the frame driver) */
static void encode_one_slice(void *ctx, int i)
{
    FFV1OEncoder *e = ctx;
    const Stream *s = &e->s;
    Slice *sl = &e->sl[i];
    Plane pl[4];
    BitW bw = { 0 };
    int n, k, y, ac_bytes = 0;

    if (e->key_frame)
        slice_reset_states(s, sl);
    if (s->version > 2)
        write_slice_header(s, sl);
    if (s->ac == AC_GOLOMB) {
        if (s->version > 2 || (!sl->x && !sl->y))
            ac_bytes = rac_enc_finish(&sl->c, s->version > 2);
        bw.start = sl->c.start + ac_bytes;
        bw.end = sl->c.end;
    }
    n = slice_planes(s, sl, pl);
    if (alloc_planes(pl, n) < 0) {
        sl->err = FFV1O_ENOMEM;
        free_planes(pl, n);
        return;
    }
    load_slice(s, sl, e->cur, pl, n);
    if (s->colorspace == 0) {
        for (k = 0; k < n; k++) {
            sl->run_index = 0;
            for (y = 0; y < pl[k].h; y++) {
                if (s->ac == AC_GOLOMB)
                    enc_line_golomb(s, sl, &bw, &pl[k], y);
                else
                    enc_line_range(s, sl, &pl[k], y);
            }
        }
    } else {
        sl->run_index = 0;
        for (y = 0; y < sl->h; y++)
            for (k = 0; k < n; k++) {
                if (s->ac == AC_GOLOMB)
                    enc_line_golomb(s, sl, &bw, &pl[k], y);
                else
                    enc_line_range(s, sl, &pl[k], y);
            }
    }
    free_planes(pl, n);
    /* termination, ffv1enc.c:1241-1247 */
    if (s->ac != AC_GOLOMB)
        sl->bytes = rac_enc_finish(&sl->c, 1);
    else
        sl->bytes = ac_bytes + (int)((bw.nbits + 7) / 8);
}

/* encode_frame, ffv1enc.c:1122-1281 */
int ffv1o_encode(FFV1OEncoder *e, const uint8_t *const planes[4], const int linesize[4],
                 uint8_t *out, int cap, int *key)
{
    Stream *s = &e->s;
    Picture pic;
    uint8_t keystate = 128;
    int i, total = 0;
    for (i = 0; i < 4; i++) {
        pic.data[i] = (uint8_t *)planes[i];
        pic.linesize[i] = linesize[i];
    }
    e->cur = &pic;
    e->key_frame = e->gop_size == 0 || e->picture_number % e->gop_size == 0;

    for (i = 0; i < e->nslices; i++) {
        Slice *sl = &e->sl[i];
        /* worst case 37 bytes per sample like the reference's packet bound; far
         * less is touched */
        int64_t need = 8192 + (int64_t)sl->w * 40 + (int64_t)sl->w * sl->h * 4 * 6;
        if (need > INT_MAX / 2)
            need = INT_MAX / 2;
        if (sl->buf_cap < need) {
            free(sl->buf);
            sl->buf = malloc(need);
            if (!sl->buf)
                return FFV1O_ENOMEM;
            sl->buf_cap = (int)need;
        }
        rac_enc_init(&sl->c, sl->buf, sl->buf_cap);
        memcpy(sl->c.one, s->def_one, 256);
        memcpy(sl->c.zero, s->def_zero, 256);
        sl->err = 0;
    }
    /* key-frame bit and (v<2) in-band header go into slice 0's coder with the
     * DEFAULT tables; the custom table is installed afterwards (ffv1enc.c:1203-1219) */
    rac_put(&e->sl[0].c, &keystate, e->key_frame);
    if (e->key_frame && s->version < 2)
        write_frame_header_v01(s, &e->sl[0].c, e->sl[0].qidx[0]);
    if (s->ac == AC_CUSTOM)
        for (i = 0; i < e->nslices; i++)
            rac_set_custom(&e->sl[i].c, s->trans);

    fan_out(e->threads, e->nslices, encode_one_slice, e);

    /* compaction + trailers, ffv1enc.c:1236-1262 */
    for (i = 0; i < e->nslices; i++) {
        Slice *sl = &e->sl[i];
        int bytes = sl->bytes;
        uint8_t *dst = out + total;
        if (sl->err)
            return sl->err;
        if (total + bytes + 8 > cap)
            return FFV1O_ENOSPC;
        memcpy(dst, sl->buf, bytes);
        if (i > 0 || s->version > 2) {
            dst[bytes] = (uint8_t)(bytes >> 16);
            dst[bytes + 1] = (uint8_t)(bytes >> 8);
            dst[bytes + 2] = (uint8_t)bytes;
            bytes += 3;
        }
        if (s->ec) {
            dst[bytes++] = 0;
            put_le32(dst + bytes, ffv1o_crc32(0, dst, bytes));
            bytes += 4;
        }
        total += bytes;
    }
    if (key)
        *key = e->key_frame;
    e->picture_number++;
    return total;
}

/* ------------------------------------------------------------------ */
/* decoder                                                              */
/* ------------------------------------------------------------------ */
struct FFV1ODecoder {
    Stream s;
    int threads;
    int have_extradata;
    int key_frame_ok;
    int key_frame;
    Slice *sl;
    int max_slices, nslices;
    uint8_t *frame[2];
    size_t frame_size;
    int cur_idx, have_last;
    Picture pic, last;
    char fmt_name[32];
    int v01_context_count;
    int damaged_count;
    const uint8_t *pkt;
};

static int read_quant_table(Rac *c, int16_t *q, int scale)
{
    /* ffv1dec.c:368-393 */
    uint8_t st[CONTEXT_SIZE];
    int v, i = 0;
    memset(st, 128, sizeof(st));
    for (v = 0; i < 128; v++) {
        unsigned len = (unsigned)sym_get(c, st, 0) + 1U;
        if (len > (unsigned)(128 - i) || !len)
            return FFV1O_INVALIDDATA;
        while (len--) {
            q[i] = (int16_t)(scale * v);
            i++;
        }
    }
    for (i = 1; i < 128; i++)
        q[256 - i] = (int16_t)-q[i];
    q[128] = (int16_t)-q[127];
    return 2 * v - 1;
}

static int read_quant_tables(Rac *c, int16_t q[MAX_CTX_INPUTS][256])
{
    /* ffv1dec.c:395-411 */
    int i, n = 1;
    for (i = 0; i < 5; i++) {
        int r = read_quant_table(c, q[i], n);
        if (r < 0)
            return r;
        n *= r;
        if ((unsigned)n > 32768U)
            return FFV1O_INVALIDDATA;
    }
    return (n + 1) / 2;
}

/* pixel format chosen by the decoder, read_header ffv1dec.c:597-739 */
static int pick_format(Stream *s)
{
    char name[32] = "";
    const int b = s->bits;
    const int sub = 16 * s->hs + s->vs;
    const char *yuv = NULL;
    s->packed_lsb = 0;
    s->use32 = 0;
    if (s->colorspace == 0) {
        switch (sub) {
        case 0x00: yuv = "444"; break;
        case 0x01: yuv = "440"; break;
        case 0x10: yuv = "422"; break;
        case 0x11: yuv = "420"; break;
        case 0x20: yuv = "411"; break;
        case 0x22: yuv = "410"; break;
        }
        if (!s->transparency && !s->chroma_planes) {
            if (b <= 8) strcpy(name, "gray");
            else if (b == 9 || b == 10 || b == 12) { s->packed_lsb = 1; strcpy(name, b == 9 ? "gray9le" : b == 10 ? "gray10le" : "gray12le"); }
            else if (b == 16) { s->packed_lsb = 1; strcpy(name, "gray16le"); }
            else if (b < 16) strcpy(name, "gray16le");
            else return FFV1O_ENOSYS;
        } else if (s->transparency && !s->chroma_planes) {
            if (b <= 8) strcpy(name, "ya8");
            else return FFV1O_ENOSYS;
        } else if (b <= 8) {
            if (yuv && (!s->transparency || sub == 0x00 || sub == 0x10 || sub == 0x11)) {
                strcpy(name, s->transparency ? "yuva" : "yuv");
                strcat(name, yuv);
                strcat(name, "p");
            }
        } else if (b == 9 || b == 10 || b == 12 || b == 14 || b == 16) {
            int ok = yuv && (sub == 0x00 || sub == 0x10 || sub == 0x11 ||
                             (sub == 0x01 && !s->transparency && (b == 10 || b == 12)));
            if (s->transparency && (b == 12 || b == 14))
                ok = 0;
            s->packed_lsb = 1;
            if (ok) {
                strcpy(name, s->transparency ? "yuva" : "yuv");
                strcat(name, yuv);
                strcat(name, b == 9 ? "p9le" : b == 10 ? "p10le" : b == 12 ? "p12le" :
                             b == 14 ? "p14le" : "p16le");
            }
        }
    } else if (s->colorspace == 1) {
        if (s->hs || s->vs)
            return FFV1O_ENOSYS;
        if (b <= 8) strcpy(name, s->transparency ? "bgra" : "bgr0");
        else if (b == 9 && !s->transparency) strcpy(name, "gbrp9le");
        else if (b == 10) strcpy(name, s->transparency ? "gbrap10le" : "gbrp10le");
        else if (b == 12) strcpy(name, s->transparency ? "gbrap12le" : "gbrp12le");
        else if (b == 14 && !s->transparency) strcpy(name, "gbrp14le");
        else if (b == 16) { strcpy(name, s->transparency ? "gbrap16le" : "gbrp16le"); s->use32 = 1; }
    } else {
        return FFV1O_ENOSYS;
    }
    s->pf = find_pixfmt(name);
    if (!s->pf)
        return FFV1O_ENOSYS;
    return 0;
}

/* read_extra_header, ffv1dec.c:413-528 */
static int parse_extradata(FFV1ODecoder *d, const uint8_t *data, int size)
{
    Stream *s = &d->s;
    Rac c;
    uint8_t st[CONTEXT_SIZE], st2[32][CONTEXT_SIZE];
    int i, j, k;
    memset(st, 128, sizeof(st));
    memset(st2, 128, sizeof(st2));
    if (size < 2)
        return FFV1O_INVALIDDATA;
    rac_dec_init(&c, data, size);
    memcpy(c.one, s->def_one, 256);
    memcpy(c.zero, s->def_zero, 256);
    s->version = sym_get(&c, st, 0);
    if (s->version < 2)
        return FFV1O_INVALIDDATA;
    if (s->version > 2) {
        c.end -= 4;
        s->micro_version = sym_get(&c, st, 0);
        if (s->micro_version < 0)
            return FFV1O_INVALIDDATA;
    }
    s->ac = sym_get(&c, st, 0);
    if (s->ac == AC_CUSTOM)
        for (i = 1; i < 256; i++)
            s->trans[i] = (uint8_t)(sym_get(&c, st, 1) + c.one[i]);
    s->colorspace = sym_get(&c, st, 0);
    s->bits = sym_get(&c, st, 0);
    s->chroma_planes = rac_get(&c, st);
    s->hs = sym_get(&c, st, 0);
    s->vs = sym_get(&c, st, 0);
    s->transparency = rac_get(&c, st);
    s->plane_sets = 1 + (s->chroma_planes || s->version < 4) + s->transparency;
    s->nh = 1 + sym_get(&c, st, 0);
    s->nv = 1 + sym_get(&c, st, 0);
    if ((unsigned)s->hs > 4U || (unsigned)s->vs > 4U)
        return FFV1O_INVALIDDATA;
    if ((unsigned)s->nh > (unsigned)s->width || !s->nh || (unsigned)s->nv > (unsigned)s->height || !s->nv)
        return FFV1O_INVALIDDATA;
    s->qt_count = sym_get(&c, st, 0);
    if ((unsigned)s->qt_count > MAX_QUANT_TABLES || !s->qt_count) {
        s->qt_count = 0;
        return FFV1O_INVALIDDATA;
    }
    for (i = 0; i < s->qt_count; i++) {
        s->ctx_count[i] = read_quant_tables(&c, s->qt[i]);
        if (s->ctx_count[i] < 0)
            return FFV1O_INVALIDDATA;
    }
    for (i = 0; i < s->qt_count; i++)
        if (rac_get(&c, st)) {
            s->initial[i] = malloc((size_t)s->ctx_count[i] * CONTEXT_SIZE);
            if (!s->initial[i])
                return FFV1O_ENOMEM;
            for (j = 0; j < s->ctx_count[i]; j++)
                for (k = 0; k < CONTEXT_SIZE; k++) {
                    int pred = j ? s->initial[i][j - 1][k] : 128;
                    s->initial[i][j][k] = (uint8_t)((pred + sym_get(&c, st2[k], 1)) & 0xFF);
                }
        }
    if (s->version > 2) {
        s->ec = sym_get(&c, st, 0);
        if (s->micro_version > 2)
            s->intra = sym_get(&c, st, 0);
        if (ffv1o_crc32(0, data, size) || size < 4)
            return FFV1O_INVALIDDATA;
    }
    if (s->version > 3)
        return FFV1O_ENOSYS;                 /* v4 not restated */
    return 0;
}

FFV1ODecoder *ffv1o_decoder_open(int width, int height, const uint8_t *extradata,
                                 int extradata_size, int threads, int *err)
{
    FFV1ODecoder *d = calloc(1, sizeof(*d));
    Stream *s;
    *err = 0;
    if (!d) {
        *err = FFV1O_ENOMEM;
        return NULL;
    }
    s = &d->s;
    if (!width || !height) {
        *err = FFV1O_INVALIDDATA;
        goto fail;
    }
    s->width = width;
    s->height = height;
    s->nh = s->nv = 1;
    d->threads = threads > 0 ? threads : 1;
    rac_default_tables(s->def_one, s->def_zero);
    if (extradata_size > 0) {
        if ((*err = parse_extradata(d, extradata, extradata_size)) < 0)
            goto fail;
        d->have_extradata = 1;
    }
    d->max_slices = s->nh * s->nv;
    d->sl = calloc(d->max_slices, sizeof(Slice));
    if (!d->sl) {
        *err = FFV1O_ENOMEM;
        goto fail;
    }
    return d;
fail:
    ffv1o_decoder_close(d);
    return NULL;
}

void ffv1o_decoder_close(FFV1ODecoder *d)
{
    int i;
    if (!d)
        return;
    for (i = 0; i < d->max_slices && d->sl; i++)
        slice_free(&d->sl[i]);
    for (i = 0; i < MAX_QUANT_TABLES; i++)
        free(d->s.initial[i]);
    free(d->sl);
    free(d->frame[0]);
    free(d->frame[1]);
    free(d);
}

int ffv1o_decoder_damaged(FFV1ODecoder *d)
{
    return d->damaged_count;
}

/* read_header, ffv1dec.c:530-816: v<2 in-band header; pixel format; slice count */
static int read_frame_header(FFV1ODecoder *d, Rac *c)
{
    Stream *s = &d->s;
    uint8_t st[CONTEXT_SIZE];
    int i, j, r;
    memset(st, 128, sizeof(st));
    if (s->version < 2) {
        int cs, bits, cp, hs, vs, tr;
        unsigned v = (unsigned)sym_get(c, st, 0);
        if (v >= 2)
            return FFV1O_INVALIDDATA;
        s->version = (int)v;
        s->ac = sym_get(c, st, 0);
        if (s->ac == AC_CUSTOM)
            for (i = 1; i < 256; i++) {
                int t = sym_get(c, st, 1) + c->one[i];
                if (t < 1 || t > 255)
                    return FFV1O_INVALIDDATA;
                s->trans[i] = (uint8_t)t;
            }
        cs = sym_get(c, st, 0);
        bits = s->version > 0 ? sym_get(c, st, 0) : 0;
        cp = rac_get(c, st);
        hs = sym_get(c, st, 0);
        vs = sym_get(c, st, 0);
        tr = rac_get(c, st);
        if (s->plane_sets &&
            (cs != s->colorspace || bits != s->bits || cp != s->chroma_planes ||
             hs != s->hs || vs != s->vs || tr != s->transparency))
            return FFV1O_INVALIDDATA;
        if ((unsigned)hs > 4U || (unsigned)vs > 4U)
            return FFV1O_INVALIDDATA;
        s->colorspace = cs;
        s->bits = bits;
        s->chroma_planes = cp;
        s->hs = hs;
        s->vs = vs;
        s->transparency = tr;
        s->plane_sets = 2 + tr;
    }
    if ((r = pick_format(s)) < 0)
        return r;
    if (s->version < 2) {
        int n = read_quant_tables(c, s->qt[0]);
        if (n < 0)
            return FFV1O_INVALIDDATA;
        s->qt_count = 1;
        s->ctx_count[0] = n;
        d->nslices = d->max_slices;
    } else {
        /* v3: walk the size trailers from the packet end, ffv1dec.c:746-756 */
        const int trailer = 3 + 5 * !!s->ec;
        const uint8_t *p = c->end;
        for (d->nslices = 0; d->nslices < MAX_SLICES && trailer < p - c->start; d->nslices++) {
            int size = (p[-trailer] << 16) | (p[-trailer + 1] << 8) | p[-trailer + 2];
            if (size + trailer > p - c->start)
                break;
            p -= size + trailer;
        }
    }
    if (d->nslices <= 0 || d->nslices > d->max_slices)
        return FFV1O_INVALIDDATA;
    for (j = 0; j < d->nslices; j++) {
        Slice *sl = &d->sl[j];
        sl->damaged = 0;
        if (s->version < 2) {
            for (i = 0; i < s->plane_sets; i++) {
                sl->qidx[i] = 0;
            }
            slice_rect(s, j, sl);
        }
    }
    return 0;
}

/* decode_slice_header, ffv1dec.c:167-244 */
static int read_slice_header(const Stream *s, Slice *sl)
{
    uint8_t st[CONTEXT_SIZE];
    Rac *c = &sl->c;
    int i;
    unsigned sx, sy, sw, sh;
    memset(st, 128, sizeof(st));
    sx = (unsigned)sym_get(c, st, 0) * s->width;
    sy = (unsigned)sym_get(c, st, 0) * s->height;
    sw = ((unsigned)sym_get(c, st, 0) + 1) * s->width + sx;
    sh = ((unsigned)sym_get(c, st, 0) + 1) * s->height + sy;
    sl->x = (int)sx / s->nh;
    sl->y = (int)sy / s->nv;
    sl->w = (int)sw / s->nh - sl->x;
    sl->h = (int)sh / s->nv - sl->y;
    if ((unsigned)sl->w > (unsigned)s->width || (unsigned)sl->h > (unsigned)s->height)
        return -1;
    if ((unsigned)sl->x + (uint64_t)sl->w > (unsigned)s->width ||
        (unsigned)sl->y + (uint64_t)sl->h > (unsigned)s->height)
        return -1;
    for (i = 0; i < s->plane_sets; i++) {
        int idx = sym_get(c, st, 0);
        if ((unsigned)idx >= (unsigned)s->qt_count)
            return -1;
        sl->qidx[i] = idx;
    }
    sl->ps = sym_get(c, st, 0);
    sl->sar_num = sym_get(c, st, 0);
    sl->sar_den = sym_get(c, st, 0);
    return 0;
}

/* decode_slice, ffv1dec.c:246-366 */
static void decode_one_slice(void *ctx, int i)
{
    FFV1ODecoder *d = ctx;
    const Stream *s = &d->s;
    Slice *sl = &d->sl[i];
    Plane pl[4];
    BitR br = { 0 };
    int n, k, y, r = 0;

    if (s->version > 2) {
        if (read_slice_header(s, sl) < 0) {
            sl->x = sl->y = sl->w = sl->h = 0;
            sl->damaged = 1;
            return;
        }
    }
    if (slice_alloc_states(s, sl) < 0) {
        sl->damaged = 1;
        return;
    }
    if (d->key_frame)
        slice_reset_states(s, sl);
    if (s->ac == AC_GOLOMB) {
        int ac_bytes;
        if ((s->version == 3 && s->micro_version > 1) || s->version > 3) {
            uint8_t t = 129;
            rac_get(&sl->c, &t);
        }
        ac_bytes = (s->version > 2 || (!sl->x && !sl->y)) ? (int)(sl->c.p - sl->c.start) - 1 : 0;
        br.buf = sl->c.start + ac_bytes;
        br.size_bits = (int64_t)(sl->c.end - sl->c.start - ac_bytes) * 8;
        br.pos = 0;
    }
    if (!sl->w || !sl->h)
        return;
    n = slice_planes(s, sl, pl);
    if (alloc_planes(pl, n) < 0) {
        free_planes(pl, n);
        sl->damaged = 1;
        return;
    }
    if (s->colorspace == 0) {
        for (k = 0; k < n; k++) {
            sl->run_index = 0;
            for (y = 0; y < pl[k].h; y++) {
                r = s->ac == AC_GOLOMB ? dec_line_golomb(s, sl, &br, &pl[k], y)
                                       : dec_line_range(s, sl, &pl[k], y);
                if (r < 0)
                    break;          /* decode_plane returns; the next plane still runs */
                store_yuv_line(s, &d->pic, &pl[k], y);
            }
        }
    } else {
        sl->run_index = 0;
        for (y = 0; y < sl->h && r >= 0; y++) {
            for (k = 0; k < n; k++) {
                r = s->ac == AC_GOLOMB ? dec_line_golomb(s, sl, &br, &pl[k], y)
                                       : dec_line_range(s, sl, &pl[k], y);
                if (r < 0)
                    break;
            }
            if (r >= 0)
                store_rgb_line(s, sl, &d->pic, pl, n, y);
        }
    }
    free_planes(pl, n);
    /* end-of-slice check, ffv1dec.c:351-359 */
    if (s->ac != AC_GOLOMB && s->version > 2) {
        uint8_t t = 129;
        int v;
        rac_get(&sl->c, &t);
        v = (int)(sl->c.end - sl->c.p) - 2 - 5 * s->ec;
        if (v)
            sl->damaged = 1;
    }
}

static int alloc_frames(FFV1ODecoder *d)
{
    const Stream *s = &d->s;
    size_t need = 0, off;
    int p, bw, rows, n = s->pf->nplanes;
    for (p = 0; p < n; p++) {
        ffv1o_plane_geometry(s->pf->name, s->width, s->height, p, &bw, &rows);
        need += (size_t)((bw + 63) & ~63) * rows + 64;
    }
    if (need > d->frame_size) {
        free(d->frame[0]);
        free(d->frame[1]);
        d->frame[0] = calloc(1, need);
        d->frame[1] = calloc(1, need);
        if (!d->frame[0] || !d->frame[1])
            return FFV1O_ENOMEM;
        d->frame_size = need;
        d->have_last = 0;
    }
    d->last = d->pic;
    d->cur_idx ^= 1;
    off = 0;
    memset(&d->pic, 0, sizeof(d->pic));
    for (p = 0; p < n; p++) {
        ffv1o_plane_geometry(s->pf->name, s->width, s->height, p, &bw, &rows);
        d->pic.data[p] = d->frame[d->cur_idx] + off;
        d->pic.linesize[p] = (bw + 63) & ~63;
        off += (size_t)d->pic.linesize[p] * rows + 64;
    }
    return 0;
}

/* decode_frame, ffv1dec.c:837-983 */
int ffv1o_decode(FFV1ODecoder *d, const uint8_t *pkt_in, int pkt_size,
                 uint8_t *planes[4], int linesize[4], const char **pix_fmt, int *key)
{
    Stream *s = &d->s;
    uint8_t keystate = 128;
    uint8_t *pkt;
    const uint8_t *end;
    Rac *c0;
    int i, r;

    if (pkt_size < 2)
        return FFV1O_INVALIDDATA;
    pkt = calloc(1, (size_t)pkt_size + 64);       /* AV_INPUT_BUFFER_PADDING_SIZE */
    if (!pkt)
        return FFV1O_ENOMEM;
    memcpy(pkt, pkt_in, pkt_size);

    c0 = &d->sl[0].c;
    rac_dec_init(c0, pkt, pkt_size);
    memcpy(c0->one, s->def_one, 256);
    memcpy(c0->zero, s->def_zero, 256);
    if (rac_get(c0, &keystate)) {
        d->key_frame = 1;
        d->key_frame_ok = 0;
        if ((r = read_frame_header(d, c0)) < 0)
            goto out;
        d->key_frame_ok = 1;
    } else {
        if (!d->key_frame_ok) {
            r = FFV1O_INVALIDDATA;
            goto out;
        }
        d->key_frame = 0;
    }
    if ((r = alloc_frames(d)) < 0)
        goto out;

    /* slice table from the packet tail, ffv1dec.c:890-931 */
    end = pkt + pkt_size;
    for (i = d->nslices - 1; i >= 0; i--) {
        Slice *sl = &d->sl[i];
        const int trailer = 3 + 5 * !!s->ec;
        int v;
        if (i || s->version > 2)
            v = ((end[-trailer] << 16) | (end[-trailer + 1] << 8) | end[-trailer + 2]) + trailer;
        else
            v = (int)(end - c0->start);
        if (end - c0->start < v) {
            r = FFV1O_INVALIDDATA;
            goto out;
        }
        end -= v;
        if (s->ec && ffv1o_crc32(0, end, v))
            sl->damaged = 1;
        if (i) {
            rac_dec_init(&sl->c, end, v);
            memcpy(sl->c.one, s->def_one, 256);
            memcpy(sl->c.zero, s->def_zero, 256);
        } else {
            sl->c.end = (uint8_t *)end + v;
        }
    }
    /* ff_ffv1_init_slice_state installs the custom table before the slice
     * header is parsed (ffv1dec.c:293-295, ffv1.c:93-102) */
    if (s->ac == AC_CUSTOM)
        for (i = 0; i < d->nslices; i++)
            rac_set_custom(&d->sl[i].c, s->trans);

    fan_out(d->threads, d->nslices, decode_one_slice, d);

    /* conceal damaged slices from the previous picture, ffv1dec.c:940-969 */
    d->damaged_count = 0;
    for (i = d->nslices - 1; i >= 0; i--) {
        Slice *sl = &d->sl[i];
        if (!sl->damaged)
            continue;
        d->damaged_count++;
        if (d->have_last) {
            int p, n = s->pf->nplanes;
            for (p = 0; p < n; p++) {
                int sh = (s->pf->layout == LAY_PLANAR && (p == 1 || p == 2)) ? s->hs : 0;
                int sv = (s->pf->layout == LAY_PLANAR && (p == 1 || p == 2)) ? s->vs : 0;
                int bpp = bytes_per_pixel(s->pf, p);
                /* the reference offsets by (x >> sh) << (depth > 8) bytes; for the
                 * packed RGB layouts that is not x*bytes_per_pixel -- kept as is */
                int pixshift = s->pf->depth > 8;
                int bw = CEIL_RSHIFT(sl->w, sh) * bpp, rows = CEIL_RSHIFT(sl->h, sv), y;
                size_t xo = (size_t)((sl->x >> sh) << pixshift);
                for (y = 0; y < rows; y++)
                    memcpy(d->pic.data[p] + (size_t)((sl->y >> sv) + y) * d->pic.linesize[p] + xo,
                           d->last.data[p] + (size_t)((sl->y >> sv) + y) * d->last.linesize[p] + xo, bw);
            }
        }
    }
    d->have_last = 1;
    for (i = 0; i < 4; i++) {
        planes[i] = d->pic.data[i];
        linesize[i] = d->pic.linesize[i];
    }
    if (pix_fmt)
        *pix_fmt = s->pf->name;
    if (key)
        *key = d->key_frame;
    r = pkt_size;
out:
    free(pkt);
    return r;
}
