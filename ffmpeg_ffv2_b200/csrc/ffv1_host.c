/*
 * ffv1_host.c -- host-side C of the B200 FFV1 codec: everything the reference
 * does once per stream or once per frame on the calling thread, i.e. option
 * resolution (encode_init), the range-coded global header (extradata), the
 * key-frame bit / in-band header / slice headers that prefix each slice's
 * arithmetic-coded stream, and packet framing for the decoder.
 *
 * No per-sample work is done here (that is ffv1_kernels.cu) and nothing in this
 * file can stand in for the GPU path.
 *
 * Reference citations are relative to the reference tree (libavcodec/...).
 */
#include "ffv1_host.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define FFMIN(a, b) ((a) < (b) ? (a) : (b))
#define FFMAX(a, b) ((a) > (b) ? (a) : (b))
#define CEIL_RSHIFT(a, b) (-((-(a)) >> (b)))

/* ------------------------------------------------------------------ */
/* pixel formats accepted by the encoder (ffv1enc.c:1333-1355) and     */
/* produced by the decoder (ffv1dec.c:597-739), little-endian names    */
/* ------------------------------------------------------------------ */
#define PLANAR(n, d, h, v, c, a) { n, FF_LAY_PLANAR, d, h, v, c, a, (c ? 3 : 1) + a }
static const FFPixFmt pixfmt_table[] = {
    PLANAR("gray", 8, 0, 0, 0, 0),       PLANAR("gray9le", 9, 0, 0, 0, 0),
    PLANAR("gray10le", 10, 0, 0, 0, 0),  PLANAR("gray12le", 12, 0, 0, 0, 0),
    PLANAR("gray16le", 16, 0, 0, 0, 0),
    { "ya8", FF_LAY_YA8, 8, 0, 0, 0, 1, 1 },
    PLANAR("yuv444p", 8, 0, 0, 1, 0),    PLANAR("yuv440p", 8, 0, 1, 1, 0),
    PLANAR("yuv422p", 8, 1, 0, 1, 0),    PLANAR("yuv420p", 8, 1, 1, 1, 0),
    PLANAR("yuv411p", 8, 2, 0, 1, 0),    PLANAR("yuv410p", 8, 2, 2, 1, 0),
    PLANAR("yuva444p", 8, 0, 0, 1, 1),   PLANAR("yuva422p", 8, 1, 0, 1, 1),
    PLANAR("yuva420p", 8, 1, 1, 1, 1),
    PLANAR("yuv444p9le", 9, 0, 0, 1, 0), PLANAR("yuv422p9le", 9, 1, 0, 1, 0),
    PLANAR("yuv420p9le", 9, 1, 1, 1, 0),
    PLANAR("yuv444p10le", 10, 0, 0, 1, 0), PLANAR("yuv440p10le", 10, 0, 1, 1, 0),
    PLANAR("yuv422p10le", 10, 1, 0, 1, 0), PLANAR("yuv420p10le", 10, 1, 1, 1, 0),
    PLANAR("yuv444p12le", 12, 0, 0, 1, 0), PLANAR("yuv440p12le", 12, 0, 1, 1, 0),
    PLANAR("yuv422p12le", 12, 1, 0, 1, 0), PLANAR("yuv420p12le", 12, 1, 1, 1, 0),
    PLANAR("yuv444p14le", 14, 0, 0, 1, 0), PLANAR("yuv422p14le", 14, 1, 0, 1, 0),
    PLANAR("yuv420p14le", 14, 1, 1, 1, 0),
    PLANAR("yuv444p16le", 16, 0, 0, 1, 0), PLANAR("yuv422p16le", 16, 1, 0, 1, 0),
    PLANAR("yuv420p16le", 16, 1, 1, 1, 0),
    PLANAR("yuva444p9le", 9, 0, 0, 1, 1),  PLANAR("yuva422p9le", 9, 1, 0, 1, 1),
    PLANAR("yuva420p9le", 9, 1, 1, 1, 1),
    PLANAR("yuva444p10le", 10, 0, 0, 1, 1), PLANAR("yuva422p10le", 10, 1, 0, 1, 1),
    PLANAR("yuva420p10le", 10, 1, 1, 1, 1),
    PLANAR("yuva444p16le", 16, 0, 0, 1, 1), PLANAR("yuva422p16le", 16, 1, 0, 1, 1),
    PLANAR("yuva420p16le", 16, 1, 1, 1, 1),
    { "bgr0", FF_LAY_BGR32, 8, 0, 0, 1, 0, 1 },   { "bgra", FF_LAY_BGR32, 8, 0, 0, 1, 1, 1 },
    { "gbrp9le", FF_LAY_GBRP, 9, 0, 0, 1, 0, 3 },  { "gbrp10le", FF_LAY_GBRP, 10, 0, 0, 1, 0, 3 },
    { "gbrp12le", FF_LAY_GBRP, 12, 0, 0, 1, 0, 3 }, { "gbrp14le", FF_LAY_GBRP, 14, 0, 0, 1, 0, 3 },
    { "gbrp16le", FF_LAY_GBRP, 16, 0, 0, 1, 0, 3 },
    { "gbrap10le", FF_LAY_GBRP, 10, 0, 0, 1, 1, 4 }, { "gbrap12le", FF_LAY_GBRP, 12, 0, 0, 1, 1, 4 },
    { "gbrap16le", FF_LAY_GBRP, 16, 0, 0, 1, 1, 4 },
    { "rgb48le", FF_LAY_RGB48, 16, 0, 0, 1, 0, 1 }, { "rgba64le", FF_LAY_RGB48, 16, 0, 0, 1, 1, 1 },
};

const FFPixFmt *ff_find_pixfmt(const char *name)
{
    size_t i;
    if (!name)
        return NULL;
    for (i = 0; i < sizeof(pixfmt_table) / sizeof(pixfmt_table[0]); i++)
        if (!strcmp(pixfmt_table[i].name, name))
            return &pixfmt_table[i];
    return NULL;
}

int ff_bytes_per_pixel(const FFPixFmt *pf)
{
    switch (pf->layout) {
    case FF_LAY_YA8:   return 2;
    case FF_LAY_BGR32: return 4;
    case FF_LAY_RGB48: return pf->alpha ? 8 : 6;
    default:           return pf->depth > 8 ? 2 : 1;
    }
}

int ff_plane_geometry(const FFPixFmt *pf, int w, int h, int plane, int *rowbytes, int *rows)
{
    int sw = 0, sh = 0;
    if (!pf || plane < 0 || plane >= pf->nplanes)
        return -1;
    if (pf->layout == FF_LAY_PLANAR && pf->chroma && (plane == 1 || plane == 2)) {
        sw = pf->hs;
        sh = pf->vs;
    }
    *rowbytes = CEIL_RSHIFT(w, sw) * ff_bytes_per_pixel(pf);
    *rows = CEIL_RSHIFT(h, sh);
    return pf->nplanes;
}

/* ------------------------------------------------------------------ */
/* tables                                                               */
/* ------------------------------------------------------------------ */

/* ff_build_rac_states(c, 0.05 * (1LL << 32), 256 - 8), rangecoder.c:68-106 */
void ff_default_tables(FFRacTables *t)
{
    const int64_t unit = (int64_t)1 << 32;
    const int factor = (int)(0.05 * (double)((int64_t)1 << 32));
    const int max_p = 248;
    int64_t p = unit / 2;
    int last = 0, i;
    memset(t, 0, sizeof(*t));
    for (i = 0; i < 128; i++) {
        int p8 = (int)((256 * p + unit / 2) >> 32);
        if (p8 <= last)
            p8 = last + 1;
        if (last && last < 256 && p8 <= max_p)
            t->one[last] = (uint8_t)p8;
        p += ((unit - p) * factor + unit / 2) >> 32;
        last = p8;
    }
    for (i = 256 - max_p; i <= max_p; i++) {
        int p8;
        if (t->one[i])
            continue;
        p = (i * unit + 128) >> 8;
        p += ((unit - p) * factor + unit / 2) >> 32;
        p8 = (int)((256 * p + unit / 2) >> 32);
        p8 = FFMAX(p8, i + 1);
        p8 = FFMIN(p8, max_p);
        t->one[i] = (uint8_t)p8;
    }
    for (i = 1; i < 255; i++)
        t->zero[i] = (uint8_t)(256 - t->one[256 - i]);
}

/* ffv1.c:95-101 / ffv1enc.c:1213-1219 */
void ff_install_custom(FFRacTables *t, const uint8_t trans[256])
{
    int j;
    for (j = 1; j < 256; j++) {
        t->one[j] = trans[j];
        t->zero[256 - j] = (uint8_t)(256 - trans[j]);
    }
}

/* "ver2_state": the custom transition table of -coder range_tab, ffv1enc.c:120-137.
 * Stored as deltas against the identity so that a transposition is visible in review:
 * value[i] = i + delta[i]. */
static const int8_t custom_delta[256] = {
      0,   9,   8,   7,   6,  11,  10,   9,  20,   7,   6,  18,  30,  36,   6,  34,
     43,   8,   8,   7,   7,  10,  11,  10,   9,   9,   8,  10,  39,   9,   9,   8,
      8,   7,   7,  44,   7,   7,   7,   6,   8,   7,  22,   7,   7,   7,  42,   5,
      5,  25,   5,   6,   6,   5,  20,   5,  45,   4,   4,  25,   6,   5,   6,   6,
     23,  17,   5,  30,   5,   4,  12,   4,  39,   4,  20,   3,  11,   4,   5,  18,
      5,   2,  12,   3,  15,   4,   4,  12,  23,   3,   3,  43,   3,   5,  11,   3,
      9,  13,   4,   9,   2,  17,   1,   3,   2,   8,   3,   5,   6,   3,   6,  14,
      3,   3,   3,   2,  10,   2,   7,   2,   1,   2,  23,   1,   2,   6,   1,   2,
     37,   1,   2,   7,   1,   2,  11,   1,   1,   2,   8,   2,   3,   1,   2,   5,
      3,  10,   5,   2,   3,   1,   2,   6,   1,   1,   2,  13,   2,   5,   3,   1,
     12,   2,   7,   1,   2,  19,   1,   3,   9,   5,   1,   2,  10,   3,   6,   3,
     -1,  12,   1,   2,   6,   2,  10,   2,  16,   2,   5,   1,   2,   8,   3,   5,
      5,   1,   1,   1,   2,   5,   1,   2,  10,   2,   5,   1,   1,   1,   2,   7,
      1,   2,  11,   1,   1,   2,  10,   1,   1,   1,   1,   1,   2,   7,   1,   2,
      2,  -1,   1,   2,  12,   1,   1,   1,   1,   1,   1,   1,   2,   2,  -1,   3,
      1,   2,   0,   1,   1,   1,   1,   1,   1,   1,   1,   1,   0,   0,   0,   0,
};

static void custom_transition(uint8_t t[256])
{
    int i;
    for (i = 0; i < 256; i++)
        t[i] = (uint8_t)(i + custom_delta[i]);
}

/* context quantisers (ffv1enc.c:44-118) are odd-symmetric step functions; they are
 * generated from their rising edges instead of being listed */
static void step_quantiser(int8_t q[256], const uint8_t *edge, int levels)
{
    int i, l = 0;
    for (i = 0; i < 128; i++) {
        while (l + 1 < levels && i >= edge[l + 1])
            l++;
        q[i] = (int8_t)l;
    }
    q[128] = (int8_t)(1 - levels);
    for (i = 129; i < 256; i++)
        q[i] = (int8_t)-q[256 - i];
}

static void build_quant_tables(FFStream *s)
{
    static const uint8_t e11[] = { 0, 1, 2, 5, 12, 35 };   /* quant11      */
    static const uint8_t e5[]  = { 0, 1, 4 };              /* quant5       */
    static const uint8_t e9h[] = { 0, 5, 13, 27, 56 };     /* quant9_10bit */
    static const uint8_t e5h[] = { 0, 11, 50 };            /* quant5_10bit */
    int8_t fine[256], coarse[256];
    int i;
    if (s->bits <= 8) {
        step_quantiser(fine, e11, 6);
        step_quantiser(coarse, e5, 3);
    } else {
        step_quantiser(fine, e9h, 5);
        step_quantiser(coarse, e5h, 3);
    }
    memset(s->qt, 0, sizeof(s->qt));
    for (i = 0; i < 256; i++) {            /* ffv1enc.c:730-752 */
        s->qt[0][0][i] = fine[i];
        s->qt[0][1][i] = 11 * fine[i];
        s->qt[0][2][i] = 121 * fine[i];
        s->qt[1][0][i] = fine[i];
        s->qt[1][1][i] = 11 * fine[i];
        s->qt[1][2][i] = 121 * coarse[i];
        s->qt[1][3][i] = 605 * coarse[i];
        s->qt[1][4][i] = 3025 * coarse[i];
    }
    s->qt_count = 2;
    s->ctx_count[0] = (11 * 11 * 11 + 1) / 2;
    s->ctx_count[1] = (11 * 11 * 5 * 5 * 5 + 1) / 2;
}

/* ------------------------------------------------------------------ */
/* CRC-32 IEEE, MSB first, init 0, no final xor (libavutil/crc.c:336),  */
/* slicing-by-4.  Register convention of av_crc() on little-endian      */
/* hosts: the value is the byte-swapped textbook register, and FFV1     */
/* stores it little-endian.                                             */
/* ------------------------------------------------------------------ */
static uint32_t crc_t[4][256];
static int crc_ready;

static void crc_init(void)
{
    int i, k;
    for (i = 0; i < 256; i++) {
        uint32_t c = (uint32_t)i << 24;
        for (k = 0; k < 8; k++)
            c = (c << 1) ^ ((c >> 31) ? 0x04C11DB7u : 0u);
        crc_t[0][i] = c;
    }
    for (i = 0; i < 256; i++)
        for (k = 1; k < 4; k++)
            crc_t[k][i] = (crc_t[k - 1][i] << 8) ^ crc_t[0][crc_t[k - 1][i] >> 24];
    __atomic_store_n(&crc_ready, 1, __ATOMIC_RELEASE);
}

static uint32_t swap32(uint32_t v)
{
    return (v >> 24) | ((v >> 8) & 0xFF00u) | ((v << 8) & 0xFF0000u) | (v << 24);
}

uint32_t ff_crc32(uint32_t crc, const uint8_t *buf, size_t len)
{
    uint32_t r = swap32(crc);
    if (!__atomic_load_n(&crc_ready, __ATOMIC_ACQUIRE))
        crc_init();
    while (len >= 4) {
        r ^= ((uint32_t)buf[0] << 24) | ((uint32_t)buf[1] << 16) | ((uint32_t)buf[2] << 8) | buf[3];
        r = crc_t[3][r >> 24] ^ crc_t[2][(r >> 16) & 0xFF] ^ crc_t[1][(r >> 8) & 0xFF] ^ crc_t[0][r & 0xFF];
        buf += 4;
        len -= 4;
    }
    while (len--)
        r = (r << 8) ^ crc_t[0][(r >> 24) ^ *buf++];
    return swap32(r);
}

static void wl32(uint8_t *p, uint32_t v)
{
    p[0] = (uint8_t)v;
    p[1] = (uint8_t)(v >> 8);
    p[2] = (uint8_t)(v >> 16);
    p[3] = (uint8_t)(v >> 24);
}

/* ------------------------------------------------------------------ */
/* encoder option resolution                                            */
/* ------------------------------------------------------------------ */
/* ------------------------------------------------------------------ */
/* second pass: tables from the first pass's statistics                  */
/* ------------------------------------------------------------------ */
/* Everything here is floating point in the reference and decides bytes of the extradata, so
 * the expressions are evaluated in the reference's order (ffv1enc.c:139-183, :469-515,
 * :793-873); the integer parts are restated freely. */

/* find_best_state, ffv1enc.c:139-183: for a true probability i/256 and k further symbols,
 * which starting state (within +-10 of i) costs the fewest expected bits */
static void best_start_states(uint8_t (*best)[256], const uint8_t one_state[256])
{
    double l2[256];
    int i, j, k, m;
    for (i = 1; i < 256; i++)
        l2[i] = log2(i / 256.0);
    for (i = 0; i < 256; i++) {
        double best_len[256];
        const double p = i / 256.0;
        for (j = 0; j < 256; j++)
            best_len[j] = 1 << 30;
        for (j = FFMAX(i - 10, 1); j < FFMIN(i + 11, 256); j++) {
            double occ[256] = { 0 };
            double len = 0;
            occ[j] = 1.0;
            if (!one_state[j])
                continue;
            for (k = 0; k < 256; k++) {
                double next[256] = { 0 };
                for (m = 1; m < 256; m++)
                    if (occ[m])
                        len -= occ[m] * (p * l2[m] + (1 - p) * l2[256 - m]);
                if (len < best_len[k]) {
                    best_len[k] = len;
                    best[i][k] = (uint8_t)j;
                }
                for (m = 1; m < 256; m++)
                    if (occ[m]) {
                        next[one_state[m]] += occ[m] * p;
                        next[256 - one_state[256 - m]] += occ[m] * (1 - p);
                    }
                memcpy(occ, next, sizeof(occ));
            }
        }
    }
}

/* the reference swaps its 64-bit counters through an int (FFSWAP(int, ...), ffv1enc.c:490) */
static void swap_through_int(uint64_t *a, uint64_t *b)
{
    const int t = (int)*b;
    *b = *a;
    *a = (uint64_t)(int64_t)t;
}

/* sort_stt, ffv1enc.c:469-515: exchange neighbouring states of the transition table while
 * that shortens the first pass's decisions */
static void sort_transitions(uint64_t rc[256][2], uint8_t stt[256])
{
    int changed, i, i2, j;
#define COST_(o, n) (rc[o][0] * -log2((256 - (n)) / 256.0) + rc[o][1] * -log2((n) / 256.0))
#define COST2_(o, n) (COST_(o, n) + COST_(256 - (o), 256 - (n)))
    do {
        changed = 0;
        for (i = 12; i < 244; i++)
            for (i2 = i + 1; i2 < 245 && i2 < i + 4; i2++) {
                const double size0 = COST2_(i, i) + COST2_(i2, i2);
                const double sizeX = COST2_(i, i2) + COST2_(i2, i);
                if (size0 - sizeX > size0 * (1e-14) && i != 128 && i2 != 128) {
                    uint8_t t = stt[i];
                    stt[i] = stt[i2];
                    stt[i2] = t;
                    swap_through_int(&rc[i][0], &rc[i2][0]);
                    swap_through_int(&rc[i][1], &rc[i2][1]);
                    if (i != 256 - i2) {
                        t = stt[256 - i];
                        stt[256 - i] = stt[256 - i2];
                        stt[256 - i2] = t;
                        swap_through_int(&rc[256 - i][0], &rc[256 - i2][0]);
                        swap_through_int(&rc[256 - i][1], &rc[256 - i2][1]);
                    }
                    for (j = 1; j < 256; j++) {
                        if (stt[j] == i)
                            stt[j] = (uint8_t)i2;
                        else if (stt[j] == i2)
                            stt[j] = (uint8_t)i;
                        if (i != 256 - i2) {
                            if (stt[256 - j] == 256 - i)
                                stt[256 - j] = (uint8_t)(256 - i2);
                            else if (stt[256 - j] == 256 - i2)
                                stt[256 - j] = (uint8_t)(256 - i);
                        }
                    }
                    changed = 1;
                }
            }
    } while (changed);
#undef COST_
#undef COST2_
}

static int clip_int(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }

/* AVCodecContext.stats_in -> s->trans (sorted) and s->initial[], ffv1enc.c:793-873 */
static int second_pass_tables(FFStream *s, const char *stats)
{
    static uint64_t rc[256][2];                 /* as in the reference the LAST block of the file counts */
    uint64_t (*rc2[FF_MAX_QUANT_TABLES])[32][2] = { 0 };
    uint8_t (*best)[256] = NULL;
    const char *p = stats;
    char *next;
    int gob_count = 0, i, j, k, m, ret = FFGPU_INVALIDDATA;

    for (i = 0; i < s->qt_count; i++) {
        rc2[i] = calloc((size_t)s->ctx_count[i], sizeof(*rc2[i]));
        if (!rc2[i]) {
            ret = FFGPU_ENOMEM;
            goto done;
        }
    }
    for (;;) {
        for (j = 0; j < 256; j++)
            for (i = 0; i < 2; i++) {
                rc[j][i] = (uint64_t)strtol(p, &next, 0);
                if (next == p)
                    goto done;                  /* "2Pass file invalid" */
                p = next;
            }
        for (i = 0; i < s->qt_count; i++)
            for (j = 0; j < s->ctx_count[i]; j++)
                for (k = 0; k < 32; k++)
                    for (m = 0; m < 2; m++) {
                        rc2[i][j][k][m] = (uint64_t)strtol(p, &next, 0);
                        if (next == p)
                            goto done;
                        p = next;
                    }
        gob_count = (int)strtol(p, &next, 0);
        if (next == p || gob_count <= 0)
            goto done;
        p = next;
        while (*p == '\n' || *p == ' ')
            p++;
        if (!p[0])
            break;
    }
    if (s->ac == FF_AC_CUSTOM)
        sort_transitions(rc, s->trans);
    best = malloc(256 * 256);
    if (!best) {
        ret = FFGPU_ENOMEM;
        goto done;
    }
    memset(best, 0, 256 * 256);
    best_start_states(best, s->trans);
    for (i = 0; i < s->qt_count; i++) {
        uint8_t *init = malloc((size_t)s->ctx_count[i] * FF_CONTEXT_SIZE);
        if (!init) {
            ret = FFGPU_ENOMEM;
            goto done;
        }
        memset(init, 128, (size_t)s->ctx_count[i] * FF_CONTEXT_SIZE);
        free(s->initial[i]);
        s->initial[i] = init;
        for (k = 0; k < 32; k++) {
            double a = 0, b = 0;
            int jp = 0;
            for (j = 0; j < s->ctx_count[i]; j++) {
                double pr = 128;
                if ((rc2[i][j][k][0] + rc2[i][j][k][1] > 200 && j) || a + b > 200) {
                    if (a + b)
                        pr = 256.0 * b / (a + b);
                    init[jp * FF_CONTEXT_SIZE + k] =
                        best[clip_int((int)round(pr), 1, 255)][clip_int((int)((a + b) / gob_count), 0, 255)];
                    for (jp++; jp < j; jp++)
                        init[jp * FF_CONTEXT_SIZE + k] = init[(jp - 1) * FF_CONTEXT_SIZE + k];
                    a = b = 0;
                }
                a += rc2[i][j][k][0];
                b += rc2[i][j][k][1];
                if (a + b)
                    pr = 256.0 * b / (a + b);
                init[j * FF_CONTEXT_SIZE + k] =
                    best[clip_int((int)round(pr), 1, 255)][clip_int((int)((a + b) / gob_count), 0, 255)];
            }
        }
    }
    ret = 0;
done:
    for (i = 0; i < FF_MAX_QUANT_TABLES; i++)
        free(rc2[i]);
    free(best);
    return ret;
}

int ff_stream_from_options(FFStream *s, const ffgpu_enc_options *o)
{
    const FFPixFmt *pf = ff_find_pixfmt(o->pix_fmt);
    int ac = o->coder, ec = o->slicecrc;

    memset(s, 0, sizeof(*s));
    if (!o->width || !o->height)           /* ff_ffv1_common_init, ffv1.c:46-47 */
        return FFGPU_INVALIDDATA;
    s->width = o->width;
    s->height = o->height;
    s->nh = s->nv = 1;
    ff_default_tables(&s->def_tab);

    /* version, ffv1enc.c:526-558 */
    if (o->slices > 1 || o->pass1 || o->pass2)
        s->version = FFMAX(s->version, 2);
    if (o->slices == 0 && o->level < 0 && o->width * o->height > 720 * 576)
        s->version = FFMAX(s->version, 2);
    if (o->level <= 0 && s->version == 2)
        s->version = 3;
    if (o->level >= 0 && o->level <= 4) {
        if (o->level < s->version)
            return FFGPU_EINVAL;           /* "Version %d needed for requested features" */
        s->version = o->level;
    }
    if (ec < 0)
        ec = s->version >= 3;
    if (ec)
        s->version = FFMAX(s->version, 3);
    s->ec = ec;
    if ((s->version == 2 || s->version > 3) && o->strict_std_compliance > -2)
        return FFGPU_INVALIDDATA;          /* experimental versions need -strict -2 */

    /* coder, ffv1enc.c:560-570 */
    if (ac == FFGPU_CODER_AC)
        ac = FF_AC_CUSTOM;
    else if (ac == FFGPU_CODER_RANGE_DEF)
        ac = FF_AC_DEFAULT;

    /* pixel format, ffv1enc.c:572-699 */
    if (!pf)
        return FFGPU_ENOSYS;               /* "format not supported" */
    s->pf = pf;
    s->transparency = pf->alpha;
    if (pf->layout == FF_LAY_PLANAR || pf->layout == FF_LAY_YA8) {
        s->colorspace = 0;
        s->chroma_planes = pf->chroma;
        if (pf->depth > 8) {
            s->bits = o->bits_per_raw_sample ? o->bits_per_raw_sample : pf->depth;
            s->packed_lsb = pf->depth < 16;
            if (s->bits <= 8)
                return FFGPU_INVALIDDATA;  /* "bits_per_raw_sample invalid" */
            s->version = FFMAX(s->version, 1);
        } else {
            s->bits = 8;
        }
    } else {
        s->colorspace = 1;
        s->chroma_planes = 1;
        if (pf->layout == FF_LAY_BGR32) {
            s->bits = 8;
        } else {
            if (pf->layout == FF_LAY_RGB48)
                s->bits = 16;
            else
                s->bits = o->bits_per_raw_sample ? o->bits_per_raw_sample : pf->depth;
            s->use32 = s->bits >= 16;
            s->version = FFMAX(s->version, 1);
        }
    }
    if (s->bits > 8 && ac == FF_AC_GOLOMB)
        ac = FF_AC_CUSTOM;                 /* "bits_per_raw_sample > 8, forcing range coder" */
    s->ac = ac;
    if ((unsigned)o->context > 1U)
        return FFGPU_EINVAL;
    s->context_model = o->context;
    if (s->version == 2)
        return FFGPU_ENOSYS;               /* the abandoned experimental version 2 bitstream */
    if (s->version > 3) {
        /* Version 4: every slice header carries the coefficients
         * choose_rct_params (ffv1enc.c:963-1043) picks -- also for YCbCr input, where that
         * function reads the planes as if they were RGB, outside the chroma planes and, for
         * 8-bit input, across row ends: its result then depends on memory the encoder does
         * not own, so no second implementation can reproduce the packets.  RGB layouts whose
         * planes it reads in bounds are supported; packed 16-bit RGB dereferences the NULL
         * plane pointers there. */
        if (s->colorspace != 1 || pf->layout == FF_LAY_RGB48)
            return FFGPU_ENOSYS;
    }

    /* transition table, ffv1enc.c:720-728 */
    if (ac == FF_AC_CUSTOM)
        custom_transition(s->trans);
    else
        memcpy(s->trans, s->def_tab.one, 256);
    s->cur_tab = s->def_tab;
    if (ac == FF_AC_CUSTOM)
        ff_install_custom(&s->cur_tab, s->trans);

    build_quant_tables(s);
    if (o->stats_in && o->stats_in[0]) {
        int r = second_pass_tables(s, o->stats_in);
        if (r < 0)
            return r;
        if (ac == FF_AC_CUSTOM) {              /* sort_stt may have reordered the table */
            s->cur_tab = s->def_tab;
            ff_install_custom(&s->cur_tab, s->trans);
        }
    }
    s->plane_sets = s->transparency ? 3 : 2;   /* v<=3: ffv1enc.c:769-772 */
    s->hs = pf->layout == FF_LAY_PLANAR && pf->chroma ? pf->hs : 0;
    s->vs = pf->layout == FF_LAY_PLANAR && pf->chroma ? pf->vs : 0;

    /* slice grid search, ffv1enc.c:875-903 */
    if (s->version > 1) {
        const int planes = 1 + 2 * s->chroma_planes + s->transparency;
        const int max_h = CEIL_RSHIFT(o->width, s->hs);
        const int max_v = CEIL_RSHIFT(o->height, s->vs);
        int nv, nh, found = 0;
        nv = (o->width > 352 || o->height > 288 || !o->slices) ? 2 : 1;
        nv = FFMIN(nv, max_v);
        for (; nv < 32 && !found; nv++)
            for (nh = nv; nh < 2 * nv; nh++) {
                const int maxw = (o->width + nh - 1) / nh;
                const int maxh = (o->height + nv - 1) / nv;
                if (nh > max_h || nv > max_v)
                    continue;
                if (maxw * maxh * (int64_t)(s->bits + 1) * planes > 8 << 24)
                    continue;
                if ((o->slices == nh * nv && o->slices <= FF_MAX_SLICES) || !o->slices) {
                    s->nh = nh;
                    s->nv = nv;
                    found = 1;
                    break;
                }
            }
        if (!found)
            return FFGPU_ENOSYS;           /* "Unsupported number %d of slices requested" */
    }
    return 0;
}

void ff_stream_free(FFStream *s)
{
    int i;
    for (i = 0; i < FF_MAX_QUANT_TABLES; i++) {
        free(s->initial[i]);
        s->initial[i] = NULL;
    }
}

void ff_slice_rect(const FFStream *s, int i, FFSliceRect *r)
{
    const int sx = i % s->nh, sy = i / s->nh;
    const int xs = s->width * sx / s->nh, xe = s->width * (sx + 1) / s->nh;
    const int ys = s->height * sy / s->nv, ye = s->height * (sy + 1) / s->nv;
    r->x = xs;
    r->w = xe - xs;
    r->y = ys;
    r->h = ye - ys;
}

/* ------------------------------------------------------------------ */
/* global header                                                        */
/* ------------------------------------------------------------------ */
static void put_quant_table(FFRacEnc *c, const FFRacTables *t, const int16_t *q)
{
    /* write_quant_table, ffv1enc.c:314-327: run lengths of the positive half */
    uint8_t st[FF_CONTEXT_SIZE];
    int i, run_start = 0;
    memset(st, 128, sizeof(st));
    for (i = 1; i < 128; i++)
        if (q[i] != q[i - 1]) {
            ffrac_put_symbol(c, t, st, i - run_start - 1, 0);
            run_start = i;
        }
    ffrac_put_symbol(c, t, st, 128 - run_start - 1, 0);
}

int ff_write_extradata(FFStream *s, int gop_size, uint8_t **data, int *size)
{
    const FFRacTables *t = &s->def_tab;
    FFRacEnc c;
    uint8_t st[FF_CONTEXT_SIZE];
    uint8_t *buf;
    uint32_t n;
    int i, j;

    *data = NULL;
    *size = 0;
    if (s->version < 2)
        return 0;
    buf = (uint8_t *)malloc(1 << 20);          /* a second pass adds up to 7563 x 32 initial states */
    if (!buf)
        return FFGPU_ENOMEM;
    memset(st, 128, sizeof(st));
    ffrac_enc_init(&c, buf, (1 << 20) - 8);

    ffrac_put_symbol(&c, t, st, s->version, 0);
    if (s->version > 2) {
        s->micro_version = s->version == 3 ? 4 : 2;
        ffrac_put_symbol(&c, t, st, s->micro_version, 0);
    }
    ffrac_put_symbol(&c, t, st, s->ac, 0);
    if (s->ac == FF_AC_CUSTOM)
        for (i = 1; i < 256; i++)
            ffrac_put_symbol(&c, t, st, s->trans[i] - t->one[i], 1);
    ffrac_put_symbol(&c, t, st, s->colorspace, 0);
    ffrac_put_symbol(&c, t, st, s->bits, 0);
    ffrac_put(&c, t, st, s->chroma_planes);
    ffrac_put_symbol(&c, t, st, s->hs, 0);
    ffrac_put_symbol(&c, t, st, s->vs, 0);
    ffrac_put(&c, t, st, s->transparency);
    ffrac_put_symbol(&c, t, st, s->nh - 1, 0);
    ffrac_put_symbol(&c, t, st, s->nv - 1, 0);
    ffrac_put_symbol(&c, t, st, s->qt_count, 0);
    for (i = 0; i < s->qt_count; i++)
        for (j = 0; j < FF_MAX_CTX_INPUTS; j++)
            put_quant_table(&c, t, s->qt[i][j]);
    for (i = 0; i < s->qt_count; i++) {
        /* contains_non_128 / the initial states of a second pass, ffv1enc.c:337-347, :442-455 */
        int any = 0;
        size_t q;
        for (q = 0; s->initial[i] && q < (size_t)s->ctx_count[i] * FF_CONTEXT_SIZE; q++)
            any |= s->initial[i][q] != 128;
        ffrac_put(&c, t, st, any);
        if (any) {
            uint8_t st2[FF_CONTEXT_SIZE][FF_CONTEXT_SIZE];
            memset(st2, 128, sizeof(st2));
            for (j = 0; j < s->ctx_count[i]; j++)
                for (int k = 0; k < FF_CONTEXT_SIZE; k++) {
                    const int pred = j ? s->initial[i][(j - 1) * FF_CONTEXT_SIZE + k] : 128;
                    ffrac_put_symbol(&c, t, st2[k], (int8_t)(s->initial[i][j * FF_CONTEXT_SIZE + k] - pred), 1);
                }
        }
    }
    if (s->version > 2) {
        s->intra = gop_size < 2;
        ffrac_put_symbol(&c, t, st, s->ec, 0);
        ffrac_put_symbol(&c, t, st, s->intra, 0);
    }
    n = ffrac_enc_finish(&c, t, 0);
    if (c.overflow) {
        free(buf);
        return FFGPU_ENOMEM;
    }
    wl32(buf + n, ff_crc32(0, buf, n));
    *data = buf;
    *size = (int)n + 4;
    return 0;
}

static int get_quant_table(FFRacDec *c, const FFRacTables *t, int16_t *q, int scale)
{
    /* read_quant_table, ffv1dec.c:368-393 */
    uint8_t st[FF_CONTEXT_SIZE];
    int level, i = 0;
    memset(st, 128, sizeof(st));
    for (level = 0; i < 128; level++) {
        int sym = ffrac_get_symbol(c, t, st, 0);
        unsigned len = (unsigned)sym + 1U;
        if (len > (unsigned)(128 - i) || !len)
            return FFGPU_INVALIDDATA;
        for (; len; len--)
            q[i++] = (int16_t)(scale * level);
    }
    for (i = 1; i < 128; i++)
        q[256 - i] = (int16_t)-q[i];
    q[128] = (int16_t)-q[127];
    return 2 * level - 1;
}

static int get_quant_tables(FFRacDec *c, const FFRacTables *t, int16_t q[FF_MAX_CTX_INPUTS][256])
{
    /* read_quant_tables, ffv1dec.c:395-411 */
    int i, n = 1;
    for (i = 0; i < FF_MAX_CTX_INPUTS; i++) {
        int r = get_quant_table(c, t, q[i], n);
        if (r < 0)
            return r;
        n *= r;
        if ((unsigned)n > 32768U)
            return FFGPU_INVALIDDATA;
    }
    return (n + 1) / 2;
}

int ff_parse_extradata(FFStream *s, const uint8_t *data, int size)
{
    const FFRacTables *t = &s->def_tab;
    FFRacDec c;
    uint8_t st[FF_CONTEXT_SIZE];
    uint8_t (*st2)[FF_CONTEXT_SIZE];
    int i, j, k;

    if (size < 2)
        return FFGPU_INVALIDDATA;
    memset(st, 128, sizeof(st));
    ffrac_dec_init(&c, data, (uint32_t)size);
    s->version = ffrac_get_symbol(&c, t, st, 0);
    if (s->version < 2)
        return FFGPU_INVALIDDATA;          /* "Invalid version in global header" */
    if (s->version > 2) {
        c.end -= 4;
        s->micro_version = ffrac_get_symbol(&c, t, st, 0);
        if (s->micro_version < 0)
            return FFGPU_INVALIDDATA;
    }
    s->ac = ffrac_get_symbol(&c, t, st, 0);
    if (s->ac == FF_AC_CUSTOM)
        for (i = 1; i < 256; i++)
            s->trans[i] = (uint8_t)(ffrac_get_symbol(&c, t, st, 1) + t->one[i]);
    s->colorspace = ffrac_get_symbol(&c, t, st, 0);
    s->bits = ffrac_get_symbol(&c, t, st, 0);
    s->chroma_planes = ffrac_get(&c, t, st);
    s->hs = ffrac_get_symbol(&c, t, st, 0);
    s->vs = ffrac_get_symbol(&c, t, st, 0);
    s->transparency = ffrac_get(&c, t, st);
    s->plane_sets = 1 + (s->chroma_planes || s->version < 4) + s->transparency;
    s->nh = 1 + ffrac_get_symbol(&c, t, st, 0);
    s->nv = 1 + ffrac_get_symbol(&c, t, st, 0);
    if ((unsigned)s->hs > 4U || (unsigned)s->vs > 4U)
        return FFGPU_INVALIDDATA;
    if ((unsigned)s->nh > (unsigned)s->width || !s->nh ||
        (unsigned)s->nv > (unsigned)s->height || !s->nv)
        return FFGPU_INVALIDDATA;          /* "slice count invalid" */
    if ((int64_t)s->nh * s->nv > FF_MAX_SLICES)
        return FFGPU_INVALIDDATA;
    s->qt_count = ffrac_get_symbol(&c, t, st, 0);
    if ((unsigned)s->qt_count > FF_MAX_QUANT_TABLES || !s->qt_count) {
        s->qt_count = 0;
        return FFGPU_INVALIDDATA;
    }
    for (i = 0; i < s->qt_count; i++) {
        s->ctx_count[i] = get_quant_tables(&c, t, s->qt[i]);
        if (s->ctx_count[i] < 0)
            return FFGPU_INVALIDDATA;
    }
    st2 = (uint8_t (*)[FF_CONTEXT_SIZE])malloc(32 * FF_CONTEXT_SIZE);
    if (!st2)
        return FFGPU_ENOMEM;
    memset(st2, 128, 32 * FF_CONTEXT_SIZE);
    for (i = 0; i < s->qt_count; i++) {
        if (!ffrac_get(&c, t, st))
            continue;
        s->initial[i] = (uint8_t *)malloc((size_t)s->ctx_count[i] * FF_CONTEXT_SIZE);
        if (!s->initial[i]) {
            free(st2);
            return FFGPU_ENOMEM;
        }
        for (j = 0; j < s->ctx_count[i]; j++)
            for (k = 0; k < FF_CONTEXT_SIZE; k++) {
                int pred = j ? s->initial[i][(j - 1) * FF_CONTEXT_SIZE + k] : 128;
                s->initial[i][j * FF_CONTEXT_SIZE + k] =
                    (uint8_t)((pred + ffrac_get_symbol(&c, t, st2[k], 1)) & 0xFF);
            }
    }
    free(st2);
    if (s->version > 2) {
        s->ec = ffrac_get_symbol(&c, t, st, 0);
        if (s->micro_version > 2)
            s->intra = ffrac_get_symbol(&c, t, st, 0);
        if (ff_crc32(0, data, (size_t)size) || size < 4)
            return FFGPU_INVALIDDATA;      /* "CRC mismatch" */
    }
    if (s->version > 4)
        return FFGPU_ENOSYS;
    s->cur_tab = s->def_tab;
    if (s->ac == FF_AC_CUSTOM)
        ff_install_custom(&s->cur_tab, s->trans);
    return 0;
}

int ff_pick_decoder_format(FFStream *s)
{
    char name[32] = "";
    const int b = s->bits;
    const int sub = 16 * s->hs + s->vs;
    const char *ss = NULL;

    s->packed_lsb = 0;
    s->use32 = 0;
    s->pf = NULL;
    if (s->colorspace == 0) {
        switch (sub) {
        case 0x00: ss = "444"; break;
        case 0x01: ss = "440"; break;
        case 0x10: ss = "422"; break;
        case 0x11: ss = "420"; break;
        case 0x20: ss = "411"; break;
        case 0x22: ss = "410"; break;
        }
        if (!s->transparency && !s->chroma_planes) {
            if (b <= 8)
                strcpy(name, "gray");
            else if (b == 9 || b == 10 || b == 12 || b == 16) {
                s->packed_lsb = 1;
                strcpy(name, b == 9 ? "gray9le" : b == 10 ? "gray10le" : b == 12 ? "gray12le" : "gray16le");
            } else if (b < 16)
                strcpy(name, "gray16le");
            else
                return FFGPU_ENOSYS;
        } else if (s->transparency && !s->chroma_planes) {
            if (b > 8)
                return FFGPU_ENOSYS;
            strcpy(name, "ya8");
        } else if (b <= 8) {
            if (ss && (!s->transparency || sub == 0x00 || sub == 0x10 || sub == 0x11)) {
                strcpy(name, s->transparency ? "yuva" : "yuv");
                strcat(name, ss);
                strcat(name, "p");
            }
        } else if (b == 9 || b == 10 || b == 12 || b == 14 || b == 16) {
            int ok = ss && (sub == 0x00 || sub == 0x10 || sub == 0x11 ||
                            (sub == 0x01 && !s->transparency && (b == 10 || b == 12)));
            if (s->transparency && (b == 12 || b == 14))
                ok = 0;
            s->packed_lsb = 1;
            if (ok) {
                strcpy(name, s->transparency ? "yuva" : "yuv");
                strcat(name, ss);
                strcat(name, b == 9 ? "p9le" : b == 10 ? "p10le" : b == 12 ? "p12le" :
                             b == 14 ? "p14le" : "p16le");
            }
        }
    } else if (s->colorspace == 1) {
        if (s->hs || s->vs)
            return FFGPU_ENOSYS;           /* "chroma subsampling not supported in this colorspace" */
        if (b <= 8)
            strcpy(name, s->transparency ? "bgra" : "bgr0");
        else if (b == 9 && !s->transparency)
            strcpy(name, "gbrp9le");
        else if (b == 10)
            strcpy(name, s->transparency ? "gbrap10le" : "gbrp10le");
        else if (b == 12)
            strcpy(name, s->transparency ? "gbrap12le" : "gbrp12le");
        else if (b == 14 && !s->transparency)
            strcpy(name, "gbrp14le");
        else if (b == 16) {
            strcpy(name, s->transparency ? "gbrap16le" : "gbrp16le");
            s->use32 = 1;
        }
    } else {
        return FFGPU_ENOSYS;               /* "colorspace not supported" */
    }
    s->pf = ff_find_pixfmt(name);
    return s->pf ? 0 : FFGPU_ENOSYS;       /* "format not supported" */
}

/* ------------------------------------------------------------------ */
/* kernel-facing description                                            */
/* ------------------------------------------------------------------ */
int ff_fill_dev_params(const FFStream *s, int encoder, FFDevParams *P, FFDevSlice *slices)
{
    const FFPixFmt *pf = s->pf;
    int i, k, n = 0, max_ctx = 0;
    size_t off = 0;
    uint32_t tok = 0, bs = 0;

    memset(P, 0, sizeof(*P));
    P->width = s->width;
    P->height = s->height;
    P->nh = s->nh;
    P->nv = s->nv;
    P->nslices = s->nh * s->nv;
    P->ac = s->ac;
    P->colorspace = s->colorspace;
    P->layout = pf->layout;
    P->sbits = s->bits <= 8 ? 8 : s->bits;
    P->packed_lsb = s->packed_lsb;
    P->use32 = s->use32;
    P->transparency = s->transparency;
    P->chroma_planes = s->chroma_planes;
    P->hs = s->hs;
    P->vs = s->vs;
    P->version = s->version;
    P->ec = s->ec;
    P->rgb_pixbytes = ff_bytes_per_pixel(pf);

    /* coded planes in coding order: encode_slice ffv1enc.c:1083-1104 */
    if (s->colorspace == 0) {
        const int step = s->bits > 8 ? 2 : 1;
        P->cbits = P->sbits;
        if (pf->layout == FF_LAY_YA8) {
            P->cp[n++] = (FFDevPlane){ 0, 0, 0, 0, 2, 0 };
            P->cp[n++] = (FFDevPlane){ 0, 0, 0, 1, 2, 1 };
        } else {
            P->cp[n++] = (FFDevPlane){ 0, 0, 0, 0, step, 0 };
            if (s->chroma_planes) {
                P->cp[n++] = (FFDevPlane){ 1, s->hs, s->vs, 1, step, 0 };
                P->cp[n++] = (FFDevPlane){ 2, s->hs, s->vs, 1, step, 0 };
            }
            if (s->transparency)
                P->cp[n++] = (FFDevPlane){ s->chroma_planes ? 3 : 1, 0, 0, 2, step, 0 };
        }
    } else {
        P->cbits = s->bits <= 8 ? 9 : s->bits + 1;
        for (k = 0; k < 3 + s->transparency; k++)
            P->cp[n++] = (FFDevPlane){ k, 0, 0, (k + 1) / 2, 2, 0 };
    }
    P->ncoded = n;
    P->nsets = s->plane_sets;
    for (k = 0; k < s->qt_count; k++)
        max_ctx = FFMAX(max_ctx, s->ctx_count[k]);
    for (k = 0; k < P->nsets; k++) {
        P->set_qidx[k] = encoder ? s->context_model : 0;
        /* encoder: every set uses the same table; decoder: sized for the largest table,
         * the table itself is chosen per slice */
        P->set_base[k] = k * (encoder ? s->ctx_count[s->context_model] : max_ctx);
    }
    P->total_ctx = P->nsets * (encoder ? s->ctx_count[s->context_model] : max_ctx);

    /* device picture layout: planes back to back, 256-byte pitch alignment */
    for (k = 0; k < pf->nplanes; k++) {
        int rb, rows;
        ff_plane_geometry(pf, s->width, s->height, k, &rb, &rows);
        P->plane_off[k] = off;
        P->pitch[k] = (rb + 255) & ~255;
        P->rows[k] = rows;
        off += (size_t)P->pitch[k] * rows;
    }
    P->frame_bytes = off;

    for (i = 0; i < P->nslices; i++) {
        FFSliceRect r;
        FFDevSlice *d = &slices[i];
        uint32_t samples = 0;
        memset(d, 0, sizeof(*d));
        ff_slice_rect(s, i, &r);
        d->x = r.x; d->y = r.y; d->w = r.w; d->h = r.h;
        if (s->colorspace == 0) {
            for (k = 0; k < n; k++) {
                int w = CEIL_RSHIFT(r.w, P->cp[k].hs), h = CEIL_RSHIFT(r.h, P->cp[k].vs);
                d->seg_lines[k] = h;
                d->seg_w[k] = w;
                samples += (uint32_t)w * h;
            }
            d->nseg = n;
        } else {
            d->seg_lines[0] = r.h * n;
            d->seg_w[0] = r.w;
            d->nseg = 1;
            samples = (uint32_t)r.w * r.h * n;
        }
        d->tok_off = tok;                     /* 16-byte aligned: stage B streams 4 tokens per cp.async */
        d->ntok = samples;
        tok += (samples + 3) & ~3u;
        /* slice bitstream arena: twice the information content of the coded samples
         * (cbits per sample) plus room for the header prefix; an overflow is reported as
         * "encoded frame too large" like the reference's own guard */
        d->bs_off = bs;
        d->bs_cap = (uint32_t)(((((uint64_t)samples * P->cbits + 7) / 8) * 2 + 2048 + 15) & ~15ULL);
        bs += d->bs_cap;
    }
    P->frame_tokens = tok;
    P->frame_bs = bs;
    P->trailer = 3 + 5 * !!s->ec;
    P->pkt_stride = ((size_t)bs + (size_t)P->nslices * 8 + 255) & ~(size_t)255;
    return 0;
}

/* ------------------------------------------------------------------ */
/* encoder: what precedes the pixel data in each slice's coder          */
/* ------------------------------------------------------------------ */
int ff_enc_slice_prefix(const FFStream *s, int i, const FFSliceRect *r, int key_frame,
                        int picture_structure, int sar_num, int sar_den,
                        FFRacPrefix *pre, uint8_t *bytes, int cap)
{
    const FFRacTables *t = &s->cur_tab;
    FFRacEnc c;
    uint8_t st[FF_CONTEXT_SIZE];
    int j;

    memset(pre, 0, sizeof(*pre));
    ffrac_enc_init(&c, bytes, (uint32_t)cap);
    if (i == 0) {
        /* the key-frame bit and the v0/v1 header are coded with the DEFAULT table; the
         * custom table is only installed afterwards (ffv1enc.c:1203-1219) */
        uint8_t keystate = 128;
        ffrac_put(&c, &s->def_tab, &keystate, key_frame);
        if (key_frame && s->version < 2) {
            /* write_header, ffv1enc.c:348-376 */
            memset(st, 128, sizeof(st));
            ffrac_put_symbol(&c, &s->def_tab, st, s->version, 0);
            ffrac_put_symbol(&c, &s->def_tab, st, s->ac, 0);
            if (s->ac == FF_AC_CUSTOM)
                for (j = 1; j < 256; j++)
                    ffrac_put_symbol(&c, &s->def_tab, st, s->trans[j] - s->def_tab.one[j], 1);
            ffrac_put_symbol(&c, &s->def_tab, st, s->colorspace, 0);
            if (s->version > 0)
                ffrac_put_symbol(&c, &s->def_tab, st, s->bits, 0);
            ffrac_put(&c, &s->def_tab, st, s->chroma_planes);
            ffrac_put_symbol(&c, &s->def_tab, st, s->hs, 0);
            ffrac_put_symbol(&c, &s->def_tab, st, s->vs, 0);
            ffrac_put(&c, &s->def_tab, st, s->transparency);
            for (j = 0; j < FF_MAX_CTX_INPUTS; j++)
                put_quant_table(&c, &s->def_tab, s->qt[s->context_model][j]);
        }
    }
    if (s->version > 2) {
        /* encode_slice_header, ffv1enc.c:930-961 */
        memset(st, 128, sizeof(st));
        ffrac_put_symbol(&c, t, st, (r->x + 1) * s->nh / s->width, 0);
        ffrac_put_symbol(&c, t, st, (r->y + 1) * s->nv / s->height, 0);
        ffrac_put_symbol(&c, t, st, (r->w + 1) * s->nh / s->width - 1, 0);
        ffrac_put_symbol(&c, t, st, (r->h + 1) * s->nv / s->height - 1, 0);
        for (j = 0; j < s->plane_sets; j++)
            ffrac_put_symbol(&c, t, st, s->context_model, 0);
        ffrac_put_symbol(&c, t, st, picture_structure, 0);
        ffrac_put_symbol(&c, t, st, sar_num, 0);
        ffrac_put_symbol(&c, t, st, sar_den, 0);
        memcpy(pre->hdr_state, st, sizeof(st));
    }
    if (s->version > 3) {
        /* the device continues the header (and closes the coder of Golomb-Rice slices) */
        pre->nbytes = c.pos;
    } else if (s->ac == FF_AC_GOLOMB) {
        /* ffv1enc.c:1076-1081: the coder is closed, Rice bits start on the next byte */
        uint32_t n = 0;
        if (s->version > 2 || (!r->x && !r->y))
            n = ffrac_enc_finish(&c, t, s->version > 2);
        pre->golomb_start = n;
        pre->nbytes = n;
    } else {
        pre->nbytes = c.pos;
    }
    pre->low = c.low;
    pre->range = c.range;
    pre->pending = c.pending;
    pre->run = c.run;
    return c.overflow ? FFGPU_ENOSPC : 0;
}

/* ------------------------------------------------------------------ */
/* decoder: packet framing                                              */
/* ------------------------------------------------------------------ */
static int rb24(const uint8_t *p)
{
    return (p[0] << 16) | (p[1] << 8) | p[2];
}

/* v0/v1 in-band header, read_header ffv1dec.c:538-590 */
static int parse_inband_header(FFStream *s, FFRacDec *c)
{
    const FFRacTables *t = &s->def_tab;
    uint8_t st[FF_CONTEXT_SIZE];
    int i, cs, bits, cp, hs, vs, tr, n;
    unsigned v;

    memset(st, 128, sizeof(st));
    v = (unsigned)ffrac_get_symbol(c, t, st, 0);
    if (v >= 2)
        return FFGPU_INVALIDDATA;          /* "invalid version %d in ver01 header" */
    s->version = (int)v;
    s->ac = ffrac_get_symbol(c, t, st, 0);
    if (s->ac == FF_AC_CUSTOM)
        for (i = 1; i < 256; i++) {
            int x = ffrac_get_symbol(c, t, st, 1) + t->one[i];
            if (x < 1 || x > 255)
                return FFGPU_INVALIDDATA;  /* "invalid state transition %d" */
            s->trans[i] = (uint8_t)x;
        }
    cs = ffrac_get_symbol(c, t, st, 0);
    bits = s->version > 0 ? ffrac_get_symbol(c, t, st, 0) : 0;
    cp = ffrac_get(c, t, st);
    hs = ffrac_get_symbol(c, t, st, 0);
    vs = ffrac_get_symbol(c, t, st, 0);
    tr = ffrac_get(c, t, st);
    if (s->plane_sets &&
        (cs != s->colorspace || bits != s->bits || cp != s->chroma_planes || hs != s->hs ||
         vs != s->vs || tr != s->transparency))
        return FFGPU_INVALIDDATA;          /* "Invalid change of global parameters" */
    if ((unsigned)hs > 4U || (unsigned)vs > 4U)
        return FFGPU_INVALIDDATA;
    s->colorspace = cs;
    s->bits = bits;
    s->chroma_planes = cp;
    s->hs = hs;
    s->vs = vs;
    s->transparency = tr;
    s->plane_sets = 2 + tr;
    if ((i = ff_pick_decoder_format(s)) < 0)
        return i;
    n = get_quant_tables(c, t, s->qt[0]);
    if (n < 0)
        return FFGPU_INVALIDDATA;
    s->qt_count = 1;
    s->ctx_count[0] = n;
    s->cur_tab = s->def_tab;
    if (s->ac == FF_AC_CUSTOM)
        ff_install_custom(&s->cur_tab, s->trans);
    return 0;
}

int ff_dec_parse_packet(FFStream *s, FFDecHostState *hs, const uint8_t *pkt, size_t size,
                        uint32_t pkt_off, FFDecSlice *out, FFDecFrameInfo *info)
{
    FFRacDec c0;
    uint8_t keystate = 128;
    const uint8_t *end;
    int i, n, r;

    memset(info, 0, sizeof(*info));
    info->sar_den = 1;
    if (size < 2 || size > (size_t)INT_MAX / 2)
        return FFGPU_INVALIDDATA;
    ffrac_dec_init(&c0, pkt, (uint32_t)size);
    if (ffrac_get(&c0, &s->def_tab, &keystate)) {
        info->key_frame = 1;
        hs->key_frame_ok = 0;
        if (s->version < 2) {
            if ((r = parse_inband_header(s, &c0)) < 0)
                return r;
            n = hs->max_slices;
        } else {
            /* v3: count slices by walking the size trailers from the end, ffv1dec.c:746-756 */
            const int trailer = 3 + 5 * !!s->ec;
            const uint8_t *p = pkt + c0.end;
            if (!s->pf && (r = ff_pick_decoder_format(s)) < 0)
                return r;
            for (n = 0; n < FF_MAX_SLICES && trailer < p - pkt; n++) {
                int sz = rb24(p - trailer);
                if (sz + trailer > p - pkt)
                    break;
                p -= sz + trailer;
            }
        }
        if (n <= 0 || n > hs->max_slices)
            return FFGPU_INVALIDDATA;      /* "slice count %d is invalid" */
        info->nslices = n;
        hs->slice_count = n;
        hs->key_frame_ok = 1;
        memset(hs->damaged, 0, sizeof(hs->damaged));
    } else {
        if (!hs->key_frame_ok)
            return FFGPU_INVALIDDATA;      /* "Cannot decode non-keyframe without valid keyframe" */
        info->key_frame = 0;
        info->nslices = n = hs->slice_count;   /* f->slice_count persists between key frames */
    }

    /* slice table from the tail, ffv1dec.c:890-931 */
    end = pkt + size;
    for (i = n - 1; i >= 0; i--) {
        FFDecSlice *d = &out[i];
        const int trailer = 3 + 5 * !!s->ec;
        int v;
        memset(d, 0, sizeof(*d));
        if (i || s->version > 2) {
            if (end - pkt < trailer)
                return FFGPU_INVALIDDATA;
            v = rb24(end - trailer) + trailer;
        } else {
            v = (int)(end - pkt);
        }
        if (end - pkt < v)
            return FFGPU_INVALIDDATA;      /* "Slice pointer chain broken" */
        end -= v;
        if (info->key_frame)
            hs->damaged[i] = 0;
        if (s->ec && !(hs->device_parse && i && s->version > 2) && ff_crc32(0, end, (size_t)v)) {
            hs->damaged[i] = 1;            /* "slice CRC mismatch" */
            info->crc_damaged++;
        }
        if (i) {
            d->pkt_off = pkt_off + (uint32_t)(end - pkt);
            d->size = (uint32_t)v;
        } else {
            /* slice 0 keeps decoding with the coder that read the key-frame bit: its stream
             * starts at the packet start and only bytestream_end moves (ffv1dec.c:927-928).
             * In an intact packet end == pkt here. */
            d->pkt_off = pkt_off;
            d->size = (uint32_t)((end - pkt) + v);
        }
        /* the slice's states are reset: key frames, and every picture of an intra-only stream,
         * whose decoder keeps no states between pictures -- a packet of such a stream with a
         * cleared key-frame bit (damage; the encoder never writes one) must not decode with
         * states nothing has initialised */
        d->key_frame = info->key_frame || (s->version > 2 && s->intra);
    }

    /* per slice: continue (slice 0) or start (others) the range decoder, parse the slice
     * header, hand the coder state to the device */
    info->interlaced_frame = 0;
    info->top_field_first = 0;
    info->sar_num = 0;
    for (i = 0; i < n; i++) {
        FFDecSlice *d = &out[i];
        const FFRacTables *t = &s->cur_tab;
        const uint8_t *base = pkt + (d->pkt_off - pkt_off);
        FFRacDec c;
        FFSliceRect rc;
        int j;

        d->rct_by = d->rct_ry = 1;             /* ffv1dec.c:290-291 */
        if (i && hs->device_parse && s->version == 3) {
            /* the decode kernel checks the CRC, parses the slice header and positions the
             * coder itself (ff_dec_slice_header); until it reports back the rectangle of
             * the regular grid stands in */
            d->parse = 1;
            ff_slice_rect(s, i, &rc);
            hs->rect[i] = rc;
            continue;
        }
        if (i == 0) {
            c = c0;
            c.end = d->size;               /* fs->c.bytestream_end = buf_p + v */
        } else {
            if (d->size < 2) {
                d->skip = 1;
                hs->damaged[i] = 1;
                continue;
            }
            ffrac_dec_init(&c, base, d->size);
        }
        if (s->version > 2) {
            /* decode_slice_header, ffv1dec.c:167-244 */
            uint8_t st[FF_CONTEXT_SIZE];
            unsigned sx, sy, sw, sh;
            int ps, bad = 0;
            memset(st, 128, sizeof(st));
            sx = (unsigned)ffrac_get_symbol(&c, t, st, 0) * (unsigned)s->width;
            sy = (unsigned)ffrac_get_symbol(&c, t, st, 0) * (unsigned)s->height;
            sw = ((unsigned)ffrac_get_symbol(&c, t, st, 0) + 1U) * (unsigned)s->width + sx;
            sh = ((unsigned)ffrac_get_symbol(&c, t, st, 0) + 1U) * (unsigned)s->height + sy;
            rc.x = (int)sx / s->nh;
            rc.y = (int)sy / s->nv;
            rc.w = (int)sw / s->nh - rc.x;
            rc.h = (int)sh / s->nv - rc.y;
            if ((unsigned)rc.w > (unsigned)s->width || (unsigned)rc.h > (unsigned)s->height)
                bad = 1;
            else if ((unsigned)rc.x + (uint64_t)rc.w > (unsigned)s->width ||
                     (unsigned)rc.y + (uint64_t)rc.h > (unsigned)s->height)
                bad = 1;
            for (j = 0; j < s->plane_sets && !bad; j++) {
                int idx = ffrac_get_symbol(&c, t, st, 0);
                if ((unsigned)idx >= (unsigned)s->qt_count)
                    bad = 1;               /* "quant_table_index out of range" */
                else
                    d->qidx[j] = idx;
            }
            if (bad) {
                rc.x = rc.y = rc.w = rc.h = 0;
                d->skip = 1;
                hs->damaged[i] = 1;
                hs->rect[i] = rc;
                continue;
            }
            ps = ffrac_get_symbol(&c, t, st, 0);
            if (ps == 1) {
                info->interlaced_frame = 1;
                info->top_field_first = 1;
            } else if (ps == 2) {
                info->interlaced_frame = 1;
                info->top_field_first = 0;
            } else if (ps == 3) {
                info->interlaced_frame = 0;
            }
            info->sar_num = ffrac_get_symbol(&c, t, st, 0);
            info->sar_den = ffrac_get_symbol(&c, t, st, 0);
            if (s->version > 3) {
                /* ffv1dec.c:230-241 */
                if (ffrac_get(&c, t, st))      /* slice_reset_contexts */
                    d->key_frame = 1;
                d->pcm = ffrac_get_symbol(&c, t, st, 0) == 1;
                if (!d->pcm) {
                    d->rct_by = ffrac_get_symbol(&c, t, st, 0);
                    d->rct_ry = ffrac_get_symbol(&c, t, st, 0);
                    if ((uint64_t)(unsigned)d->rct_by + (uint64_t)(unsigned)d->rct_ry > 4) {
                        /* "slice_rct_y_coef out of range": decode_slice_header fails */
                        rc.x = rc.y = rc.w = rc.h = 0;
                        d->skip = 1;
                        hs->damaged[i] = 1;
                        hs->rect[i] = rc;
                        continue;
                    }
                }
            }
        } else {
            ff_slice_rect(s, i, &rc);
        }
        hs->rect[i] = rc;
        d->x = rc.x; d->y = rc.y; d->w = rc.w; d->h = rc.h;
        if (s->ac == FF_AC_GOLOMB) {
            /* ffv1dec.c:312-319 */
            if ((s->version == 3 && s->micro_version > 1) || s->version > 3) {
                uint8_t term = 129;
                ffrac_get(&c, t, &term);
            }
            d->golomb_start = (s->version > 2 || (!rc.x && !rc.y)) ? c.pos - 1 : 0;
        }
        d->low = c.low;
        d->range = c.range;
        d->pos = c.pos;
        d->overread = c.overread;
        if (c.end < d->size)
            d->size = c.end;               /* low >= 0xFF00 special case of ff_init_range_decoder */
        if (!rc.w || !rc.h)
            d->skip = 1;
    }
    /* av_image_check_sar, ffv1dec.c:221-227 */
    if (info->sar_den <= 0 || info->sar_num < 0) {
        info->sar_num = 0;
        info->sar_den = 1;
    }
    return n;
}
