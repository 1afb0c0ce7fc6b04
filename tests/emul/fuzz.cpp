/*
 * tests/emul/fuzz.cpp -- TEST INFRASTRUCTURE: damaged-input robustness of the product's device
 * functions.  Links the CPU emulation (emul.cpp + ffv1_host.c) into one executable built with
 * -fsanitize=address,undefined and feeds the decoder mutated packets and mutated extradata:
 * every out-of-bounds access the parsers or the slice decoders could make on hostile input
 * -- on the GPU that is a fault that takes the context down -- stops the run here.
 * Both forms of the slice decoders are driven (FFV1_EMUL_LONE selects the straight-line one).
 *
 *   make -C tests/emul fuzz && tests/emul/fuzz [iterations] [seed]
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

struct Params {
    int width, height; const char *pix_fmt;
    int slices, level, gop_size, coder, context, slicecrc, strict, threads, bits_per_raw_sample;
};
extern "C" {
void *ffv1emul_encoder_open(const Params *p, int *err);
int ffv1emul_encoder_extradata(void *h, const uint8_t **d);
int ffv1emul_encode(void *h, const uint8_t *const planes[4], const int ls[4], uint8_t *out, int cap, int *key);
void ffv1emul_encoder_close(void *h);
void *ffv1emul_decoder_open(int w, int h, const uint8_t *ex, int exsize, int threads, int *err);
int ffv1emul_decode(void *h, const uint8_t *pkt, int size, uint8_t *planes[4], int ls[4], const char **fmt, int *key);
void ffv1emul_decoder_close(void *h);
int ffv1emul_plane_geometry(const char *fmt, int w, int h, int plane, int *bw, int *rows);
}

static uint64_t g_rng = 0x9E3779B97F4A7C15ull;
static uint32_t rnd(void)
{
    g_rng ^= g_rng << 13; g_rng ^= g_rng >> 7; g_rng ^= g_rng << 17;
    return (uint32_t)(g_rng >> 16);
}

struct Case { const char *fmt; int w, h, slices, level, gop, coder, context, crc, strict; };
static const Case CASES[] = {
    { "yuv420p",      64, 48,  4, -99, 12,  0, 0, -1, 0 },   /* Golomb-Rice, carried states */
    { "yuv420p",      37, 29,  9,   3,  1,  1, 0,  1, 0 },
    { "yuv420p10le",  64, 48, 12,   3,  1,  0, 0,  1, 0 },   /* the C2 shape in small */
    { "yuv444p16le",  40, 24,  4,   3,  1,  2, 1,  0, 0 },
    { "gray",         33, 17,  0,   1,  1,  1, 0, -1, 0 },   /* version 1: header in the key frame */
    { "ya8",          32, 32,  4,   3, 12,  1, 1,  1, 0 },
    { "yuva420p",     48, 32,  6,   3,  1,  0, 0,  1, 0 },
    { "bgr0",         48, 32,  4,   3,  1,  1, 1,  1, 0 },
    { "bgra",         40, 30,  4,   4,  1,  1, 0,  1, -2 },  /* version 4 */
    { "gbrp16le",     36, 20,  4,   3,  1,  0, 0,  1, 0 },
    { "rgb48le",      30, 22,  6,   3, 12,  0, 0,  0, 0 },
    { "yuv410p",      64, 40,  0,   0,  1,  0, 0, -1, 0 },   /* version 0 */
};

static void mutate(std::vector<uint8_t> &p)
{
    const int kind = rnd() % 8;
    const size_t n = p.size();
    if (!n)
        return;
    switch (kind) {
    case 0: for (int k = 1 + rnd() % 4; k > 0; k--) p[rnd() % n] ^= 1u << (rnd() % 8); break;
    case 1: for (int k = 1 + rnd() % 8; k > 0; k--) p[rnd() % n] = (uint8_t)rnd(); break;
    case 2: p.resize(rnd() % n); break;                                  /* truncated */
    case 3: {                                                            /* the size trailers */
        const size_t tail = n < 64 ? n : 64;
        for (int k = 1 + rnd() % 3; k > 0; k--) p[n - 1 - rnd() % tail] = (uint8_t)rnd();
        break;
    }
    case 4: {                                                            /* the first bytes: headers */
        const size_t head = n < 24 ? n : 24;
        for (int k = 1 + rnd() % 3; k > 0; k--) p[rnd() % head] = (uint8_t)rnd();
        break;
    }
    case 5: { const size_t a = rnd() % n, len = rnd() % (n - a) % 64; memset(&p[a], rnd() & 1 ? 0xFF : 0, len); break; }
    case 6: for (int k = 0, e = rnd() % 32; k < e; k++) p.push_back((uint8_t)rnd()); break;
    default: {                                                           /* a slice swapped for noise */
        const size_t a = rnd() % n, len = rnd() % (n - a);
        for (size_t i = 0; i < len; i++) p[a + i] = (uint8_t)rnd();
        break;
    }
    }
}

int main(int argc, char **argv)
{
    const int iters = argc > 1 ? atoi(argv[1]) : 300;
    if (argc > 2)
        g_rng ^= strtoull(argv[2], NULL, 0) * 0x2545F4914F6CDD1Dull;
    long decoded = 0, rejected = 0, opened = 0;
    for (size_t ci = 0; ci < sizeof(CASES) / sizeof(CASES[0]); ci++) {
        const Case &c = CASES[ci];
        Params p = { c.w, c.h, c.fmt, c.slices, c.level, c.gop, c.coder, c.context, c.crc, c.strict, 1, 0 };
        int err = 0;
        void *enc = ffv1emul_encoder_open(&p, &err);
        if (!enc) {
            fprintf(stderr, "case %zu (%s): encoder_open failed %d\n", ci, c.fmt, err);
            return 1;
        }
        const uint8_t *ex = NULL;
        const int exsize = ffv1emul_encoder_extradata(enc, &ex);
        /* three pictures: flat-ish gradient, noise, extremes */
        std::vector<std::vector<uint8_t>> pkts;
        for (int f = 0; f < 3; f++) {
            std::vector<uint8_t> planes[4];
            const uint8_t *pp[4] = { 0, 0, 0, 0 };
            int ls[4] = { 0, 0, 0, 0 };
            int bw = 0, rows = 0;
            const int np = ffv1emul_plane_geometry(c.fmt, c.w, c.h, 0, &bw, &rows);
            for (int k = 0; k < np; k++) {
                ffv1emul_plane_geometry(c.fmt, c.w, c.h, k, &bw, &rows);
                planes[k].resize((size_t)bw * rows);
                for (size_t i = 0; i < planes[k].size(); i++)
                    planes[k][i] = f == 0 ? (uint8_t)((i % bw) * 3 + i / bw) : f == 1 ? (uint8_t)rnd()
                                                                                       : (rnd() & 1 ? 0xFF : 0);
                pp[k] = planes[k].data();
                ls[k] = bw;
            }
            std::vector<uint8_t> out((size_t)c.w * c.h * 16 + 65536);
            int key = 0;
            const int n = ffv1emul_encode(enc, pp, ls, out.data(), (int)out.size(), &key);
            if (n <= 0) {
                fprintf(stderr, "case %zu (%s): encode failed %d\n", ci, c.fmt, n);
                return 1;
            }
            out.resize(n);
            pkts.push_back(out);
        }
        for (int form = 0; form < 2; form++) {
            if (form) setenv("FFV1_EMUL_LONE", "1", 1); else unsetenv("FFV1_EMUL_LONE");
            for (int it = 0; it < iters; it++) {
                /* damaged extradata now and then, damaged packets always */
                std::vector<uint8_t> x(ex, ex + (exsize > 0 ? exsize : 0));
                if (exsize > 0 && it % 5 == 4)
                    mutate(x);
                void *dec = ffv1emul_decoder_open(c.w, c.h, x.empty() ? NULL : x.data(), (int)x.size(), 1, &err);
                if (!dec) {
                    rejected++;
                    continue;
                }
                opened++;
                for (size_t f = 0; f < pkts.size(); f++) {
                    std::vector<uint8_t> m = pkts[f];
                    if (!(f == 0 && it % 3 == 0))             /* sometimes a sound key frame first */
                        mutate(m);
                    if (it % 7 == 6)
                        mutate(m);
                    /* an exact-size heap copy: reads past the packet's 64 padding bytes are caught */
                    uint8_t *planes[4];
                    int ls[4], key = 0;
                    const char *fmt = NULL;
                    const int r = ffv1emul_decode(dec, m.empty() ? (const uint8_t *)"" : m.data(), (int)m.size(),
                                                  planes, ls, &fmt, &key);
                    if (r >= 0) decoded++; else rejected++;
                }
                ffv1emul_decoder_close(dec);
            }
        }
        ffv1emul_encoder_close(enc);
    }
    printf("fuzz ok: %ld decoder opens, %ld packets decoded, %ld inputs rejected\n", opened, decoded, rejected);
    return 0;
}
