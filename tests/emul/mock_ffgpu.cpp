/*
 * tests/emul/mock_ffgpu.cpp -- TEST INFRASTRUCTURE: a stand-in for libffgpu.so that lets the
 * libavcodec glue (integration/ffv1_gpu.c) and the reference's own `ffmpeg` program built
 * with it (oracle/_ref/ffmpeg) run in a container without a GPU.
 *
 * It exports the entry points of include/ffgpu.h that the glue calls and keeps the
 * CONTRACT of the real library's launch-group pipeline -- pictures are held by pointer and
 * read only when their group is launched, results come back in order, FFGPU_EAGAIN /
 * FFGPU_EOF exactly where ffgpu_api.cu returns them, a group "runs" for a few polls before
 * it is finished -- while the pictures themselves go through the CPU emulation of the
 * product's device functions (emul.cpp).  What it checks is the host logic on the FFmpeg
 * side of the boundary: frame ownership, packet order and timestamps, drain and flush.
 *
 * It is built into tests/emul/mock/libffgpu.so and reaches a process only through
 * LD_LIBRARY_PATH set by tests/test_glue_cpu.py.  It is NOT a fallback: the product never
 * loads it, and the real library fails loudly without a CUDA device.
 */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <deque>
#include <string>
#include <vector>

#include "../../include/ffgpu.h"

struct Params {
    int width, height; const char *pix_fmt;
    int slices, level, gop_size, coder, context, slicecrc, strict, threads, bits_per_raw_sample;
};
extern "C" {
void *ffv1emul_encoder_open2(const Params *p, int pass1, int pass2, const char *stats_in, int *err);
int ffv1emul_encoder_extradata(void *h, const uint8_t **d);
void ffv1emul_encoder_info(void *h, int info[8]);
void ffv1emul_encoder_frame_props(void *h, int ps, int sar_num, int sar_den);
int ffv1emul_encode(void *h, const uint8_t *const planes[4], const int ls[4], uint8_t *out, int cap, int *key);
int ffv1emul_encoder_stats_out(void *h, char *buf, int cap);
void ffv1emul_encoder_close(void *h);
void *ffv1emul_decoder_open(int w, int h, const uint8_t *ex, int exsize, int threads, int *err);
int ffv1emul_decoder_damaged(void *h);
const char *ffv1emul_decoder_pix_fmt(void *h);
void ffv1emul_decoder_info(void *h, int info[8]);
int ffv1emul_decoder_probe(void *h, const uint8_t *pkt, int size);
void ffv1emul_decoder_frame_props(void *h, int props[5]);
int ffv1emul_decode(void *h, const uint8_t *pkt, int size, uint8_t *planes[4], int ls[4], const char **fmt, int *key);
void ffv1emul_decoder_close(void *h);
int ffv1emul_plane_geometry(const char *fmt, int w, int h, int plane, int *bw, int *rows);
}

static thread_local char g_err[256];
static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
extern "C" const char *ffgpu_last_error(void) { return g_err; }
extern "C" int ffgpu_cuda_push_context(void *) { return 0; }
extern "C" int ffgpu_cuda_pop_context(void) { return 0; }

enum { G_FREE, G_FILLING, G_RUNNING, G_DRAINING };
static const int RUN_POLLS = 3;                 /* a launched group is "busy" for this many polls */

static int env_int(const char *name, int def)
{
    const char *v = getenv(name);
    return v && *v ? atoi(v) : def;
}

/* ---------------------------------------------------------------------- encoder */
struct EncItem {
    ffgpu_picture pic;                           /* planes by POINTER: read at launch time */
    std::vector<uint8_t> pkt;
    int key = 0, err = 0;
};
struct EncGroup {
    int state = G_FREE, polls = 0;
    size_t drained = 0;
    std::vector<EncItem> items;
};
struct ffgpu_encoder {
    void *emul = nullptr;
    int w = 0, h = 0, batch = 1, depth = 1, fill = 0, head = 0, flushing = 0;
    size_t max_packet = 0;
    std::vector<EncGroup> groups;
};

extern "C" int ffgpu_ffv1_encode_init(ffgpu_encoder **penc, const ffgpu_enc_options *o)
{
    Params p = { o->width, o->height, o->pix_fmt, o->slices, o->level, o->gop_size, o->coder, o->context,
                 o->slicecrc, o->strict_std_compliance, 1, o->bits_per_raw_sample };
    int err = 0;
    ffgpu_encoder *e = new ffgpu_encoder;
    *penc = nullptr;
    e->emul = ffv1emul_encoder_open2(&p, o->pass1, o->pass2, o->stats_in, &err);
    if (!e->emul) {
        delete e;
        return fail(err, "encode_init: options rejected (%d)", err);
    }
    e->w = o->width; e->h = o->height;
    /* like the product: carried adaptive state = one picture per group, in order */
    e->batch = o->gop_size <= 1 ? (o->max_batch > 0 ? o->max_batch : env_int("MOCK_FFGPU_BATCH", 4)) : 1;
    e->depth = o->gop_size <= 1 ? (o->pipeline_depth > 0 ? o->pipeline_depth : 3) : 1;
    e->groups.resize(e->depth);
    e->max_packet = (size_t)o->width * o->height * 16 + 65536;
    *penc = e;
    return 0;
}

extern "C" int ffgpu_ffv1_encoder_extradata(const ffgpu_encoder *e, const uint8_t **data)
{
    return ffv1emul_encoder_extradata(e->emul, data);
}
extern "C" void ffgpu_ffv1_encoder_info(const ffgpu_encoder *e, int info[8]) { ffv1emul_encoder_info(e->emul, info); }
extern "C" size_t ffgpu_ffv1_encoder_max_packet(const ffgpu_encoder *e) { return e->max_packet; }

static void enc_launch(ffgpu_encoder *e, EncGroup *g)
{
    for (EncItem &it : g->items) {               /* the planes are read NOW, as the GPU would */
        const ffgpu_picture &p = it.pic;
        const int ps = !p.interlaced_frame ? 3 : 1 + !p.top_field_first;   /* ffv1enc.c:944-947 */
        it.pkt.resize(e->max_packet);
        ffv1emul_encoder_frame_props(e->emul, ps, p.sar_num, p.sar_den);
        const int n = ffv1emul_encode(e->emul, p.data, p.linesize, it.pkt.data(), (int)it.pkt.size(), &it.key);
        it.err = n < 0 ? n : 0;
        it.pkt.resize(n < 0 ? 0 : n);
    }
    g->state = G_RUNNING;
    g->polls = RUN_POLLS;
    g->drained = 0;
    e->fill = (e->fill + 1) % e->depth;
}

extern "C" int ffgpu_ffv1_encode_send_frame(ffgpu_encoder *e, const ffgpu_picture *pic)
{
    if (!e)
        return fail(FFGPU_EINVAL, "null encoder");
    EncGroup *g = &e->groups[e->fill];
    if (!pic) {
        e->flushing = 1;
        if (g->state == G_FILLING && !g->items.empty())
            enc_launch(e, g);
        return 0;
    }
    if (e->flushing)
        return fail(FFGPU_EOF, "send_frame after flush");
    if (g->state == G_RUNNING || g->state == G_DRAINING)
        return FFGPU_EAGAIN;                     /* every group is busy: receive first */
    if (g->state == G_FREE) {
        g->state = G_FILLING;
        g->items.clear();
    }
    EncItem it;
    it.pic = *pic;
    g->items.push_back(it);
    if ((int)g->items.size() == e->batch)
        enc_launch(e, g);
    return 0;
}

static int enc_receive(ffgpu_encoder *e, uint8_t *pkt, size_t cap, size_t *size, int *key, int64_t *pts)
{
    if (!e)
        return fail(FFGPU_EINVAL, "null encoder");
    EncGroup *g = &e->groups[e->head];
    if (g->state == G_FREE || g->state == G_FILLING) {
        if (e->flushing) {
            e->flushing = 0;                     /* drained: the handle accepts pictures again */
            return FFGPU_EOF;
        }
        return FFGPU_EAGAIN;
    }
    if (g->state == G_RUNNING) {
        const EncGroup *f = &e->groups[e->fill];
        const int must_wait = e->flushing || f->state == G_RUNNING || f->state == G_DRAINING;
        if (!must_wait && g->polls-- > 0)
            return FFGPU_EAGAIN;
        g->state = G_DRAINING;
    }
    EncItem &it = g->items[g->drained];
    int r = 0;
    if (it.err < 0)
        r = fail(it.err, "encoded frame too large (%d)", it.err);
    else if (!pkt) {                             /* peek */
        if (size) *size = it.pkt.size();
        return 0;
    } else if (it.pkt.size() > cap)
        return fail(FFGPU_ENOSPC, "packet buffer too small: need %zu bytes", it.pkt.size());
    else {
        memcpy(pkt, it.pkt.data(), it.pkt.size());
        if (size) *size = it.pkt.size();
        if (key) *key = it.key;
        if (pts) *pts = it.pic.pts;
    }
    if (++g->drained == g->items.size()) {
        g->state = G_FREE;
        g->items.clear();
        e->head = (e->head + 1) % e->depth;
    }
    return r;
}

extern "C" int ffgpu_ffv1_encode_receive_packet(ffgpu_encoder *e, uint8_t *pkt, size_t cap, size_t *size,
                                                int *key, int64_t *pts)
{
    if (!pkt)
        return fail(FFGPU_EINVAL, "null packet buffer");
    return enc_receive(e, pkt, cap, size, key, pts);
}
extern "C" int ffgpu_ffv1_encode_packet_ready(ffgpu_encoder *e, size_t *size)
{
    return enc_receive(e, nullptr, 0, size, nullptr, nullptr);
}

extern "C" int ffgpu_ffv1_encode_frame(ffgpu_encoder *e, const ffgpu_picture *pic, uint8_t *pkt, size_t cap,
                                       size_t *size, int *key)
{
    int r;
    int64_t pts;
    if ((r = ffgpu_ffv1_encode_send_frame(e, pic)) < 0)
        return r;
    if (e->groups[(e->fill)].state == G_FILLING)
        enc_launch(e, &e->groups[e->fill]);
    e->groups[e->head].polls = 0;
    return ffgpu_ffv1_encode_receive_packet(e, pkt, cap, size, key, &pts);
}

extern "C" int ffgpu_ffv1_encoder_stats_out(ffgpu_encoder *e, char *buf, size_t cap)
{
    const int r = ffv1emul_encoder_stats_out(e->emul, buf, (int)cap);
    return r < 0 ? fail(r, "stats_out (%d)", r) : r;
}

extern "C" int ffgpu_ffv1_encode_close(ffgpu_encoder *e)
{
    if (!e)
        return 0;
    ffv1emul_encoder_close(e->emul);
    delete e;
    return 0;
}

/* ---------------------------------------------------------------------- decoder */
struct DecItem {
    std::vector<uint8_t> pkt;
    int64_t pts = 0;
    int has_dst = 0, err = 0, damaged = 0;
    int props[5] = { 0, 0, 0, 0, 1 };
    ffgpu_picture_out dst;                       /* planes by POINTER: written at launch time */
    std::vector<uint8_t> held[4];                /* no destination yet: the picture waits here */
    int held_ls[4] = { 0, 0, 0, 0 };
};
struct DecGroup {
    int state = G_FREE, polls = 0;
    size_t drained = 0;
    std::vector<DecItem> items;
};
struct ffgpu_decoder {
    void *emul = nullptr;
    int w = 0, h = 0, batch = 1, depth = 1, fill = 0, head = 0, flushing = 0, max_batch_opt = 0, depth_opt = 0;
    std::vector<DecGroup> groups;
};

extern "C" int ffgpu_ffv1_decode_close(ffgpu_decoder *d);
extern "C" int ffgpu_ffv1_decode_init(ffgpu_decoder **pdec, const ffgpu_dec_options *o)
{
    int err = 0;
    ffgpu_decoder *d = new ffgpu_decoder;
    *pdec = nullptr;
    d->emul = ffv1emul_decoder_open(o->width, o->height, o->extradata, o->extradata_size, 1, &err);
    if (!d->emul) {
        delete d;
        return fail(err, "decode_init: invalid extradata (%d)", err);
    }
    d->w = o->width; d->h = o->height;
    d->max_batch_opt = o->max_batch; d->depth_opt = o->pipeline_depth;
    /* streams with extradata know their output format from here on (dec_setup_stream) */
    if (o->extradata_size > 0 && (err = ffv1emul_decoder_probe(d->emul, (const uint8_t *)"", 0)) < 0) {
        ffgpu_ffv1_decode_close(d);
        return fail(err, "format not supported");
    }
    *pdec = d;
    return 0;
}

extern "C" const char *ffgpu_ffv1_decoder_pix_fmt(const ffgpu_decoder *d) { return d ? ffv1emul_decoder_pix_fmt(d->emul) : nullptr; }
extern "C" void ffgpu_ffv1_decoder_info(const ffgpu_decoder *d, int info[8]) { ffv1emul_decoder_info(d->emul, info); }
extern "C" int ffgpu_ffv1_decoder_probe(ffgpu_decoder *d, const uint8_t *pkt, size_t size)
{
    if (!d || !pkt)
        return fail(FFGPU_EINVAL, "null argument");
    const int r = ffv1emul_decoder_probe(d->emul, pkt, (int)size);
    return r < 0 ? fail(r, "invalid packet (%d)", r) : 0;
}

/* streams whose every frame is a key frame run in groups; carried state runs one by one */
static void dec_shape(ffgpu_decoder *d)
{
    int info[8];
    if (!d->groups.empty())
        return;
    ffv1emul_decoder_info(d->emul, info);
    /* info[0] = version; only version 3+ extradata can promise intra-only, and the emulation
     * does not export that flag: the mock groups version >= 3 streams only when asked to */
    const int grouped = info[0] >= 3 && env_int("MOCK_FFGPU_DEC_GROUPS", 1);
    d->batch = grouped ? (d->max_batch_opt > 0 ? d->max_batch_opt : env_int("MOCK_FFGPU_BATCH", 4)) : 1;
    d->depth = grouped ? (d->depth_opt > 0 ? d->depth_opt : 3) : 1;
    d->groups.resize(d->depth);
}

static void copy_out(const ffgpu_picture_out *dst, const char *fmt, int w, int h, uint8_t *const planes[4], const int ls[4])
{
    int bw = 0, rows = 0;
    const int np = ffv1emul_plane_geometry(fmt, w, h, 0, &bw, &rows);
    for (int k = 0; k < np; k++) {
        ffv1emul_plane_geometry(fmt, w, h, k, &bw, &rows);
        for (int y = 0; y < rows; y++)
            memcpy(dst->data[k] + (ptrdiff_t)y * dst->linesize[k], planes[k] + (ptrdiff_t)y * ls[k], bw);
    }
}

static void dec_launch(ffgpu_decoder *d, DecGroup *g)
{
    for (DecItem &it : g->items) {
        uint8_t *planes[4];
        int ls[4], key = 0;
        const char *fmt = nullptr;
        const int r = ffv1emul_decode(d->emul, it.pkt.data(), (int)it.pkt.size(), planes, ls, &fmt, &key);
        it.err = r < 0 ? r : 0;
        if (r < 0)
            continue;
        it.damaged = ffv1emul_decoder_damaged(d->emul);
        ffv1emul_decoder_frame_props(d->emul, it.props);
        if (it.has_dst) {
            copy_out(&it.dst, fmt, d->w, d->h, planes, ls);      /* written NOW, as the download would */
        } else {
            int bw = 0, rows = 0;
            const int np = ffv1emul_plane_geometry(fmt, d->w, d->h, 0, &bw, &rows);
            for (int k = 0; k < np; k++) {
                ffv1emul_plane_geometry(fmt, d->w, d->h, k, &bw, &rows);
                it.held[k].resize((size_t)ls[k] * rows);
                memcpy(it.held[k].data(), planes[k], it.held[k].size());
                it.held_ls[k] = ls[k];
            }
        }
    }
    g->state = G_RUNNING;
    g->polls = RUN_POLLS;
    g->drained = 0;
    d->fill = (d->fill + 1) % d->depth;
}

extern "C" int ffgpu_ffv1_decode_send_packet(ffgpu_decoder *d, const uint8_t *pkt, size_t size, int64_t pts,
                                              const ffgpu_picture_out *dst)
{
    if (!d)
        return fail(FFGPU_EINVAL, "null decoder");
    if (!pkt) {
        d->flushing = 1;
        if (!d->groups.empty()) {
            DecGroup *g = &d->groups[d->fill];
            if (g->state == G_FILLING && !g->items.empty())
                dec_launch(d, g);
        }
        return 0;
    }
    if (d->flushing)
        return fail(FFGPU_EOF, "send_packet after flush");
    if (d->groups.empty()) {
        const int r = ffv1emul_decoder_probe(d->emul, pkt, (int)size);
        if (r < 0)
            return fail(r, "invalid packet (%d)", r);
        dec_shape(d);
    }
    DecGroup *g = &d->groups[d->fill];
    if (g->state == G_RUNNING || g->state == G_DRAINING)
        return FFGPU_EAGAIN;
    if (g->state == G_FREE) {
        g->state = G_FILLING;
        g->items.clear();
    }
    g->items.emplace_back();
    DecItem &it = g->items.back();
    it.pkt.assign(pkt, pkt + size);              /* the real library copies the packet as well */
    it.pts = pts;
    if (dst) {
        it.dst = *dst;
        it.has_dst = 1;
    }
    if ((int)g->items.size() == d->batch)
        dec_launch(d, g);
    return 0;
}

extern "C" int ffgpu_ffv1_decode_receive_frame(ffgpu_decoder *d, ffgpu_picture_out *out)
{
    if (!d)
        return fail(FFGPU_EINVAL, "null decoder");
    if (d->groups.empty()) {
        if (d->flushing) {
            d->flushing = 0;
            return FFGPU_EOF;
        }
        return FFGPU_EAGAIN;
    }
    DecGroup *g = &d->groups[d->head];
    if (g->state == G_FREE || g->state == G_FILLING) {
        if (d->flushing) {
            d->flushing = 0;
            return FFGPU_EOF;
        }
        return FFGPU_EAGAIN;
    }
    if (g->state == G_RUNNING) {
        const DecGroup *f = &d->groups[d->fill];
        const int must_wait = d->flushing || f->state == G_RUNNING || f->state == G_DRAINING;
        if (!must_wait && g->polls-- > 0)
            return FFGPU_EAGAIN;
        g->state = G_DRAINING;
    }
    DecItem &it = g->items[g->drained];
    int r = 0;
    if (it.err < 0) {
        r = fail(it.err, "invalid packet (%d)", it.err);
    } else {
        if (!it.has_dst) {
            if (!out)
                return fail(FFGPU_EINVAL, "no destination picture");
            uint8_t *planes[4] = { it.held[0].data(), it.held[1].data(), it.held[2].data(), it.held[3].data() };
            copy_out(out, ffv1emul_decoder_pix_fmt(d->emul), d->w, d->h, planes, it.held_ls);
        }
        if (out) {
            out->key_frame = it.props[0];
            out->interlaced_frame = it.props[1];
            out->top_field_first = it.props[2];
            out->sar_num = it.props[3];
            out->sar_den = it.props[4];
            out->damaged_slices = it.damaged;
            out->pts = it.pts;
        }
    }
    if (++g->drained == g->items.size()) {
        g->state = G_FREE;
        g->items.clear();
        d->head = (d->head + 1) % d->depth;
    }
    return r;
}

extern "C" int ffgpu_ffv1_decode_frame(ffgpu_decoder *d, const uint8_t *pkt, size_t size, ffgpu_picture_out *out,
                                       int *got_frame)
{
    int r;
    if (got_frame)
        *got_frame = 0;
    if (!d || !pkt || !out)
        return fail(FFGPU_EINVAL, "null argument");
    if ((r = ffgpu_ffv1_decode_send_packet(d, pkt, size, 0, out)) < 0)
        return r;
    if (d->groups[d->fill].state == G_FILLING)
        dec_launch(d, &d->groups[d->fill]);
    d->groups[d->head].polls = 0;
    if ((r = ffgpu_ffv1_decode_receive_frame(d, out)) < 0)
        return r;
    if (got_frame)
        *got_frame = 1;
    return (int)size;
}

extern "C" int ffgpu_ffv1_decode_close(ffgpu_decoder *d)
{
    if (!d)
        return 0;
    ffv1emul_decoder_close(d->emul);
    delete d;
    return 0;
}
