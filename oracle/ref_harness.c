/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Drives the UNMODIFIED reference FFV1 codec objects (ff_ffv1_encoder /
 * ff_ffv1_decoder, /root/reference/libavcodec/ffv1enc.c:1323, ffv1dec.c:1087)
 * through their own AVCodec vtable (.init/.encode2/.decode/.close), the same
 * boundary the product replaces.  The reference sources
 * are compiled where they lie under /root/reference by oracle/Makefile; nothing
 * is copied.  This file only supplies the handful of libavcodec/libavutil
 * services those objects import (frame/packet allocation, logging, the
 * slice-thread fan-out avctx->execute) so that no part of the reference's own
 * build system has to run.
 *
 * Exposes a tiny C API (ffv1ref_*) used by tests/ and by bench.py's
 * cpu_baseline / --impl reference legs only.
 */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "config.h"
#include "libavutil/avassert.h"
#include "libavutil/common.h"
#include "libavutil/frame.h"
#include "libavutil/imgutils.h"
#include "libavutil/log.h"
#include "libavutil/mem.h"
#include "libavutil/pixdesc.h"
#include "libavcodec/avcodec.h"
#include "libavcodec/internal.h"
#include "libavcodec/thread.h"
#include "libavcodec/ffv1.h"

/* -DHARNESS_GPU builds the same harness around the product's libavcodec glue
 * (integration/ffv1_gpu.c -> libffgpu.so) instead of the reference codec objects: both are
 * then driven through the identical AVCodec boundary and can be compared packet by packet. */
#ifdef HARNESS_GPU
extern AVCodec ff_ffv1_gpu_encoder;
extern AVCodec ff_ffv1_gpu_decoder;
void ff_ffv1_gpu_set_options(void *priv, int slicecrc, int coder, int context);
#define ff_ffv1_encoder ff_ffv1_gpu_encoder
#define ff_ffv1_decoder ff_ffv1_gpu_decoder
#define ffv1ref_encoder_open      ffv1glue_encoder_open
#define ffv1ref_encoder_extradata ffv1glue_encoder_extradata
#define ffv1ref_encoder_info      ffv1glue_encoder_info
#define ffv1ref_encode            ffv1glue_encode
#define ffv1ref_encoder_close     ffv1glue_encoder_close
#define ffv1ref_decoder_open      ffv1glue_decoder_open
#define ffv1ref_decode            ffv1glue_decode
#define ffv1ref_decoder_close     ffv1glue_decoder_close
#define ffv1ref_plane_geometry    ffv1glue_plane_geometry
#define ffv1ref_last_error        ffv1glue_last_error
#define ffv1ref_set_log_level     ffv1glue_set_log_level
#else
extern AVCodec ff_ffv1_encoder;
extern AVCodec ff_ffv1_decoder;
#endif

/* ------------------------------------------------------------------ */
/* libavutil / libavcodec services the FFV1 objects import             */
/* ------------------------------------------------------------------ */

/* libavcodec/bitstream.c:39 -- run-length table of the Golomb run mode. It is
 * data, declared in libavcodec/mathops.h; bitstream.c drags in the VLC builder,
 * so the 41 constants are restated here (checked against golden packets). */
const uint8_t ff_log2_run[41] = {
     0,  0,  0,  0,  1,  1,  1,  1,  2,  2,  2,  2,  3,  3,  3,  3,
     4,  4,  5,  5,  6,  6,  7,  7,  8,  9, 10, 11, 12, 13, 14, 15,
    16, 17, 18, 19, 20, 21, 22, 23, 24,
};

static int g_log_level = AV_LOG_ERROR;
static char g_last_log[512];

void av_log(void *avcl, int level, const char *fmt, ...)
{
    va_list ap;
    (void)avcl;
    va_start(ap, fmt);
    if (level <= AV_LOG_ERROR)
        vsnprintf(g_last_log, sizeof(g_last_log), fmt, ap);
    va_end(ap);
    if (level <= g_log_level) {
        va_start(ap, fmt);
        vfprintf(stderr, fmt, ap);
        va_end(ap);
    }
}

const char *av_default_item_name(void *ptr)
{
    return (*(AVClass **)ptr)->class_name;
}

size_t av_strlcpy(char *dst, const char *src, size_t size)
{
    size_t len = 0;
    while (++len < size && *src)
        *dst++ = *src++;
    if (len <= size)
        *dst = 0;
    return len + strlen(src) - 1;
}

int av_match_name(const char *name, const char *names)
{
    size_t n;
    if (!name || !names)
        return 0;
    n = strlen(name);
    while (*names) {
        const char *e = strchr(names, ',');
        size_t l = e ? (size_t)(e - names) : strlen(names);
        if (l == n && !strncmp(name, names, n))
            return 1;
        names += l + !!e;
    }
    return 0;
}

AVFrame *av_frame_alloc(void)
{
    AVFrame *f = calloc(1, sizeof(*f));
    if (f) {
        f->format = -1;
        f->pts = f->pkt_dts = AV_NOPTS_VALUE;
        f->sample_aspect_ratio = (AVRational){ 0, 1 };
    }
    return f;
}

/* Frames in this harness never own their planes: the harness owns every
 * buffer, so ref/unref are shallow. */
void av_frame_unref(AVFrame *f)
{
    if (!f)
        return;
    memset(f, 0, sizeof(*f));
    f->format = -1;
    f->pts = f->pkt_dts = AV_NOPTS_VALUE;
    f->sample_aspect_ratio = (AVRational){ 0, 1 };
}

void av_frame_free(AVFrame **f)
{
    if (f && *f) {
        free(*f);
        *f = NULL;
    }
}

int av_frame_ref(AVFrame *dst, const AVFrame *src)
{
    *dst = *src;
    return 0;
}

int av_image_check_sar(unsigned int w, unsigned int h, AVRational sar)
{
    int64_t scaled_dim;
    if (sar.den <= 0 || sar.num < 0)
        return AVERROR(EINVAL);
    if (!sar.num || sar.num == sar.den)
        return 0;
    if (sar.num < sar.den)
        scaled_dim = ((int64_t)w * sar.num + sar.den - 1) / sar.den;
    else
        scaled_dim = ((int64_t)h * sar.den + sar.num - 1) / sar.num;
    if (scaled_dim > INT_MAX)
        return AVERROR(EINVAL);
    return 0;
}

static void plane_geometry(const AVPixFmtDescriptor *d, int plane, int w, int h,
                           int *bytewidth, int *rows)
{
    int step = 0, c, sw = 0, sh = 0;
    for (c = 0; c < d->nb_components; c++)
        if (d->comp[c].plane == plane && d->comp[c].step > step)
            step = d->comp[c].step;
    if (plane == 1 || plane == 2) {
        sw = d->log2_chroma_w;
        sh = d->log2_chroma_h;
    }
    *bytewidth = AV_CEIL_RSHIFT(w, sw) * step;
    *rows      = AV_CEIL_RSHIFT(h, sh);
}

static int plane_count(const AVPixFmtDescriptor *d)
{
    int c, n = 0;
    for (c = 0; c < d->nb_components; c++)
        n = FFMAX(n, d->comp[c].plane + 1);
    return n;
}

void av_image_copy(uint8_t *dst_data[4], int dst_linesizes[4],
                   const uint8_t *src_data[4], const int src_linesizes[4],
                   enum AVPixelFormat pix_fmt, int width, int height)
{
    const AVPixFmtDescriptor *d = av_pix_fmt_desc_get(pix_fmt);
    int p, y, n = plane_count(d);
    for (p = 0; p < n; p++) {
        int bw, rows;
        plane_geometry(d, p, width, height, &bw, &rows);
        for (y = 0; y < rows; y++)
            memcpy(dst_data[p] + (ptrdiff_t)y * dst_linesizes[p],
                   src_data[p] + (ptrdiff_t)y * src_linesizes[p], bw);
    }
}

/* packet buffer: one persistent worst-case allocation per encoder, the role
 * avctx->internal->byte_buffer plays in libavcodec/encode.c:32-71 */
struct Pool;
typedef struct Harness {
    AVCodecContext *avctx;
    uint8_t *pktbuf;
    int64_t  pktbuf_size;
    AVFrame *in;
    uint8_t *framebuf[2];      /* decoder output double buffer */
    size_t   framebuf_size;
    int      framebuf_idx;
    int      threads;
    struct Pool *pool;
} Harness;

int ff_alloc_packet2(AVCodecContext *avctx, AVPacket *avpkt, int64_t size, int64_t min_size)
{
    Harness *h = avctx->opaque;
    (void)min_size;
    if (size < 0 || size > INT_MAX - AV_INPUT_BUFFER_PADDING_SIZE)
        return AVERROR(EINVAL);
    if (h->pktbuf_size < size + AV_INPUT_BUFFER_PADDING_SIZE) {
        free(h->pktbuf);
        h->pktbuf = malloc(size + AV_INPUT_BUFFER_PADDING_SIZE);
        if (!h->pktbuf) {
            h->pktbuf_size = 0;
            return AVERROR(ENOMEM);
        }
        h->pktbuf_size = size + AV_INPUT_BUFFER_PADDING_SIZE;
    }
    memset(avpkt, 0, sizeof(*avpkt));
    avpkt->pts = avpkt->dts = AV_NOPTS_VALUE;
    avpkt->pos = -1;
    avpkt->data = h->pktbuf;
    avpkt->size = size;
    return 0;
}

int ff_thread_get_buffer(AVCodecContext *avctx, ThreadFrame *tf, int flags)
{
    Harness *h = avctx->opaque;
    const AVPixFmtDescriptor *d = av_pix_fmt_desc_get(avctx->pix_fmt);
    AVFrame *f = tf->f;
    int p, n;
    size_t off = 0, need = 0;
    (void)flags;
    if (!d)
        return AVERROR(EINVAL);
    n = plane_count(d);
    for (p = 0; p < n; p++) {
        int bw, rows;
        plane_geometry(d, p, avctx->width, avctx->height, &bw, &rows);
        need += (size_t)FFALIGN(bw, 64) * rows + 64;
    }
    if (need > h->framebuf_size) {
        free(h->framebuf[0]);
        free(h->framebuf[1]);
        h->framebuf[0] = calloc(1, need);   /* zeroed: rows no slice covers stay deterministic */
        h->framebuf[1] = calloc(1, need);
        h->framebuf_size = need;
        if (!h->framebuf[0] || !h->framebuf[1])
            return AVERROR(ENOMEM);
    }
    h->framebuf_idx ^= 1;
    tf->owner[0] = tf->owner[1] = avctx;
    f->width  = avctx->width;
    f->height = avctx->height;
    f->format = avctx->pix_fmt;
    for (p = 0; p < n; p++) {
        int bw, rows;
        plane_geometry(d, p, avctx->width, avctx->height, &bw, &rows);
        f->data[p]     = h->framebuf[h->framebuf_idx] + off;
        f->linesize[p] = FFALIGN(bw, 64);
        off += (size_t)f->linesize[p] * rows + 64;
    }
    return 0;
}

void ff_thread_release_buffer(AVCodecContext *avctx, ThreadFrame *f)
{
    (void)avctx;
    if (f->f)
        av_frame_unref(f->f);
}

int ff_thread_ref_frame(ThreadFrame *dst, ThreadFrame *src)
{
    dst->owner[0] = src->owner[0];
    dst->owner[1] = src->owner[1];
    return av_frame_ref(dst->f, src->f);
}

#ifdef HARNESS_GPU
static AVCodecContext *g_glue_avctx;       /* av_new_packet has no context argument */
int ff_get_buffer(AVCodecContext *avctx, AVFrame *frame, int flags)
{
    ThreadFrame tf = { frame, { avctx, avctx }, NULL };
    return ff_thread_get_buffer(avctx, &tf, flags);
}
void av_packet_unref(AVPacket *pkt)
{
    memset(pkt, 0, sizeof(*pkt));
}
void avpriv_report_missing_feature(void *avc, const char *msg, ...)
{
    av_log(avc, AV_LOG_WARNING, "%s is not implemented\n", msg);
}
/* what the glue's send_frame / receive_packet / receive_frame callbacks use of libavcodec
 * and libavutil, in the harness's shallow-ownership model (frames never own their planes) */
AVFrame *av_frame_clone(const AVFrame *src)
{
    AVFrame *f = av_frame_alloc();
    if (f)
        *f = *src;
    return f;
}
void av_frame_move_ref(AVFrame *dst, AVFrame *src)
{
    *dst = *src;
    av_frame_unref(src);
}
int av_new_packet(AVPacket *pkt, int size)
{
    return ff_alloc_packet2(g_glue_avctx, pkt, size, 0);
}
AVPacket *av_packet_alloc(void)
{
    return calloc(1, sizeof(AVPacket));
}
void av_packet_free(AVPacket **pkt)
{
    if (pkt && *pkt) {
        free(*pkt);
        *pkt = NULL;
    }
}
/* the synchronous .decode callback is what this harness drives; receive_frame's packet
 * source is never reached */
int ff_decode_get_packet(AVCodecContext *avctx, AVPacket *pkt)
{
    (void)avctx; (void)pkt;
    return AVERROR(EAGAIN);
}
#endif
void ff_thread_finish_setup(AVCodecContext *avctx) { (void)avctx; }
void ff_thread_report_progress(ThreadFrame *f, int progress, int field) { (void)f; (void)progress; (void)field; }
void ff_thread_await_progress(ThreadFrame *f, int progress, int field) { (void)f; (void)progress; (void)field; }

/* avctx->execute: the slice fan-out of libavcodec/pthread_slice.c:95-112 and
 * libavutil/slicethread.c:65 -- persistent workers pull job indices from one
 * shared atomic counter; the calling thread takes part as well. */
typedef struct Pool {
    pthread_t *thr;
    int nthr;                 /* worker threads (callers excluded) */
    pthread_mutex_t mu;
    pthread_cond_t  go, done;
    unsigned gen;
    int stop, active;
    /* current batch */
    AVCodecContext *c;
    int (*func)(AVCodecContext *, void *);
    char *arg;
    int *ret;
    int count, size;
    int next;                 /* atomic job cursor */
} Pool;

static void pool_run_jobs(Pool *p)
{
    for (;;) {
        int i = __atomic_fetch_add(&p->next, 1, __ATOMIC_RELAXED);
        int r;
        if (i >= p->count)
            break;
        r = p->func(p->c, p->arg + (size_t)i * p->size);
        if (p->ret)
            p->ret[i] = r;
    }
}

static void *pool_worker(void *v)
{
    Pool *p = v;
    unsigned seen = 0;
    pthread_mutex_lock(&p->mu);
    for (;;) {
        while (!p->stop && p->gen == seen)
            pthread_cond_wait(&p->go, &p->mu);
        if (p->stop)
            break;
        seen = p->gen;
        pthread_mutex_unlock(&p->mu);
        pool_run_jobs(p);
        pthread_mutex_lock(&p->mu);
        if (--p->active == 0)
            pthread_cond_signal(&p->done);
    }
    pthread_mutex_unlock(&p->mu);
    return NULL;
}

static Pool *pool_new(int threads)
{
    Pool *p = calloc(1, sizeof(*p));
    int i;
    pthread_mutex_init(&p->mu, NULL);
    pthread_cond_init(&p->go, NULL);
    pthread_cond_init(&p->done, NULL);
    p->nthr = threads - 1;
    p->thr = calloc(FFMAX(p->nthr, 1), sizeof(*p->thr));
    for (i = 0; i < p->nthr; i++)
        pthread_create(&p->thr[i], NULL, pool_worker, p);
    return p;
}

static void pool_free(Pool *p)
{
    int i;
    if (!p)
        return;
    pthread_mutex_lock(&p->mu);
    p->stop = 1;
    pthread_cond_broadcast(&p->go);
    pthread_mutex_unlock(&p->mu);
    for (i = 0; i < p->nthr; i++)
        pthread_join(p->thr[i], NULL);
    free(p->thr);
    free(p);
}

static int harness_execute(AVCodecContext *c, int (*func)(AVCodecContext *c2, void *arg),
                           void *arg, int *ret, int count, int size)
{
    Harness *h = c->opaque;
    Pool *p = h->pool;
    int i;
    if (!p || count <= 1) {
        for (i = 0; i < count; i++) {
            int r = func(c, (char *)arg + (size_t)i * size);
            if (ret)
                ret[i] = r;
        }
        return 0;
    }
    pthread_mutex_lock(&p->mu);
    p->c = c; p->func = func; p->arg = arg; p->ret = ret;
    p->count = count; p->size = size; p->next = 0;
    p->active = p->nthr;
    p->gen++;
    pthread_cond_broadcast(&p->go);
    pthread_mutex_unlock(&p->mu);
    pool_run_jobs(p);
    pthread_mutex_lock(&p->mu);
    while (p->active)
        pthread_cond_wait(&p->done, &p->mu);
    pthread_mutex_unlock(&p->mu);
    return 0;
}

/* ------------------------------------------------------------------ */
/* public harness API                                                   */
/* ------------------------------------------------------------------ */

typedef struct FFV1RefParams {
    int width, height;
    const char *pix_fmt;        /* libavutil pixdesc name, e.g. "yuv420p10le" */
    int slices;                 /* -slices  (0 = encoder default)              */
    int level;                  /* -level   (-99 = unknown/default)            */
    int gop_size;               /* -g       (libavcodec default 12)            */
    int coder;                  /* -coder   0 rice, -2 range_def, 2 range_tab, 1 ac */
    int context;                /* -context 0/1                                */
    int slicecrc;               /* -slicecrc -1 auto, 0, 1                     */
    int strict;                 /* -strict  (0 normal, -2 experimental)        */
    int threads;                /* slice threads for avctx->execute            */
    int bits_per_raw_sample;    /* 0 = derive from pix_fmt                     */
} FFV1RefParams;

const char *ffv1ref_last_error(void) { return g_last_log; }
void ffv1ref_set_log_level(int l) { g_log_level = l; }

static Pool *pool_new(int threads);
static void pool_free(Pool *p);
static Harness *harness_new(const AVCodec *codec, int threads)
{
    Harness *h = calloc(1, sizeof(*h));
    AVCodecContext *a = calloc(1, sizeof(*a));
    if (!h || !a)
        return NULL;
    h->avctx = a;
    h->threads = threads;
    h->pool = threads > 1 ? pool_new(threads) : NULL;
    a->opaque = h;
    a->codec = codec;
    a->codec_id = codec->id;
    a->codec_type = AVMEDIA_TYPE_VIDEO;
    a->priv_data = calloc(1, codec->priv_data_size);
    *(const AVClass **)a->priv_data = codec->priv_class;
    a->internal = calloc(1, sizeof(*a->internal));
    a->execute = harness_execute;
    a->level = FF_LEVEL_UNKNOWN;
    a->gop_size = 12;
    a->thread_count = threads;
    a->active_thread_type = threads > 1 ? FF_THREAD_SLICE : 0;
    a->time_base = (AVRational){ 1, 25 };
    a->sample_aspect_ratio = (AVRational){ 0, 1 };
#if FF_API_CODER_TYPE
    a->coder_type = -1;
#endif
#if FF_API_CODED_FRAME
    a->coded_frame = av_frame_alloc();
#endif
    return h;
}

static void harness_free(Harness *h)
{
    if (!h)
        return;
    pool_free(h->pool);
#if FF_API_CODED_FRAME
    av_frame_free(&h->avctx->coded_frame);
#endif
    av_frame_free(&h->in);
    free(h->avctx->extradata);
    free(h->avctx->internal);
    free(h->avctx->priv_data);
    free(h->avctx);
    free(h->pktbuf);
    free(h->framebuf[0]);
    free(h->framebuf[1]);
    free(h);
}

/* two-pass coding: AV_CODEC_FLAG_PASS1 / _PASS2 and AVCodecContext.stats_in for the next open */
static int g_open_flags;
static const char *g_open_stats_in;
void *ffv1ref_encoder_open(const FFV1RefParams *p, int *err);
void *ffv1ref_encoder_open2(const FFV1RefParams *p, int pass1, int pass2, const char *stats_in, int *err)
{
    void *h;
    g_open_flags = (pass1 ? AV_CODEC_FLAG_PASS1 : 0) | (pass2 ? AV_CODEC_FLAG_PASS2 : 0);
    g_open_stats_in = stats_in;
    h = ffv1ref_encoder_open(p, err);
    g_open_flags = 0;
    g_open_stats_in = NULL;
    return h;
}

void *ffv1ref_encoder_open(const FFV1RefParams *p, int *err)
{
    Harness *h = harness_new(&ff_ffv1_encoder, p->threads);
    AVCodecContext *a;
    FFV1Context *s;
    int ret;
    if (!h) {
        *err = AVERROR(ENOMEM);
        return NULL;
    }
    a = h->avctx;
    s = a->priv_data;
    a->width  = p->width;
    a->height = p->height;
    a->pix_fmt = av_get_pix_fmt(p->pix_fmt);
    a->slices = p->slices;
    a->level = p->level;
    a->gop_size = p->gop_size;
    a->strict_std_compliance = p->strict;
    a->bits_per_raw_sample = p->bits_per_raw_sample;
    a->flags |= g_open_flags;
    a->stats_in = (char *)g_open_stats_in;
    /* AVOption defaults of ffv1enc.c:1291-1307, then the caller's values */
#ifdef HARNESS_GPU
    (void)s;
    ff_ffv1_gpu_set_options(a->priv_data, p->slicecrc, p->coder, p->context);
#else
    s->ec = p->slicecrc;
    s->ac = p->coder;
    s->context_model = p->context;
#endif
    g_last_log[0] = 0;
    ret = ff_ffv1_encoder.init(a);
    if (ret < 0) {
        *err = ret;
        ff_ffv1_encoder.close(a);
        harness_free(h);
        return NULL;
    }
    h->in = av_frame_alloc();
    *err = 0;
    return h;
}

#ifndef HARNESS_GPU
/* the flush call of a first pass (encode2 with a NULL frame, ffv1enc.c:1134-1177) leaves the
 * statistics in AVCodecContext.stats_out */
int ffv1ref_encoder_stats_out(void *hh, char *buf, int cap)
{
    Harness *h = hh;
    AVCodecContext *a = h->avctx;
    AVPacket pkt;
    int got = 0, n;
    if (!(a->flags & AV_CODEC_FLAG_PASS1) || !a->stats_out)
        return AVERROR(EINVAL);
    memset(&pkt, 0, sizeof(pkt));
    ff_ffv1_encoder.encode2(a, &pkt, NULL, &got);
    n = (int)strlen(a->stats_out);
    if (n >= cap)
        return AVERROR(ENOSPC);
    memcpy(buf, a->stats_out, n + 1);
    return n;
}
#endif

int ffv1ref_encoder_extradata(void *hh, const uint8_t **data)
{
    Harness *h = hh;
    *data = h->avctx->extradata;
    return h->avctx->extradata_size;
}

/* geometry chosen by encode_init, for tests */
void ffv1ref_encoder_info(void *hh, int info[8])
{
    Harness *h = hh;
#ifdef HARNESS_GPU
    memset(info, 0, 8 * sizeof(int));
    info[6] = h->avctx->bits_per_raw_sample;
#else
    FFV1Context *s = h->avctx->priv_data;
    info[0] = s->version;
    info[1] = s->micro_version;
    info[2] = s->ac;
    info[3] = s->num_h_slices;
    info[4] = s->num_v_slices;
    info[5] = s->ec;
    info[6] = s->bits_per_raw_sample;
    info[7] = s->colorspace;
#endif
}

int ffv1ref_encode(void *hh, const uint8_t *const planes[4], const int linesize[4],
                   uint8_t *out, int cap, int *key)
{
    Harness *h = hh;
    AVCodecContext *a = h->avctx;
    AVPacket pkt;
    int got = 0, ret, i;
    av_frame_unref(h->in);
    for (i = 0; i < 4; i++) {
        h->in->data[i] = (uint8_t *)planes[i];
        h->in->linesize[i] = linesize[i];
    }
    h->in->width = a->width;
    h->in->height = a->height;
    h->in->format = a->pix_fmt;
    h->in->pts = 0;
    memset(&pkt, 0, sizeof(pkt));
    g_last_log[0] = 0;
#ifdef HARNESS_GPU
    g_glue_avctx = a;
    /* send_frame / receive_packet: one picture, flush, one packet (the EOF re-arms the handle) */
    if ((ret = ff_ffv1_encoder.send_frame(a, h->in)) < 0)
        return ret;
    if ((ret = ff_ffv1_encoder.send_frame(a, NULL)) < 0)
        return ret;
    ret = ff_ffv1_encoder.receive_packet(a, &pkt);
    if (ret >= 0) {
        AVPacket eof;
        got = 1;
        memset(&eof, 0, sizeof(eof));
        ff_ffv1_encoder.receive_packet(a, &eof);      /* AVERROR_EOF */
    }
#else
    ret = ff_ffv1_encoder.encode2(a, &pkt, h->in, &got);
#endif
    if (ret < 0)
        return ret;
    if (!got)
        return 0;
    if (key)
        *key = !!(pkt.flags & AV_PKT_FLAG_KEY);
    if (out) {
        if (pkt.size > cap)
            return AVERROR(ENOSPC);
        memcpy(out, pkt.data, pkt.size);
    }
    return pkt.size;
}

void ffv1ref_encoder_close(void *hh)
{
    Harness *h = hh;
    if (!h)
        return;
    ff_ffv1_encoder.close(h->avctx);
    harness_free(h);
}

void *ffv1ref_decoder_open(int width, int height, const uint8_t *extradata, int extradata_size,
                           int threads, int *err)
{
    Harness *h = harness_new(&ff_ffv1_decoder, threads);
    AVCodecContext *a;
    int ret;
    if (!h) {
        *err = AVERROR(ENOMEM);
        return NULL;
    }
    a = h->avctx;
    a->width = width;
    a->height = height;
    a->pix_fmt = AV_PIX_FMT_NONE;
    if (extradata_size > 0) {
        a->extradata = calloc(1, extradata_size + AV_INPUT_BUFFER_PADDING_SIZE);
        memcpy(a->extradata, extradata, extradata_size);
        a->extradata_size = extradata_size;
    }
    g_last_log[0] = 0;
    ret = ff_ffv1_decoder.init(a);
    if (ret < 0) {
        *err = ret;
        ff_ffv1_decoder.close(a);
        harness_free(h);
        return NULL;
    }
    *err = 0;
    return h;
}

/* Decodes one packet. On success fills planes/linesize (pointing into a buffer
 * owned by the harness, valid until the next-but-one call), returns consumed
 * bytes; *pix_fmt_name receives the pixdesc name chosen by the decoder. */
int ffv1ref_decode(void *hh, const uint8_t *pkt_data, int pkt_size,
                   uint8_t *planes[4], int linesize[4], const char **pix_fmt_name,
                   int *key)
{
    Harness *h = hh;
    AVCodecContext *a = h->avctx;
    AVFrame *out = av_frame_alloc();
    AVPacket pkt;
    uint8_t *padded;
    int got = 0, ret, i;
    memset(&pkt, 0, sizeof(pkt));
    padded = calloc(1, pkt_size + AV_INPUT_BUFFER_PADDING_SIZE);
    memcpy(padded, pkt_data, pkt_size);
    pkt.data = padded;
    pkt.size = pkt_size;
    pkt.pts = pkt.dts = AV_NOPTS_VALUE;
    g_last_log[0] = 0;
    ret = ff_ffv1_decoder.decode(a, out, &got, &pkt);
    free(padded);
    if (ret >= 0 && got) {
        for (i = 0; i < 4; i++) {
            planes[i] = out->data[i];
            linesize[i] = out->linesize[i];
        }
        if (pix_fmt_name)
            *pix_fmt_name = av_get_pix_fmt_name(a->pix_fmt);
        if (key)
            *key = out->key_frame;
    } else if (ret >= 0) {
        ret = AVERROR(EAGAIN);
    }
    av_frame_free(&out);
    return ret;
}

void ffv1ref_decoder_close(void *hh)
{
    Harness *h = hh;
    if (!h)
        return;
    ff_ffv1_decoder.close(h->avctx);
    harness_free(h);
}

/* plane geometry helper shared with the Python tests */
int ffv1ref_plane_geometry(const char *pix_fmt, int w, int h, int plane, int *bytewidth, int *rows)
{
    const AVPixFmtDescriptor *d = av_pix_fmt_desc_get(av_get_pix_fmt(pix_fmt));
    if (!d || plane >= plane_count(d))
        return -1;
    plane_geometry(d, plane, w, h, bytewidth, rows);
    return plane_count(d);
}
