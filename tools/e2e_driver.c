/*
 * e2e_driver.c -- the host side of bench.py's end-to-end measurement, in C.
 *
 * It is what a transcoder built on the C ABI does: one host thread feeds pictures to
 * ffgpu_ffv1_encode_send_frame and collects packets, a second one hands those packets to
 * ffgpu_ffv1_decode_send_packet and collects pictures, all buffers in (pinned) HOST memory.
 * Only include/ffgpu.h entry points are called; nothing of the codec lives here.  Written
 * in C so that the timed region measures the library, not an interpreter loop around it.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "ffgpu.h"

typedef struct E2ERun {
    ffgpu_encoder *enc;
    ffgpu_decoder *dec;
    int nframes;                    /* pictures pushed through encode + decode          */
    int nsrc;
    const ffgpu_picture *src;       /* picture i is src[i % nsrc]                        */
    int ndst;
    const ffgpu_picture_out *dst;   /* picture i is decoded into dst[i % ndst]           */
    double timeout_s;
    /* results */
    uint8_t **pkt;                  /* [nframes] malloc'ed packets, in order             */
    size_t *pkt_size;               /* [nframes]                                         */
    int decoded;
    int damaged;
    int error;                      /* first failing return code, 0 if none              */
    char message[256];
    /* internal */
    _Atomic int produced;
    _Atomic int enc_finished;
    _Atomic int failed;
    double t0;
} E2ERun;

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

static void idle(void)
{
    struct timespec ts = { 0, 100000 };
    nanosleep(&ts, NULL);
}

static void set_error(E2ERun *r, int code, const char *where)
{
    int expected = 0;
    if (atomic_compare_exchange_strong(&r->failed, &expected, 1)) {
        r->error = code;
        snprintf(r->message, sizeof(r->message), "%s: %d: %s", where, code, ffgpu_last_error());
    }
}

static void *enc_thread(void *arg)
{
    E2ERun *r = (E2ERun *)arg;
    const size_t cap = ffgpu_ffv1_encoder_max_packet(r->enc);
    uint8_t *scratch = (uint8_t *)malloc(cap);
    int sent = 0, got = 0, done = 0;
    if (!scratch) {
        set_error(r, FFGPU_ENOMEM, "encoder scratch");
        done = 1;
    }
    while (!done && !atomic_load(&r->failed)) {
        int progressed = 0, ret;
        if (now_s() - r->t0 > r->timeout_s) {
            set_error(r, FFGPU_EXTERNAL, "encoder stalled");
            break;
        }
        if (sent < r->nframes) {
            ffgpu_picture pic = r->src[sent % r->nsrc];
            pic.pts = sent;
            ret = ffgpu_ffv1_encode_send_frame(r->enc, &pic);
            if (ret == 0) {
                progressed = 1;
                if (++sent == r->nframes && (ret = ffgpu_ffv1_encode_send_frame(r->enc, NULL)) < 0) {
                    set_error(r, ret, "encode flush");
                    break;
                }
            } else if (ret != FFGPU_EAGAIN) {
                set_error(r, ret, "encode_send_frame");
                break;
            }
        }
        for (;;) {
            size_t size = 0;
            int key = 0;
            int64_t pts = 0;
            ret = ffgpu_ffv1_encode_receive_packet(r->enc, scratch, cap, &size, &key, &pts);
            if (ret == FFGPU_EOF) {
                done = 1;
                break;
            }
            if (ret == FFGPU_EAGAIN)
                break;
            if (ret < 0 || got >= r->nframes || pts != got) {
                set_error(r, ret < 0 ? ret : FFGPU_EXTERNAL, "encode_receive_packet");
                done = 1;
                break;
            }
            r->pkt[got] = (uint8_t *)malloc(size ? size : 1);
            if (!r->pkt[got]) {
                set_error(r, FFGPU_ENOMEM, "packet");
                done = 1;
                break;
            }
            memcpy(r->pkt[got], scratch, size);
            r->pkt_size[got] = size;
            got++;
            atomic_store(&r->produced, got);
            progressed = 1;
        }
        if (!progressed && !done)
            idle();
    }
    free(scratch);
    atomic_store(&r->enc_finished, 1);
    return NULL;
}

static void *dec_thread(void *arg)
{
    E2ERun *r = (E2ERun *)arg;
    int sent = 0, got = 0, eos = 0, done = 0, damaged = 0;
    while (!done && !atomic_load(&r->failed)) {
        int progressed = 0, ret;
        if (now_s() - r->t0 > r->timeout_s) {
            set_error(r, FFGPU_EXTERNAL, "decoder stalled");
            break;
        }
        if (!eos && sent < atomic_load(&r->produced)) {
            ret = ffgpu_ffv1_decode_send_packet(r->dec, r->pkt[sent], r->pkt_size[sent], sent,
                                                &r->dst[sent % r->ndst]);
            if (ret == 0) {
                sent++;
                progressed = 1;
            } else if (ret != FFGPU_EAGAIN) {
                set_error(r, ret, "decode_send_packet");
                break;
            }
        } else if (!eos && atomic_load(&r->enc_finished) && sent == atomic_load(&r->produced)) {
            if ((ret = ffgpu_ffv1_decode_send_packet(r->dec, NULL, 0, 0, NULL)) < 0) {
                set_error(r, ret, "decode flush");
                break;
            }
            eos = 1;
            progressed = 1;
        }
        for (;;) {
            ffgpu_picture_out out;
            memset(&out, 0, sizeof(out));
            ret = ffgpu_ffv1_decode_receive_frame(r->dec, &out);
            if (ret == FFGPU_EOF) {
                done = 1;
                break;
            }
            if (ret == FFGPU_EAGAIN)
                break;
            if (ret < 0 || out.pts != got) {
                set_error(r, ret < 0 ? ret : FFGPU_EXTERNAL, "decode_receive_frame");
                done = 1;
                break;
            }
            damaged += out.damaged_slices;
            got++;
            progressed = 1;
        }
        if (!progressed && !done)
            idle();
    }
    r->decoded = got;
    r->damaged = damaged;
    return NULL;
}

/* runs nframes pictures through encoder and decoder; returns 0 or the first error code */
int ffgpu_e2e_run(E2ERun *r)
{
    pthread_t te, td;
    r->decoded = r->damaged = r->error = 0;
    r->message[0] = 0;
    atomic_store(&r->produced, 0);
    atomic_store(&r->enc_finished, 0);
    atomic_store(&r->failed, 0);
    r->t0 = now_s();
    if (pthread_create(&te, NULL, enc_thread, r))
        return FFGPU_EXTERNAL;
    if (pthread_create(&td, NULL, dec_thread, r)) {
        atomic_store(&r->failed, 1);
        pthread_join(te, NULL);
        return FFGPU_EXTERNAL;
    }
    pthread_join(te, NULL);
    pthread_join(td, NULL);
    return r->error;
}

void ffgpu_e2e_free_packets(E2ERun *r)
{
    for (int i = 0; i < r->nframes; i++) {
        free(r->pkt[i]);
        r->pkt[i] = NULL;
    }
}
