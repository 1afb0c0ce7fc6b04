/*
 * oracle/ffv1_oracle.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Public interface of the CPU restatement of the reference's FFV1 slice pixel
 * path (see ffv1_oracle.c for the per-function reference citations).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library; the product (libffgpu.so) never links or calls it.
 */
#ifndef FFV1_ORACLE_H
#define FFV1_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes: the values libavutil uses, so tests can compare with the reference */
#define FFV1O_EINVAL       (-22)
#define FFV1O_ENOMEM       (-12)
#define FFV1O_ENOSYS       (-38)
#define FFV1O_ENOSPC       (-28)
#define FFV1O_INVALIDDATA  (-1094995529)   /* AVERROR_INVALIDDATA = -MKTAG('I','N','D','A') */

/* user-facing encoder options: the AVOptions/AVCodecContext fields that
 * encode_init (ffv1enc.c:517-928) reads */
typedef struct FFV1OOptions {
    int width, height;
    const char *pix_fmt;      /* pixdesc name (little-endian spellings): "yuv420p10le", "bgr0", ... */
    int slices;               /* -slices, 0 = automatic                                   */
    int level;                /* -level, -99 = unknown                                    */
    int gop_size;             /* -g (libavcodec default 12)                               */
    int coder;                /* -coder: 0 rice, -2 range_def, 2 range_tab, 1 ac          */
    int context;              /* -context 0/1                                             */
    int slicecrc;             /* -slicecrc -1 auto / 0 / 1                                */
    int strict;               /* -strict (0 normal, -2 experimental)                      */
    int threads;              /* slice threads (pthread fan-out)                          */
    int bits_per_raw_sample;  /* 0 = from pix_fmt                                         */
} FFV1OOptions;

typedef struct FFV1OEncoder FFV1OEncoder;
typedef struct FFV1ODecoder FFV1ODecoder;

FFV1OEncoder *ffv1o_encoder_open(const FFV1OOptions *opt, int *err);
int  ffv1o_encoder_extradata(FFV1OEncoder *e, const uint8_t **data);
/* info: version, micro_version, ac, num_h_slices, num_v_slices, ec, bits_per_raw_sample, colorspace */
void ffv1o_encoder_info(FFV1OEncoder *e, int info[8]);
/* returns packet size (>0) or a negative error; *key = 1 for key frames */
int  ffv1o_encode(FFV1OEncoder *e, const uint8_t *const planes[4], const int linesize[4],
                  uint8_t *out, int cap, int *key);
void ffv1o_encoder_close(FFV1OEncoder *e);

FFV1ODecoder *ffv1o_decoder_open(int width, int height, const uint8_t *extradata,
                                 int extradata_size, int threads, int *err);
/* decodes into a frame owned by the decoder (double-buffered like the reference's
 * picture/last_picture); returns bytes consumed or a negative error */
int  ffv1o_decode(FFV1ODecoder *d, const uint8_t *pkt, int pkt_size,
                  uint8_t *planes[4], int linesize[4], const char **pix_fmt, int *key);
/* number of slices flagged damaged by the last ffv1o_decode call */
int  ffv1o_decoder_damaged(FFV1ODecoder *d);
void ffv1o_decoder_close(FFV1ODecoder *d);

/* plane geometry for a pix_fmt name: returns plane count, fills bytes-per-row / rows */
int  ffv1o_plane_geometry(const char *pix_fmt, int w, int h, int plane, int *bytewidth, int *rows);

/* small pieces exported for unit tests */
uint32_t ffv1o_crc32(uint32_t crc, const uint8_t *buf, int len);
void     ffv1o_default_state_transition(uint8_t one_state[256]);

#ifdef __cplusplus
}
#endif
#endif
