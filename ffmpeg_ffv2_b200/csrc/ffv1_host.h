/*
 * ffv1_host.h -- host-side (CPU, plain C) part of the FFV1 codec handle: option
 * resolution, global/slice headers, packet framing.  No pixel work happens here;
 * the per-sample path lives in ffv1_kernels.cu.
 */
#ifndef FFGPU_FFV1_HOST_H
#define FFGPU_FFV1_HOST_H

#include "../../include/ffgpu.h"
#include "ffv1_types.h"

#ifdef __cplusplus
extern "C" {
#endif

const FFPixFmt *ff_find_pixfmt(const char *name);
int  ff_bytes_per_pixel(const FFPixFmt *pf);
int  ff_plane_geometry(const FFPixFmt *pf, int w, int h, int plane, int *rowbytes, int *rows);

void ff_default_tables(FFRacTables *t);
void ff_install_custom(FFRacTables *t, const uint8_t trans[256]);
uint32_t ff_crc32(uint32_t crc, const uint8_t *buf, size_t len);

/* encode_init (ffv1enc.c:517-928) minus the 2-pass paths */
int  ff_stream_from_options(FFStream *s, const ffgpu_enc_options *o);
/* write_extradata (ffv1enc.c:396-467); *data is malloc'ed */
int  ff_write_extradata(FFStream *s, int gop_size, uint8_t **data, int *size);
/* read_extra_header (ffv1dec.c:413-528) */
int  ff_parse_extradata(FFStream *s, const uint8_t *data, int size);
/* output format selection of read_header (ffv1dec.c:597-739) */
int  ff_pick_decoder_format(FFStream *s);
void ff_stream_free(FFStream *s);

void ff_slice_rect(const FFStream *s, int i, FFSliceRect *r);

/* fills the kernel-facing description of the stream and its slices; returns 0 */
int  ff_fill_dev_params(const FFStream *s, int encoder, FFDevParams *P, FFDevSlice *slices);

/* Encoder: code everything that precedes the pixel data of slice i into its coder
 * (key-frame bit + v0/v1 header for slice 0, ffv1enc.c:1203-1219; slice header,
 * :930-961; Golomb: header termination, :1076-1081) and hand back the coder state. */
int  ff_enc_slice_prefix(const FFStream *s, int i, const FFSliceRect *r, int key_frame,
                         int picture_structure, int sar_num, int sar_den,
                         FFRacPrefix *pre, uint8_t *bytes, int cap);

/* Decoder: frame-level parse of one packet (decode_frame ffv1dec.c:837-931, read_header
 * :530-816, decode_slice_header :167-244, the Golomb hand-over :312-319).
 * out[] receives one work item per slice.  Returns the slice count or a negative error. */
typedef struct FFDecFrameInfo {
    int key_frame;
    int nslices;
    int interlaced_frame, top_field_first;
    int sar_num, sar_den;
    int crc_damaged;           /* slices whose CRC failed */
} FFDecFrameInfo;

typedef struct FFDecHostState {
    int key_frame_ok;
    int max_slices;
    int slice_count;                    /* of the last key frame */
    int device_parse;                   /* slices 1..n-1: CRC + header are left to the device */
    uint8_t damaged[FF_MAX_SLICES];     /* CRC / header damage of the current packet */
    FFSliceRect rect[FF_MAX_SLICES];    /* geometry of the current packet's slices   */
} FFDecHostState;

int  ff_dec_parse_packet(FFStream *s, FFDecHostState *hs, const uint8_t *pkt, size_t size,
                         uint32_t pkt_off, FFDecSlice *out, FFDecFrameInfo *info);

#ifdef __cplusplus
}
#endif
#endif
