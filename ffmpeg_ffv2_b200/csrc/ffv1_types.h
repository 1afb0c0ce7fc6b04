/*
 * ffv1_types.h -- data shared between the host C layer (ffv1_host.c), the CUDA
 * kernels (ffv1_kernels.cu) and the API glue (ffgpu_api.cu).  Plain C.
 */
#ifndef FFGPU_FFV1_TYPES_H
#define FFGPU_FFV1_TYPES_H

#include <stddef.h>
#include <stdint.h>

#include "ffv1_rac.h"

#define FF_MAX_SLICES        1024   /* libavcodec/ffv1.h:77 */
#define FF_CONTEXT_SIZE      32     /* libavcodec/ffv1.h:52 */
#define FF_MAX_QUANT_TABLES  8
#define FF_MAX_CTX_INPUTS    5
#define FF_MAX_PLANES        4

#define FF_AC_GOLOMB   0
#define FF_AC_DEFAULT  1
#define FF_AC_CUSTOM   2

enum FFLayout {
    FF_LAY_PLANAR = 0,   /* gray / yuv / yuva planar, 1 or 2 bytes per sample */
    FF_LAY_YA8,          /* gray+alpha interleaved, 2 bytes per pixel         */
    FF_LAY_BGR32,        /* bgr0 / bgra: one little-endian u32 per pixel      */
    FF_LAY_GBRP,         /* planar G,B,R[,A] in 16-bit containers             */
    FF_LAY_RGB48,        /* packed R,G,B[,A] 16-bit little-endian             */
};

typedef struct FFPixFmt {
    const char *name;
    int layout;
    int depth;           /* native sample depth  */
    int hs, vs;          /* log2 chroma subsampling */
    int chroma, alpha;
    int nplanes;         /* memory planes        */
} FFPixFmt;

/* the resolved stream: what FFV1Context holds after encode_init / read_extra_header
 * (libavcodec/ffv1.h:79-141) */
typedef struct FFStream {
    int width, height;
    int version, micro_version;
    int ac;
    int colorspace;
    int bits;                        /* bits_per_raw_sample as coded (0 for v0) */
    int chroma_planes, hs, vs, transparency;
    int nh, nv;
    int ec, intra;
    int qt_count;
    int16_t qt[FF_MAX_QUANT_TABLES][FF_MAX_CTX_INPUTS][256];
    int ctx_count[FF_MAX_QUANT_TABLES];
    uint8_t *initial[FF_MAX_QUANT_TABLES];   /* [ctx_count][32] or NULL (=128) */
    uint8_t trans[256];              /* state_transition */
    int plane_sets;                  /* f->plane_count */
    int use32, packed_lsb;
    int context_model;               /* encoder: quant table used by every plane */
    const FFPixFmt *pf;
    FFRacTables def_tab;             /* ff_build_rac_states(0.05*2^32, 248) */
    FFRacTables cur_tab;             /* tables the slice coders run with    */
} FFStream;

/* luma rectangle of a slice (ffv1.c:122-141) */
typedef struct FFSliceRect {
    int x, y, w, h;
} FFSliceRect;

/* ---------------------------------------------------------------------- */
/* device-side descriptors                                                 */
/* ---------------------------------------------------------------------- */

/* one coded plane of a slice: where its samples live and how they are coded */
typedef struct FFDevPlane {
    int mem;           /* memory plane index (YCbCr layouts)              */
    int hs, vs;        /* subsampling relative to the luma rectangle      */
    int set;           /* plane-context set: state + quant table selector */
    int step, off;     /* bytes between samples / offset inside a pixel   */
} FFDevPlane;

/* constant for the lifetime of a codec handle; passed to kernels by value */
typedef struct FFDevParams {
    int width, height;
    int nslices, nh, nv;
    int ac;
    int colorspace;
    int layout;
    int sbits;              /* s->bits: raw sample depth (<=8 means 8)              */
    int cbits;              /* depth handed to fold(): 8 / sbits / 9 / sbits+1      */
    int packed_lsb, use32;
    int transparency, chroma_planes, hs, vs;
    int ncoded;             /* coded planes per slice                               */
    FFDevPlane cp[FF_MAX_PLANES];
    int nsets;              /* plane-context sets                                   */
    int set_base[FF_MAX_PLANES];   /* first global context index of each set (encoder)   */
    int set_qidx[FF_MAX_PLANES];   /* quant table of each set (encoder)                  */
    int total_ctx;          /* contexts per slice over all sets                     */
    int rgb_pixbytes;       /* bytes per pixel of packed RGB layouts                */
    /* device picture layout */
    size_t plane_off[FF_MAX_PLANES];
    int    pitch[FF_MAX_PLANES];
    int    rows[FF_MAX_PLANES];
    size_t frame_bytes;
    /* per-frame strides of the work arrays */
    size_t frame_tokens;    /* tokens per frame (sum over slices)                   */
    size_t frame_bs;        /* bitstream arena bytes per frame                      */
    size_t pkt_stride;      /* packet buffer bytes per frame                        */
    int    trailer;         /* 3 (+5 with ec) bytes after each slice payload, 0 for v<2 slice 0 */
    int    version, ec;
} FFDevParams;

/* per-slice constants (device array, nslices entries) */
typedef struct FFDevSlice {
    int x, y, w, h;         /* luma rectangle                                       */
    uint32_t tok_off;       /* first token of the slice inside a frame's token array */
    uint32_t ntok;
    uint32_t bs_off;        /* start of the slice's bitstream arena inside a frame   */
    uint32_t bs_cap;
    /* line table of the token stream for the Golomb run mode: up to 4 segments of
     * nlines lines of width w each; run_index is reset at the start of a segment */
    int seg_lines[FF_MAX_PLANES];
    int seg_w[FF_MAX_PLANES];
    int nseg;
} FFDevSlice;

/* range-coder state a slice starts from: the key-frame bit, the v0/v1 in-band header and
 * the v3 slice header are coded on the host (a few symbols per slice, ffv1enc.c:1203-1219,
 * :930-961); the pixel stream continues in the same coder on the device */
typedef struct FFRacPrefix {
    int32_t low, range, pending, run;
    uint32_t nbytes;        /* bytes already emitted (copied in front of the payload) */
    uint32_t byte_off;      /* offset of those bytes in the prefix byte arena         */
    uint32_t golomb_start;  /* Golomb: ac_byte_count, where the bit writer starts     */
    uint32_t pad;
    /* version 4: the slice header goes on with per-slice values only the device knows
     * (slice_rct_by/ry_coef from choose_rct_params, ffv1enc.c:951-959), so the host stops in
     * front of them and hands over the header's adaptive states as well; the device also
     * closes the coder of Golomb-Rice slices (ffv1enc.c:1076-1081) then */
    uint8_t hdr_state[FF_CONTEXT_SIZE];
} FFRacPrefix;

/* decoder: per (frame, slice) work item produced by the host packet parser */
typedef struct FFDecSlice {
    uint32_t pkt_off;       /* slice start inside the job's packet arena            */
    uint32_t size;          /* bytestream_end - bytestream_start                     */
    int32_t  low, range;    /* coder state after the host parsed the slice header    */
    uint32_t pos;
    int32_t  overread;
    int x, y, w, h;
    int qidx[FF_MAX_PLANES];
    int key_frame;          /* reset the adaptive states first                       */
    int skip;               /* header invalid: leave the rectangle untouched         */
    uint32_t golomb_start;  /* Golomb: bit reader starts here (ac_byte_count)        */
    int parse;              /* 1: the device checks the CRC and parses the slice     */
                            /* header itself (slices 1..n-1 of v3 packets)           */
    /* version 4 slice header, ffv1dec.c:230-241 (slice_reset_contexts is folded into
     * key_frame: both mean "reset the adaptive states first", ffv1dec.c:304) */
    int pcm;                /* slice_coding_mode == 1: raw bits, no prediction, no RCT */
    int rct_by, rct_ry;     /* slice_rct_by_coef / slice_rct_ry_coef (1, 1 before v4) */
} FFDecSlice;

#define FF_RES_NOT_DECODED 1  /* absent or skipped work item                             */
#define FF_RES_CRC_BAD     2  /* slice CRC mismatch (ffv1dec.c:905-922)                  */
#define FF_RES_HDR_BAD     4  /* decode_slice_header failed (ffv1dec.c:296-300)          */

typedef struct FFDecResult {
    uint32_t end_pos;       /* range coder: bytestream position after the terminator */
    int32_t  overread;
    int32_t  error;         /* decode_line returned AVERROR_INVALIDDATA              */
    int32_t  flags;         /* FF_RES_*                                              */
    int32_t  x, y, w, h;    /* rectangle from the slice header (device-parsed slices) */
    uint32_t size;          /* bytestream_end - bytestream_start actually used       */
    uint32_t pad[3];
} FFDecResult;

/* stream constants the device-side slice header parser needs */
typedef struct FFDecHdr {
    int micro_version;
    int qt_count;
    int ctx_cap;            /* contexts per set the state arena was sized for        */
    int ctx_count[FF_MAX_QUANT_TABLES];
} FFDecHdr;

#endif
