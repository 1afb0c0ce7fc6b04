/*
 * ffv1_launch.h -- host-callable launchers of the CUDA kernels in ffv1_kernels.cu.
 * Every function enqueues on `stream` and returns the number of kernels it launched
 * (>= 0) or a negative FFGPU_EXTERNAL on a launch error.
 */
#ifndef FFGPU_FFV1_LAUNCH_H
#define FFGPU_FFV1_LAUNCH_H

#include "ffv1_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef void *ffk_stream;      /* cudaStream_t */

/* Which slice a coder thread takes.  The slices of a launch are sorted by weight (binary
 * decisions / bytes), heaviest first.  A launch lasts as long as its heaviest slices, and a
 * warp whose 32 lanes all carry heavy, unrelated slices is the slowest way to run them (every
 * iteration executes the divergent paths of 32 coders).  The heaviest slices -- those above
 * heavy_factor x the mean weight, at most a quarter of all -- therefore run `heavy_stride`
 * lanes apart (8 per warp for stride 4), the rest one per lane.  k_sched derives n_heavy on
 * the device from the sorted weights; the grid is sized for the worst case and surplus
 * threads leave at once. */
typedef struct FFSched {
    uint32_t n_heavy;               /* sorted items [0, n_heavy) are the heavy class     */
    uint32_t heavy_threads;         /* threads reserved for them (a multiple of 32)      */
} FFSched;

/* encoder work arrays of one launch group (all device pointers) */
typedef struct FFEncDev {
    const FFDevSlice *slices;       /* [nslices]                                        */
    const int16_t *qt;              /* [FF_MAX_QUANT_TABLES][FF_QT_STRIDE]              */
    const FFRacTables *tab;         /* transition tables of the slice coders            */
    const uint8_t *frames;          /* [nframes][frame_bytes]                           */
    uint32_t *tokens;               /* [nframes][frame_tokens]                          */
    uint8_t *state;                 /* [nstate_frames][nslices][total_ctx][32 | 8]      */
    const FFRacPrefix *prefix;      /* [nprefix_sets][nslices]                          */
    const uint8_t *prefix_bytes;    /* arena addressed by FFRacPrefix.byte_off          */
    const uint8_t *frame_prefix_set;/* [nframes] which prefix set a frame uses          */
    const uint8_t *frame_key;       /* [nframes] 1 = reset the adaptive states first    */
    uint8_t *bs;                    /* [nframes][frame_bs] slice bitstream arenas        */
    uint32_t *slice_bytes;          /* [nframes][nslices]                               */
    uint32_t *slice_off;            /* [nframes][nslices] offset inside the frame packet */
    uint32_t *pkt_size;             /* [nframes]                                        */
    uint32_t *pkt_off;              /* [nframes + 1] offset inside the packed output, in 16-byte units */
    uint32_t *overflow;             /* [1] sticky flag                                  */
    uint8_t *pkt;                   /* packed output of the whole group                 */
    int state_per_frame;            /* 1: state[frame][slice] (intra), 0: state[slice]   */
    int five;                       /* the encoder's quant table uses 5 context inputs   */
    /* longest-first scheduling of stage B: per-slice decision counts from stage A, sorted */
    uint32_t *weight;               /* [nframes][nslices]                               */
    uint32_t *weight_sorted;        /* scratch                                          */
    const uint32_t *iota;           /* 0,1,2,...                                        */
    uint32_t *order;                /* [nframes*nslices] slice ids, heaviest first      */
    void *sort_tmp;
    size_t sort_tmp_bytes;
    int lane_stride;                /* slice coders: 1 = every lane codes a slice ... 32 = one per warp */
    int legacy_stage_a;             /* 1: always k_symbolize_planar, -1: always the bulk-copy kernel, 0: by slice width */
    FFSched *sched;                 /* heavy / light classes of this launch, or NULL      */
    int heavy_stride;               /* lanes between two heavy slices (power of two)      */
    float heavy_factor;
    /* stage B in two halves (few-slice launches, see ff_chain_token): decision records of the
     * group, where each slice's records start, and whether the records fit (decided on the
     * device: incompressible pictures take the one-kernel coder instead) */
    uint16_t *rec;                  /* [rec_cap] or NULL: the split form is not configured */
    unsigned long long rec_cap;     /* records the arena holds                           */
    unsigned long long *rec_off;    /* [nframes*nslices + 1]                             */
    uint32_t *guard_rec;            /* [nframes*nslices] version 4 overflow guard position */
    int *split_ok;                  /* [1]                                               */
    /* two-pass coding: counters of a first pass (or NULL), initial states of a second pass
     * as one row per context of the slice, [total_ctx][32] (or NULL: every state starts at 128) */
    unsigned long long *rc_stat, *rc_stat2;
    int stat_ctx_count;
    const uint8_t *initial;
    /* version 4: slice_rct_by/ry_coef per (frame, slice), chosen on the device */
    int *rct;                       /* [nframes][nslices][2] or NULL (version <= 3)      */
    int32_t *rct_stat;              /* [nframes][nslices][16] scratch of the reduction   */
    void **events;                  /* optional cudaEvent_t[FFK_ENC_KERNELS + 1]: recorded    */
                                    /* before the first and after every kernel (profiling)  */
} FFEncDev;

/* temp bytes cub's radix sort needs for n (key,value) pairs */
size_t ffk_sort_tmp_bytes(int n);

/* kernels of one encode group, in launch order */
enum { FFK_SYMBOLIZE = 0, FFK_FILL_STATE, FFK_SORT, FFK_CODE, FFK_PACK_SLICE_SCAN,
       FFK_PACK_FRAME_SCAN, FFK_PACK_GATHER, FFK_ENC_KERNELS };
/* kernels of one decode group */
enum { FFK_DEC_INIT_STATE = 0, FFK_DEC_SORT, FFK_DECODE, FFK_DEC_KERNELS };

int ffk_encode_group(const FFDevParams *P, const FFEncDev *E, int nframes, ffk_stream stream);

typedef struct FFDecDev {
    const FFDecSlice *work;         /* [nframes][max_slices]                            */
    const int *nslices;             /* [nframes] slices present in each packet          */
    const int16_t *qt;
    const FFRacTables *tab;
    const uint8_t *initial;         /* [FF_MAX_QUANT_TABLES][max_ctx][32] or NULL        */
    const uint8_t *has_initial;     /* [FF_MAX_QUANT_TABLES] (host pointer)             */
    const uint8_t *pkt;             /* packet arena of the group                        */
    uint8_t *state;                 /* [nstate_frames][max_slices][total_ctx][32 | 8]   */
    int32_t *lines;                 /* [nframes*max_slices][ncoded][2][line_stride]     */
    int line_stride;
    /* a v3 slice header may name any rectangle of the picture (ffv1dec.c:176-184), wider than
     * the grid cell `lines` was sized for: such slices take a picture-wide scratch from this
     * small pool (wide_used is zeroed per launch); when it runs dry the slice is reported
     * damaged instead of writing past its scratch */
    int32_t *wide_lines;            /* [wide_count][ncoded][2][wide_stride]             */
    int wide_stride, wide_count;
    uint32_t *wide_used;
    uint8_t *frames;                /* [nframes][frame_bytes] output pictures           */
    FFDecResult *result;            /* [nframes][max_slices]                            */
    int max_slices;
    int max_ctx;                    /* contexts per set (largest quant table)           */
    int state_per_frame;
    int qt_count;                   /* quant table sets of the stream                   */
    FFDecHdr hdr;                   /* constants for device-side slice header parsing   */
    int gate_wait;                  /* sample set-up gating of the decode loop          */
    /* lazily created adaptive states (intra-only streams without initial-state tables):
     * one bit per (work item, context), zeroed per launch instead of filling the state arena */
    uint32_t *touched;              /* [nframes*max_slices][touched_words] or NULL      */
    int touched_words;
    int any_five;                   /* some quant table uses 5 context inputs           */
    int lane_stride;                /* 1: every lane decodes a slice; 32: one slice per warp */
    FFSched *sched;                 /* heavy / light classes of this launch, or NULL     */
    int heavy_stride;
    float heavy_factor;
    int generic;                    /* 1: use the generic slice decoder even where the planar one applies */
    uint32_t *weight;               /* [nframes*max_slices] slice byte counts            */
    uint32_t *weight_sorted;
    const uint32_t *iota;
    uint32_t *order;                /* work items, largest slice first                  */
    void *sort_tmp;
    size_t sort_tmp_bytes;
    void **events;                  /* optional cudaEvent_t[FFK_DEC_KERNELS + 1]          */
} FFDecDev;

int ffk_decode_group(const FFDevParams *P, const FFDecDev *D, int nframes, ffk_stream stream);

/* concealment: copy the rectangle of a damaged slice from the previous picture
 * (ffv1dec.c:940-969); rect in luma samples */
int ffk_conceal_rect(const FFDevParams *P, uint8_t *dst_frame, const uint8_t *src_frame,
                     int x, int y, int w, int h, int depth_gt8, ffk_stream stream);

/* copies done by the SMs between mapped pinned host memory and device memory (both ways):
 * up to 4 segments per launch, 16-byte aligned pointers.  If dyn_bytes is set, the size of
 * segment 0 is read from device memory at run time (0 bytes are copied when it exceeds
 * dyn_cap, and the caller falls back to a plain copy). */
typedef struct FFCopySeg {
    void *dst;
    const void *src;
    size_t bytes;
} FFCopySeg;
typedef struct FFCopyArgs {
    FFCopySeg seg[4];
    int nseg;
    const uint32_t *dyn_bytes;      /* size of segment 0 in units of 1 << dyn_shift bytes */
    int dyn_shift;
    size_t dyn_cap;
} FFCopyArgs;
int ffk_copy_segments(const FFCopyArgs *a, ffk_stream stream);

#ifdef __cplusplus
}
#endif
#endif
