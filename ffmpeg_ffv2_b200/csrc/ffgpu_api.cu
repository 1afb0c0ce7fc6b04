/*
 * ffgpu_api.cu -- the C ABI of include/ffgpu.h: handle management, device memory,
 * CUDA streams and the launch-group pipeline around the kernels of ffv1_kernels.cu.
 *
 * Execution model
 *   A handle owns `pipeline_depth` launch groups.  A group collects up to `max_batch`
 *   pictures (encoder) or packets (decoder) and enqueues its kernel chain once, on its
 *   own CUDA stream; the host thread meanwhile fills the next group, so H2D, kernels and
 *   D2H of different groups overlap.  Pictures cross PCIe on one upload stream (encoder)
 *   and one download stream (decoder) per handle, ordered against the group streams by
 *   events, so the copy engines serve the groups in order; packets, work items and
 *   result tables are moved by the SMs through mapped pinned memory (k_copy_segments)
 *   and do not queue behind the picture DMA.  Streams whose adaptive states carry from
 *   frame to frame (gop_size > 1) run one picture per group, in order.
 *
 * There is no CPU pixel path anywhere in this file: without a CUDA device every
 * pixel-path entry point fails with FFGPU_EXTERNAL.
 */
#include <cuda_runtime.h>

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/ffgpu.h"
#include "ffv1_host.h"
#include "ffv1_launch.h"
#include "ffv1_slice.cuh"

#define NPREFIX_SETS 8
#define MAX_DEPTH    8
#define DEC_WIDE_POOL 16        /* picture-wide line scratches per decode group (see FFDecDev) */

static thread_local char g_err[512];

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char *ffgpu_last_error(void) { return g_err; }

/* Frames that live on the GPU (AV_PIX_FMT_CUDA) come from libavutil's CUDA hw device context,
 * which owns a driver-API CUcontext of its own (hwcontext_cuda.c creates one with
 * cuCtxCreate).  Device pointers are only valid inside the context that allocated them, and
 * the CUDA runtime this library is written against adopts whichever context is current on
 * the calling thread: the glue therefore makes the hw device's context current around every
 * call into the library (the way nvenc.c:1329-1340 brackets its work), and all memory and
 * streams of the handle then live in that context too. */
typedef int (*cu_ctx_push_fn)(void *);
typedef int (*cu_ctx_pop_fn)(void **);

static void *driver_entry(const char *name)
{
    void *fn = NULL;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &st) != cudaSuccess ||
        st != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return NULL;
    }
    return fn;
}

extern "C" int ffgpu_cuda_push_context(void *cu_context)
{
    static cu_ctx_push_fn push;
    if (!push)
        push = (cu_ctx_push_fn)driver_entry("cuCtxPushCurrent");
    if (!push)
        return fail(FFGPU_EXTERNAL, "CUDA driver entry point cuCtxPushCurrent not available");
    return push(cu_context) ? fail(FFGPU_EXTERNAL, "cuCtxPushCurrent failed") : 0;
}

extern "C" int ffgpu_cuda_pop_context(void)
{
    static cu_ctx_pop_fn pop;
    void *old = NULL;
    if (!pop)
        pop = (cu_ctx_pop_fn)driver_entry("cuCtxPopCurrent");
    if (!pop)
        return fail(FFGPU_EXTERNAL, "CUDA driver entry point cuCtxPopCurrent not available");
    return pop(&old) ? fail(FFGPU_EXTERNAL, "cuCtxPopCurrent failed") : 0;
}
extern "C" int ffgpu_abi_version(void) { return FFGPU_ABI_VERSION; }

#define CK(call)                                                                            \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess)                                                              \
            return fail(FFGPU_EXTERNAL, "CUDA: %s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

/* FFGPU_TRACE=1: device timeline of the launch groups.  Every mark is a CUDA event recorded
 * on the group's stream; the list is printed (times relative to the first mark of the
 * process) when a handle is closed. */
#include <pthread.h>
static int trace_on(void)
{
    static int v = -1;
    if (v < 0)
        v = getenv("FFGPU_TRACE") != NULL;
    return v;
}
struct TraceMark {
    cudaEvent_t ev;
    const char *what;
    int group, n;
};
static TraceMark g_marks[1 << 14];
static int g_nmarks;
static cudaEvent_t g_trace_base;
static pthread_mutex_t g_trace_lock = PTHREAD_MUTEX_INITIALIZER;

static void trace_mark(cudaStream_t st, const char *what, int group, int n)
{
    if (!trace_on())
        return;
    pthread_mutex_lock(&g_trace_lock);
    if (!g_trace_base) {
        cudaEventCreate(&g_trace_base);
        cudaEventRecord(g_trace_base, st);
    }
    if (g_nmarks < (int)(sizeof(g_marks) / sizeof(g_marks[0]))) {
        TraceMark *m = &g_marks[g_nmarks++];
        cudaEventCreate(&m->ev);
        cudaEventRecord(m->ev, st);
        m->what = what;
        m->group = group;
        m->n = n;
    }
    pthread_mutex_unlock(&g_trace_lock);
}
static void trace_dump(void)
{
    if (!trace_on())
        return;
    pthread_mutex_lock(&g_trace_lock);
    for (int i = 0; i < g_nmarks; i++) {
        float ms = 0;
        cudaEventSynchronize(g_marks[i].ev);
        cudaEventElapsedTime(&ms, g_trace_base, g_marks[i].ev);
        fprintf(stderr, "[ffgpu] %10.2f ms  %-10s group %d (%d)\n", ms, g_marks[i].what, g_marks[i].group,
                g_marks[i].n);
        cudaEventDestroy(g_marks[i].ev);
    }
    g_nmarks = 0;
    pthread_mutex_unlock(&g_trace_lock);
}

extern "C" size_t ffgpu_ffv1_frame_layout(const char *pix_fmt, int width, int height,
                                          size_t plane_offset[4], int plane_pitch[4],
                                          int plane_rows[4], int plane_rowbytes[4])
{
    const FFPixFmt *pf = ff_find_pixfmt(pix_fmt);
    size_t off = 0;
    if (!pf || width <= 0 || height <= 0)
        return 0;
    for (int k = 0; k < 4; k++) {
        plane_offset[k] = 0;
        plane_pitch[k] = plane_rows[k] = plane_rowbytes[k] = 0;
    }
    for (int k = 0; k < pf->nplanes; k++) {
        int rb, rows;
        ff_plane_geometry(pf, width, height, k, &rb, &rows);
        plane_offset[k] = off;
        plane_pitch[k] = (rb + 255) & ~255;
        plane_rows[k] = rows;
        plane_rowbytes[k] = rb;
        off += (size_t)plane_pitch[k] * rows;
    }
    return off;
}

/* Slice coders per warp.  A launch with fewer work items than the GPU has resident lanes
 * (148 SMs x 16 warps x 32 lanes) spreads them out, down to one slice per warp: unrelated
 * slices in one warp execute each other's divergent paths, which only pays off when the
 * lanes are needed.  FFGPU_LANE_STRIDE overrides (1, 2, 4, ... 32). */
static int coder_lane_stride(long items)
{
    const char *env = getenv("FFGPU_LANE_STRIDE");
    int s = 32;
    if (env && atoi(env) >= 1 && atoi(env) <= 32)
        return atoi(env);
    while (s > 1 && items * s > 148L * 16 * 32)
        s >>= 1;
    return s;
}

/* heavy / light schedule (FFSched): FFGPU_HEAVY_STRIDE (0 switches the classes off) and
 * FFGPU_HEAVY_FACTOR (percent of the mean weight) are tuning hooks */
static int heavy_stride_opt(void)
{
    const char *env = getenv("FFGPU_HEAVY_STRIDE");
    const int v = env ? atoi(env) : 4;
    return v == 2 || v == 4 || v == 8 || v == 16 || v == 32 ? v : (v == 0 ? 0 : 4);
}
static float heavy_factor_opt(void)
{
    const char *env = getenv("FFGPU_HEAVY_FACTOR");
    const int v = env ? atoi(env) : 350;
    return v > 0 ? v / 100.0f : 3.5f;
}

/* quant tables in the layout the kernels index: [table][5*256 + flag] */
static void flatten_qt(const FFStream *s, int16_t *q)
{
    memset(q, 0, sizeof(int16_t) * FF_MAX_QUANT_TABLES * FF_QT_STRIDE);
    for (int i = 0; i < s->qt_count; i++) {
        memcpy(q + (size_t)i * FF_QT_STRIDE, s->qt[i], sizeof(s->qt[i]));
        q[(size_t)i * FF_QT_STRIDE + FF_MAX_CTX_INPUTS * 256] = s->qt[i][3][127] || s->qt[i][4][127];
    }
}

/* Copy a picture between caller memory and the device layout.  Planes whose caller linesize
 * equals the device pitch go as ONE linear copy (a pitched copy is one DMA descriptor per
 * row and measurably slower over PCIe), and neighbouring planes that are contiguous on both
 * sides without padding are merged into the same copy. */
static int copy_picture(const FFDevParams *P, const FFPixFmt *pf, int w, int h, uint8_t *const data[4],
                        const int linesize[4], uint8_t *d_frame, int to_device, cudaStream_t st)
{
    const cudaMemcpyKind kind = cudaMemcpyDefault;      /* host or device planes: taken from the pointers */
    (void)to_device;
    int k = 0;
    while (k < pf->nplanes) {
        int rb, rows;
        ff_plane_geometry(pf, w, h, k, &rb, &rows);
        if (!data[k] || linesize[k] < rb)
            return fail(FFGPU_EINVAL, "picture plane %d missing or linesize too small", k);
        uint8_t *dev = d_frame + P->plane_off[k];
        if (linesize[k] == P->pitch[k] && rows > 0) {
            size_t bytes = (size_t)(rows - 1) * P->pitch[k] + rb;
            int last = k;
            /* extend over following planes while both sides stay gap-free */
            while (last + 1 < pf->nplanes && data[last + 1] && rb == P->pitch[last] &&
                   linesize[last + 1] == P->pitch[last + 1] &&
                   P->plane_off[last + 1] == P->plane_off[last] + (size_t)rows * P->pitch[last] &&
                   data[last + 1] == data[last] + (size_t)rows * P->pitch[last]) {
                last++;
                ff_plane_geometry(pf, w, h, last, &rb, &rows);
                if (rows <= 0)
                    break;
                bytes = (size_t)(P->plane_off[last] - P->plane_off[k]) + (size_t)(rows - 1) * P->pitch[last] + rb;
            }
            if (to_device)
                CK(cudaMemcpyAsync(dev, data[k], bytes, kind, st));
            else
                CK(cudaMemcpyAsync(data[k], dev, bytes, kind, st));
            k = last + 1;
        } else {
            if (to_device)
                CK(cudaMemcpy2DAsync(dev, P->pitch[k], data[k], linesize[k], rb, rows, kind, st));
            else
                CK(cudaMemcpy2DAsync(data[k], linesize[k], dev, P->pitch[k], rb, rows, kind, st));
            k++;
        }
    }
    return 0;
}


/* ====================================================================== */
/* pageable caller memory: pinned staging + a small pool of copy threads   */
/* ====================================================================== */
/* AVFrames are ordinary (pageable) memory.  A cudaMemcpyAsync on such a pointer is staged by
 * the driver and blocks the calling thread -- for a download until the kernels before it have
 * finished -- which would serialise the launch-group pipeline.  Pictures in pageable memory
 * therefore cross through pinned staging buffers owned by the handle: the host side of the
 * copy is a plain memcpy, split over a few worker threads (one core moves ~10 GB/s, a 4K
 * 10-bit picture is 25 MB), and the PCIe side is one linear asynchronous copy per picture.
 * Pinned, registered or device pointers skip all of this.  FFGPU_COPY_THREADS (default 8,
 * 1 = the calling thread only) sizes the pool. */
enum { MEM_PAGEABLE = 0, MEM_PINNED, MEM_DEVICE };

static int mem_kind(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return MEM_PAGEABLE;
    }
    if (a.type == cudaMemoryTypeUnregistered)
        return MEM_PAGEABLE;
    return a.type == cudaMemoryTypeHost ? MEM_PINNED : MEM_DEVICE;
}

/* does a picture in caller memory go through the pinned staging of the handle?  Pageable
 * memory always; bottom-up pictures (negative linesize) in any HOST memory as well, because a
 * pitched DMA cannot walk backwards.  -1: bottom-up planes in device memory, not supported. */
static int needs_staging(const void *plane0, const int linesize[4])
{
    const int kind = mem_kind(plane0);
    int bottom_up = 0;
    for (int k = 0; k < 4; k++)
        bottom_up |= linesize[k] < 0;
    if (bottom_up && kind == MEM_DEVICE)
        return -1;
    return kind == MEM_PAGEABLE || bottom_up;
}

struct CopyTask {
    uint8_t *dst;
    const uint8_t *src;
    ptrdiff_t dst_pitch, src_pitch;     /* signed: a bottom-up picture has a negative linesize */
    size_t rowbytes;
    int rows;
};

#define POOL_MAX_THREADS 16
#define POOL_MAX_CHUNKS  64
static struct {
    pthread_once_t once;
    pthread_mutex_t use;            /* one parallel copy at a time; others copy inline */
    pthread_mutex_t lock;
    pthread_cond_t wake, done;
    pthread_t thr[POOL_MAX_THREADS];
    int nthreads;
    CopyTask chunk[POOL_MAX_CHUNKS];
    int nchunks, next, finished;
    uint64_t generation;
} g_pool = { PTHREAD_ONCE_INIT, PTHREAD_MUTEX_INITIALIZER, PTHREAD_MUTEX_INITIALIZER,
             PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER, {0}, 0, {{0}}, 0, 0, 0, 0 };

static void copy_rows(const CopyTask *t)
{
    if (t->dst_pitch == t->src_pitch && t->dst_pitch > 0 && t->rowbytes == (size_t)t->dst_pitch) {
        memcpy(t->dst, t->src, (size_t)t->rows * t->rowbytes);
        return;
    }
    for (int y = 0; y < t->rows; y++)
        memcpy(t->dst + (ptrdiff_t)y * t->dst_pitch, t->src + (ptrdiff_t)y * t->src_pitch, t->rowbytes);
}

static void *pool_worker(void *)
{
    uint64_t seen = 0;
    pthread_mutex_lock(&g_pool.lock);
    for (;;) {
        while (g_pool.generation == seen || g_pool.next >= g_pool.nchunks) {
            seen = g_pool.generation;
            pthread_cond_wait(&g_pool.wake, &g_pool.lock);
        }
        const int c = g_pool.next++;
        pthread_mutex_unlock(&g_pool.lock);
        copy_rows(&g_pool.chunk[c]);
        pthread_mutex_lock(&g_pool.lock);
        if (++g_pool.finished == g_pool.nchunks)
            pthread_cond_signal(&g_pool.done);
    }
    return NULL;
}

static void pool_start(void)
{
    const char *env = getenv("FFGPU_COPY_THREADS");
    int n = env ? atoi(env) : 8;
    if (n < 1) n = 1;
    if (n > POOL_MAX_THREADS) n = POOL_MAX_THREADS;
    for (int i = 0; i < n - 1; i++) {              /* the calling thread is one of the n */
        if (pthread_create(&g_pool.thr[g_pool.nthreads], NULL, pool_worker, NULL) != 0)
            break;
        pthread_detach(g_pool.thr[g_pool.nthreads]);
        g_pool.nthreads++;
    }
}

/* copy the planes described by tasks[], in parallel when the pool is free */
static void par_copy(const CopyTask *tasks, int ntasks)
{
    pthread_once(&g_pool.once, pool_start);
    if (g_pool.nthreads == 0) {
        for (int i = 0; i < ntasks; i++)
            copy_rows(&tasks[i]);
        return;
    }
    /* one picture at a time through the pool: an encoder and a decoder thread that both
     * stage pictures queue here, each copy then runs at the pool's full width */
    pthread_mutex_lock(&g_pool.use);
    pthread_mutex_lock(&g_pool.lock);
    g_pool.nchunks = 0;
    for (int i = 0; i < ntasks; i++) {
        /* about 1 MB per chunk, at most POOL_MAX_CHUNKS / ntasks chunks per plane */
        const size_t bytes = (size_t)tasks[i].rows * tasks[i].rowbytes;
        int parts = (int)(bytes >> 20) + 1;
        const int most = POOL_MAX_CHUNKS / ntasks;
        if (parts > most) parts = most;
        if (parts > tasks[i].rows) parts = tasks[i].rows > 0 ? tasks[i].rows : 1;
        for (int k = 0; k < parts; k++) {
            CopyTask c = tasks[i];
            const int y0 = (int)((long)tasks[i].rows * k / parts), y1 = (int)((long)tasks[i].rows * (k + 1) / parts);
            c.dst += (ptrdiff_t)y0 * c.dst_pitch;
            c.src += (ptrdiff_t)y0 * c.src_pitch;
            c.rows = y1 - y0;
            g_pool.chunk[g_pool.nchunks++] = c;
        }
    }
    g_pool.next = 0;
    g_pool.finished = 0;
    g_pool.generation++;
    pthread_cond_broadcast(&g_pool.wake);
    while (g_pool.next < g_pool.nchunks) {         /* the caller works too */
        const int c = g_pool.next++;
        pthread_mutex_unlock(&g_pool.lock);
        copy_rows(&g_pool.chunk[c]);
        pthread_mutex_lock(&g_pool.lock);
        g_pool.finished++;
    }
    while (g_pool.finished < g_pool.nchunks)
        pthread_cond_wait(&g_pool.done, &g_pool.lock);
    pthread_mutex_unlock(&g_pool.lock);
    pthread_mutex_unlock(&g_pool.use);
}

/* a picture between caller planes and a staging buffer in the device layout */
static int stage_picture(const FFDevParams *P, const FFPixFmt *pf, int w, int h, uint8_t *const data[4],
                         const int linesize[4], uint8_t *stage, int to_stage)
{
    CopyTask t[4];
    int n = 0;
    for (int k = 0; k < pf->nplanes; k++) {
        int rb, rows;
        ff_plane_geometry(pf, w, h, k, &rb, &rows);
        /* a negative linesize (bottom-up picture, e.g. after vflip) is as good as a positive one,
         * like for the reference's pointer arithmetic (ffv1enc.c:283, ffv1dec.c:141) */
        if (!data[k] || (linesize[k] < 0 ? -(long)linesize[k] : (long)linesize[k]) < rb)
            return fail(FFGPU_EINVAL, "picture plane %d missing or linesize too small", k);
        if (to_stage) {
            t[n].dst = stage + P->plane_off[k]; t[n].dst_pitch = (ptrdiff_t)P->pitch[k];
            t[n].src = data[k];                 t[n].src_pitch = (ptrdiff_t)linesize[k];
        } else {
            t[n].dst = data[k];                 t[n].dst_pitch = (ptrdiff_t)linesize[k];
            t[n].src = stage + P->plane_off[k]; t[n].src_pitch = (ptrdiff_t)P->pitch[k];
        }
        t[n].rowbytes = (size_t)rb;
        t[n].rows = rows;
        n++;
    }
    par_copy(t, n);
    return 0;
}

static int upload_picture(const FFDevParams *P, const FFPixFmt *pf, int w, int h,
                          const ffgpu_picture *pic, uint8_t *d_frame, cudaStream_t st)
{
    uint8_t *data[4];
    for (int k = 0; k < 4; k++)
        data[k] = (uint8_t *)pic->data[k];
    return copy_picture(P, pf, w, h, data, pic->linesize, d_frame, 1, st);
}

/* ====================================================================== */
/* encoder                                                                 */
/* ====================================================================== */
enum { JOB_FREE = 0, JOB_FILLING, JOB_RUNNING, JOB_DRAINING };

struct EncJob {
    cudaStream_t stream;
    cudaEvent_t done;
    cudaEvent_t uploaded;       /* the group's pictures have arrived (recorded on up_stream) */
    int n, state, drained, fetched;
    uint8_t *d_frames;
    uint32_t *d_tokens;
    uint8_t *d_state;
    uint8_t *d_bs;
    uint32_t *d_slice_bytes, *d_slice_off, *d_pkt_size, *d_pkt_off, *d_overflow;
    uint8_t *d_pkt;
    uint8_t *d_frame_set, *d_frame_key;
    uint32_t *d_weight, *d_weight_sorted, *d_order;
    void *d_sort_tmp;
    size_t sort_tmp_bytes;
    FFSched *d_sched;
    /* stage B in two halves (few-slice streams): decision records */
    uint16_t *d_rec;
    unsigned long long rec_cap, *d_rec_off;
    uint32_t *d_guard_rec;
    int *d_split_ok;
    int *d_rct;                 /* version 4: RCT coefficients per (picture, slice) */
    int32_t *d_rct_stat;
    int pkt_owned;
    /* pinned host */
    uint8_t *h_frame_set, *h_frame_key;
    uint32_t *h_pkt_size, *h_pkt_off, *h_overflow;
    uint8_t *h_pkt;
    size_t h_pkt_cap;
    uint8_t *h_stage;           /* pinned staging for pictures in pageable memory (lazily) */
    int64_t *pts;
    int *key;
};

struct PrefixSet {
    int valid, key, ps, sar_num, sar_den;
};

struct ffgpu_encoder {
    ffgpu_enc_options opt;
    char pix_fmt[32];
    FFStream s;
    FFDevParams P;
    FFDevSlice *h_slices;
    uint8_t *extradata;
    int extradata_size;
    int intra;                  /* every frame is a key frame -> groups of many pictures */
    int max_batch, depth;
    int prefix_stride;
    /* device */
    int dev_ready;
    FFDevSlice *d_slices;
    int16_t *d_qt;
    FFRacTables *d_tab;
    FFRacPrefix *d_prefix;
    uint8_t *d_prefix_bytes;
    uint32_t *d_iota;           /* 0,1,2,... values for the slice sort */
    uint8_t *d_state_shared;    /* carried-state streams: one arena for all groups */
    PrefixSet sets[NPREFIX_SETS];
    EncJob jobs[MAX_DEPTH];
    /* all picture uploads go through ONE stream: the H2D engine then serves the groups in
     * order (with one stream per group it interleaves them, the oldest group finishes last
     * and in-order packet delivery stalls behind it) */
    cudaStream_t up_stream;
    int fill, head;             /* ring positions */
    int flushing, eof;
    int64_t picture_number;
    uint64_t launches;
    int profile;                        /* record events around every kernel of device batches */
    void *events[FFK_ENC_KERNELS + 1];
    /* two-pass coding */
    unsigned long long *d_rc_stat, *d_rc_stat2;   /* first pass: decision counters (device) */
    uint8_t *d_initial;                           /* second pass: [total_ctx][32]           */
    uint64_t gob_count;                           /* key frames coded (ffv1enc.c:1206)      */
    char *stats_in;
    /* more than one GPU: this handle only routes; sub[i] is a complete encoder on devices[i].
     * Picture k belongs to sub[(k / chunk) % nsub]; packets are returned in that order. */
    int nsub, chunk, send_blocked;
    ffgpu_encoder *sub[FFGPU_MAX_DEVICES];
    uint64_t in_count, out_count;
};

static int enc_free_job(EncJob *j)
{
    cudaFree(j->d_frames); cudaFree(j->d_tokens); cudaFree(j->d_state); cudaFree(j->d_bs);
    cudaFree(j->d_slice_bytes); cudaFree(j->d_slice_off); cudaFree(j->d_pkt_size);
    cudaFree(j->d_pkt_off); cudaFree(j->d_overflow);
    if (j->pkt_owned)
        cudaFree(j->d_pkt);
    cudaFree(j->d_frame_set); cudaFree(j->d_frame_key);
    cudaFree(j->d_weight); cudaFree(j->d_weight_sorted); cudaFree(j->d_order); cudaFree(j->d_sort_tmp);
    cudaFree(j->d_rct); cudaFree(j->d_rct_stat); cudaFree(j->d_sched);
    cudaFree(j->d_rec); cudaFree(j->d_rec_off); cudaFree(j->d_guard_rec); cudaFree(j->d_split_ok);
    cudaFreeHost(j->h_frame_set); cudaFreeHost(j->h_frame_key); cudaFreeHost(j->h_pkt_size);
    cudaFreeHost(j->h_pkt_off); cudaFreeHost(j->h_overflow); cudaFreeHost(j->h_pkt);
    cudaFreeHost(j->h_stage);
    free(j->pts); free(j->key);
    if (j->done) cudaEventDestroy(j->done);
    if (j->uploaded) cudaEventDestroy(j->uploaded);
    if (j->stream) cudaStreamDestroy(j->stream);
    memset(j, 0, sizeof(*j));
    return 0;
}

/* a host array that exists only to be uploaded: device allocation + copy, then free(host) */
static int upload_and_free(void **dev, void *host, size_t bytes)
{
    cudaError_t ce;
    if (!host)
        return fail(FFGPU_ENOMEM, "out of memory");
    ce = cudaMalloc(dev, bytes);
    if (ce == cudaSuccess)
        ce = cudaMemcpy(*dev, host, bytes, cudaMemcpyHostToDevice);
    free(host);
    if (ce != cudaSuccess)
        return fail(FFGPU_EXTERNAL, "CUDA: table upload failed: %s", cudaGetErrorString(ce));
    return 0;
}

/* everything enc_device_create() made, complete or not */
static void enc_device_release(ffgpu_encoder *e)
{
    cudaSetDevice(e->opt.device);
    cudaDeviceSynchronize();
    for (int i = 0; i <= FFK_ENC_KERNELS; i++)
        if (e->events[i]) {
            cudaEventDestroy((cudaEvent_t)e->events[i]);
            e->events[i] = NULL;
        }
    for (int i = 0; i < MAX_DEPTH; i++)
        enc_free_job(&e->jobs[i]);
    if (e->up_stream) cudaStreamDestroy(e->up_stream);
    cudaFree(e->d_slices); cudaFree(e->d_qt); cudaFree(e->d_tab); cudaFree(e->d_prefix);
    cudaFree(e->d_prefix_bytes); cudaFree(e->d_state_shared); cudaFree(e->d_iota);
    cudaFree(e->d_rc_stat); cudaFree(e->d_rc_stat2); cudaFree(e->d_initial);
    e->up_stream = NULL;
    e->d_slices = NULL; e->d_qt = NULL; e->d_tab = NULL; e->d_prefix = NULL;
    e->d_prefix_bytes = NULL; e->d_state_shared = NULL; e->d_iota = NULL;
    e->d_rc_stat = e->d_rc_stat2 = NULL; e->d_initial = NULL;
    e->dev_ready = 0;
}

static int enc_device_create(ffgpu_encoder *e)
{
    const FFDevParams *P = &e->P;
    const int golomb = P->ac == FF_AC_GOLOMB;
    const size_t state_frame = (size_t)P->nslices * P->total_ctx * (golomb ? 8 : FF_CONTEXT_SIZE);
    int16_t *qt;
    int r;

    CK(cudaSetDevice(e->opt.device));
    CK(cudaMalloc(&e->d_slices, sizeof(FFDevSlice) * P->nslices));
    CK(cudaMemcpy(e->d_slices, e->h_slices, sizeof(FFDevSlice) * P->nslices, cudaMemcpyHostToDevice));
    qt = (int16_t *)malloc(sizeof(int16_t) * FF_MAX_QUANT_TABLES * FF_QT_STRIDE);
    if (qt)
        flatten_qt(&e->s, qt);
    if ((r = upload_and_free((void **)&e->d_qt, qt, sizeof(int16_t) * FF_MAX_QUANT_TABLES * FF_QT_STRIDE)) < 0)
        return r;
    CK(cudaMalloc(&e->d_tab, sizeof(FFRacTables)));
    CK(cudaMemcpy(e->d_tab, &e->s.cur_tab, sizeof(FFRacTables), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&e->d_prefix, sizeof(FFRacPrefix) * NPREFIX_SETS * P->nslices));
    CK(cudaMalloc(&e->d_prefix_bytes, (size_t)NPREFIX_SETS * P->nslices * e->prefix_stride));
    if (!e->intra) {
        CK(cudaMalloc(&e->d_state_shared, state_frame));
    }
    if (e->opt.pass1 && !golomb) {
        const size_t n2 = (size_t)e->s.ctx_count[e->s.context_model] * 32 * 2 * sizeof(unsigned long long);
        CK(cudaMalloc(&e->d_rc_stat, 256 * 2 * sizeof(unsigned long long)));
        CK(cudaMemset(e->d_rc_stat, 0, 256 * 2 * sizeof(unsigned long long)));
        CK(cudaMalloc(&e->d_rc_stat2, n2));
        CK(cudaMemset(e->d_rc_stat2, 0, n2));
    }
    if (e->s.initial[e->s.context_model] && !golomb) {
        /* one row per context of a slice: the table's initial states, repeated per plane set */
        const size_t per = (size_t)e->s.ctx_count[e->s.context_model] * FF_CONTEXT_SIZE;
        CK(cudaMalloc(&e->d_initial, per * P->nsets));
        for (int k = 0; k < P->nsets; k++)
            CK(cudaMemcpy(e->d_initial + per * k, e->s.initial[e->s.context_model], per, cudaMemcpyHostToDevice));
    }
    {
        const size_t n = (size_t)e->max_batch * P->nslices;
        uint32_t *iota = (uint32_t *)malloc(n * sizeof(uint32_t));
        for (size_t i = 0; iota && i < n; i++)
            iota[i] = (uint32_t)i;
        if ((r = upload_and_free((void **)&e->d_iota, iota, n * sizeof(uint32_t))) < 0)
            return r;
    }
    for (int i = 0; i < e->depth; i++) {
        EncJob *j = &e->jobs[i];
        const size_t B = (size_t)e->max_batch;
        CK(cudaStreamCreateWithFlags(&j->stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&j->done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&j->uploaded, cudaEventDisableTiming));
        if (!e->up_stream)
            CK(cudaStreamCreateWithFlags(&e->up_stream, cudaStreamNonBlocking));
        CK(cudaMalloc(&j->d_frames, B * P->frame_bytes));
        CK(cudaMalloc(&j->d_tokens, B * P->frame_tokens * sizeof(uint32_t)));
        if (e->intra)
            CK(cudaMalloc(&j->d_state, B * state_frame));
        CK(cudaMalloc(&j->d_bs, B * P->frame_bs));
        CK(cudaMalloc(&j->d_slice_bytes, B * P->nslices * sizeof(uint32_t)));
        CK(cudaMalloc(&j->d_slice_off, B * P->nslices * sizeof(uint32_t)));
        CK(cudaMalloc(&j->d_pkt_size, B * sizeof(uint32_t)));
        CK(cudaMalloc(&j->d_pkt_off, (B + 1) * sizeof(uint32_t)));
        CK(cudaMalloc(&j->d_overflow, sizeof(uint32_t)));
        /* the packed output reuses the token array: tokens are dead once stage B has run,
         * and a frame's tokens (4 bytes per sample) are larger than its bitstream arena */
        if (P->frame_tokens * sizeof(uint32_t) >= P->pkt_stride) {
            j->d_pkt = (uint8_t *)j->d_tokens;
        } else {                                    /* tiny slices / 17-bit samples */
            CK(cudaMalloc(&j->d_pkt, B * P->pkt_stride));
            j->pkt_owned = 1;
        }
        CK(cudaMalloc(&j->d_weight, B * P->nslices * sizeof(uint32_t)));
        CK(cudaMalloc(&j->d_weight_sorted, B * P->nslices * sizeof(uint32_t)));
        CK(cudaMalloc(&j->d_order, B * P->nslices * sizeof(uint32_t)));
        j->sort_tmp_bytes = ffk_sort_tmp_bytes((int)(B * P->nslices));
        CK(cudaMalloc(&j->d_sort_tmp, j->sort_tmp_bytes));
        CK(cudaMalloc(&j->d_sched, sizeof(FFSched)));
        CK(cudaMemset(j->d_sched, 0, sizeof(FFSched)));
        const int stride_max = coder_lane_stride((long)B * P->nslices);
        const int lone_off = getenv("FFGPU_LONE") && !atoi(getenv("FFGPU_LONE"));
        if (!golomb && !e->opt.pass1 && stride_max >= 8 && (stride_max < 32 || lone_off) &&
            !(getenv("FFGPU_SPLIT") && !atoi(getenv("FFGPU_SPLIT")))) {
            /* a group never has enough slices to fill the GPU's lanes, but more than one per
             * warp (those take the straight-line coder, k_code_range<true>): stage B in two halves.
             * Room for two decisions per sample on average (4 bytes per sample, like the
             * tokens); pictures that need more take the one-kernel coder, decided per group on
             * the device.  FFGPU_SPLIT=0 switches the split form off. */
            j->rec_cap = (unsigned long long)B * P->frame_tokens * 2;
            CK(cudaMalloc(&j->d_rec, j->rec_cap * sizeof(uint16_t)));
            CK(cudaMalloc(&j->d_rec_off, (B * P->nslices + 1) * sizeof(unsigned long long)));
            CK(cudaMalloc(&j->d_guard_rec, B * P->nslices * sizeof(uint32_t)));
            CK(cudaMalloc(&j->d_split_ok, sizeof(int)));
            CK(cudaMemset(j->d_split_ok, 0, sizeof(int)));
        }
        if (P->version > 3) {
            CK(cudaMalloc(&j->d_rct, B * P->nslices * 2 * sizeof(int)));
            CK(cudaMalloc(&j->d_rct_stat, B * P->nslices * 16 * sizeof(int32_t)));
        }
        CK(cudaMalloc(&j->d_frame_set, B));
        CK(cudaMalloc(&j->d_frame_key, B));
        CK(cudaHostAlloc(&j->h_frame_set, B, cudaHostAllocDefault));
        CK(cudaHostAlloc(&j->h_frame_key, B, cudaHostAllocDefault));
        CK(cudaHostAlloc(&j->h_pkt_size, B * sizeof(uint32_t), cudaHostAllocDefault));
        CK(cudaHostAlloc(&j->h_pkt_off, (B + 1) * sizeof(uint32_t), cudaHostAllocDefault));
        CK(cudaHostAlloc(&j->h_overflow, sizeof(uint32_t), cudaHostAllocDefault));
        j->h_pkt_cap = align_up(B * (P->frame_bytes / 4 + 65536), 4096);
        CK(cudaHostAlloc(&j->h_pkt, j->h_pkt_cap, cudaHostAllocDefault));
        j->pts = (int64_t *)calloc(B, sizeof(int64_t));
        j->key = (int *)calloc(B, sizeof(int));
        if (!j->pts || !j->key)
            return fail(FFGPU_ENOMEM, "out of memory");
    }
    return 0;
}

/* the device side of a handle appears with the first picture.  A failure part-way (out of
 * device or page-locked memory) leaves nothing behind: the next call starts over. */
static int enc_device_init(ffgpu_encoder *e)
{
    int ndev = 0, r;
    if (e->dev_ready)
        return 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(FFGPU_EXTERNAL, "no CUDA device: the FFV1 pixel path has no CPU fallback");
    if ((r = enc_device_create(e)) < 0) {
        enc_device_release(e);
        return r;
    }
    e->dev_ready = 1;
    return 0;
}

extern "C" int ffgpu_ffv1_encode_init(ffgpu_encoder **penc, const ffgpu_enc_options *opt)
{
    ffgpu_encoder *e;
    int r;
    if (!penc || !opt)
        return fail(FFGPU_EINVAL, "null argument");
    *penc = NULL;
    g_err[0] = 0;
    e = (ffgpu_encoder *)calloc(1, sizeof(*e));
    if (!e)
        return fail(FFGPU_ENOMEM, "out of memory");
    e->opt = *opt;
    snprintf(e->pix_fmt, sizeof(e->pix_fmt), "%s", opt->pix_fmt ? opt->pix_fmt : "");
    e->opt.pix_fmt = e->pix_fmt;
    if (opt->stats_in) {
        e->stats_in = strdup(opt->stats_in);
        e->opt.stats_in = e->stats_in;
    }
    if ((r = ff_stream_from_options(&e->s, &e->opt)) < 0) {
        ffgpu_ffv1_encode_close(e);             /* frees what exists so far (stats_in, state tables) */
        return fail(r, "encode_init: options rejected (%d)", r);
    }
    if ((r = ff_write_extradata(&e->s, opt->gop_size, &e->extradata, &e->extradata_size)) < 0) {
        ffgpu_ffv1_encode_close(e);
        return fail(r, "encode_init: extradata");
    }
    e->h_slices = (FFDevSlice *)calloc((size_t)e->s.nh * e->s.nv, sizeof(FFDevSlice));
    if (!e->h_slices) {
        ffgpu_ffv1_encode_close(e);
        return fail(FFGPU_ENOMEM, "out of memory");
    }
    ff_fill_dev_params(&e->s, 1, &e->P, e->h_slices);
    e->intra = opt->gop_size <= 1;
    e->prefix_stride = e->s.version > 2 ? 64 : 4096;
    if (e->intra) {
        /* enough pictures per group for ~48k resident slice coders, bounded by memory */
        const size_t per_frame = e->P.frame_bytes + e->P.frame_tokens * 4 + e->P.frame_bs +
                                 (size_t)e->P.nslices * e->P.total_ctx * FF_CONTEXT_SIZE;
        /* enough pictures per group to keep ~64k slice coders (2k warps) resident */
        int b = opt->max_batch > 0 ? opt->max_batch : (65536 + e->P.nslices - 1) / e->P.nslices;
        size_t cap = ((size_t)16 << 30) / (per_frame ? per_frame : 1);
        if (b > 1024) b = 1024;                   /* one block scans a group's packet sizes */
        if (opt->max_batch <= 0 && (size_t)b > cap) b = (int)cap;
        if (b < 1) b = 1;
        e->max_batch = b;
        e->depth = opt->pipeline_depth > 0 ? opt->pipeline_depth : 3;
    } else {
        e->max_batch = 1;
        e->depth = 1;
    }
    if (e->depth > MAX_DEPTH)
        e->depth = MAX_DEPTH;
    if (opt->ndevices > 1) {
        ffgpu_enc_options so = e->opt;
        if (opt->ndevices > FFGPU_MAX_DEVICES) {
            ffgpu_ffv1_encode_close(e);
            return fail(FFGPU_EINVAL, "ndevices %d exceeds %d", opt->ndevices, FFGPU_MAX_DEVICES);
        }
        so.ndevices = 0;
        e->chunk = e->intra ? 1 : opt->gop_size;
        for (int i = 0; i < opt->ndevices; i++) {
            so.device = opt->devices[i];
            if ((r = ffgpu_ffv1_encode_init(&e->sub[i], &so)) < 0) {
                ffgpu_ffv1_encode_close(e);
                return r;
            }
            e->nsub = i + 1;
        }
    }
    *penc = e;
    return 0;
}

extern "C" int ffgpu_ffv1_encoder_extradata(const ffgpu_encoder *e, const uint8_t **data)
{
    if (data)
        *data = e->extradata;
    return e->extradata_size;
}

extern "C" void ffgpu_ffv1_encoder_info(const ffgpu_encoder *e, int info[8])
{
    info[0] = e->s.version; info[1] = e->s.micro_version; info[2] = e->s.ac; info[3] = e->s.nh;
    info[4] = e->s.nv; info[5] = e->s.ec; info[6] = e->s.bits; info[7] = e->s.colorspace;
}

extern "C" size_t ffgpu_ffv1_encoder_max_packet(const ffgpu_encoder *e)
{
    return e->P.pkt_stride;
}

extern "C" uint64_t ffgpu_ffv1_encoder_launches(const ffgpu_encoder *e)
{
    uint64_t n = e->launches;
    for (int i = 0; i < e->nsub; i++)
        n += e->sub[i]->launches;
    return n;
}

struct EncJob;
static int enc_launch(ffgpu_encoder *e, EncJob *j);

/* find or build the prefix set for (key, picture structure, SAR) */
static int enc_prefix_set(ffgpu_encoder *e, int key, int ps, int sn, int sd)
{
    const FFDevParams *P = &e->P;
    int slot = -1;
    for (int i = 0; i < NPREFIX_SETS; i++) {
        PrefixSet *p = &e->sets[i];
        if (p->valid && p->key == key && p->ps == ps && p->sar_num == sn && p->sar_den == sd)
            return i;
        if (!p->valid && slot < 0)
            slot = i;
    }
    if (slot < 0) {
        /* cache full.  Pictures already queued in the group being filled hold indices into
         * the current sets: launch that group first, then wait for everything in flight and
         * start over */
        EncJob *f = &e->jobs[e->fill];
        if (f->state == JOB_FILLING && f->n > 0) {
            int r = enc_launch(e, f);
            if (r < 0)
                return r;
            e->fill = (e->fill + 1) % e->depth;
        }
        CK(cudaDeviceSynchronize());
        memset(e->sets, 0, sizeof(e->sets));
        slot = 0;
    }
    {
        FFRacPrefix *pre = (FFRacPrefix *)calloc(P->nslices, sizeof(FFRacPrefix));
        uint8_t *bytes = (uint8_t *)calloc((size_t)P->nslices, e->prefix_stride);
        int r = 0;
        if (!pre || !bytes) {
            free(pre); free(bytes);
            return fail(FFGPU_ENOMEM, "out of memory");
        }
        for (int i = 0; i < P->nslices && r >= 0; i++) {
            FFSliceRect rc = { e->h_slices[i].x, e->h_slices[i].y, e->h_slices[i].w, e->h_slices[i].h };
            r = ff_enc_slice_prefix(&e->s, i, &rc, key, ps, sn, sd, &pre[i],
                                    bytes + (size_t)i * e->prefix_stride, e->prefix_stride);
            pre[i].byte_off = (uint32_t)(((size_t)slot * P->nslices + i) * e->prefix_stride);
        }
        if (r >= 0) {
            cudaError_t ce = cudaMemcpy(e->d_prefix + (size_t)slot * P->nslices, pre,
                                        sizeof(FFRacPrefix) * P->nslices, cudaMemcpyHostToDevice);
            if (ce == cudaSuccess)
                ce = cudaMemcpy(e->d_prefix_bytes + (size_t)slot * P->nslices * e->prefix_stride, bytes,
                                (size_t)P->nslices * e->prefix_stride, cudaMemcpyHostToDevice);
            if (ce != cudaSuccess)
                r = fail(FFGPU_EXTERNAL, "CUDA: prefix upload: %s", cudaGetErrorString(ce));
        } else {
            fail(r, "slice header does not fit its prefix buffer");
        }
        free(pre);
        free(bytes);
        if (r < 0)
            return r;
    }
    e->sets[slot].valid = 1;
    e->sets[slot].key = key;
    e->sets[slot].ps = ps;
    e->sets[slot].sar_num = sn;
    e->sets[slot].sar_den = sd;
    return slot;
}

static void enc_fill_dev(const ffgpu_encoder *e, const EncJob *j, const uint8_t *frames, FFEncDev *E)
{
    memset(E, 0, sizeof(*E));
    E->slices = e->d_slices;
    E->qt = e->d_qt;
    E->tab = e->d_tab;
    E->frames = frames;
    E->tokens = j->d_tokens;
    E->state = e->intra ? j->d_state : e->d_state_shared;
    E->prefix = e->d_prefix;
    E->prefix_bytes = e->d_prefix_bytes;
    E->frame_prefix_set = j->d_frame_set;
    E->frame_key = j->d_frame_key;
    E->bs = j->d_bs;
    E->slice_bytes = j->d_slice_bytes;
    E->slice_off = j->d_slice_off;
    E->pkt_size = j->d_pkt_size;
    E->pkt_off = j->d_pkt_off;
    E->overflow = j->d_overflow;
    E->pkt = j->d_pkt;
    E->state_per_frame = e->intra;
    {
        const int q = e->P.set_qidx[0];
        E->five = e->s.qt[q][3][127] || e->s.qt[q][4][127];
    }
    E->weight = j->d_weight;
    E->weight_sorted = j->d_weight_sorted;
    E->iota = e->d_iota;
    E->order = j->d_order;
    E->sort_tmp = j->d_sort_tmp;
    E->sort_tmp_bytes = j->sort_tmp_bytes;
    E->lane_stride = coder_lane_stride((long)j->n * e->P.nslices);
    E->rct = j->d_rct;
    E->rct_stat = j->d_rct_stat;
    E->rec = j->d_rec;
    E->rec_cap = j->rec_cap;
    E->rec_off = j->d_rec_off;
    E->guard_rec = j->d_guard_rec;
    E->split_ok = j->d_split_ok;
    E->rc_stat = e->d_rc_stat;
    E->rc_stat2 = e->d_rc_stat2;
    E->stat_ctx_count = e->s.ctx_count[e->s.context_model];
    E->initial = e->d_initial;
    {
        const char *env = getenv("FFGPU_STAGE_A");
        E->legacy_stage_a = env && !strcmp(env, "legacy") ? 1 : env && !strcmp(env, "bulk") ? -1 : 0;
    }
    E->heavy_stride = heavy_stride_opt();
    E->heavy_factor = heavy_factor_opt();
    E->sched = E->heavy_stride ? j->d_sched : NULL;
}

/* enqueue the kernel chain + result download of a filled group */
static int enc_launch(ffgpu_encoder *e, EncJob *j)
{
    FFEncDev E;
    int r;
    enc_fill_dev(e, j, j->d_frames, &E);
    CK(cudaEventRecord(j->uploaded, e->up_stream));
    CK(cudaStreamWaitEvent(j->stream, j->uploaded, 0));
    CK(cudaMemcpyAsync(j->d_frame_set, j->h_frame_set, j->n, cudaMemcpyHostToDevice, j->stream));
    CK(cudaMemcpyAsync(j->d_frame_key, j->h_frame_key, j->n, cudaMemcpyHostToDevice, j->stream));
    CK(cudaMemsetAsync(j->d_overflow, 0, sizeof(uint32_t), j->stream));
    trace_mark(j->stream, "enc k0", (int)(j - e->jobs), j->n);
    r = ffk_encode_group(&e->P, &E, j->n, j->stream);
    if (r < 0)
        return fail(r, "kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    e->launches += r;
    trace_mark(j->stream, "enc k1", (int)(j - e->jobs), j->n);
    {
        /* packets and their size tables go to pinned host memory through the SMs: as
         * copy-engine transfers they would queue behind the decoder's picture downloads */
        FFCopyArgs c;
        memset(&c, 0, sizeof(c));
        c.seg[0].dst = j->h_pkt;       c.seg[0].src = j->d_pkt;
        c.dyn_bytes = j->d_pkt_off + j->n;       /* total of the group, in 16-byte units */
        c.dyn_shift = 4;
        c.dyn_cap = j->h_pkt_cap;
        c.seg[1].dst = j->h_pkt_size;  c.seg[1].src = j->d_pkt_size;  c.seg[1].bytes = sizeof(uint32_t) * j->n;
        c.seg[2].dst = j->h_pkt_off;   c.seg[2].src = j->d_pkt_off;   c.seg[2].bytes = sizeof(uint32_t) * (j->n + 1);
        c.seg[3].dst = j->h_overflow;  c.seg[3].src = j->d_overflow;  c.seg[3].bytes = sizeof(uint32_t);
        c.nseg = 4;
        if ((r = ffk_copy_segments(&c, j->stream)) < 0)
            return fail(r, "kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        e->launches += r;
    }
    CK(cudaEventRecord(j->done, j->stream));
    j->state = JOB_RUNNING;
    j->fetched = 0;
    j->drained = 0;
    return 0;
}

/* wait for a running group and bring its packets to pinned host memory */
static int enc_fetch(ffgpu_encoder *e, EncJob *j)
{
    size_t total;
    (void)e;
    if (j->fetched)
        return 0;
    CK(cudaEventSynchronize(j->done));
    if (*j->h_overflow)
        return fail(FFGPU_INVALIDDATA, "encoded frame too large");   /* ffv1enc_template.c:34-44 */
    total = (size_t)j->h_pkt_off[j->n] << 4;
    if (total > j->h_pkt_cap) {
        /* the pinned packet buffer was too small for the SM copy: grow it, copy the plain way */
        cudaFreeHost(j->h_pkt);
        j->h_pkt = NULL;
        j->h_pkt_cap = align_up(total + total / 2, 4096);
        CK(cudaHostAlloc(&j->h_pkt, j->h_pkt_cap, cudaHostAllocDefault));
        CK(cudaMemcpyAsync(j->h_pkt, j->d_pkt, total, cudaMemcpyDeviceToHost, j->stream));
        CK(cudaStreamSynchronize(j->stream));
    }
    trace_mark(j->stream, "enc pkts", (int)(j - e->jobs), j->n);
    j->fetched = 1;
    j->state = JOB_DRAINING;
    return 0;
}

/* launch the group being filled although it is not full (the caller cannot send more) */
static int enc_kick(ffgpu_encoder *e)
{
    if (e->dev_ready) {
        EncJob *j = &e->jobs[e->fill];
        if (j->state == JOB_FILLING && j->n > 0) {
            int r;
            cudaSetDevice(e->opt.device);
            if ((r = enc_launch(e, j)) < 0)
                return r;
            e->fill = (e->fill + 1) % e->depth;
        }
    }
    return 0;
}

static int enc_receive(ffgpu_encoder *e, uint8_t *pkt, size_t cap, size_t *size, int *key_frame,
                       int64_t *pts, int block);

/* ---- routing over several GPUs ---- */
static ffgpu_encoder *menc_owner(ffgpu_encoder *p, uint64_t k)
{
    return p->sub[(k / (uint64_t)p->chunk) % (uint64_t)p->nsub];
}

static int menc_send(ffgpu_encoder *p, const ffgpu_picture *pic)
{
    int r;
    if (!pic) {
        for (int i = 0; i < p->nsub; i++)
            if ((r = ffgpu_ffv1_encode_send_frame(p->sub[i], NULL)) < 0)
                return r;
        p->flushing = 1;
        return 0;
    }
    if (p->flushing)
        return fail(FFGPU_EOF, "send_frame after flush");
    r = ffgpu_ffv1_encode_send_frame(menc_owner(p, p->in_count), pic);
    if (r == 0)
        p->in_count++;
    else if (r == FFGPU_EAGAIN)
        p->send_blocked = 1;
    return r;
}

static int menc_receive(ffgpu_encoder *p, uint8_t *pkt, size_t cap, size_t *size, int *key_frame, int64_t *pts)
{
    ffgpu_encoder *s;
    int r;
    if (p->out_count == p->in_count) {
        if (!p->flushing)
            return FFGPU_EAGAIN;
        for (int i = 0; i < p->nsub; i++)          /* every sub-encoder ends its own flush */
            enc_receive(p->sub[i], pkt, cap, size, key_frame, pts, 0);
        p->flushing = 0;
        return FFGPU_EOF;
    }
    s = menc_owner(p, p->out_count);
    r = enc_receive(s, pkt, cap, size, key_frame, pts, 0);
    if (r == FFGPU_EAGAIN && (p->send_blocked || p->flushing)) {
        /* the caller cannot send (another GPU's queue is full) and the packet that is due
         * sits in a group that waits for more pictures: launch it and wait */
        if ((r = enc_kick(s)) < 0)
            return r;
        r = enc_receive(s, pkt, cap, size, key_frame, pts, 1);
    }
    if (r == 0 && pkt) {
        p->out_count++;
        p->send_blocked = 0;
    }
    return r;
}

extern "C" int ffgpu_ffv1_encode_send_frame(ffgpu_encoder *e, const ffgpu_picture *pic)
{
    if (e && e->nsub)
        return menc_send(e, pic);
    if (e)
        cudaSetDevice(e->opt.device);        /* the current device is per host thread */
    EncJob *j;
    int r, key, set, ps;
    if (!e)
        return fail(FFGPU_EINVAL, "null encoder");
    if ((r = enc_device_init(e)) < 0)
        return r;
    j = &e->jobs[e->fill];
    if (!pic) {
        e->flushing = 1;
        if (j->state == JOB_FILLING && j->n > 0) {
            if ((r = enc_launch(e, j)) < 0)
                return r;
            e->fill = (e->fill + 1) % e->depth;
        }
        return 0;
    }
    if (e->flushing)
        return fail(FFGPU_EOF, "send_frame after flush");
    if (j->state == JOB_RUNNING || j->state == JOB_DRAINING)
        return FFGPU_EAGAIN;                       /* every group is busy: receive first */
    key = e->opt.gop_size == 0 || e->picture_number % e->opt.gop_size == 0;
    ps = !pic->interlaced_frame ? 3 : 1 + !pic->top_field_first;   /* ffv1enc.c:944-947 */
    set = enc_prefix_set(e, key, ps, pic->sar_num, pic->sar_den);
    if (set < 0)
        return set;
    j = &e->jobs[e->fill];                         /* a full prefix cache launches the filling group */
    if (j->state == JOB_RUNNING || j->state == JOB_DRAINING)
        return FFGPU_EAGAIN;
    /* a free group becomes the filling one only when its first picture is in: a failure on
     * the way (staging allocation, a bad plane) leaves the handle as it was */
    if (j->state == JOB_FREE)
        j->n = 0;
    if (j->n == 0)
        trace_mark(e->up_stream, "enc h2d", (int)(j - e->jobs), 0);
    const int staged = needs_staging(pic->data[0], pic->linesize);
    if (staged < 0)
        return fail(FFGPU_EINVAL, "bottom-up pictures (negative linesize) in device memory are not supported");
    if (staged) {
        /* an ordinary AVFrame: host memcpy into pinned staging, one linear DMA from there */
        uint8_t *data[4], *st;
        /* the first pageable picture creates the staging of EVERY launch group, so that no
         * later group stops the pipeline for a page-locked allocation */
        for (int g = 0; g < e->depth && !j->h_stage; g++)
            if (!e->jobs[g].h_stage)
                CK(cudaHostAlloc(&e->jobs[g].h_stage, (size_t)e->max_batch * e->P.frame_bytes, cudaHostAllocDefault));
        st = j->h_stage + (size_t)j->n * e->P.frame_bytes;
        for (int k = 0; k < 4; k++)
            data[k] = (uint8_t *)pic->data[k];
        if ((r = stage_picture(&e->P, e->s.pf, e->s.width, e->s.height, data, pic->linesize, st, 1)) < 0)
            return r;
        CK(cudaMemcpyAsync(j->d_frames + (size_t)j->n * e->P.frame_bytes, st, e->P.frame_bytes,
                           cudaMemcpyHostToDevice, e->up_stream));
    } else if ((r = upload_picture(&e->P, e->s.pf, e->s.width, e->s.height, pic,
                                   j->d_frames + (size_t)j->n * e->P.frame_bytes, e->up_stream)) < 0)
        return r;
    j->state = JOB_FILLING;
    j->h_frame_set[j->n] = (uint8_t)set;
    j->h_frame_key[j->n] = (uint8_t)key;
    e->gob_count += key;
    j->pts[j->n] = pic->pts;
    j->key[j->n] = key;
    j->n++;
    e->picture_number++;
    if (j->n == e->max_batch) {
        if ((r = enc_launch(e, j)) < 0)
            return r;
        e->fill = (e->fill + 1) % e->depth;
    }
    return 0;
}

extern "C" int ffgpu_ffv1_encode_receive_packet(ffgpu_encoder *e, uint8_t *pkt, size_t cap,
                                                size_t *size, int *key_frame, int64_t *pts)
{
    if (e && e->nsub)
        return menc_receive(e, pkt, cap, size, key_frame, pts);
    return enc_receive(e, pkt, cap, size, key_frame, pts, 0);
}

extern "C" int ffgpu_ffv1_encode_packet_ready(ffgpu_encoder *e, size_t *size)
{
    if (e && e->nsub)
        return menc_receive(e, NULL, 0, size, NULL, NULL);
    return enc_receive(e, NULL, 0, size, NULL, NULL, 0);
}

/* block != 0: wait for the oldest running group instead of returning FFGPU_EAGAIN */
static int enc_receive(ffgpu_encoder *e, uint8_t *pkt, size_t cap, size_t *size, int *key_frame,
                       int64_t *pts, int block)
{
    if (e)
        cudaSetDevice(e->opt.device);        /* the current device is per host thread */
    EncJob *j;
    int r, i;
    if (!e)
        return fail(FFGPU_EINVAL, "null encoder");
    if (!e->dev_ready) {
        if (e->flushing) {
            e->flushing = 0;
            return FFGPU_EOF;
        }
        return FFGPU_EAGAIN;
    }
    j = &e->jobs[e->head];
    if (j->state == JOB_FREE || j->state == JOB_FILLING) {
        if (e->flushing) {
            e->flushing = 0;                       /* drained: the handle accepts pictures again */
            return FFGPU_EOF;
        }
        return FFGPU_EAGAIN;
    }
    if (j->state == JOB_RUNNING) {
        /* block only if the caller cannot make progress otherwise */
        const EncJob *f = &e->jobs[e->fill];
        const int must_wait = block || e->flushing || f->state == JOB_RUNNING || f->state == JOB_DRAINING;
        if (!must_wait && cudaEventQuery(j->done) == cudaErrorNotReady)
            return FFGPU_EAGAIN;
        if ((r = enc_fetch(e, j)) < 0) {
            j->state = JOB_FREE;
            e->head = (e->head + 1) % e->depth;
            return r;
        }
    }
    i = j->drained;
    if (!pkt) {                                    /* peek: size of the packet that is due */
        if (size) *size = j->h_pkt_size[i];
        return 0;
    }
    if (j->h_pkt_size[i] > cap)
        return fail(FFGPU_ENOSPC, "packet buffer too small: need %u bytes", j->h_pkt_size[i]);
    memcpy(pkt, j->h_pkt + ((size_t)j->h_pkt_off[i] << 4), j->h_pkt_size[i]);
    if (size) *size = j->h_pkt_size[i];
    if (key_frame) *key_frame = j->key[i];
    if (pts) *pts = j->pts[i];
    if (++j->drained == j->n) {
        j->state = JOB_FREE;
        j->n = 0;
        e->head = (e->head + 1) % e->depth;
    }
    return 0;
}

extern "C" int ffgpu_ffv1_encode_frame(ffgpu_encoder *e, const ffgpu_picture *pic, uint8_t *pkt,
                                       size_t cap, size_t *size, int *key_frame)
{
    EncJob *j;
    int r;
    if (!e || !pic)
        return fail(FFGPU_EINVAL, "null argument");
    if (e->nsub) {
        if (e->in_count != e->out_count)
            return fail(FFGPU_EINVAL, "encode_frame while send/receive pictures are pending");
        r = ffgpu_ffv1_encode_frame(menc_owner(e, e->in_count), pic, pkt, cap, size, key_frame);
        if (r >= 0) {
            e->in_count++;
            e->out_count++;
        }
        return r;
    }
    if ((r = enc_device_init(e)) < 0)
        return r;
    for (int i = 0; i < e->depth; i++)
        if (e->jobs[i].state != JOB_FREE)
            return fail(FFGPU_EINVAL, "encode_frame while send/receive pictures are pending");
    if ((r = ffgpu_ffv1_encode_send_frame(e, pic)) < 0)
        return r;
    j = &e->jobs[e->fill];
    if (j->state == JOB_FILLING) {
        if ((r = enc_launch(e, j)) < 0)
            return r;
        e->fill = (e->fill + 1) % e->depth;
    }
    /* the group is running: force completion */
    j = &e->jobs[e->head];
    if ((r = enc_fetch(e, j)) < 0) {
        j->state = JOB_FREE;
        j->n = 0;
        e->head = (e->head + 1) % e->depth;
        return r;
    }
    return ffgpu_ffv1_encode_receive_packet(e, pkt, cap, size, key_frame, NULL);
}

extern "C" int ffgpu_ffv1_encode_device(ffgpu_encoder *e, const void *d_frames, int nframes,
                                        void *cuda_stream)
{
    if (e)
        cudaSetDevice(e->opt.device);        /* the current device is per host thread */
    EncJob *j;
    FFEncDev E;
    cudaStream_t st;
    int r, set;
    if (!e || !d_frames || nframes <= 0)
        return fail(FFGPU_EINVAL, "bad argument");
    if (e->nsub)
        return fail(FFGPU_EINVAL, "device-resident batches live on one GPU: open one handle per device");
    if ((r = enc_device_init(e)) < 0)
        return r;
    if (nframes > e->max_batch)
        return fail(FFGPU_EINVAL, "nframes %d exceeds max_batch %d", nframes, e->max_batch);
    if (nframes > 1 && !e->intra)
        return fail(FFGPU_EINVAL, "device batches need gop_size <= 1");
    j = &e->jobs[0];
    if (j->state != JOB_FREE)
        return fail(FFGPU_EINVAL, "encode_device while send/receive pictures are pending");
    st = cuda_stream ? (cudaStream_t)cuda_stream : j->stream;
    /* the previous device batch reads the pinned per-frame tables asynchronously */
    CK(cudaEventSynchronize(j->done));
    for (int i = 0; i < nframes; i++) {
        const int key = e->opt.gop_size == 0 || (e->picture_number + i) % e->opt.gop_size == 0;
        set = enc_prefix_set(e, key, 3, 0, 1);
        if (set < 0)
            return set;
        j->h_frame_set[i] = (uint8_t)set;
        j->h_frame_key[i] = (uint8_t)key;
        j->key[i] = key;
        e->gob_count += key;
    }
    e->picture_number += nframes;
    j->n = nframes;
    enc_fill_dev(e, j, (const uint8_t *)d_frames, &E);
    if (e->profile)
        E.events = e->events;
    CK(cudaMemcpyAsync(j->d_frame_set, j->h_frame_set, nframes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(j->d_frame_key, j->h_frame_key, nframes, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(j->d_overflow, 0, sizeof(uint32_t), st));
    r = ffk_encode_group(&e->P, &E, nframes, st);
    if (r < 0)
        return fail(r, "kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    e->launches += r;
    CK(cudaEventRecord(j->done, st));
    return 0;
}

static int enc_device_sizes(ffgpu_encoder *e, EncJob *j)
{
    (void)e;
    /* the batch may have run on the handle's own non-blocking stream, which the legacy
     * default stream of the copies below does not wait for */
    CK(cudaEventSynchronize(j->done));
    CK(cudaMemcpy(j->h_pkt_size, j->d_pkt_size, sizeof(uint32_t) * j->n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(j->h_pkt_off, j->d_pkt_off, sizeof(uint32_t) * (j->n + 1), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(j->h_overflow, j->d_overflow, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (*j->h_overflow)
        return fail(FFGPU_INVALIDDATA, "encoded frame too large");
    return 0;
}

extern "C" int ffgpu_ffv1_encode_device_result(ffgpu_encoder *e, int frame, const void **d_pkt,
                                               size_t *pkt_size)
{
    if (e)
        cudaSetDevice(e->opt.device);        /* the current device is per host thread */
    EncJob *j;
    int r;
    if (!e || !e->dev_ready)
        return fail(FFGPU_EINVAL, "no device batch");
    j = &e->jobs[0];
    if (frame < 0 || frame >= j->n)
        return fail(FFGPU_EINVAL, "frame index out of range");
    if ((r = enc_device_sizes(e, j)) < 0)
        return r;
    if (d_pkt) *d_pkt = j->d_pkt + ((size_t)j->h_pkt_off[frame] << 4);
    if (pkt_size) *pkt_size = j->h_pkt_size[frame];
    return 0;
}

extern "C" int ffgpu_ffv1_encode_device_fetch(ffgpu_encoder *e, int frame, uint8_t *pkt, size_t cap,
                                              size_t *pkt_size)
{
    if (e)
        cudaSetDevice(e->opt.device);        /* the current device is per host thread */
    const void *d;
    size_t n;
    int r = ffgpu_ffv1_encode_device_result(e, frame, &d, &n);
    if (r < 0)
        return r;
    if (n > cap)
        return fail(FFGPU_ENOSPC, "packet buffer too small: need %zu bytes", n);
    CK(cudaMemcpy(pkt, d, n, cudaMemcpyDeviceToHost));
    if (pkt_size) *pkt_size = n;
    return 0;
}

/* the counters of this handle (and of its sub-handles) added into host arrays */
static int enc_collect_stats(ffgpu_encoder *e, uint64_t *rc, uint64_t *rc2, size_t n2, uint64_t *gob)
{
    for (int i = 0; i < e->nsub; i++) {
        int r = enc_collect_stats(e->sub[i], rc, rc2, n2, gob);
        if (r < 0)
            return r;
    }
    *gob += e->gob_count;
    if (e->dev_ready && e->d_rc_stat) {
        uint64_t *tmp = (uint64_t *)malloc((512 + n2) * sizeof(uint64_t));
        if (!tmp)
            return fail(FFGPU_ENOMEM, "out of memory");
        cudaSetDevice(e->opt.device);
        if (cudaDeviceSynchronize() != cudaSuccess ||
            cudaMemcpy(tmp, e->d_rc_stat, 512 * sizeof(uint64_t), cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(tmp + 512, e->d_rc_stat2, n2 * sizeof(uint64_t), cudaMemcpyDeviceToHost) != cudaSuccess) {
            free(tmp);
            return fail(FFGPU_EXTERNAL, "CUDA: copy of the first-pass counters failed");
        }
        for (size_t k = 0; k < 512; k++)
            rc[k] += tmp[k];
        for (size_t k = 0; k < n2; k++)
            rc2[k] += tmp[512 + k];
        free(tmp);
    }
    return 0;
}

extern "C" int ffgpu_ffv1_encoder_stats_out(ffgpu_encoder *e, char *buf, size_t cap)
{
    /* the text of ffv1enc.c:1160-1175: rc_stat, newline, rc_stat2 of every quant table (only
     * the encoder's own table was ever counted), the number of key frames, newline */
    if (!e || !buf)
        return fail(FFGPU_EINVAL, "null argument");
    if (!e->opt.pass1)
        return fail(FFGPU_EINVAL, "not a first pass");
    const FFStream *s = &e->s;
    const size_t n2 = (size_t)s->ctx_count[s->context_model] * 64;
    uint64_t *rc = (uint64_t *)calloc(512 + n2, sizeof(uint64_t)), gob = 0;
    size_t pos = 0;
    int r;
    if (!rc)
        return fail(FFGPU_ENOMEM, "out of memory");
    if ((r = enc_collect_stats(e, rc, rc + 512, n2, &gob)) < 0) {
        free(rc);
        return r;
    }
#define OUT_(...) do { int w_ = snprintf(buf + pos, cap - pos, __VA_ARGS__); \
                       if (w_ < 0 || (size_t)w_ >= cap - pos) { free(rc); return fail(FFGPU_ENOSPC, "stats buffer too small"); } \
                       pos += (size_t)w_; } while (0)
    for (int j = 0; j < 256; j++)
        OUT_("%llu %llu ", (unsigned long long)rc[2 * j], (unsigned long long)rc[2 * j + 1]);
    /* (the reference prints a newline here without advancing its pointer, ffv1enc.c:1165: the
     * next number overwrites it, so there is none) */
    for (int i = 0; i < s->qt_count; i++)
        for (int j = 0; j < s->ctx_count[i]; j++)
            for (int m = 0; m < 32; m++) {
                const uint64_t *v = i == s->context_model ? rc + 512 + ((size_t)j * 32 + m) * 2 : NULL;
                OUT_("%llu %llu ", (unsigned long long)(v ? v[0] : 0), (unsigned long long)(v ? v[1] : 0));
            }
    OUT_("%d\n", (int)gob);
#undef OUT_
    free(rc);
    return (int)pos;
}

extern "C" int ffgpu_ffv1_encoder_decisions(ffgpu_encoder *e, uint64_t *total, uint32_t *heaviest)
{
    EncJob *j;
    uint32_t *w;
    size_t n;
    uint64_t sum = 0;
    uint32_t mx = 0;
    if (!e || !e->dev_ready)
        return fail(FFGPU_EINVAL, "no device batch");
    cudaSetDevice(e->opt.device);
    j = &e->jobs[0];
    n = (size_t)j->n * e->P.nslices;
    CK(cudaEventSynchronize(j->done));
    w = (uint32_t *)malloc(n * sizeof(uint32_t) + 4);
    if (!w)
        return fail(FFGPU_ENOMEM, "out of memory");
    if (cudaMemcpy(w, j->d_weight, n * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess) {
        free(w);
        return fail(FFGPU_EXTERNAL, "CUDA: copy of the decision counts failed");
    }
    for (size_t i = 0; i < n && e->P.ac != FF_AC_GOLOMB; i++) {
        sum += w[i];
        if (w[i] > mx)
            mx = w[i];
    }
    free(w);
    if (total) *total = sum;
    if (heaviest) *heaviest = mx;
    return 0;
}

/* profiling of the device-batch path: per-kernel CUDA-event times of the last group */
extern "C" int ffgpu_ffv1_encoder_profile(ffgpu_encoder *e, int enable)
{
    int r;
    if (!e)
        return fail(FFGPU_EINVAL, "null encoder");
    if ((r = enc_device_init(e)) < 0)
        return r;
    if (enable && !e->events[0])
        for (int i = 0; i <= FFK_ENC_KERNELS; i++) {
            cudaEvent_t ev;
            CK(cudaEventCreate(&ev));
            e->events[i] = ev;
        }
    e->profile = enable;
    return 0;
}

extern "C" int ffgpu_ffv1_encoder_kernel_ms(ffgpu_encoder *e, float *ms, int n)
{
    if (!e || !e->profile || !e->events[0])
        return fail(FFGPU_EINVAL, "profiling not enabled");
    for (int i = 0; i < n && i < FFK_ENC_KERNELS; i++)
        CK(cudaEventElapsedTime(&ms[i], (cudaEvent_t)e->events[i], (cudaEvent_t)e->events[i + 1]));
    return FFK_ENC_KERNELS;
}

extern "C" int ffgpu_ffv1_encode_close(ffgpu_encoder *e)
{
    if (!e)
        return 0;
    for (int i = 0; i < e->nsub; i++)
        ffgpu_ffv1_encode_close(e->sub[i]);
    if (e->dev_ready) {
        cudaSetDevice(e->opt.device);
        cudaDeviceSynchronize();
        trace_dump();
        enc_device_release(e);
    }
    free(e->stats_in);
    free(e->h_slices);
    free(e->extradata);
    ff_stream_free(&e->s);
    free(e);
    return 0;
}

/* ====================================================================== */
/* decoder                                                                 */
/* ====================================================================== */
struct DecFrameMeta {
    FFDecFrameInfo info;
    int64_t pts;
    int nslices;
    int has_dst;
    int staged;                 /* dst is pageable memory: the picture comes through h_stage */
    ffgpu_picture_out dst;
    uint8_t damaged[FF_MAX_SLICES];
    FFSliceRect rect[FF_MAX_SLICES];
};

struct DecJob {
    cudaStream_t stream;
    cudaEvent_t done;
    cudaEvent_t decoded;        /* kernels finished: the download stream may start */
    int n, state, drained, fetched;
    int n_device;               /* pictures of the last ffgpu_ffv1_decode_device() batch */
    uint8_t *h_pkt, *d_pkt;
    size_t pkt_cap, pkt_used;
    FFDecSlice *h_work, *d_work;
    int *h_nslices, *d_nslices;
    uint8_t *d_state;
    int32_t *d_lines;
    int32_t *d_wide_lines;      /* picture-wide scratch pool for slices wider than their grid cell */
    uint32_t *d_wide_used;
    uint32_t *d_touched;        /* lazily created states: one bit per (work item, context) */
    FFSched *d_sched;
    uint8_t *d_frames;
    uint8_t *h_stage;           /* pinned staging for destinations in pageable memory (lazily) */
    FFDecResult *d_result, *h_result;
    uint32_t *d_weight, *d_weight_sorted, *d_order;
    void *d_sort_tmp;
    size_t sort_tmp_bytes;
    DecFrameMeta *meta;
};

struct ffgpu_decoder {
    ffgpu_dec_options opt;
    FFStream s;
    FFDecHostState hs;
    FFDevParams P;
    FFDevSlice *h_slices;
    int have_params;
    int intra;
    int max_batch, depth, max_slices, max_ctx, line_stride;
    int lazy_states;            /* state rows are created on first touch (see FFDecDev.touched) */
    int generic;
    int any_five;
    int dev_ready;
    int16_t *d_qt;
    FFRacTables *d_tab;
    uint8_t *d_initial;
    uint8_t *d_state_shared;
    uint8_t *d_prev;            /* last output picture, for concealment */
    cudaEvent_t prev_ready;
    uint32_t *d_iota;
    int have_prev;
    DecJob jobs[MAX_DEPTH];
    cudaStream_t down_stream;   /* all picture downloads, in group order (see up_stream) */
    int fill, head, flushing;
    uint64_t launches;
    int profile, profile_next;
    void *events[FFK_DEC_KERNELS + 1];
    /* more than one GPU (intra-only streams): packet k is decoded by sub[k % nsub] */
    int nsub, send_blocked;
    ffgpu_decoder *sub[FFGPU_MAX_DEVICES];
    uint64_t in_count, out_count;
};

static void dec_free_job(DecJob *j)
{
    cudaFreeHost(j->h_pkt); cudaFree(j->d_pkt); cudaFreeHost(j->h_work); cudaFree(j->d_work);
    cudaFreeHost(j->h_nslices); cudaFree(j->d_nslices); cudaFree(j->d_state); cudaFree(j->d_lines);
    cudaFree(j->d_wide_lines); cudaFree(j->d_wide_used); cudaFree(j->d_touched); cudaFree(j->d_sched);
    cudaFree(j->d_frames); cudaFree(j->d_result); cudaFreeHost(j->h_result); cudaFreeHost(j->h_stage);
    cudaFree(j->d_weight); cudaFree(j->d_weight_sorted); cudaFree(j->d_order); cudaFree(j->d_sort_tmp);
    free(j->meta);
    if (j->done) cudaEventDestroy(j->done);
    if (j->decoded) cudaEventDestroy(j->decoded);
    if (j->stream) cudaStreamDestroy(j->stream);
    memset(j, 0, sizeof(*j));
}

/* (re)derive the kernel parameters once the stream is known (after the extradata, or after
 * the first v0/v1 key frame) */
static int dec_setup_stream(ffgpu_decoder *d)
{
    int r;
    if (!d->s.pf && (r = ff_pick_decoder_format(&d->s)) < 0)
        return fail(r, "format not supported");
    free(d->h_slices);
    d->h_slices = (FFDevSlice *)calloc((size_t)d->s.nh * d->s.nv, sizeof(FFDevSlice));
    if (!d->h_slices)
        return fail(FFGPU_ENOMEM, "out of memory");
    ff_fill_dev_params(&d->s, 0, &d->P, d->h_slices);
    d->max_slices = d->s.nh * d->s.nv;
    d->max_ctx = d->P.total_ctx / d->P.nsets;
    d->line_stride = 8;
    for (int i = 0; i < d->max_slices; i++)
        if (d->h_slices[i].w + 8 > d->line_stride)
            d->line_stride = d->h_slices[i].w + 8;
    /* a slice header may describe any rectangle of the picture; size the scratch for the
     * worst case only when the stream is not trusted to match the grid */
    if (d->s.version > 2 && d->line_stride < d->s.width + 8 && d->max_slices <= 16)
        d->line_stride = d->s.width + 8;
    d->intra = d->s.version > 2 && d->s.intra;
    if (d->intra) {
        /* the arena is sized later from the tables the stream uses (dec_size_states): budget
         * with the smallest table here */
        int min_ctx = d->s.ctx_count[0];
        for (int i = 1; i < d->s.qt_count; i++)
            if (d->s.ctx_count[i] < min_ctx)
                min_ctx = d->s.ctx_count[i];
        const size_t per_frame = d->P.frame_bytes * 2 +
                                 (size_t)d->max_slices * d->P.nsets * min_ctx * FF_CONTEXT_SIZE;
        int b = d->opt.max_batch > 0 ? d->opt.max_batch : (65536 + d->max_slices - 1) / d->max_slices;
        size_t cap = ((size_t)16 << 30) / (per_frame ? per_frame : 1);
        if (b > 1024) b = 1024;
        if (d->opt.max_batch <= 0 && (size_t)b > cap) b = (int)cap;
        if (b < 1) b = 1;
        d->max_batch = b;
        d->depth = d->opt.pipeline_depth > 0 ? d->opt.pipeline_depth : 3;
    } else {
        d->max_batch = 1;
        d->depth = 1;
    }
    if (d->depth > MAX_DEPTH)
        d->depth = MAX_DEPTH;
    d->have_params = 1;
    return 0;
}

/* The adaptive-state arena is sized for the quant tables the stream really uses (the slice
 * headers name them), not for the largest table of the global header: -context 0 streams
 * carry the 7563-context table too but never touch it. */
static int dec_size_states(ffgpu_decoder *d, const uint8_t *pkt, size_t size)
{
    FFDecHostState hs = d->hs;
    FFDecFrameInfo info;
    FFDecSlice *tmp = (FFDecSlice *)calloc((size_t)d->max_slices, sizeof(FFDecSlice));
    int n, need = 1;
    if (!tmp)
        return fail(FFGPU_ENOMEM, "out of memory");
    n = ff_dec_parse_packet(&d->s, &hs, pkt, size, 0, tmp, &info);
    if (n < 0) {
        free(tmp);
        return fail(n, "invalid packet (%d)", n);
    }
    for (int i = 0; i < n; i++)
        for (int k = 0; k < d->P.nsets && !tmp[i].skip; k++)
            if (d->s.ctx_count[tmp[i].qidx[k]] > need)
                need = d->s.ctx_count[tmp[i].qidx[k]];
    free(tmp);
    /* from now on only slice 0 is parsed on the host -- unless the stream carries initial
     * state tables: the state reset picks them by the slice's quant_table_index
     * (ffv1.c:182-207), which must then be known before the kernels run */
    d->hs.device_parse = 1;
    for (int i = 0; i < d->s.qt_count; i++)
        if (d->s.initial[i])
            d->hs.device_parse = 0;
    d->max_ctx = need;
    d->P.total_ctx = d->P.nsets * need;
    for (int k = 0; k < d->P.nsets; k++)
        d->P.set_base[k] = k * need;
    return 0;
}

/* everything dec_device_create() made, complete or not */
static void dec_device_release(ffgpu_decoder *d)
{
    cudaSetDevice(d->opt.device);
    cudaDeviceSynchronize();
    for (int i = 0; i < MAX_DEPTH; i++)
        dec_free_job(&d->jobs[i]);
    if (d->down_stream) cudaStreamDestroy(d->down_stream);
    cudaFree(d->d_qt); cudaFree(d->d_tab); cudaFree(d->d_initial); cudaFree(d->d_state_shared);
    cudaFree(d->d_prev); cudaFree(d->d_iota);
    if (d->prev_ready) cudaEventDestroy(d->prev_ready);
    d->down_stream = NULL;
    d->d_qt = NULL; d->d_tab = NULL; d->d_initial = NULL; d->d_state_shared = NULL;
    d->d_prev = NULL; d->d_iota = NULL;
    d->prev_ready = NULL;
    d->dev_ready = 0;
}

static int dec_device_create(ffgpu_decoder *d)
{
    const FFDevParams *P = &d->P;
    const int golomb = P->ac == FF_AC_GOLOMB;
    const size_t state_frame = (size_t)d->max_slices * P->total_ctx * (golomb ? 8 : FF_CONTEXT_SIZE);
    int16_t *qt;
    int any_initial = 0, r;
    CK(cudaSetDevice(d->opt.device));
    qt = (int16_t *)malloc(sizeof(int16_t) * FF_MAX_QUANT_TABLES * FF_QT_STRIDE);
    if (qt)
        flatten_qt(&d->s, qt);
    if ((r = upload_and_free((void **)&d->d_qt, qt, sizeof(int16_t) * FF_MAX_QUANT_TABLES * FF_QT_STRIDE)) < 0)
        return r;
    CK(cudaMalloc(&d->d_tab, sizeof(FFRacTables)));
    CK(cudaMemcpy(d->d_tab, &d->s.cur_tab, sizeof(FFRacTables), cudaMemcpyHostToDevice));
    for (int i = 0; i < d->s.qt_count; i++)
        any_initial |= d->s.initial[i] != NULL;
    if (any_initial) {
        const size_t per = (size_t)d->max_ctx * FF_CONTEXT_SIZE;
        uint8_t *tmp = (uint8_t *)malloc(per * FF_MAX_QUANT_TABLES);
        if (tmp)
            memset(tmp, 128, per * FF_MAX_QUANT_TABLES);
        for (int i = 0; tmp && i < d->s.qt_count; i++)
            if (d->s.initial[i])
                memcpy(tmp + per * i, d->s.initial[i],
                       (size_t)(d->s.ctx_count[i] < d->max_ctx ? d->s.ctx_count[i] : d->max_ctx) * FF_CONTEXT_SIZE);
        if ((r = upload_and_free((void **)&d->d_initial, tmp, per * FF_MAX_QUANT_TABLES)) < 0)
            return r;
    }
    /* The carried states start zeroed, like the reference's av_mallocz'ed ones (ffv1.c:70-90):
     * a damaged stream can reach a state no key frame has reset (a slice that was skipped in
     * the key frame), and get_vlc_symbol's k search (ffv1dec.c:77-81) does not end on a
     * garbage state with count 0.  Intra-only streams reset the states of every picture
     * whatever its key-frame bit says (ff_dec_parse_packet). */
    if (!d->intra) {
        CK(cudaMalloc(&d->d_state_shared, state_frame));
        CK(cudaMemset(d->d_state_shared, 0, state_frame));
    }
    /* every frame a key frame, all states start at 128, and the stream runs through the
     * specialised planar decoder: no arena fill, rows are created on first touch */
    d->generic = getenv("FFGPU_DEC_GENERIC") != NULL;      /* development: A/B against the generic decoder */
    d->lazy_states = d->intra && !any_initial && !d->generic && ff_decode_planar_mode(P) != 0;
    /* 5-input contexts: only tables the arena was sized for can be named by a slice header
     * (-context 0 streams carry the large 5-input table too, but never use it) */
    d->any_five = 0;
    for (int i = 0; i < d->s.qt_count; i++)
        if (d->s.ctx_count[i] <= d->max_ctx)
            d->any_five |= d->s.qt[i][3][127] || d->s.qt[i][4][127];
    CK(cudaMalloc(&d->d_prev, P->frame_bytes));
    CK(cudaMemset(d->d_prev, 0, P->frame_bytes));
    CK(cudaEventCreateWithFlags(&d->prev_ready, cudaEventDisableTiming));
    CK(cudaEventRecord(d->prev_ready, 0));
    {
        const size_t n = (size_t)d->max_batch * d->max_slices;
        uint32_t *iota = (uint32_t *)malloc(n * sizeof(uint32_t));
        for (size_t i = 0; iota && i < n; i++)
            iota[i] = (uint32_t)i;
        if ((r = upload_and_free((void **)&d->d_iota, iota, n * sizeof(uint32_t))) < 0)
            return r;
    }
    for (int i = 0; i < d->depth; i++) {
        DecJob *j = &d->jobs[i];
        const size_t B = (size_t)d->max_batch;
        CK(cudaStreamCreateWithFlags(&j->stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&j->done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&j->decoded, cudaEventDisableTiming));
        if (!d->down_stream)
            CK(cudaStreamCreateWithFlags(&d->down_stream, cudaStreamNonBlocking));
        j->pkt_cap = align_up(B * (P->frame_bytes / 2 + 65536) + 256, 4096);
        if (j->pkt_cap > 0xFFFF0000u)              /* work items address the arena with 32 bits; */
            j->pkt_cap = 0xFFFF0000u;              /* a full arena launches the group early      */
        CK(cudaHostAlloc(&j->h_pkt, j->pkt_cap, cudaHostAllocDefault));
        CK(cudaMalloc(&j->d_pkt, j->pkt_cap));
        CK(cudaHostAlloc(&j->h_work, B * d->max_slices * sizeof(FFDecSlice), cudaHostAllocDefault));
        CK(cudaMalloc(&j->d_work, B * d->max_slices * sizeof(FFDecSlice)));
        CK(cudaHostAlloc(&j->h_nslices, B * sizeof(int), cudaHostAllocDefault));
        CK(cudaMalloc(&j->d_nslices, B * sizeof(int)));
        if (d->intra)
            CK(cudaMalloc(&j->d_state, B * state_frame));     /* reset for every picture, see ff_dec_parse_packet */
        CK(cudaMalloc(&j->d_sched, sizeof(FFSched)));
        CK(cudaMemset(j->d_sched, 0, sizeof(FFSched)));
        if (d->lazy_states)
            CK(cudaMalloc(&j->d_touched, B * d->max_slices * (size_t)((P->total_ctx + 31) / 32) * sizeof(uint32_t)));
        CK(cudaMalloc(&j->d_lines, B * d->max_slices * P->ncoded * 2 * d->line_stride * sizeof(int32_t)));
        if (d->line_stride < d->s.width + 8) {
            CK(cudaMalloc(&j->d_wide_lines, (size_t)DEC_WIDE_POOL * P->ncoded * 2 * (d->s.width + 8) * sizeof(int32_t)));
            CK(cudaMalloc(&j->d_wide_used, sizeof(uint32_t)));
        }
        CK(cudaMalloc(&j->d_frames, B * P->frame_bytes));
        CK(cudaMemset(j->d_frames, 0, B * P->frame_bytes));
        CK(cudaMalloc(&j->d_result, B * d->max_slices * sizeof(FFDecResult)));
        CK(cudaHostAlloc(&j->h_result, B * d->max_slices * sizeof(FFDecResult), cudaHostAllocDefault));
        CK(cudaMalloc(&j->d_weight, B * d->max_slices * sizeof(uint32_t)));
        CK(cudaMalloc(&j->d_weight_sorted, B * d->max_slices * sizeof(uint32_t)));
        CK(cudaMalloc(&j->d_order, B * d->max_slices * sizeof(uint32_t)));
        j->sort_tmp_bytes = ffk_sort_tmp_bytes((int)(B * d->max_slices));
        CK(cudaMalloc(&j->d_sort_tmp, j->sort_tmp_bytes));
        j->meta = (DecFrameMeta *)calloc(B, sizeof(DecFrameMeta));
        if (!j->meta)
            return fail(FFGPU_ENOMEM, "out of memory");
    }
    /* the fills above ran on the default stream; the launch groups use streams of their own */
    CK(cudaDeviceSynchronize());
    return 0;
}

/* the device side appears with the first packet; a failure part-way leaves nothing behind */
static int dec_device_init(ffgpu_decoder *d)
{
    int ndev = 0, r;
    if (d->dev_ready)
        return 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(FFGPU_EXTERNAL, "no CUDA device: the FFV1 pixel path has no CPU fallback");
    if ((r = dec_device_create(d)) < 0) {
        dec_device_release(d);
        return r;
    }
    d->dev_ready = 1;
    return 0;
}

extern "C" int ffgpu_ffv1_decode_init(ffgpu_decoder **pdec, const ffgpu_dec_options *opt)
{
    ffgpu_decoder *d;
    int r;
    if (!pdec || !opt)
        return fail(FFGPU_EINVAL, "null argument");
    *pdec = NULL;
    g_err[0] = 0;
    if (!opt->width || !opt->height)
        return fail(FFGPU_INVALIDDATA, "decode_init: zero dimensions");
    d = (ffgpu_decoder *)calloc(1, sizeof(*d));
    if (!d)
        return fail(FFGPU_ENOMEM, "out of memory");
    d->opt = *opt;
    d->opt.extradata = NULL;
    d->s.width = opt->width;
    d->s.height = opt->height;
    d->s.nh = d->s.nv = 1;
    ff_default_tables(&d->s.def_tab);
    d->s.cur_tab = d->s.def_tab;
    if (opt->extradata_size > 0) {
        if ((r = ff_parse_extradata(&d->s, opt->extradata, opt->extradata_size)) < 0) {
            ffgpu_ffv1_decode_close(d);
            return fail(r, "decode_init: invalid extradata (%d)", r);
        }
        if ((r = dec_setup_stream(d)) < 0) {
            ffgpu_ffv1_decode_close(d);
            return r;
        }
    }
    d->hs.max_slices = d->s.nh * d->s.nv;
    if (opt->ndevices > 1) {
        if (opt->ndevices > FFGPU_MAX_DEVICES) {
            ffgpu_ffv1_decode_close(d);
            return fail(FFGPU_EINVAL, "ndevices %d exceeds %d", opt->ndevices, FFGPU_MAX_DEVICES);
        }
        d->opt.device = opt->devices[0];
        if (d->have_params && d->intra) {
            ffgpu_dec_options so = *opt;
            so.ndevices = 0;
            for (int i = 0; i < opt->ndevices; i++) {
                so.device = opt->devices[i];
                if ((r = ffgpu_ffv1_decode_init(&d->sub[i], &so)) < 0) {
                    ffgpu_ffv1_decode_close(d);
                    return r;
                }
                d->nsub = i + 1;
            }
        }
    }
    *pdec = d;
    return 0;
}

/* Streams without extradata (v0/v1) announce their parameters in the first key frame: parse
 * that header (host only, nothing is decoded) so the caller can size its output picture. */
extern "C" int ffgpu_ffv1_decoder_probe(ffgpu_decoder *d, const uint8_t *pkt, size_t size)
{
    FFDecSlice tmp[1];
    FFDecFrameInfo info;
    FFDecHostState hs;
    int r;
    if (!d || !pkt)
        return fail(FFGPU_EINVAL, "null argument");
    if (d->have_params)
        return 0;
    hs = d->hs;
    hs.max_slices = 1;
    r = ff_dec_parse_packet(&d->s, &hs, pkt, size, 0, tmp, &info);
    if (r < 0)
        return fail(r, "invalid packet (%d)", r);
    if ((r = dec_setup_stream(d)) < 0)
        return r;
    d->hs.max_slices = d->max_slices;
    return 0;
}

extern "C" const char *ffgpu_ffv1_decoder_pix_fmt(const ffgpu_decoder *d)
{
    return d && d->s.pf ? d->s.pf->name : NULL;
}

extern "C" void ffgpu_ffv1_decoder_info(const ffgpu_decoder *d, int info[8])
{
    info[0] = d->s.version; info[1] = d->s.micro_version; info[2] = d->s.ac; info[3] = d->s.nh;
    info[4] = d->s.nv; info[5] = d->s.ec; info[6] = d->s.bits; info[7] = d->s.colorspace;
}

extern "C" uint64_t ffgpu_ffv1_decoder_launches(const ffgpu_decoder *d)
{
    uint64_t n = d->launches;
    for (int i = 0; i < d->nsub; i++)
        n += d->sub[i]->launches;
    return n;
}

static void dec_fill_dev(const ffgpu_decoder *d, const DecJob *j, uint8_t *frames, FFDecDev *D)
{
    memset(D, 0, sizeof(*D));
    D->work = j->d_work;
    D->nslices = j->d_nslices;
    D->qt = d->d_qt;
    D->tab = d->d_tab;
    D->initial = d->d_initial;
    D->pkt = j->d_pkt;
    D->state = d->intra ? j->d_state : d->d_state_shared;
    D->lines = j->d_lines;
    D->line_stride = d->line_stride;
    D->wide_lines = j->d_wide_lines;
    D->wide_stride = d->s.width + 8;
    D->wide_count = j->d_wide_lines ? DEC_WIDE_POOL : 0;
    D->wide_used = j->d_wide_used;
    D->frames = frames;
    D->result = j->d_result;
    D->max_slices = d->max_slices;
    D->max_ctx = d->max_ctx;
    D->state_per_frame = d->intra;
    D->qt_count = d->s.qt_count;
    D->gate_wait = getenv("FFGPU_GATE_WAIT") ? atoi(getenv("FFGPU_GATE_WAIT")) : FF_NEW_WAIT;
    D->touched = j->d_touched;
    D->touched_words = (d->P.total_ctx + 31) / 32;
    D->any_five = d->any_five;
    D->generic = d->generic;
    D->lane_stride = 1;                            /* set per launch from the number of work items */
    D->heavy_stride = heavy_stride_opt();
    D->heavy_factor = heavy_factor_opt();
    D->sched = D->heavy_stride ? j->d_sched : NULL;
    D->hdr.micro_version = d->s.micro_version;
    D->hdr.qt_count = d->s.qt_count;
    D->hdr.ctx_cap = d->max_ctx;
    for (int i = 0; i < FF_MAX_QUANT_TABLES; i++)
        D->hdr.ctx_count[i] = i < d->s.qt_count ? d->s.ctx_count[i] : 0;
    D->weight = j->d_weight;
    D->weight_sorted = j->d_weight_sorted;
    D->iota = d->d_iota;
    D->order = j->d_order;
    D->sort_tmp = j->d_sort_tmp;
    D->sort_tmp_bytes = j->sort_tmp_bytes;
}

static int download_picture(const ffgpu_decoder *d, const uint8_t *d_frame, const ffgpu_picture_out *dst,
                            cudaStream_t st)
{
    uint8_t *data[4];
    for (int k = 0; k < 4; k++)
        data[k] = (uint8_t *)dst->data[k];
    return copy_picture(&d->P, d->s.pf, d->s.width, d->s.height, data, dst->linesize, (uint8_t *)d_frame, 0, st);
}

static int dec_launch(ffgpu_decoder *d, DecJob *j, uint8_t *frames, cudaStream_t st, int download)
{
    FFDecDev D;
    int r;
    dec_fill_dev(d, j, frames, &D);
    D.lane_stride = coder_lane_stride((long)j->n * d->max_slices);
    if (d->profile_next)
        D.events = d->events;
    {
        /* packets and work items are read from pinned host memory by the SMs: as copy-engine
         * transfers they would queue behind the encoder's picture uploads */
        FFCopyArgs c;
        memset(&c, 0, sizeof(c));
        c.seg[0].dst = j->d_pkt;      c.seg[0].src = j->h_pkt;      c.seg[0].bytes = align_up(j->pkt_used + 64, 16);
        c.seg[1].dst = j->d_work;     c.seg[1].src = j->h_work;
        c.seg[1].bytes = (size_t)j->n * d->max_slices * sizeof(FFDecSlice);
        c.seg[2].dst = j->d_nslices;  c.seg[2].src = j->h_nslices;  c.seg[2].bytes = j->n * sizeof(int);
        c.nseg = 3;
        if ((r = ffk_copy_segments(&c, st)) < 0)
            return fail(r, "kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        d->launches += r;
    }
    r = ffk_decode_group(&d->P, &D, j->n, st);
    if (r < 0)
        return fail(r, "kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    d->launches += r;
    trace_mark(st, "dec k1", (int)(j - d->jobs), j->n);
    {
        FFCopyArgs c;                              /* per-slice results, also by the SMs */
        memset(&c, 0, sizeof(c));
        c.seg[0].dst = j->h_result;   c.seg[0].src = j->d_result;
        c.seg[0].bytes = (size_t)j->n * d->max_slices * sizeof(FFDecResult);
        c.nseg = 1;
        if ((r = ffk_copy_segments(&c, st)) < 0)
            return fail(r, "kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        d->launches += r;
    }
    if (download) {
        CK(cudaEventRecord(j->decoded, st));
        CK(cudaStreamWaitEvent(d->down_stream, j->decoded, 0));
        for (int i = 0; i < j->n; i++) {
            if (!j->meta[i].has_dst)
                continue;
            if (j->meta[i].staged) {
                /* an ordinary AVFrame: one linear DMA into pinned staging now, the host
                 * memcpy into the frame when the picture is handed back */
                if (!j->h_stage)
                    CK(cudaHostAlloc(&j->h_stage, (size_t)d->max_batch * d->P.frame_bytes, cudaHostAllocDefault));
                CK(cudaMemcpyAsync(j->h_stage + (size_t)i * d->P.frame_bytes,
                                   j->d_frames + (size_t)i * d->P.frame_bytes, d->P.frame_bytes,
                                   cudaMemcpyDeviceToHost, d->down_stream));
            } else if ((r = download_picture(d, j->d_frames + (size_t)i * d->P.frame_bytes, &j->meta[i].dst,
                                             d->down_stream)) < 0)
                return r;
        }
    }
    return 0;
}

/* append one packet to the group being filled; parses its frame/slice headers on the host */
static int dec_add_packet(ffgpu_decoder *d, DecJob *j, const uint8_t *pkt, size_t size, int64_t pts,
                          const ffgpu_picture_out *dst)
{
    DecFrameMeta *m;
    size_t off = align_up(j->pkt_used, 16);
    int n;
    if (off + size + 64 > j->pkt_cap) {
        if (j->n > 0)
            return FFGPU_EAGAIN;                   /* caller launches the group first */
        cudaFreeHost(j->h_pkt);
        cudaFree(j->d_pkt);
        j->h_pkt = j->d_pkt = NULL;
        j->pkt_cap = align_up(size * 2 + 4096, 4096);
        CK(cudaHostAlloc(&j->h_pkt, j->pkt_cap, cudaHostAllocDefault));
        CK(cudaMalloc(&j->d_pkt, j->pkt_cap));
        off = 0;
    }
    memcpy(j->h_pkt + off, pkt, size);
    memset(j->h_pkt + off + size, 0, 64);          /* AV_INPUT_BUFFER_PADDING_SIZE */
    m = &j->meta[j->n];
    memset(m, 0, sizeof(*m));
    n = ff_dec_parse_packet(&d->s, &d->hs, j->h_pkt + off, size, (uint32_t)off,
                            j->h_work + (size_t)j->n * d->max_slices, &m->info);
    if (n < 0)
        return fail(n, "invalid packet (%d)", n);
    for (int i = 0; i < n; i++) {
        const FFDecSlice *w = &j->h_work[(size_t)j->n * d->max_slices + i];
        for (int k = 0; k < d->P.nsets && !w->skip && !w->parse; k++)
            if (d->s.ctx_count[w->qidx[k]] > d->max_ctx)
                return fail(FFGPU_ENOSYS, "slice switches to a larger quant table mid-stream");
    }
    m->nslices = n;
    m->pts = pts;
    memcpy(m->damaged, d->hs.damaged, sizeof(m->damaged));
    memcpy(m->rect, d->hs.rect, sizeof(FFSliceRect) * (n > 0 ? n : 0));
    if (dst) {
        m->dst = *dst;
        m->has_dst = 1;
        const int staged = needs_staging(dst->data[0], dst->linesize);
        if (staged < 0)
            return fail(FFGPU_EINVAL, "bottom-up pictures (negative linesize) in device memory are not supported");
        m->staged = staged;
        /* the first pageable destination creates the staging of every launch group */
        for (int g = 0; g < d->depth && m->staged && !j->h_stage; g++)
            if (!d->jobs[g].h_stage)
                CK(cudaHostAlloc(&d->jobs[g].h_stage, (size_t)d->max_batch * d->P.frame_bytes, cudaHostAllocDefault));
    }
    j->h_nslices[j->n] = n;
    j->pkt_used = off + size;
    j->n++;
    return 0;
}

/* per-slice damage of picture i of a finished group: CRC / header failures found on the
 * device and the end-of-slice check of ffv1dec.c:351-359.  Returns the damaged-slice count. */
static int dec_mark_damage(ffgpu_decoder *d, DecJob *j, int i)
{
    DecFrameMeta *m = &j->meta[i];
    const FFDevParams *P = &d->P;
    int damaged = 0;
    for (int s = 0; s < m->nslices; s++) {
        const FFDecSlice *w = &j->h_work[(size_t)i * d->max_slices + s];
        const FFDecResult *res = &j->h_result[(size_t)i * d->max_slices + s];
        if (w->parse) {
            /* CRC, header and rectangle were established on the device */
            if (res->flags & (FF_RES_CRC_BAD | FF_RES_HDR_BAD)) {
                m->damaged[s] = 1;
                d->hs.damaged[s] = 1;
            }
            m->rect[s].x = res->x; m->rect[s].y = res->y;
            m->rect[s].w = res->w; m->rect[s].h = res->h;
        } else if (!w->skip && (res->flags & FF_RES_HDR_BAD)) {
            m->damaged[s] = 1;                     /* no picture-wide scratch left for it */
            d->hs.damaged[s] = 1;
        }
        if (!(res->flags & FF_RES_NOT_DECODED) && P->ac != FF_AC_GOLOMB && P->version > 2) {
            const int v = (int)res->size - (int)res->end_pos - 2 - 5 * P->ec;
            if (v) {
                m->damaged[s] = 1;                 /* "bytestream end mismatching by %d" */
                d->hs.damaged[s] = 1;
            }
        }
        damaged += m->damaged[s];
    }
    return damaged;
}

/* after a group has finished: damage bookkeeping (ffv1dec.c:351-359, :940-969) */
static int dec_finish_frame(ffgpu_decoder *d, DecJob *j, int i, ffgpu_picture_out *out)
{
    DecFrameMeta *m = &j->meta[i];
    const FFDevParams *P = &d->P;
    uint8_t *frame = j->d_frames + (size_t)i * P->frame_bytes;
    int damaged, r;
    damaged = dec_mark_damage(d, j, i);
    if (damaged && d->have_prev) {
        const uint8_t *prev = i > 0 ? j->d_frames + (size_t)(i - 1) * P->frame_bytes : d->d_prev;
        if (i == 0)
            CK(cudaEventSynchronize(d->prev_ready));
        for (int s = m->nslices - 1; s >= 0; s--)
            if (m->damaged[s]) {
                r = ffk_conceal_rect(P, frame, prev, m->rect[s].x, m->rect[s].y, m->rect[s].w, m->rect[s].h,
                                     d->s.pf->depth > 8, j->stream);
                if (r < 0)
                    return fail(r, "conceal launch failed");
                d->launches += r;
            }
        if (m->has_dst && m->staged)
            CK(cudaMemcpyAsync(j->h_stage + (size_t)i * P->frame_bytes, frame, P->frame_bytes,
                               cudaMemcpyDeviceToHost, j->stream));
        else if (m->has_dst && (r = download_picture(d, frame, &m->dst, j->stream)) < 0)
            return r;
        CK(cudaStreamSynchronize(j->stream));
    }
    if (m->has_dst && m->staged &&
        (r = stage_picture(P, d->s.pf, d->s.width, d->s.height, m->dst.data, m->dst.linesize,
                           j->h_stage + (size_t)i * P->frame_bytes, 0)) < 0)
        return r;
    if (i == j->n - 1) {
        /* keep the last picture of the group for the next group's concealment; the copy is
         * only awaited by the (rare) concealment path itself */
        CK(cudaMemcpyAsync(d->d_prev, frame, P->frame_bytes, cudaMemcpyDeviceToDevice, j->stream));
        CK(cudaEventRecord(d->prev_ready, j->stream));
    }
    d->have_prev = 1;
    if (out) {
        out->key_frame = m->info.key_frame;
        out->interlaced_frame = m->info.interlaced_frame;
        out->top_field_first = m->info.top_field_first;
        out->sar_num = m->info.sar_num;
        out->sar_den = m->info.sar_den;
        out->damaged_slices = damaged;
        out->pts = m->pts;
    }
    return 0;
}

static int dec_launch_group(ffgpu_decoder *d, DecJob *j)
{
    int r;
    trace_mark(j->stream, "dec k0", (int)(j - d->jobs), j->n);
    r = dec_launch(d, j, j->d_frames, j->stream, 1);
    if (r < 0)
        return r;
    trace_mark(d->down_stream, "dec done", (int)(j - d->jobs), j->n);
    CK(cudaEventRecord(j->done, d->down_stream));
    j->state = JOB_RUNNING;
    j->drained = 0;
    return 0;
}

static int dec_receive(ffgpu_decoder *d, ffgpu_picture_out *out, int block);

/* launch the group being filled although it is not full (the caller cannot send more) */
static int dec_kick(ffgpu_decoder *d)
{
    if (d->dev_ready) {
        DecJob *j = &d->jobs[d->fill];
        if (j->state == JOB_FILLING && j->n > 0) {
            int r;
            cudaSetDevice(d->opt.device);
            if ((r = dec_launch_group(d, j)) < 0)
                return r;
            d->fill = (d->fill + 1) % d->depth;
        }
    }
    return 0;
}

/* ---- routing over several GPUs ---- */
static int mdec_send(ffgpu_decoder *p, const uint8_t *pkt, size_t size, int64_t pts, const ffgpu_picture_out *dst)
{
    int r;
    if (!pkt) {
        for (int i = 0; i < p->nsub; i++)
            if ((r = ffgpu_ffv1_decode_send_packet(p->sub[i], NULL, 0, 0, NULL)) < 0)
                return r;
        p->flushing = 1;
        return 0;
    }
    if (p->flushing)
        return fail(FFGPU_EOF, "send_packet after flush");
    r = ffgpu_ffv1_decode_send_packet(p->sub[p->in_count % (uint64_t)p->nsub], pkt, size, pts, dst);
    if (r == 0)
        p->in_count++;
    else if (r == FFGPU_EAGAIN)
        p->send_blocked = 1;
    return r;
}

static int mdec_receive(ffgpu_decoder *p, ffgpu_picture_out *out)
{
    ffgpu_decoder *s;
    int r;
    if (p->out_count == p->in_count) {
        if (!p->flushing)
            return FFGPU_EAGAIN;
        for (int i = 0; i < p->nsub; i++)
            dec_receive(p->sub[i], out, 0);        /* every sub-decoder ends its own flush */
        p->flushing = 0;
        return FFGPU_EOF;
    }
    s = p->sub[p->out_count % (uint64_t)p->nsub];
    r = dec_receive(s, out, 0);
    if (r == FFGPU_EAGAIN && (p->send_blocked || p->flushing)) {
        if ((r = dec_kick(s)) < 0)
            return r;
        r = dec_receive(s, out, 1);
    }
    if (r == 0) {
        p->out_count++;
        p->send_blocked = 0;
    }
    return r;
}

extern "C" int ffgpu_ffv1_decode_send_packet(ffgpu_decoder *d, const uint8_t *pkt, size_t size,
                                              int64_t pts, const ffgpu_picture_out *dst)
{
    if (d && d->nsub)
        return mdec_send(d, pkt, size, pts, dst);
    if (d)
        cudaSetDevice(d->opt.device);        /* the current device is per host thread */
    DecJob *j;
    int r;
    if (!d)
        return fail(FFGPU_EINVAL, "null decoder");
    if (!pkt) {
        d->flushing = 1;
        if (d->dev_ready) {
            j = &d->jobs[d->fill];
            if (j->state == JOB_FILLING && j->n > 0) {
                if ((r = dec_launch_group(d, j)) < 0)
                    return r;
                d->fill = (d->fill + 1) % d->depth;
            }
        }
        return 0;
    }
    if (d->flushing)
        return fail(FFGPU_EOF, "send_packet after flush");
    if (d->dev_ready) {
        j = &d->jobs[d->fill];
        if (j->state == JOB_RUNNING || j->state == JOB_DRAINING)
            return FFGPU_EAGAIN;
    }
    /* the first packet may have to create the device side */
    if (!d->dev_ready) {
        if (!d->have_params) {
            FFDecSlice tmp[1];
            FFDecFrameInfo info;
            FFDecHostState hs = d->hs;
            hs.max_slices = 1;
            r = ff_dec_parse_packet(&d->s, &hs, pkt, size, 0, tmp, &info);
            if (r < 0)
                return fail(r, "invalid packet (%d)", r);
            if ((r = dec_setup_stream(d)) < 0)
                return r;
            d->hs.max_slices = d->max_slices;
        }
        if ((r = dec_size_states(d, pkt, size)) < 0)
            return r;
        if ((r = dec_device_init(d)) < 0)
            return r;
    }
    j = &d->jobs[d->fill];
    if (j->state == JOB_FREE) {
        j->state = JOB_FILLING;
        j->n = 0;
        j->pkt_used = 0;
    }
    r = dec_add_packet(d, j, pkt, size, pts, dst);
    if (r == FFGPU_EAGAIN) {
        /* packet arena full: launch what we have, the caller retries */
        if ((r = dec_launch_group(d, j)) < 0)
            return r;
        d->fill = (d->fill + 1) % d->depth;
        return ffgpu_ffv1_decode_send_packet(d, pkt, size, pts, dst);
    }
    if (r < 0) {
        if (j->n == 0)
            j->state = JOB_FREE;
        return r;
    }
    if (j->n == d->max_batch) {
        if ((r = dec_launch_group(d, j)) < 0)
            return r;
        d->fill = (d->fill + 1) % d->depth;
    }
    return 0;
}

extern "C" int ffgpu_ffv1_decode_receive_frame(ffgpu_decoder *d, ffgpu_picture_out *out)
{
    if (d && d->nsub)
        return mdec_receive(d, out);
    return dec_receive(d, out, 0);
}

/* block != 0: wait for the oldest running group instead of returning FFGPU_EAGAIN */
static int dec_receive(ffgpu_decoder *d, ffgpu_picture_out *out, int block)
{
    if (d)
        cudaSetDevice(d->opt.device);        /* the current device is per host thread */
    DecJob *j;
    int r, i;
    if (!d)
        return fail(FFGPU_EINVAL, "null decoder");
    if (!d->dev_ready) {
        if (d->flushing) {
            d->flushing = 0;
            return FFGPU_EOF;
        }
        return FFGPU_EAGAIN;
    }
    j = &d->jobs[d->head];
    if (j->state == JOB_FREE || j->state == JOB_FILLING) {
        if (d->flushing) {
            d->flushing = 0;                       /* drained: the handle accepts packets again */
            return FFGPU_EOF;
        }
        return FFGPU_EAGAIN;
    }
    if (j->state == JOB_RUNNING) {
        const DecJob *f = &d->jobs[d->fill];
        const int must_wait = block || d->flushing || f->state == JOB_RUNNING || f->state == JOB_DRAINING;
        if (!must_wait && cudaEventQuery(j->done) == cudaErrorNotReady)
            return FFGPU_EAGAIN;
        CK(cudaEventSynchronize(j->done));
        j->state = JOB_DRAINING;
    }
    i = j->drained;
    if (!j->meta[i].has_dst) {
        /* destination supplied only now: copy this picture out synchronously */
        if (!out)
            return fail(FFGPU_EINVAL, "no destination picture");
        j->meta[i].dst = *out;
        j->meta[i].has_dst = 1;
        if ((r = download_picture(d, j->d_frames + (size_t)i * d->P.frame_bytes, out, j->stream)) < 0)
            return r;
        CK(cudaStreamSynchronize(j->stream));
    }
    if ((r = dec_finish_frame(d, j, i, out)) < 0)
        return r;
    if (++j->drained == j->n) {
        j->state = JOB_FREE;
        j->n = 0;
        d->head = (d->head + 1) % d->depth;
    }
    return 0;
}

extern "C" int ffgpu_ffv1_decode_frame(ffgpu_decoder *d, const uint8_t *pkt, size_t size,
                                       ffgpu_picture_out *out, int *got_frame)
{
    DecJob *j;
    int r;
    if (got_frame)
        *got_frame = 0;
    if (!d || !pkt || !out)
        return fail(FFGPU_EINVAL, "null argument");
    if (d->nsub) {
        if (d->in_count != d->out_count)
            return fail(FFGPU_EINVAL, "decode_frame while send/receive packets are pending");
        r = ffgpu_ffv1_decode_frame(d->sub[d->in_count % (uint64_t)d->nsub], pkt, size, out, got_frame);
        if (r >= 0) {
            d->in_count++;
            d->out_count++;
        }
        return r;
    }
    if (d->dev_ready)
        for (int i = 0; i < d->depth; i++)
            if (d->jobs[i].state != JOB_FREE)
                return fail(FFGPU_EINVAL, "decode_frame while send/receive packets are pending");
    if ((r = ffgpu_ffv1_decode_send_packet(d, pkt, size, 0, out)) < 0)
        return r;
    j = &d->jobs[d->fill];
    if (j->state == JOB_FILLING) {
        if ((r = dec_launch_group(d, j)) < 0)
            return r;
        d->fill = (d->fill + 1) % d->depth;
    }
    j = &d->jobs[d->head];
    CK(cudaEventSynchronize(j->done));
    j->state = JOB_DRAINING;
    if ((r = ffgpu_ffv1_decode_receive_frame(d, out)) < 0)
        return r;
    if (got_frame)
        *got_frame = 1;
    return (int)size;
}

extern "C" int ffgpu_ffv1_decode_device(ffgpu_decoder *d, const uint8_t *const *pkts,
                                        const size_t *sizes, int nframes, void *d_frames,
                                        void *cuda_stream)
{
    if (d)
        cudaSetDevice(d->opt.device);        /* the current device is per host thread */
    DecJob *j;
    cudaStream_t st;
    int r;
    if (!d || !pkts || !sizes || nframes <= 0 || !d_frames)
        return fail(FFGPU_EINVAL, "bad argument");
    if (d->nsub)
        return fail(FFGPU_EINVAL, "device-resident batches live on one GPU: open one handle per device");
    if (d->dev_ready)
        for (int i = 0; i < d->depth; i++)
            if (d->jobs[i].state != JOB_FREE)
                return fail(FFGPU_EINVAL, "decode_device while send/receive packets are pending");
    if (!d->have_params || !d->dev_ready) {
        if (!d->have_params) {
            FFDecSlice tmp[1];
            FFDecFrameInfo info;
            FFDecHostState hs = d->hs;
            hs.max_slices = 1;
            r = ff_dec_parse_packet(&d->s, &hs, pkts[0], sizes[0], 0, tmp, &info);
            if (r < 0)
                return fail(r, "invalid packet (%d)", r);
            if ((r = dec_setup_stream(d)) < 0)
                return r;
            d->hs.max_slices = d->max_slices;
        }
        if ((r = dec_size_states(d, pkts[0], sizes[0])) < 0)
            return r;
        if ((r = dec_device_init(d)) < 0)
            return r;
    }
    if (nframes > d->max_batch)
        return fail(FFGPU_EINVAL, "nframes %d exceeds max_batch %d", nframes, d->max_batch);
    j = &d->jobs[0];
    /* the previous device batch reads the pinned staging buffers asynchronously */
    CK(cudaEventSynchronize(j->done));
    j->n = 0;
    j->pkt_used = 0;
    {
        /* the whole batch goes into one launch: size the packet arena for it (incompressible
         * pictures pack to about their raw size, the default arena assumes half of it) */
        size_t need = 256;
        for (int i = 0; i < nframes; i++)
            need += align_up(sizes[i], 16) + 64;
        if (need > 0xFFFF0000u)
            return fail(FFGPU_ENOSPC, "packets of the batch exceed 4 GiB: use smaller batches");
        if (need > j->pkt_cap) {
            cudaFreeHost(j->h_pkt);
            cudaFree(j->d_pkt);
            j->h_pkt = j->d_pkt = NULL;
            j->pkt_cap = align_up(need + need / 8, 4096);
            if (j->pkt_cap > 0xFFFF0000u)
                j->pkt_cap = 0xFFFF0000u;
            CK(cudaHostAlloc(&j->h_pkt, j->pkt_cap, cudaHostAllocDefault));
            CK(cudaMalloc(&j->d_pkt, j->pkt_cap));
        }
    }
    for (int i = 0; i < nframes; i++) {
        r = dec_add_packet(d, j, pkts[i], sizes[i], 0, NULL);
        if (r == FFGPU_EAGAIN)
            return fail(FFGPU_ENOSPC, "packets of the batch exceed the packet arena");
        if (r < 0) {
            j->n = 0;
            return r;
        }
    }
    st = cuda_stream ? (cudaStream_t)cuda_stream : j->stream;
    d->profile_next = d->profile;
    r = dec_launch(d, j, (uint8_t *)d_frames, st, 0);
    if (r >= 0)
        CK(cudaEventRecord(j->done, st));
    d->profile_next = 0;
    j->n_device = r >= 0 ? nframes : 0;
    j->n = 0;
    j->state = JOB_FREE;
    return r;
}

extern "C" int ffgpu_ffv1_decode_device_status(ffgpu_decoder *d, int *damaged_slices, int nframes)
{
    DecJob *j;
    int total = 0;
    if (!d || !d->dev_ready)
        return fail(FFGPU_EINVAL, "no device batch");
    cudaSetDevice(d->opt.device);
    j = &d->jobs[0];
    if (nframes > j->n_device)
        nframes = j->n_device;
    CK(cudaEventSynchronize(j->done));             /* kernels + result table copy have finished */
    for (int i = 0; i < nframes; i++) {
        const int n = dec_mark_damage(d, j, i);
        if (damaged_slices)
            damaged_slices[i] = n;
        total += n;
    }
    return total;
}

extern "C" int ffgpu_ffv1_decoder_profile(ffgpu_decoder *d, int enable)
{
    if (!d)
        return fail(FFGPU_EINVAL, "null decoder");
    if (enable && !d->events[0])
        for (int i = 0; i <= FFK_DEC_KERNELS; i++) {
            cudaEvent_t ev;
            CK(cudaEventCreate(&ev));
            d->events[i] = ev;
        }
    d->profile = enable;
    return 0;
}

extern "C" int ffgpu_ffv1_decoder_kernel_ms(ffgpu_decoder *d, float *ms, int n)
{
    if (!d || !d->profile || !d->events[0])
        return fail(FFGPU_EINVAL, "profiling not enabled");
    for (int i = 0; i < n && i < FFK_DEC_KERNELS; i++)
        CK(cudaEventElapsedTime(&ms[i], (cudaEvent_t)d->events[i], (cudaEvent_t)d->events[i + 1]));
    return FFK_DEC_KERNELS;
}

extern "C" int ffgpu_ffv1_decode_close(ffgpu_decoder *d)
{
    if (!d)
        return 0;
    for (int i = 0; i < d->nsub; i++)
        ffgpu_ffv1_decode_close(d->sub[i]);
    for (int i = 0; i <= FFK_DEC_KERNELS; i++)
        if (d->events[i])
            cudaEventDestroy((cudaEvent_t)d->events[i]);
    if (d->dev_ready) {
        cudaSetDevice(d->opt.device);
        cudaDeviceSynchronize();
        trace_dump();
        dec_device_release(d);
    }
    free(d->h_slices);
    ff_stream_free(&d->s);
    free(d);
    return 0;
}
