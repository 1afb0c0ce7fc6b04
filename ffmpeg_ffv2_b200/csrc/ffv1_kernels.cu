/*
 * ffv1_kernels.cu -- sm_100a kernels of the FFV1 slice pixel path.
 *
 * The work is integer, byte-granular and, inside a slice, a strict dependence chain
 * (adaptive probability states + range coder low/range).  Parallelism therefore comes from
 *   - samples            (stage A: one lane per sample column of a strip, coalesced along x)
 *   - slices x pictures  (stage B / decode: one thread per slice; a launch group holds
 *                         max_batch pictures so that tens of thousands of independent
 *                         coders are resident on the 148 SMs)
 * No tensor cores: nothing here is a contraction.
 *
 * Kernel        reference span it replaces
 * k_symbolize   encode_plane / encode_rgb_frame sample loads + get_context + predict
 *               (ffv1enc.c:274-312, ffv1enc_template.c:125-201, ffv1_template.c:23-52)
 * k_fill_state  ff_ffv1_clear_slice_state (ffv1.c:182-207)
 * k_code_*      encode_line / put_symbol_inline / put_vlc_symbol per slice, the job
 *               avctx->execute(encode_slice) fans out (ffv1enc.c:1233, :1045-1120)
 * k_pack_*      the serial compaction loop of encode_frame (ffv1enc.c:1236-1262): device
 *               prefix sums over slice and packet sizes + gather + size/CRC trailers
 * k_decode      decode_slice after its header (ffv1dec.c:304-359) for every slice of every
 *               packet of the group
 */
#include <stdlib.h>
#include <cuda_runtime.h>
#include <cub/block/block_reduce.cuh>
#include <cub/block/block_scan.cuh>
#include <cub/device/device_radix_sort.cuh>

#include "../../include/ffgpu.h"
#include "ffv1_launch.h"
#include "ffv1_slice.cuh"

#define CODE_THREADS FF_CODE_THREADS
#define SYM_THREADS  256
#define SYM_ROWS     4      /* rows a warp symbolizes per step (stage A, planar) */
#define SYM_CHUNK    16     /* rows of a work unit (a multiple of SYM_ROWS)      */

static inline int launch_ok(void)
{
    return cudaGetLastError() == cudaSuccess;
}

static inline void mark(void **events, int i, cudaStream_t st)
{
    if (events && events[i])
        cudaEventRecord((cudaEvent_t)events[i], st);
}

/* ---------------- version 4: per-slice RCT coefficients ---------------- */
/* choose_rct_params (ffv1enc.c:963-1043) as a reduction: every thread sums the 15 candidate
 * magnitudes of its pixels, the block reduces them, one atomic per candidate and block. */
__global__ void __launch_bounds__(SYM_THREADS)
k_rct_stat(const FFDevParams P, const FFDevSlice *__restrict__ slices,
           const uint8_t *__restrict__ frames, int32_t *__restrict__ stat)
{
    typedef cub::BlockReduce<int32_t, SYM_THREADS> Reduce;
    __shared__ typename Reduce::TempStorage tmp;
    const FFDevSlice sl = slices[blockIdx.x];
    const uint8_t *frame = frames + (size_t)blockIdx.y * P.frame_bytes;
    const int w1 = sl.w - 1, h1 = sl.h - 1;          /* pixels with x >= 1 and y >= 1 */
    int32_t acc[FF_RCT_CANDIDATES];
    for (int i = 0; i < FF_RCT_CANDIDATES; i++)
        acc[i] = 0;
    if (w1 > 0 && h1 > 0) {
        const uint32_t n = (uint32_t)w1 * h1;
        for (uint32_t i = blockIdx.z * SYM_THREADS + threadIdx.x; i < n; i += gridDim.z * SYM_THREADS) {
            int32_t v[FF_RCT_CANDIDATES];
            const int y = (int)(i / (uint32_t)w1), x = (int)(i - (uint32_t)y * w1);
            ff_rct_pixel_stat(P, frame, sl.x, sl.y, x + 1, y + 1, v);
            for (int k = 0; k < FF_RCT_CANDIDATES; k++)
                acc[k] = (int32_t)((uint32_t)acc[k] + (uint32_t)v[k]);
        }
    }
    int32_t *out = stat + ((size_t)blockIdx.y * P.nslices + blockIdx.x) * 16;
    for (int k = 0; k < FF_RCT_CANDIDATES; k++) {
        const int32_t sum = Reduce(tmp).Sum(acc[k]);
        __syncthreads();
        if (threadIdx.x == 0 && sum)
            atomicAdd(&out[k], sum);
    }
}

__global__ void k_rct_pick(const int32_t *__restrict__ stat, int *__restrict__ rct, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    int by, ry;
    ff_rct_pick(stat + (size_t)i * 16, &by, &ry);
    rct[2 * i] = by;
    rct[2 * i + 1] = ry;
}

/* ---------------- stage A ---------------- */
__global__ void __launch_bounds__(SYM_THREADS)
k_symbolize(const FFDevParams P, const FFDevSlice *__restrict__ slices,
            const uint8_t *__restrict__ frames, const int16_t *__restrict__ qt,
            uint32_t *__restrict__ tokens, uint32_t *__restrict__ weight, const int *__restrict__ rct)
{
    typedef cub::BlockReduce<uint32_t, SYM_THREADS> Reduce;
    __shared__ typename Reduce::TempStorage tmp;
    const FFDevSlice sl = slices[blockIdx.x];
    const uint8_t *frame = frames + (size_t)blockIdx.y * P.frame_bytes;
    uint32_t *tok = tokens + (size_t)blockIdx.y * P.frame_tokens + sl.tok_off;
    uint32_t wsum = 0;
    const int *rc = rct ? rct + 2 * ((size_t)blockIdx.y * P.nslices + blockIdx.x) : (const int *)0;
    const int by = rc ? rc[0] : 1, ry = rc ? rc[1] : 1;
    for (uint32_t i = blockIdx.z * SYM_THREADS + threadIdx.x; i < sl.ntok; i += gridDim.z * SYM_THREADS) {
        const uint32_t t = ff_symbolize_index(P, sl, frame, qt, i, by, ry);
        tok[i] = t;
        wsum += ff_token_weight(t);
    }
    wsum = Reduce(tmp).Sum(wsum);
    if (threadIdx.x == 0 && weight)
        atomicAdd(&weight[(size_t)blockIdx.y * P.nslices + blockIdx.x], wsum);
}

/* Stage A for the planar YCbCr / gray (+alpha) layouts.
 *
 * Work unit = a strip of SYM_CHUNK rows x TW samples of one plane of a slice, taken by one
 * warp; the units of a slice (all planes) form one flat list that the warps of its CTA(s)
 * pull from a shared counter, so luma and chroma keep every warp busy.  A lane owns one
 * sample COLUMN of the strip and walks down it SYM_ROWS rows at a time: the row loads of a
 * step are independent (all in flight together, coalesced along x), the two rows above are
 * carried in registers, so T and TT are the lane's own values and L / LT / RT / LL come from
 * the neighbouring lanes by shuffle.  The outermost lanes of the warp are halo columns (HL on
 * the left, one on the right): they only feed the shuffles, which removes every seam special
 * case; what is left of the border rules of ffv1enc.c:287-288 are three selects (x == 0,
 * x == 1, x == w-1).  The context quantiser and the per-plane geometry live in shared memory. */
struct SymPlane {
    const uint8_t *origin;      /* sample (0,0) of the plane's slice rectangle */
    int w, h, ntiles, ucount;
    float inv_tiles;
    uint32_t base;              /* first token of the plane inside the slice */
    int ctx_base, pitch, step;
};

template <bool WIDE, bool FIVE>
__global__ void __launch_bounds__(SYM_THREADS, 4)
k_symbolize_planar(const FFDevParams P, const FFDevSlice *__restrict__ slices,
                   const uint8_t *__restrict__ frames, const int16_t *__restrict__ qt,
                   uint32_t *__restrict__ tokens, uint32_t *__restrict__ weight)
{
    constexpr int HL = FIVE ? 2 : 1;
    constexpr int TW = 32 - HL - 1;
    __shared__ int16_t sq[FF_QT_STRIDE];
    __shared__ SymPlane sp[FF_MAX_PLANES];
    __shared__ int s_next, s_nunits;
    {
        const uint32_t *src = (const uint32_t *)(qt + (size_t)P.set_qidx[0] * FF_QT_STRIDE);
        for (int i = threadIdx.x; i < FF_QT_STRIDE / 2; i += SYM_THREADS)
            ((uint32_t *)sq)[i] = src[i];
    }
    if (threadIdx.x == 0) {
        const FFDevSlice *slp = &slices[blockIdx.x];
        const uint8_t *frame = frames + (size_t)blockIdx.y * P.frame_bytes;
        uint32_t base = 0;
        int n = 0;
        for (int k = 0; k < FF_MAX_PLANES; k++) {
            SymPlane q;
            memset(&q, 0, sizeof(q));
            if (k < P.ncoded) {
                const int mem = P.cp[k].mem;
                q.w = slp->seg_w[k];
                q.h = slp->seg_lines[k];
                q.ntiles = (q.w + TW - 1) / TW;
                q.ucount = q.ntiles * ((q.h + SYM_CHUNK - 1) / SYM_CHUNK);
                q.inv_tiles = 1.0f / (float)(q.ntiles > 0 ? q.ntiles : 1);
                q.base = base;
                q.ctx_base = P.set_base[P.cp[k].set];
                q.pitch = P.pitch[mem];
                q.step = WIDE ? 2 : P.cp[k].step;
                q.origin = frame + P.plane_off[mem] + P.cp[k].off +
                           (size_t)(slp->y >> P.cp[k].vs) * q.pitch + (size_t)(slp->x >> P.cp[k].hs) * q.step;
                base += (uint32_t)q.w * q.h;
                n += q.ucount;
            }
            sp[k] = q;
        }
        s_nunits = n;
        s_next = 0;
    }
    __syncthreads();
    uint32_t *tok = tokens + (size_t)blockIdx.y * P.frame_tokens + slices[blockIdx.x].tok_off;
    const int lane = threadIdx.x & 31;
    const int shift = P.packed_lsb ? 0 : 16 - P.sbits;
    const int cbits = P.cbits;
    const int nunits = s_nunits;
    uint32_t wsum = 0;

#define SYM_LOAD(ptr) (WIDE ? (int)(int16_t)(*(const uint16_t *)(ptr) >> shift) : (int)*(ptr))
    for (;;) {
        int u = 0;
        if (lane == 0)
            u = atomicAdd(&s_next, 1);
        u = __shfl_sync(0xffffffffu, u, 0) * gridDim.z + blockIdx.z;
        if (u >= nunits)
            break;
        int k = 0;
#pragma unroll
        for (int q = 0; q < FF_MAX_PLANES - 1; q++)
            if (k == q && u >= sp[q].ucount) {
                u -= sp[q].ucount;
                k = q + 1;
            }
        const int w = sp[k].w, h = sp[k].h, ntiles = sp[k].ntiles;
        const int chunk = (int)(((float)u + 0.5f) * sp[k].inv_tiles);
        const int tile = u - chunk * ntiles;
        const int y0 = chunk * SYM_CHUNK;
        const int yend = min(y0 + SYM_CHUNK, h);
        const int x = tile * TW - HL + lane;
        const int ctx_base = sp[k].ctx_base;
        const ptrdiff_t pitch = sp[k].pitch;
        const bool in_x = x >= 0 && x < w;
        /* column x, row y0-2; only dereferenced where the sample exists */
        const uint8_t *p = sp[k].origin + (ptrdiff_t)(y0 - 2) * pitch + (ptrdiff_t)x * sp[k].step;
        int v[SYM_ROWS + 2];
        /* row y0-2 is only needed by TT (5-input contexts) and by LT at x == 0 */
        v[0] = (in_x && y0 >= 2 && (FIVE || tile == 0)) ? SYM_LOAD(p) : 0;
        v[1] = (in_x && y0 >= 1) ? SYM_LOAD(p + pitch) : 0;
        p += 2 * pitch;
        const bool first = x == 0, last = x == w - 1;
        const bool out = lane >= HL && lane < 31 && x < w;
        uint32_t *trow = tok + sp[k].base + (uint32_t)y0 * w + x;
        for (int y = y0; y < yend; y += SYM_ROWS) {
#pragma unroll
            for (int j = 0; j < SYM_ROWS; j++)
                v[j + 2] = (in_x && y + j < yend) ? SYM_LOAD(p + j * pitch) : 0;
#pragma unroll
            for (int j = 0; j < SYM_ROWS; j++) {
                if (y + j >= yend)
                    break;
                const int cur = v[j + 2], T = v[j + 1], TT = v[j];
                int L = __shfl_up_sync(0xffffffffu, cur, 1);
                int LT = __shfl_up_sync(0xffffffffu, T, 1);
                int RT = __shfl_down_sync(0xffffffffu, T, 1);
                int LL = FIVE ? __shfl_up_sync(0xffffffffu, cur, 2) : 0;
                if (FIVE)                          /* x == 1: sample[-1] = T of column 0; x == 0: 0 */
                    LL = x >= 2 ? LL : (x == 1 ? LT : 0);
                L = first ? T : L;                 /* left border: ffv1enc.c:287 */
                LT = first ? TT : LT;
                RT = last ? T : RT;                /* right border: ffv1enc.c:288 */
                int ctx = sq[(L - LT) & 0xFF] + sq[256 + ((LT - T) & 0xFF)] + sq[512 + ((T - RT) & 0xFF)];
                if (FIVE)
                    ctx += sq[768 + ((LL - L) & 0xFF)] + sq[1024 + ((TT - T) & 0xFF)];
                int diff = cur - ff_median3(L, L + T - LT, T);
                const int neg = ctx < 0;
                ctx = neg ? -ctx : ctx;
                diff = ff_fold(neg ? -diff : diff, cbits);
                const uint32_t t = ((uint32_t)diff << FF_TOKEN_CTX_BITS) | (uint32_t)(ctx_base + ctx);
                if (out) {
                    trow[j * w] = t;
                    wsum += ff_token_weight(t);
                }
            }
            v[0] = v[SYM_ROWS];
            v[1] = v[SYM_ROWS + 1];
            p += SYM_ROWS * pitch;
            trow += SYM_ROWS * w;
        }
    }
#undef SYM_LOAD
    /* per-warp reduction + one atomic per warp: no block-wide barrier at the end */
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    if (lane == 0 && weight && wsum)
        atomicAdd(&weight[(size_t)blockIdx.y * P.nslices + blockIdx.x], wsum);
}


/* ---------------- stage A, bulk-copy form (sm_90+ async proxy, SASS: UBLKCP) ----------------
 * Same result as k_symbolize_planar.  The strip a CTA works on -- up to SB_TX x SB_TY samples
 * of one plane of a slice plus its halo (two rows above, two columns left, one right) -- is
 * brought into shared memory by the bulk-copy engine (cp.async.bulk, one 16-byte-aligned row
 * segment per copy, completion counted on an mbarrier), double buffered: while the threads
 * turn one strip into tokens the next one is already in flight, and no thread's issue slots
 * or registers are spent on global loads.  A thread owns a sample column of (a row segment of)
 * the strip and walks down it, so T / LT / RT / TT are last row's values carried in registers
 * and only the current sample and its left / right neighbours are read from shared memory.
 * The border rules of ffv1enc.c:284-289 are the same selects as in k_symbolize_planar. */
#define SB_TX      256
#define SB_TY      16
#define SB_THREADS 256

__device__ __forceinline__ void sb_mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void sb_mbar_expect(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
                 "r"(bytes) : "memory");
}
__device__ __forceinline__ void sb_bulk_load(void *smem, const void *gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem), "r"(bytes),
                   "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void sb_mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; spins++) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (spins > (1u << 24))
            __trap();                                /* a lost copy must not hang the GPU */
    }
}

struct SbPlane {
    const uint8_t *plane;       /* row 0, byte 0 of the memory plane in this picture */
    int X0, Y0, w, h;           /* slice rectangle in the plane's sample grid        */
    int ntx, nunits, pitch, ctx_base;
    uint32_t base;              /* first token of the plane inside the slice         */
};

template <bool WIDE, bool FIVE>
__global__ void __launch_bounds__(SB_THREADS)
k_symbolize_bulk(const FFDevParams P, const FFDevSlice *__restrict__ slices,
                 const uint8_t *__restrict__ frames, const int16_t *__restrict__ qt,
                 uint32_t *__restrict__ tokens, uint32_t *__restrict__ weight)
{
    constexpr int BPS = WIDE ? 2 : 1;
    constexpr int ROWB = ((SB_TX + 3) * BPS + 30) & ~15;     /* aligned span of a halo'd row */
    __shared__ __align__(128) uint8_t tile[2][(SB_TY + 2) * ROWB + 16];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ int16_t sq[FF_QT_STRIDE];
    __shared__ SbPlane sp[FF_MAX_PLANES];
    __shared__ int s_nunits;
    {
        const uint32_t *src = (const uint32_t *)(qt + (size_t)P.set_qidx[0] * FF_QT_STRIDE);
        for (int i = threadIdx.x; i < FF_QT_STRIDE / 2; i += SB_THREADS)
            ((uint32_t *)sq)[i] = src[i];
    }
    if (threadIdx.x == 0) {
        const FFDevSlice *slp = &slices[blockIdx.x];
        const uint8_t *frame = frames + (size_t)blockIdx.y * P.frame_bytes;
        uint32_t base = 0;
        int n = 0;
        for (int k = 0; k < FF_MAX_PLANES; k++) {
            SbPlane q;
            memset(&q, 0, sizeof(q));
            if (k < P.ncoded) {
                const int mem = P.cp[k].mem;
                q.w = slp->seg_w[k];
                q.h = slp->seg_lines[k];
                q.X0 = slp->x >> P.cp[k].hs;
                q.Y0 = slp->y >> P.cp[k].vs;
                q.ntx = (q.w + SB_TX - 1) / SB_TX;
                q.nunits = q.ntx * ((q.h + SB_TY - 1) / SB_TY);
                q.pitch = P.pitch[mem];
                q.plane = frame + P.plane_off[mem];
                q.ctx_base = P.set_base[P.cp[k].set];
                q.base = base;
                base += (uint32_t)q.w * q.h;
                n += q.nunits;
            }
            sp[k] = q;
        }
        s_nunits = n;
        sb_mbar_init(&bar[0], 1);
        sb_mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t *tok = tokens + (size_t)blockIdx.y * P.frame_tokens + slices[blockIdx.x].tok_off;
    const int cbits = P.cbits;
    const int nunits = s_nunits;
    uint32_t wsum = 0, phase0 = 0, phase1 = 0;

    /* unit u -> plane k, tile origin (x0, y0), extent (txe, tye), byte window [b0, b0 + nb) */
#define SB_UNIT(u, k, x0, y0, txe, tye, b0, nb)                                                   \
    int k = 0, x0, y0, txe, tye, b0, nb;                                                         \
    {                                                                                             \
        int uu = (u);                                                                             \
        for (int q = 0; q < FF_MAX_PLANES - 1; q++)                                               \
            if (k == q && uu >= sp[q].nunits) {                                                   \
                uu -= sp[q].nunits;                                                               \
                k = q + 1;                                                                        \
            }                                                                                     \
        const int ty_ = uu / sp[k].ntx, tx_ = uu - ty_ * sp[k].ntx;                               \
        x0 = tx_ * SB_TX;                                                                         \
        y0 = ty_ * SB_TY;                                                                         \
        txe = min(SB_TX, sp[k].w - x0);                                                           \
        tye = min(SB_TY, sp[k].h - y0);                                                           \
        const int c0_ = (sp[k].X0 + x0 - 2) * BPS, c1_ = (sp[k].X0 + x0 + txe + 1) * BPS;         \
        b0 = max(c0_, 0) & ~15;                                                                   \
        nb = min((c1_ + 15) & ~15, sp[k].pitch) - b0;                                             \
    }

    /* warp 0 starts the copies of unit u into buffer b */
#define SB_ISSUE(u, b)                                                                            \
    if (threadIdx.x < 32) {                                                                       \
        SB_UNIT(u, k_, x0_, y0_, txe_, tye_, b0_, nb_)                                            \
        const int first_ = max(y0_ - 2, 0), nrows_ = y0_ + tye_ - first_;                         \
        (void)txe_;                                                                               \
        if (threadIdx.x == 0) {                                                                   \
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                          \
            sb_mbar_expect(&bar[b], (uint32_t)nrows_ * (uint32_t)nb_);                            \
        }                                                                                         \
        __syncwarp();                                                                             \
        if ((int)threadIdx.x < nrows_) {                                                          \
            const int r_ = first_ + (int)threadIdx.x;                                             \
            sb_bulk_load(&tile[b][(r_ - (y0_ - 2)) * ROWB],                                       \
                         sp[k_].plane + (size_t)(sp[k_].Y0 + r_) * sp[k_].pitch + b0_,            \
                         (uint32_t)nb_, &bar[b]);                                                 \
        }                                                                                         \
    }

    int u = blockIdx.z;
    if (u < nunits) {
        SB_ISSUE(u, 0)
    }
    for (int it = 0; u < nunits; it++, u += gridDim.z) {
        const int b = it & 1;
        __syncthreads();                             /* buffer b^1 is no longer read */
        if (u + (int)gridDim.z < nunits) {
            if (b) { SB_ISSUE(u + gridDim.z, 0) } else { SB_ISSUE(u + gridDim.z, 1) }
        }
        sb_mbar_wait(&bar[b], b ? phase1 : phase0);
        if (b) phase1 ^= 1; else phase0 ^= 1;
        {
            SB_UNIT(u, k, x0, y0, txe, tye, b0, nb)
            (void)nb;
            /* threads: `cols` sample columns x `nseg` row segments */
            int cols = 32;
            while (cols < txe)
                cols <<= 1;
            const int nseg = SB_THREADS / cols, rps = (tye + nseg - 1) / nseg;
            const int lx = threadIdx.x & (cols - 1), seg = threadIdx.x / cols;
            const int x = x0 + lx, ys = y0 + seg * rps, ye = min(ys + rps, y0 + tye);
            if (lx < txe && ys < ye) {
                const int w = sp[k].w;
                const uint8_t *tb = tile[b];
                /* byte offset of column x (and its neighbours, clamped into the copied window) */
                const int cx = (sp[k].X0 + x) * BPS - b0;
                const int cl = max(cx - BPS, 0), cll = max(cx - 2 * BPS, 0), cr = cx + BPS;
#define SB_LD(row, off) (WIDE ? (int)*(const int16_t *)(tb + ((row) - (y0 - 2)) * ROWB + (off))    \
                              : (int)tb[((row) - (y0 - 2)) * ROWB + (off)])
                int Tm = ys >= 1 ? SB_LD(ys - 1, cx) : 0;
                int Lm = ys >= 1 ? SB_LD(ys - 1, cl) : 0;
                int Rm = ys >= 1 ? SB_LD(ys - 1, cr) : 0;
                int TTm = ys >= 2 ? SB_LD(ys - 2, cx) : 0;
                const bool first = x == 0, last = x == w - 1;
                uint32_t *trow = tok + sp[k].base + (uint32_t)ys * w + x;
                const int ctx_base = sp[k].ctx_base;
                for (int y = ys; y < ye; y++) {
                    const int cur = SB_LD(y, cx), Lr = SB_LD(y, cl), Rr = SB_LD(y, cr);
                    const int T = Tm, TT = TTm;
                    const int L = first ? T : Lr;
                    const int LT = first ? TT : Lm;
                    const int RT = last ? T : Rm;
                    int ctx = sq[(L - LT) & 0xFF] + sq[256 + ((LT - T) & 0xFF)] + sq[512 + ((T - RT) & 0xFF)];
                    if (FIVE) {
                        const int LLr = SB_LD(y, cll);
                        const int LL = x >= 2 ? LLr : (x == 1 ? Lm : 0);
                        ctx += sq[768 + ((LL - L) & 0xFF)] + sq[1024 + ((TT - T) & 0xFF)];
                    }
                    int diff = cur - ff_median3(L, L + T - LT, T);
                    const int neg = ctx < 0;
                    ctx = neg ? -ctx : ctx;
                    diff = ff_fold(neg ? -diff : diff, cbits);
                    const uint32_t t = ((uint32_t)diff << FF_TOKEN_CTX_BITS) | (uint32_t)(ctx_base + ctx);
                    *trow = t;
                    trow += w;
                    wsum += ff_token_weight(t);
                    TTm = T;
                    Tm = cur;
                    Lm = Lr;
                    Rm = Rr;
                }
#undef SB_LD
            }
        }
    }
#undef SB_UNIT
#undef SB_ISSUE
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    if ((threadIdx.x & 31) == 0 && weight && wsum)
        atomicAdd(&weight[(size_t)blockIdx.y * P.nslices + blockIdx.x], wsum);
}

/* ---------------- adaptive state reset ---------------- */
/* fills the state arenas of the frames flagged as key frames with a 64-bit pattern
 * (range coder: 128 in every byte; Golomb: {drift 0, error_sum 4, bias 0, count 1}) */
__global__ void k_fill_state(uint2 *__restrict__ state, size_t words_per_frame,
                             const uint8_t *__restrict__ frame_key, int state_per_frame,
                             uint32_t lo, uint32_t hi)
{
    const int f = blockIdx.y;
    if (!frame_key[f])
        return;
    uint2 *p = state + (state_per_frame ? (size_t)f * words_per_frame : 0);
    const uint2 v = make_uint2(lo, hi);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words_per_frame;
         i += (size_t)gridDim.x * blockDim.x)
        p[i] = v;
}

/* ---------------- heavy / light schedule of the slice coders ---------------- */
__global__ void __launch_bounds__(1024)
k_sched(const uint32_t *__restrict__ sorted, int n, float factor, int heavy_stride, FFSched *out)
{
    typedef cub::BlockReduce<unsigned long long, 1024> Reduce;
    __shared__ typename Reduce::TempStorage tmp;
    __shared__ unsigned long long s_sum;
    unsigned long long sum = 0;
    for (int i = threadIdx.x; i < n; i += 1024)
        sum += sorted[i];
    sum = Reduce(tmp).Sum(sum);
    if (threadIdx.x == 0)
        s_sum = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        /* first sorted index whose weight is not above factor x mean (the array descends) */
        const double limit = (double)factor * (double)s_sum / (double)(n > 0 ? n : 1);
        int lo = 0, hi = n / 4;                      /* at most a quarter of the items */
        while (lo < hi) {
            const int mid = (lo + hi) / 2;
            if ((double)sorted[mid] > limit)
                lo = mid + 1;
            else
                hi = mid;
        }
        out->n_heavy = (uint32_t)lo;
        out->heavy_threads = ((uint32_t)lo * (uint32_t)heavy_stride + 31u) & ~31u;
    }
}

/* sorted work item of coder thread `tid`, or -1 */
__device__ __forceinline__ int ff_sched_item(int tid, int total, int lane_stride, const FFSched *sched,
                                             int heavy_stride)
{
    if (lane_stride > 1) {                           /* few items: spread them over the warps */
        if (tid % lane_stride)
            return -1;
        tid /= lane_stride;
        return tid < total ? tid : -1;
    }
    if (sched) {
        const uint32_t nh = sched->n_heavy, ht = sched->heavy_threads;
        if ((uint32_t)tid < ht) {
            const uint32_t i = (uint32_t)tid / (uint32_t)heavy_stride;
            return ((uint32_t)tid % (uint32_t)heavy_stride) == 0 && i < nh ? (int)i : -1;
        }
        tid = (int)(nh + ((uint32_t)tid - ht));
    }
    return tid < total ? tid : -1;
}

/* threads a coder launch needs in the worst case (a quarter of the items heavy) */
static long ff_sched_threads(long total, int lane_stride, int sched, int heavy_stride)
{
    if (lane_stride > 1)
        return total * lane_stride;
    if (sched)
        return ((total / 4 * heavy_stride + 31) & ~31L) + total;
    return total;
}

/* second pass: the slices of key frames start from the initial states of the extradata
 * (ff_ffv1_clear_slice_state with initial_states, ffv1.c:192-197) */
__global__ void k_fill_state_initial(uint4 *__restrict__ state, size_t vec_per_slice, int nslices,
                                     const uint8_t *__restrict__ frame_key, int state_per_frame,
                                     const uint4 *__restrict__ initial)
{
    const int f = blockIdx.y;
    if (!frame_key[f])
        return;
    uint4 *p = state + (state_per_frame ? (size_t)f * vec_per_slice * nslices : 0);
    const size_t n = vec_per_slice * nslices;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = initial[i % vec_per_slice];
}

/* ---------------- stage B ---------------- */
template <bool LONE>
__global__ void __launch_bounds__(CODE_THREADS)
k_code_range(const FFDevParams P, const FFEncDev E, int nframes)
{
    for (int i = threadIdx.x; i < (int)sizeof(FFRacTables) / 4; i += CODE_THREADS)
        ((uint32_t *)&ff_s_tab)[i] = ((const uint32_t *)E.tab)[i];
    for (int i = threadIdx.x; i < FF_STAB_ROWS * FF_STAB_STRIDE; i += CODE_THREADS)
        ff_s_stab[i] = (uint8_t)ff_slot_of(i / FF_STAB_STRIDE, i % FF_STAB_STRIDE);
    __syncthreads();
    ff_fill_tab16(threadIdx.x, CODE_THREADS);
    __syncthreads();
    if (!LONE && E.rec && *E.split_ok)                       /* k_chain_states + k_code_records did it */
        return;
    const int tid = ff_sched_item(blockIdx.x * CODE_THREADS + threadIdx.x, nframes * P.nslices,
                                  E.lane_stride, E.sched, E.heavy_stride);
    if (tid < 0)
        return;
    const int gid = E.order ? (int)E.order[tid] : tid;      /* heaviest slices first */
    const int f = gid / P.nslices, s = gid - f * P.nslices;
    const FFDevSlice sl = E.slices[s];
    const size_t st_slot = (size_t)(E.state_per_frame ? f : 0) * P.nslices + s;
    uint32_t ovf = 0;
    FFPassStats pass;
    pass.rc_stat = E.rc_stat;
    pass.rc_stat2 = E.rc_stat2;
    pass.ctx_count = E.stat_ctx_count;
    uint32_t n;
#define CODE_ARGS_  \
        sl, E.tokens + (size_t)f * P.frame_tokens + sl.tok_off, \
        E.state + st_slot * P.total_ctx * FF_CONTEXT_SIZE, &ff_s_tab, \
        E.prefix[(size_t)E.frame_prefix_set[f] * P.nslices + s], E.prefix_bytes, \
        E.bs + (size_t)f * P.frame_bs + sl.bs_off, &ovf, 0, E.rct ? E.rct + 2 * (size_t)gid : (const int *)0, \
        E.rct ? (uint32_t)((16384 + (int64_t)P.width * P.height * 12) / P.nslices) : 0u, \
        E.rc_stat ? &pass : (const FFPassStats *)0
#define CODE_ARGS_LONE_  \
        sl, E.tokens + (size_t)f * P.frame_tokens + sl.tok_off, \
        E.state + st_slot * P.total_ctx * FF_CONTEXT_SIZE, &ff_s_tab, \
        E.prefix[(size_t)E.frame_prefix_set[f] * P.nslices + s], E.prefix_bytes, \
        E.bs + (size_t)f * P.frame_bs + sl.bs_off, &ovf, 0, E.rct ? E.rct + 2 * (size_t)gid : (const int *)0, \
        E.rct ? (uint32_t)((16384 + (int64_t)P.width * P.height * 12) / P.nslices) : 0u
    if (LONE)                           /* one slice per warp, no first-pass counters */
        n = ff_encode_slice_range_lone(CODE_ARGS_LONE_);
    else if (E.rc_stat)
        n = ff_encode_slice_range<true>(CODE_ARGS_);
    else
        n = ff_encode_slice_range<false>(CODE_ARGS_);
#undef CODE_ARGS_
#undef CODE_ARGS_LONE_
    E.slice_bytes[gid] = n;
    if (ovf)
        atomicOr(E.overflow, 1u);
}


/* ---------------- stage B in two halves (few-slice launches) ---------------- */
/* where each slice's decision records start: an exclusive sum of the decision counts stage A
 * produced; also decides whether the records fit the arena.  The split form only runs for
 * launches of at most a few thousand slices, a single thread adds them up. */
__global__ void k_rec_offsets(const FFEncDev E, int n)
{
    if (blockIdx.x || threadIdx.x)
        return;
    unsigned long long sum = 0;
    for (int i = 0; i < n; i++) {
        E.rec_off[i] = sum;
        sum += E.weight[i];
    }
    E.rec_off[n] = sum;
    *E.split_ok = sum <= E.rec_cap;
}

/* One warp per slice.  32 tokens per step, in coding order; tokens that share a context are
 * taken one after the other (the member of rank r in round r), all others at once.  The
 * state rows live in global memory (the arena of the one-kernel coder, L1-resident for the
 * contexts in use); __syncwarp() orders the rounds' accesses to a shared row. */
__global__ void __launch_bounds__(CODE_THREADS)
k_chain_states(const FFDevParams P, const FFEncDev E, int nframes)
{
    __shared__ FFRacTables s_tab;
    for (int i = threadIdx.x; i < (int)sizeof(FFRacTables) / 4; i += CODE_THREADS)
        ((uint32_t *)&s_tab)[i] = ((const uint32_t *)E.tab)[i];
    __syncthreads();
    if (!*E.split_ok)
        return;
    const int gid = blockIdx.x * (CODE_THREADS / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (gid >= nframes * P.nslices)
        return;
    const int f = gid / P.nslices, s = gid - f * P.nslices;
    const FFDevSlice sl = E.slices[s];
    const uint32_t *tokens = E.tokens + (size_t)f * P.frame_tokens + sl.tok_off;
    const size_t st_slot = (size_t)(E.state_per_frame ? f : 0) * P.nslices + s;
    uint8_t *rows = E.state + st_slot * P.total_ctx * FF_CONTEXT_SIZE;
    uint16_t *rec = E.rec + E.rec_off[gid];
    const uint32_t n = sl.ntok;
    const uint32_t guard_tok = E.rct ? n - (uint32_t)sl.seg_w[sl.nseg - 1] : 0xFFFFFFFFu;
    uint32_t base = 0;
    for (uint32_t i0 = 0; i0 < n; i0 += 32) {
        const uint32_t i = i0 + lane;
        const bool valid = i < n;
        const uint32_t tok = valid ? tokens[i] : 0u;
        uint32_t w = valid ? ff_token_weight(tok) : 0u, incl = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o)
                incl += v;
        }
        const uint32_t off = base + incl - w;
        base += __shfl_sync(0xffffffffu, incl, 31);
        if (valid && i == guard_tok)
            E.guard_rec[gid] = off;
        /* lanes without a token get a key no context has */
        const unsigned grp = __match_any_sync(0xffffffffu, valid ? (tok & FF_TOKEN_CTX_MASK) : 0x10000u + lane);
        const int rank = __popc(grp & ((1u << lane) - 1u));
        const int rounds = __reduce_max_sync(0xffffffffu, valid ? __popc(grp) : 0);
        for (int r = 0; r < rounds; r++) {
            if (valid && rank == r)
                ff_chain_token(tok, rows, &s_tab, rec + off);
            __syncwarp();
        }
    }
}

__global__ void __launch_bounds__(CODE_THREADS)
k_code_records(const FFDevParams P, const FFEncDev E, int nframes)
{
    for (int i = threadIdx.x; i < (int)sizeof(FFRacTables) / 4; i += CODE_THREADS)
        ((uint32_t *)&ff_s_tab)[i] = ((const uint32_t *)E.tab)[i];
    __syncthreads();
    if (!*E.split_ok)
        return;
    const int tid = ff_sched_item(blockIdx.x * CODE_THREADS + threadIdx.x, nframes * P.nslices,
                                  E.lane_stride, (const FFSched *)0, E.heavy_stride);
    if (tid < 0)
        return;
    const int gid = E.order ? (int)E.order[tid] : tid;
    const int f = gid / P.nslices, s = gid - f * P.nslices;
    const FFDevSlice sl = E.slices[s];
    uint32_t ovf = 0;
    const uint32_t n = ff_encode_slice_records(
        sl, E.rec + E.rec_off[gid], (uint32_t)(E.rec_off[gid + 1] - E.rec_off[gid]),
        E.rct ? E.guard_rec[gid] : 0xFFFFFFFFu, &ff_s_tab,
        E.prefix[(size_t)E.frame_prefix_set[f] * P.nslices + s], E.prefix_bytes,
        E.bs + (size_t)f * P.frame_bs + sl.bs_off, &ovf, E.rct ? E.rct + 2 * (size_t)gid : (const int *)0,
        E.rct ? (uint32_t)((16384 + (int64_t)P.width * P.height * 12) / P.nslices) : 0u);
    E.slice_bytes[gid] = n;
    if (ovf)
        atomicOr(E.overflow, 1u);
}

__global__ void __launch_bounds__(CODE_THREADS)
k_code_golomb(const FFDevParams P, const FFEncDev E, int nframes)
{
    const int tid = ff_sched_item(blockIdx.x * CODE_THREADS + threadIdx.x, nframes * P.nslices,
                                  E.lane_stride, E.sched, E.heavy_stride);
    if (tid < 0)
        return;
    const int gid = E.order ? (int)E.order[tid] : tid;
    const int f = gid / P.nslices, s = gid - f * P.nslices;
    const FFDevSlice sl = E.slices[s];
    const size_t st_slot = (size_t)(E.state_per_frame ? f : 0) * P.nslices + s;
    uint32_t ovf = 0;
    const uint32_t n = ff_encode_slice_golomb(
        P, sl, E.tokens + (size_t)f * P.frame_tokens + sl.tok_off,
        (uint2 *)E.state + st_slot * P.total_ctx,
        E.prefix[(size_t)E.frame_prefix_set[f] * P.nslices + s], E.prefix_bytes,
        E.bs + (size_t)f * P.frame_bs + sl.bs_off, &ovf, E.tab,
        E.rct ? E.rct + 2 * (size_t)gid : (const int *)0);
    E.slice_bytes[gid] = n;
    if (ovf)
        atomicOr(E.overflow, 1u);
}

/* ---------------- packet assembly ---------------- */
/* per picture: exclusive scan of the packed slice sizes -> slice offsets + packet size */
__global__ void __launch_bounds__(1024)
k_pack_slice_scan(const FFDevParams P, const FFEncDev E)
{
    typedef cub::BlockScan<uint32_t, 1024> Scan;
    __shared__ typename Scan::TempStorage tmp;
    const int f = blockIdx.x, s = threadIdx.x;
    uint32_t v = s < P.nslices ? ff_slice_packed_size(P, s, E.slice_bytes[(size_t)f * P.nslices + s]) : 0;
    uint32_t off, total;
    Scan(tmp).ExclusiveSum(v, off, total);
    if (s < P.nslices)
        E.slice_off[(size_t)f * P.nslices + s] = off;
    if (s == 0)
        E.pkt_size[f] = total;
}

/* over the pictures of the group: exclusive scan of packet sizes.  Packets start 16-byte
 * aligned and pkt_off counts 16-BYTE UNITS: a group of incompressible 4K pictures packs to
 * more than 4 GiB */
__global__ void __launch_bounds__(1024)
k_pack_frame_scan(const FFEncDev E, int nframes)
{
    typedef cub::BlockScan<uint32_t, 1024> Scan;
    __shared__ typename Scan::TempStorage tmp;
    const int f = threadIdx.x;
    uint32_t v = f < nframes ? ((E.pkt_size[f] + 15u) >> 4) : 0;
    uint32_t off, total;
    Scan(tmp).ExclusiveSum(v, off, total);
    if (f < nframes)
        E.pkt_off[f] = off;
    if (f == 0)
        E.pkt_off[nframes] = total;
}

__global__ void __launch_bounds__(CODE_THREADS)
k_pack_gather(const FFDevParams P, const FFEncDev E)
{
    __shared__ uint32_t crc_tab[256];
    for (int i = threadIdx.x; i < 256; i += CODE_THREADS)
        crc_tab[i] = ff_crc_table_entry(i);
    __syncthreads();
    const int f = blockIdx.y, s = blockIdx.x * CODE_THREADS + threadIdx.x;
    if (s >= P.nslices)
        return;
    const FFDevSlice sl = E.slices[s];
    ff_pack_slice(P, s, E.bs + (size_t)f * P.frame_bs + sl.bs_off,
                  E.slice_bytes[(size_t)f * P.nslices + s],
                  E.pkt + ((size_t)E.pkt_off[f] << 4) + E.slice_off[(size_t)f * P.nslices + s], crc_tab);
}

extern "C" int ffk_encode_group(const FFDevParams *P, const FFEncDev *E, int nframes, ffk_stream stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    int launches = 0;
    uint32_t max_tok = 0;
    if (nframes <= 0 || nframes > 1024 || P->nslices > 1024)
        return FFGPU_EINVAL;

    mark(E->events, 0, st);
    /* stage A */
    {
        /* enough z-blocks to cover large slices; small slices get one block each */
        max_tok = (uint32_t)(P->frame_tokens / (size_t)P->nslices) + 1;
        int z = (int)((max_tok + SYM_THREADS * 16 - 1) / (SYM_THREADS * 16));
        if (z < 1) z = 1;
        if (z > 64) z = 64;
        dim3 grid(P->nslices, nframes, z);
        if (E->weight)
            cudaMemsetAsync(E->weight, 0, sizeof(uint32_t) * (size_t)nframes * P->nslices, st);
        if (P->colorspace == 0) {
            /* about 4 units per warp: fine enough for the shared counter to balance the
             * warps, coarse enough to amortise the CTA prologue (a C2 slice has 32 units ->
             * one CTA of 8 warps) */
            const int five = E->five;
            const int tw = five ? 29 : 30;
            const int sw = (P->width + P->nh - 1) / P->nh + 1, sh = (P->height + P->nv - 1) / P->nv + 1;
            long units = 0;
            for (int k = 0; k < P->ncoded; k++) {
                const int wk = ((sw - 1) >> P->cp[k].hs) + 1, hk = ((sh - 1) >> P->cp[k].vs) + 1;
                units += (long)((wk + tw - 1) / tw) * ((hk + SYM_CHUNK - 1) / SYM_CHUNK);
            }
            int zr = (int)((units + 8 * 4 - 1) / (8 * 4));
            if (zr < 1) zr = 1;
            if (zr > 1024) zr = 1024;
            dim3 g2(P->nslices, nframes, zr);
            /* the bulk-copy form needs samples that are dense in their memory plane (every
             * planar layout; not the interleaved gray+alpha one).  FFGPU_STAGE_A=legacy keeps
             * the register/shuffle kernel for A/B runs. */
            int dense = 1;
            for (int k = 0; k < P->ncoded; k++)
                dense &= P->cp[k].step == (P->sbits > 8 ? 2 : 1) && P->cp[k].off == 0;
            /* measured (r02, same box): slices a strip wide or wider gain 1.3-1.9x from the
             * bulk-copy form (C1 10.9 -> 5.8 ms, C5 3.7 -> 2.8 ms); the 117-sample slices of
             * C2 fill only half of a strip's columns and stay on the shuffle kernel (7.9 vs
             * 9.8 ms).  FFGPU_STAGE_A=bulk forces the bulk form. */
            if (dense && !E->legacy_stage_a && (sw - 1 >= 192 || E->legacy_stage_a < 0)) {
                long tiles = 0;
                for (int k = 0; k < P->ncoded; k++) {
                    const int wk = ((sw - 1) >> P->cp[k].hs) + 1, hk = ((sh - 1) >> P->cp[k].vs) + 1;
                    tiles += (long)((wk + SB_TX - 1) / SB_TX) * ((hk + SB_TY - 1) / SB_TY);
                }
                /* a few strips per CTA so that the double buffering has something to overlap */
                int zb = (int)((tiles + 3) / 4);
                if (zb < 1) zb = 1;
                if (zb > 1024) zb = 1024;
                dim3 g3(P->nslices, nframes, zb);
#define SB_LAUNCH(W, F) k_symbolize_bulk<W, F><<<g3, SB_THREADS, 0, st>>>(*P, E->slices, E->frames, E->qt, \
                                                                          E->tokens, E->weight)
                if (P->sbits > 8) {
                    if (five) SB_LAUNCH(true, true); else SB_LAUNCH(true, false);
                } else {
                    if (five) SB_LAUNCH(false, true); else SB_LAUNCH(false, false);
                }
#undef SB_LAUNCH
            } else {
#define SYM_LAUNCH(W, F) k_symbolize_planar<W, F><<<g2, SYM_THREADS, 0, st>>>(*P, E->slices, E->frames, E->qt, \
                                                                              E->tokens, E->weight)
            if (P->sbits > 8) {
                if (five) SYM_LAUNCH(true, true); else SYM_LAUNCH(true, false);
            } else {
                if (five) SYM_LAUNCH(false, true); else SYM_LAUNCH(false, false);
            }
#undef SYM_LAUNCH
            }
        } else {
            if (E->rct) {
                /* version 4: the RCT coefficients of every slice first */
                const int n = nframes * P->nslices;
                cudaMemsetAsync(E->rct_stat, 0, sizeof(int32_t) * 16 * (size_t)n, st);
                k_rct_stat<<<grid, SYM_THREADS, 0, st>>>(*P, E->slices, E->frames, E->rct_stat);
                k_rct_pick<<<(n + 255) / 256, 256, 0, st>>>(E->rct_stat, E->rct, n);
                launches += 2;
            }
            k_symbolize<<<grid, SYM_THREADS, 0, st>>>(*P, E->slices, E->frames, E->qt, E->tokens, E->weight, E->rct);
        }
        mark(E->events, FFK_SYMBOLIZE + 1, st);
        launches++;
        if (!launch_ok()) return FFGPU_EXTERNAL;
    }
    /* state reset for key frames */
    {
        const int golomb = P->ac == FF_AC_GOLOMB;
        const size_t bytes = (size_t)P->nslices * P->total_ctx * (golomb ? 8 : FF_CONTEXT_SIZE);
        const size_t words = bytes / 8;
        int bx = (int)((words + 255) / 256);
        if (bx > 2048) bx = 2048;
        dim3 grid(bx, nframes);
        if (E->initial && !golomb)
            k_fill_state_initial<<<grid, 256, 0, st>>>((uint4 *)E->state, (size_t)P->total_ctx * FF_CONTEXT_SIZE / 16,
                                                       P->nslices, E->frame_key, E->state_per_frame,
                                                       (const uint4 *)E->initial);
        else
            k_fill_state<<<grid, 256, 0, st>>>((uint2 *)E->state, words, E->frame_key, E->state_per_frame,
                                               golomb ? FF_VLC_INIT_LO : 0x80808080u,
                                               golomb ? FF_VLC_INIT_HI : 0x80808080u);
        mark(E->events, FFK_FILL_STATE + 1, st);
        launches++;
        if (!launch_ok()) return FFGPU_EXTERNAL;
    }
    /* longest slices first: sort (decision count, slice id) descending */
    if (E->weight && E->order) {
        size_t tmp = E->sort_tmp_bytes;
        cub::DeviceRadixSort::SortPairsDescending(E->sort_tmp, tmp, E->weight, E->weight_sorted, E->iota,
                                                  E->order, nframes * P->nslices, 0, 32, st);
    }
    if (E->sched && E->weight && E->order && E->lane_stride <= 1)
        k_sched<<<1, 1024, 0, st>>>(E->weight_sorted, nframes * P->nslices, E->heavy_factor, E->heavy_stride,
                                    E->sched);
    mark(E->events, FFK_SORT + 1, st);
    /* stage B */
    {
        const long total = ff_sched_threads((long)nframes * P->nslices, E->lane_stride,
                                            E->sched && E->weight && E->order, E->heavy_stride);
        const int blocks = (int)((total + CODE_THREADS - 1) / CODE_THREADS);
        if (P->ac == FF_AC_GOLOMB) {
            k_code_golomb<<<blocks, CODE_THREADS, 0, st>>>(*P, *E, nframes);
        } else {
            /* one slice per warp: the straight-line coder (FFGPU_LONE=0: tuning hook) */
            const int lone_ok = !(getenv("FFGPU_LONE") && !atoi(getenv("FFGPU_LONE")));
            const bool lone = E->lane_stride == 32 && !E->rc_stat && lone_ok;
            if (lone) {
                k_code_range<true><<<blocks, CODE_THREADS, 0, st>>>(*P, *E, nframes);
            } else {
            if (E->rec && E->weight) {
                /* few, large slices: state chains by warps, then the bare arithmetic coder */
                const int n = nframes * P->nslices;
                k_rec_offsets<<<1, 32, 0, st>>>(*E, n);
                k_chain_states<<<(n + CODE_THREADS / 32 - 1) / (CODE_THREADS / 32), CODE_THREADS, 0, st>>>(*P, *E, nframes);
                k_code_records<<<blocks, CODE_THREADS, 0, st>>>(*P, *E, nframes);
                launches += 3;
            }
            k_code_range<false><<<blocks, CODE_THREADS, 0, st>>>(*P, *E, nframes);
            }
        }
        mark(E->events, FFK_CODE + 1, st);
        launches++;
        if (!launch_ok()) return FFGPU_EXTERNAL;
    }
    /* packet assembly */
    k_pack_slice_scan<<<nframes, 1024, 0, st>>>(*P, *E);
    mark(E->events, FFK_PACK_SLICE_SCAN + 1, st);
    k_pack_frame_scan<<<1, 1024, 0, st>>>(*E, nframes);
    mark(E->events, FFK_PACK_FRAME_SCAN + 1, st);
    {
        dim3 grid((P->nslices + CODE_THREADS - 1) / CODE_THREADS, nframes);
        k_pack_gather<<<grid, CODE_THREADS, 0, st>>>(*P, *E);
    }
    mark(E->events, FFK_PACK_GATHER + 1, st);
    launches += 3;
    if (!launch_ok()) return FFGPU_EXTERNAL;
    return launches;
}

/* ---------------- decoder ---------------- */
/* reset the adaptive states of the slices of key frames: 128 or the stream's initial
 * states per quant table (ffv1.c:182-207) */
__global__ void k_dec_init_state(const FFDevParams P, const FFDecDev D, int golomb)
{
    const int f = blockIdx.y, s = blockIdx.x;
    if (s >= D.nslices[f])
        return;
    const FFDecSlice w = D.work[(size_t)f * D.max_slices + s];
    if (!w.key_frame || w.skip)
        return;
    const size_t slot = (size_t)(D.state_per_frame ? f : 0) * D.max_slices + s;
    if (golomb) {
        uint2 *p = (uint2 *)D.state + slot * P.total_ctx;
        for (int i = threadIdx.x; i < P.total_ctx; i += blockDim.x)
            p[i] = make_uint2(FF_VLC_INIT_LO, FF_VLC_INIT_HI);
    } else {
        uint32_t *p = (uint32_t *)(D.state + slot * P.total_ctx * FF_CONTEXT_SIZE);
        const int words_per_set = D.max_ctx * FF_CONTEXT_SIZE / 4;
        for (int set = 0; set < P.nsets; set++) {
            const uint32_t *init = D.initial
                ? (const uint32_t *)(D.initial + (size_t)w.qidx[set] * D.max_ctx * FF_CONTEXT_SIZE) : 0;
            for (int i = threadIdx.x; i < words_per_set; i += blockDim.x)
                p[set * words_per_set + i] = init ? init[i] : 0x80808080u;
        }
    }
}

/* sort key of a decode work item: its byte count (0 for absent / skipped slices) */
__global__ void k_dec_keys(const FFDecDev D, int nframes)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= nframes * D.max_slices)
        return;
    const int f = gid / D.max_slices, s = gid - f * D.max_slices;
    uint32_t k = 0;
    if (s < D.nslices[f] && !D.work[gid].skip)
        k = D.work[gid].size;
    D.weight[gid] = k;
}

/* whole-arena reset for streams without initial-state tables: every key frame's slices get
 * the same 16-byte pattern */
__global__ void k_dec_fill_state(uint4 *__restrict__ state, size_t vec_per_frame, const FFDecDev D,
                                 uint4 pattern)
{
    const int f = blockIdx.y;
    if (!D.work[(size_t)f * D.max_slices].key_frame)
        return;
    uint4 *p = state + (D.state_per_frame ? (size_t)f * vec_per_frame : 0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < vec_per_frame;
         i += (size_t)gridDim.x * blockDim.x)
        p[i] = pattern;
}

/* SMODE 0: every stream (Golomb-Rice, RGB, MSB-aligned containers) through the generic slice
 * functions; SMODE 1 / 2: range-coded planar YCbCr / gray, 8-bit / LSB-packed 16-bit samples,
 * through ff_decode_slice_range_planar.  D.lane_stride > 1 spreads the work items over the
 * warps (one slice per `lane_stride` lanes): streams with a handful of large slices then run
 * one slice per warp, without the cost of 32 unrelated slices diverging in one warp. */
#ifndef FF_DEC_MINBLOCKS
#define FF_DEC_MINBLOCKS 1
#endif
template <int SMODE, bool FIVE, bool LONE = false>
__global__ void __launch_bounds__(CODE_THREADS, FF_DEC_MINBLOCKS)
k_decode(const FFDevParams P, const FFDecDev D, int nframes)
{
    for (int i = threadIdx.x; i < (int)sizeof(FFRacTables) / 4; i += CODE_THREADS)
        ((uint32_t *)&ff_s_tab)[i] = ((const uint32_t *)D.tab)[i];
    for (int i = threadIdx.x; i < D.qt_count * FF_QT_STRIDE / 2; i += CODE_THREADS)
        ((uint32_t *)ff_s_qt)[i] = ((const uint32_t *)D.qt)[i];
    __shared__ uint32_t s_crc[256];
    for (int i = threadIdx.x; i < 256; i += CODE_THREADS)
        s_crc[i] = ff_crc_table_entry(i);
    __syncthreads();
    ff_fill_tab16(threadIdx.x, CODE_THREADS);
    __syncthreads();
    /* the planar path keeps every lane of a warp inside the decode loop (lanes without a
     * slice idle there), so nothing returns before the call */
    const int tid = ff_sched_item(blockIdx.x * CODE_THREADS + threadIdx.x, nframes * D.max_slices,
                                  D.lane_stride, D.sched, D.heavy_stride);
    const bool have = tid >= 0;
    if ((SMODE == 0 || LONE) && !have)
        return;
    const int gid = have ? (D.order ? (int)D.order[tid] : tid) : 0;   /* largest slices first */
    const int f = gid / D.max_slices, s = gid - f * D.max_slices;
    FFDecResult r;
    FFDecSlice w;
    FFDecCtx C;
    bool live = false;
    r.end_pos = 0; r.overread = 0; r.error = 0; r.flags = FF_RES_NOT_DECODED;
    r.x = r.y = r.w = r.h = 0; r.size = 0; r.pad[0] = r.pad[1] = r.pad[2] = 0;
    C.lines = 0;
    if (have && s < D.nslices[f]) {
        w = D.work[gid];
        r.x = w.x; r.y = w.y; r.w = w.w; r.h = w.h;
        if (w.parse && !w.skip) {
            ff_dec_slice_header(P, D.hdr, &w, D.pkt, &ff_s_tab, s_crc, &r);
            r.x = w.x; r.y = w.y; r.w = w.w; r.h = w.h;
        }
        r.size = w.size;
        if (!w.skip) {
            r.flags &= ~FF_RES_NOT_DECODED;
            const size_t slot = (size_t)(D.state_per_frame ? f : 0) * D.max_slices + s;
            C.qt_all = D.qt;
            C.tab = &ff_s_tab;
            C.rstate = D.state + slot * P.total_ctx * FF_CONTEXT_SIZE;
            C.vstate = (uint2 *)D.state + slot * P.total_ctx;
            C.lines = D.lines + (size_t)gid * P.ncoded * 2 * D.line_stride;
            C.line_stride = D.line_stride;
            C.frame = D.frames + (size_t)f * P.frame_bytes;
            C.gate_wait = D.gate_wait;
            C.touched = D.touched ? D.touched + (size_t)gid * D.touched_words : (uint32_t *)0;
            C.any_five = D.any_five;
            C.lone = LONE;
            if (w.w + 8 > D.line_stride) {
                /* rectangle wider than the grid cell: picture-wide scratch from the pool */
                const uint32_t slot_w = D.wide_used ? atomicAdd(D.wide_used, 1u) : 0xFFFFFFFFu;
                if (slot_w < (uint32_t)D.wide_count) {
                    C.lines = D.wide_lines + (size_t)slot_w * P.ncoded * 2 * D.wide_stride;
                    C.line_stride = D.wide_stride;
                } else {
                    C.lines = 0;
                    r.flags |= FF_RES_HDR_BAD | FF_RES_NOT_DECODED;
                }
            }
            live = C.lines != 0;
        }
    }
    if (SMODE == 0 && LONE) {
        /* range-coded RGB, one live lane per warp: the straight-line decoder */
        if (live && w.pcm)
            ff_decode_slice_pcm(P, w, D.pkt, C, &r);
        else if (live)
            ff_decode_slice_range_rgb_lone<FIVE>(P, w, D.pkt, C, &r, 0);
    } else if (SMODE == 0) {
        if (live)
            ff_decode_slice(P, w, D.pkt, C, &r, 0);
    } else if (LONE) {
        /* one live lane per warp (D.lane_stride == 32): the straight-line decoder */
        if (live && w.pcm)
            ff_decode_slice_pcm(P, w, D.pkt, C, &r);
        else if (live)
            ff_decode_slice_range_planar_lone<SMODE ? SMODE : 1, FIVE>(P, w, D.pkt, C, &r, 0);
    } else {
        if (live && w.pcm) {                         /* version 4 PCM slice: the plain way */
            ff_decode_slice_pcm(P, w, D.pkt, C, &r);
            live = false;
        }
        __syncwarp();
        ff_decode_slice_range_planar<SMODE ? SMODE : 1, FIVE>(P, w, D.pkt, C, &r, 0, live);
    }
    if (have)
        D.result[gid] = r;
}

extern "C" size_t ffk_sort_tmp_bytes(int n)
{
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairsDescending((void *)0, bytes, (const uint32_t *)0, (uint32_t *)0,
                                              (const uint32_t *)0, (uint32_t *)0, n);
    return bytes + 256;
}

extern "C" int ffk_decode_group(const FFDevParams *P, const FFDecDev *D, int nframes, ffk_stream stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (nframes <= 0)
        return FFGPU_EINVAL;
    dim3 g0(D->max_slices, nframes);
    const int total = nframes * D->max_slices;
    mark(D->events, 0, st);
    if (D->touched) {
        /* states are created on first touch inside k_decode: clear the touched bits only */
        cudaMemsetAsync(D->touched, 0, (size_t)total * D->touched_words * sizeof(uint32_t), st);
    } else if (!D->initial && P->version < 4 &&
        ((size_t)D->max_slices * P->total_ctx * (P->ac == FF_AC_GOLOMB ? 8 : FF_CONTEXT_SIZE)) % 16 == 0) {
        const int golomb = P->ac == FF_AC_GOLOMB;
        const size_t vec = (size_t)D->max_slices * P->total_ctx * (golomb ? 8 : FF_CONTEXT_SIZE) / 16;
        int bx = (int)((vec + 255) / 256);
        if (bx > 1024) bx = 1024;
        dim3 g1(bx, nframes);
        const uint4 pat = golomb ? make_uint4(FF_VLC_INIT_LO, FF_VLC_INIT_HI, FF_VLC_INIT_LO, FF_VLC_INIT_HI)
                                 : make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
        k_dec_fill_state<<<g1, 256, 0, st>>>((uint4 *)D->state, vec, *D, pat);
    } else {
        k_dec_init_state<<<g0, 256, 0, st>>>(*P, *D, P->ac == FF_AC_GOLOMB);
    }
    mark(D->events, FFK_DEC_INIT_STATE + 1, st);
    if (D->wide_used)
        cudaMemsetAsync(D->wide_used, 0, sizeof(uint32_t), st);
    if (D->weight && D->order) {
        size_t tmp = D->sort_tmp_bytes;
        k_dec_keys<<<(total + 255) / 256, 256, 0, st>>>(*D, nframes);
        cub::DeviceRadixSort::SortPairsDescending(D->sort_tmp, tmp, D->weight, D->weight_sorted, D->iota,
                                                  D->order, total, 0, 32, st);
    }
    if (D->sched && D->weight && D->order && D->lane_stride <= 1)
        k_sched<<<1, 1024, 0, st>>>(D->weight_sorted, total, D->heavy_factor, D->heavy_stride, D->sched);
    mark(D->events, FFK_DEC_SORT + 1, st);
    {
        const long threads = ff_sched_threads(total, D->lane_stride, D->sched && D->weight && D->order,
                                              D->heavy_stride);
        const unsigned blocks = (unsigned)((threads + CODE_THREADS - 1) / CODE_THREADS);
        const size_t smem = (size_t)D->qt_count * FF_QT_STRIDE * sizeof(int16_t);
        const int planar = D->generic ? 0 : ff_decode_planar_mode(P);
#define DEC_LAUNCH(M, F) k_decode<M, F><<<blocks, CODE_THREADS, smem, st>>>(*P, *D, nframes)
#define DEC_LAUNCH_LONE(M, F) k_decode<M, F, true><<<blocks, CODE_THREADS, smem, st>>>(*P, *D, nframes)
        /* FFGPU_LONE=0: tuning hook, the warp-wide form also for one slice per warp */
        const int lone_ok = !(getenv("FFGPU_LONE") && !atoi(getenv("FFGPU_LONE")));
        if (planar && D->lane_stride == 32 && lone_ok) {
            if (planar == 1) {
                if (D->any_five) DEC_LAUNCH_LONE(1, true); else DEC_LAUNCH_LONE(1, false);
            } else {
                if (D->any_five) DEC_LAUNCH_LONE(2, true); else DEC_LAUNCH_LONE(2, false);
            }
        } else if (!planar && !D->generic && P->colorspace && P->ac != FF_AC_GOLOMB &&
                   D->lane_stride == 32 && lone_ok) {
            if (D->any_five) DEC_LAUNCH_LONE(0, true); else DEC_LAUNCH_LONE(0, false);
        } else if (planar == 1) {
            if (D->any_five) DEC_LAUNCH(1, true); else DEC_LAUNCH(1, false);
        } else if (planar == 2) {
            if (D->any_five) DEC_LAUNCH(2, true); else DEC_LAUNCH(2, false);
        } else {
            DEC_LAUNCH(0, false);
        }
#undef DEC_LAUNCH
#undef DEC_LAUNCH_LONE
    }
    mark(D->events, FFK_DECODE + 1, st);
    if (!launch_ok()) return FFGPU_EXTERNAL;
    return 3;
}

/* ---------------- concealment ---------------- */
__global__ void k_conceal(const FFDevParams P, uint8_t *dst, const uint8_t *src,
                          int x, int y, int w, int h, int pixshift, int nplanes, int planar_chroma)
{
    for (int p = 0; p < nplanes; p++) {
        const int sh = (planar_chroma && (p == 1 || p == 2)) ? P.hs : 0;
        const int sv = (planar_chroma && (p == 1 || p == 2)) ? P.vs : 0;
        const int bpp = P.layout == FF_LAY_PLANAR ? (P.sbits > 8 ? 2 : 1) :
                        P.layout == FF_LAY_GBRP ? 2 : P.layout == FF_LAY_YA8 ? 2 : P.rgb_pixbytes;
        const int bw = ff_crshift(w, sh) * bpp, rows = ff_crshift(h, sv);
        const size_t xo = (size_t)((x >> sh) << pixshift);
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < bw * rows; i += gridDim.x * blockDim.x) {
            const int r = i / bw, c = i - r * bw;
            const size_t o = P.plane_off[p] + (size_t)((y >> sv) + r) * P.pitch[p] + xo + c;
            dst[o] = src[o];
        }
    }
}

extern "C" int ffk_conceal_rect(const FFDevParams *P, uint8_t *dst_frame, const uint8_t *src_frame,
                                int x, int y, int w, int h, int depth_gt8, ffk_stream stream)
{
    int nplanes = 0;
    for (int p = 0; p < FF_MAX_PLANES; p++)
        if (P->rows[p])
            nplanes = p + 1;
    k_conceal<<<64, 256, 0, (cudaStream_t)stream>>>(*P, dst_frame, src_frame, x, y, w, h, depth_gt8,
                                                    nplanes, P->layout == FF_LAY_PLANAR && P->chroma_planes);
    if (!launch_ok()) return FFGPU_EXTERNAL;
    return 1;
}

/* ---------------- SM-driven copies between mapped host memory and HBM ---------------- */
/* Bitstreams, work items and result tables are small next to the pictures, but as copy-engine
 * transfers they queue behind the picture traffic of the other launch groups (≈90 ms per group
 * of 4K pictures) and stall the kernels that wait for them.  Moved by the SMs through the
 * mapped pinned buffers they bypass that queue.  All pointers are 16-byte aligned. */
__global__ void __launch_bounds__(256)
k_copy_segments(const FFCopyArgs a)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    for (int s = 0; s < a.nseg; s++) {
        size_t bytes = a.seg[s].bytes;
        if (s == 0 && a.dyn_bytes) {
            bytes = (size_t)*a.dyn_bytes << a.dyn_shift;
            if (bytes > a.dyn_cap)
                bytes = 0;                       /* does not fit: the host falls back to a plain copy */
        }
        const uint4 *src = (const uint4 *)a.seg[s].src;
        uint4 *dst = (uint4 *)a.seg[s].dst;
        const size_t nvec = bytes >> 4;
        for (size_t i = tid; i < nvec; i += nthreads)
            dst[i] = src[i];
        const size_t tail = bytes & 15;
        if (tid < tail)
            ((uint8_t *)a.seg[s].dst)[(nvec << 4) + tid] = ((const uint8_t *)a.seg[s].src)[(nvec << 4) + tid];
    }
}

extern "C" int ffk_copy_segments(const FFCopyArgs *a, ffk_stream stream)
{
    size_t most = a->dyn_bytes ? a->dyn_cap : 0;
    for (int s = 0; s < a->nseg; s++)
        if (a->seg[s].bytes > most)
            most = a->seg[s].bytes;
    size_t blocks = (most / 16 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_copy_segments<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*a);
    if (!launch_ok()) return FFGPU_EXTERNAL;
    return 1;
}
