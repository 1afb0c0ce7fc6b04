/*
 * ffv1_slice.cuh -- the FFV1 slice pixel path as __host__ __device__ functions.
 *
 * The CUDA kernels in ffv1_kernels.cu are thin wrappers that map CUDA threads onto
 * these functions; tests/emul compiles the very same functions with g++ so the
 * device logic can be checked against the oracle on a machine without a GPU.
 * (The emulation build is test infrastructure; the product never runs it.)
 *
 * Encoder, two stages:
 *   A  symbolize   one thread per sample: load (+RCT), neighbourhood, context,
 *                  median prediction, fold -> 32-bit token (residual:17 | context:15)
 *                  ffv1enc.c:274-312 encode_plane, ffv1enc_template.c:125-201
 *                  encode_rgb_frame, ffv1_template.c:23-52 predict/get_context
 *   B  code        one thread per slice: adaptive range coder / Golomb-Rice coder over
 *                  the slice's token stream, ffv1enc_template.c:23-123 encode_line,
 *                  ffv1enc.c:185-262 put_symbol_inline / put_vlc_symbol
 * Decoder: one thread per slice, ffv1dec.c:119-165 decode_plane,
 *   ffv1dec_template.c:23-193 decode_line / decode_rgb_frame.
 */
#ifndef FFGPU_FFV1_SLICE_CUH
#define FFGPU_FFV1_SLICE_CUH

#include "ffv1_types.h"

#define FF_TOKEN_CTX_BITS 15
#define FF_TOKEN_CTX_MASK 0x7FFF

#if !defined(__CUDACC__) && !defined(__VECTOR_TYPES_H__)
typedef struct uint2 { uint32_t x, y; } uint2;    /* CPU emulation builds without the CUDA headers */
#endif

/* ff_log2_run[41], libavcodec/bitstream.c:39-46: 0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,
 * 4,4,5,5,6,6,7,7,8,9,...,24 -- four entries per step, then two, then one */
FFGPU_HD int ff_run_log2(int i)
{
    return i < 16 ? i >> 2 : (i < 24 ? 4 + ((i - 16) >> 1) : i - 16);
}

/* a quant table set: int16 [5][256] followed by the "uses 5 inputs" flag */
#define FF_QT_STRIDE (FF_MAX_CTX_INPUTS * 256 + 8)

FFGPU_HD int ff_min(int a, int b) { return a < b ? a : b; }
FFGPU_HD int ff_max(int a, int b) { return a > b ? a : b; }
FFGPU_HD int ff_crshift(int a, int b) { return -((-a) >> b); }

/* mid_pred, mathops.h:98-112 */
FFGPU_HD int ff_median3(int a, int b, int c)
{
    return ff_max(ff_min(a, b), ff_min(ff_max(a, b), c));
}

/* fold, ffv1.h:151-160: sign-extend from `bits` bits */
FFGPU_HD int ff_fold(int diff, int bits)
{
    const int sh = 32 - bits;
    return (int)((uint32_t)diff << sh) >> sh;
}

/* ------------------------------------------------------------------ */
/* stage A: sample access                                               */
/* ------------------------------------------------------------------ */

FFGPU_HD int ff_rd16(const uint8_t *p)
{
    return p[0] | (p[1] << 8);
}

/* coded sample of plane k at absolute position (X,Y) of that plane's grid, already wrapped
 * to the coder's sample type (int16 unless use32).  encode_plane ffv1enc.c:291-305,
 * encode_rgb_frame ffv1enc_template.c:150-186. */
FFGPU_HD int ff_coded_sample(const FFDevParams &P, const uint8_t *frame, int k, int X, int Y,
                             int rct_by = 1, int rct_ry = 1)
{
    if (P.colorspace == 0) {
        const FFDevPlane cp = P.cp[k];
        const uint8_t *p = frame + P.plane_off[cp.mem] + (size_t)Y * P.pitch[cp.mem] +
                           (size_t)X * cp.step + cp.off;
        int v;
        if (P.sbits <= 8)
            v = p[0];
        else if (P.packed_lsb)
            v = ff_rd16(p);
        else
            v = ff_rd16(p) >> (16 - P.sbits);
        return (int16_t)v;
    } else {
        int r, g, b, a = 0, v;
        if (P.layout == FF_LAY_BGR32) {
            const uint8_t *p = frame + P.plane_off[0] + (size_t)Y * P.pitch[0] + (size_t)X * 4;
            b = p[0]; g = p[1]; r = p[2]; a = p[3];
        } else if (P.layout == FF_LAY_RGB48) {
            const uint8_t *p = frame + P.plane_off[0] + (size_t)Y * P.pitch[0] +
                               (size_t)X * P.rgb_pixbytes;
            r = ff_rd16(p); g = ff_rd16(p + 2); b = ff_rd16(p + 4);
            if (P.transparency)
                a = ff_rd16(p + 6);
        } else {
            const int p0 = ff_rd16(frame + P.plane_off[0] + (size_t)Y * P.pitch[0] + (size_t)X * 2);
            const int p1 = ff_rd16(frame + P.plane_off[1] + (size_t)Y * P.pitch[1] + (size_t)X * 2);
            r = ff_rd16(frame + P.plane_off[2] + (size_t)Y * P.pitch[2] + (size_t)X * 2);
            if (P.use32 || P.transparency) {
                g = p0; b = p1;                 /* ffv1enc_template.c:163-168 */
            } else {
                b = p0; g = p1;                 /* ffv1enc_template.c:169-173 */
            }
            if (P.transparency)
                a = ff_rd16(frame + P.plane_off[3] + (size_t)Y * P.pitch[3] + (size_t)X * 2);
        }
        if (k == 3) {
            v = a;
        } else {
            b -= g;
            r -= g;
            g += (b * rct_by + r * rct_ry) >> 2;   /* slice_rct_by/ry_coef: 1, 1 before version 4 */
            v = k == 0 ? g : (k == 1 ? b : r) + (1 << P.sbits);
        }
        return P.use32 ? v : (int)(int16_t)v;
    }
}

/* The reference's rotating line buffers (ffv1enc.c:284-289), unrolled: for sample (x,y)
 * of a w-wide plane whose top-left is (X0,Y0),
 *   T  = S(x,y-1)                        0 above the slice
 *   L  = S(x-1,y)    ; x==0   -> T
 *   LT = S(x-1,y-1)  ; x==0   -> S(0,y-2)
 *   RT = S(x+1,y-1)  ; x==w-1 -> T
 *   LL = S(x-2,y)    ; x==1   -> S(0,y-1) ; x==0 -> 0
 *   TT = S(x,y-2)
 * Neighbours never cross the slice border.  Returns the token. */
FFGPU_HD uint32_t ff_symbolize_sample(const FFDevParams &P, const uint8_t *frame,
                                      const int16_t *qt, int k, int X0, int Y0, int w,
                                      int x, int y, int ctx_base, int rct_by = 1, int rct_ry = 1)
{
#define S_(xx, yy) ((yy) < 0 ? 0 : ff_coded_sample(P, frame, k, X0 + (xx), Y0 + (yy), rct_by, rct_ry))
    const int cur = S_(x, y);
    const int T   = S_(x, y - 1);
    const int L   = x ? S_(x - 1, y) : T;
    const int LT  = x ? S_(x - 1, y - 1) : S_(0, y - 2);
    const int RT  = (x + 1 < w) ? S_(x + 1, y - 1) : T;
    int ctx = qt[(L - LT) & 0xFF] + qt[256 + ((LT - T) & 0xFF)] + qt[512 + ((T - RT) & 0xFF)];
    int diff;
    if (qt[FF_MAX_CTX_INPUTS * 256]) {      /* quant_table[3][127] || quant_table[4][127] */
        const int LL = x >= 2 ? S_(x - 2, y) : (x == 1 ? S_(0, y - 1) : 0);
        const int TT = S_(x, y - 2);
        ctx += qt[768 + ((LL - L) & 0xFF)] + qt[1024 + ((TT - T) & 0xFF)];
    }
#undef S_
    diff = cur - ff_median3(L, L + T - LT, T);
    if (ctx < 0) {
        ctx = -ctx;
        diff = -diff;
    }
    diff = ff_fold(diff, P.cbits);
    return ((uint32_t)diff << FF_TOKEN_CTX_BITS) | (uint32_t)(ctx_base + ctx);
}

/* choose_rct_params, ffv1enc.c:963-1043 (version 4): for every pixel with x >= 1 and y >= 1 of
 * the slice the second difference (horizontal, then vertical) of the three components, and
 * for each of 15 candidate coefficient pairs the magnitude the luma-like component would
 * have.  The reference walks the slice with a line buffer; the differences only involve the
 * pixel and its left, upper and upper-left neighbours, so every pixel is independent here.
 * NOTE the component naming is the reference's: for the planar layouts it calls plane 0
 * "b", plane 1 "g" and plane 2 "r" (:1000-1002) whatever the planes hold.
 * Returns the 15 magnitudes of pixel (x, y) of the slice at (X0, Y0). */
#define FF_RCT_CANDIDATES 15
FFGPU_HD void ff_rct_candidates(int i, int *ry, int *by)
{
    /* rct_y_coeff[15][2] of ffv1enc.c:966-984, {ry, by} */
    const int t[FF_RCT_CANDIDATES][2] = { {0, 0}, {1, 1}, {2, 2}, {0, 2}, {2, 0}, {4, 0}, {0, 4}, {0, 3},
                                          {3, 0}, {3, 1}, {1, 3}, {1, 2}, {2, 1}, {0, 1}, {1, 0} };
    *ry = t[i][0];
    *by = t[i][1];
}

FFGPU_HD void ff_rct_load_bgr(const FFDevParams &P, const uint8_t *frame, int X, int Y, int *b, int *g, int *r)
{
    if (P.layout == FF_LAY_BGR32) {
        const uint8_t *p = frame + P.plane_off[0] + (size_t)Y * P.pitch[0] + (size_t)X * 4;
        *b = p[0]; *g = p[1]; *r = p[2];
    } else {
        *b = ff_rd16(frame + P.plane_off[0] + (size_t)Y * P.pitch[0] + (size_t)X * 2);
        *g = ff_rd16(frame + P.plane_off[1] + (size_t)Y * P.pitch[1] + (size_t)X * 2);
        *r = ff_rd16(frame + P.plane_off[2] + (size_t)Y * P.pitch[2] + (size_t)X * 2);
    }
}

FFGPU_HD void ff_rct_pixel_stat(const FFDevParams &P, const uint8_t *frame, int X0, int Y0, int x, int y,
                                int32_t stat[FF_RCT_CANDIDATES])
{
    int b[4], g[4], r[4];                  /* (x,y) (x-1,y) (x,y-1) (x-1,y-1) */
    int bg, bb, br;
    /* the left neighbour of x == 1 is pixel 0 of the line: lastr/lastg/lastb only restart
     * from 0 for x == 0 itself (:990), which contributes nothing */
    ff_rct_load_bgr(P, frame, X0 + x, Y0 + y, &b[0], &g[0], &r[0]);
    ff_rct_load_bgr(P, frame, X0 + x - 1, Y0 + y, &b[1], &g[1], &r[1]);
    ff_rct_load_bgr(P, frame, X0 + x, Y0 + y - 1, &b[2], &g[2], &r[2]);
    ff_rct_load_bgr(P, frame, X0 + x - 1, Y0 + y - 1, &b[3], &g[3], &r[3]);
    /* the previous line's differences went through the int16_t line buffer (:987, :1023-1025):
     * they wrap for 16-bit samples */
    bg = (g[0] - g[1]) - (int)(int16_t)(g[2] - g[3]);
    bb = (b[0] - b[1]) - (int)(int16_t)(b[2] - b[3]);
    br = (r[0] - r[1]) - (int)(int16_t)(r[2] - r[3]);
    br -= bg;
    bb -= bg;
    for (int i = 0; i < FF_RCT_CANDIDATES; i++) {
        int cy, cb, v;
        ff_rct_candidates(i, &cy, &cb);
        v = bg + ((br * cy + bb * cb) >> 2);
        stat[i] = v < 0 ? -v : v;
    }
}

/* the winner: the first candidate with the smallest sum (the sums are `int` in the
 * reference and wrap the same way here) */
FFGPU_HD void ff_rct_pick(const int32_t stat[FF_RCT_CANDIDATES], int *by, int *ry)
{
    int best = 0;
    for (int i = 1; i < FF_RCT_CANDIDATES; i++)
        if (stat[i] < stat[best])
            best = i;
    ff_rct_candidates(best, ry, by);
}

/* token index -> (plane, x, y) and the token itself; tokens are laid out in CODING order:
 * YCbCr plane after plane, RGB line-interleaved G,B,R[,A] (ffv1enc_template.c:188-198) */
FFGPU_HD uint32_t ff_symbolize_index(const FFDevParams &P, const FFDevSlice &sl,
                                     const uint8_t *frame, const int16_t *qt_all, uint32_t idx,
                                     int rct_by = 1, int rct_ry = 1)
{
    int k, x, y, w;
    if (P.colorspace == 0) {
        uint32_t base = 0;
        for (k = 0; k < P.ncoded - 1; k++) {
            const uint32_t n = (uint32_t)sl.seg_w[k] * sl.seg_lines[k];
            if (idx < base + n)
                break;
            base += n;
        }
        w = sl.seg_w[k];
        y = (int)((idx - base) / (uint32_t)w);
        x = (int)((idx - base) - (uint32_t)y * w);
    } else {
        const uint32_t line = idx / (uint32_t)sl.w;
        w = sl.w;
        x = (int)(idx - line * (uint32_t)w);
        y = (int)(line / (uint32_t)P.ncoded);
        k = (int)(line - (uint32_t)y * P.ncoded);
    }
    {
        const FFDevPlane cp = P.cp[k];
        const int16_t *qt = qt_all + (size_t)P.set_qidx[cp.set] * FF_QT_STRIDE;
        return ff_symbolize_sample(P, frame, qt, k, sl.x >> cp.hs, sl.y >> cp.vs, w, x, y,
                                   P.set_base[cp.set], rct_by, rct_ry);
    }
}

/* ------------------------------------------------------------------ */
/* stage B: range coder over the token stream                           */
/* ------------------------------------------------------------------ */

FFGPU_HD void ff_enc_resume(FFRacEnc *c, const FFRacPrefix &pre, const uint8_t *pre_bytes,
                            uint8_t *out, uint32_t cap)
{
    uint32_t i;
    ffrac_enc_init(c, out, cap);
    for (i = 0; i < pre.nbytes && i < cap; i++)
        out[i] = pre_bytes[pre.byte_off + i];
    c->pos = pre.nbytes;
    c->low = pre.low;
    c->range = pre.range;
    c->pending = pre.pending;
    c->run = pre.run;
}

/* The part of a version 4 slice header that depends on the picture (ffv1enc.c:951-959):
 * slice_coding_mode == 1 flag, the coding mode (always 0: PCM slices are the reference's
 * answer to an overflowing packet buffer, which this encoder reports as an error instead)
 * and the two RCT coefficients, coded with the header's adaptive states the host handed
 * over.  Golomb-Rice slices then close the coder (ffv1enc.c:1076-1081); returns where the
 * Rice bits start in that case. */
FFGPU_HD uint32_t ff_enc_v4_header_tail(FFRacEnc *c, const FFRacTables *t, const FFRacPrefix &pre,
                                        int rct_by, int rct_ry, int golomb)
{
    uint8_t st[FF_CONTEXT_SIZE];
    for (int i = 0; i < FF_CONTEXT_SIZE; i++)
        st[i] = pre.hdr_state[i];
    ffrac_put(c, t, st, 0);
    ffrac_put_symbol(c, t, st, 0, 0);
    ffrac_put_symbol(c, t, st, rct_by, 0);
    ffrac_put_symbol(c, t, st, rct_ry, 0);
    return golomb ? ffrac_enc_finish(c, t, 1) : 0;
}

/* ---- per-thread cache of the current context's 32 adaptive state bytes ----
 * The row lives in shared memory (9-word stride: conflict-free when all lanes touch the
 * same slot) so that the per-decision state read/update is an LDS/STS instead of a global
 * round trip; it is written back only when the context changes. */
#define FF_ROW_WORDS 9
#define FF_NEW_WAIT 24          /* iterations a lane may wait for the rest of its warp before the (divergent,
                                 * expensive) per-sample set-up runs; measured optimum on B200 */

typedef struct
#if defined(__CUDACC__)
__align__(16)
#else
__attribute__((aligned(16)))
#endif
FFU128 { uint32_t x, y, z, w; } FFU128;

#if defined(__CUDACC__)
__device__ __forceinline__ void ff_cp_async16(void *smem, const void *gmem)
{
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void ff_cp_async_commit(void)
{
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void ff_cp_async_wait(void)
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
#endif
#define FF_CODE_THREADS 128      /* threads per block of the slice-coder kernels */

#if defined(__CUDACC__)
/* statically named shared memory, so that the device code below addresses it with LDS/STS
 * and immediate offsets instead of generic pointers */
static __shared__ FFRacTables ff_s_tab;
static __shared__ uint32_t ff_s_rows[FF_CODE_THREADS * FF_ROW_WORDS];
extern __shared__ int16_t ff_s_qt[];             /* decoder: qt_count quant table sets */
/* stage B token stream: per lane a ring of 16-byte chunks filled by cp.async (LDGSTS) six
 * chunks ahead of the read position, so a lane's DRAM/L2 miss never stalls its warp */
#define FF_TOK_CHUNKS 8
#define FF_TOK_AHEAD  6
static __shared__ FFU128 ff_s_tok[FF_TOK_CHUNKS * FF_CODE_THREADS];
#endif
#if defined(__CUDACC__)
/* both successors of an adaptive state in one 16-bit entry: one_state | zero_state << 8 */
static __shared__ uint16_t ff_s_tab16[256];
__device__ __forceinline__ void ff_fill_tab16(int tid, int nthreads)
{
    for (int i = tid; i < 256; i += nthreads)
        ff_s_tab16[i] = (uint16_t)(ff_s_tab.one[i] | (ff_s_tab.zero[i] << 8));
}
#endif
#if defined(__CUDA_ARCH__)
#define FF_TAB(i)   (((const uint8_t *)&ff_s_tab)[i])
#define FF_ROWB(i)  (((uint8_t *)ff_s_rows)[threadIdx.x * (FF_ROW_WORDS * 4) + (i)])
#define FF_ROWW     (&ff_s_rows[threadIdx.x * FF_ROW_WORDS])
#define FF_QT(set_off, i) (ff_s_qt[(set_off) + (i)])
#else
#define FF_TAB(i)   (((const uint8_t *)tab_)[i])
#define FF_ROWB(i)  (((uint8_t *)row_)[i])
#define FF_ROWW     (row_)
#define FF_QT(set_off, i) (qt_all_[(set_off) + (i)])
#endif



FFGPU_HD void ff_row_load(uint32_t *row, const uint8_t *g)
{
    const FFU128 a = ((const FFU128 *)g)[0], b = ((const FFU128 *)g)[1];
    row[0] = a.x; row[1] = a.y; row[2] = a.z; row[3] = a.w;
    row[4] = b.x; row[5] = b.y; row[6] = b.z; row[7] = b.w;
}

FFGPU_HD void ff_row_store(const uint32_t *row, uint8_t *g)
{
    FFU128 a, b;
    a.x = row[0]; a.y = row[1]; a.z = row[2]; a.w = row[3];
    b.x = row[4]; b.y = row[5]; b.z = row[6]; b.w = row[7];
    ((FFU128 *)g)[0] = a;
    ((FFU128 *)g)[1] = b;
}

/* one renormalisation step of renorm_encoder (rangecoder.h:71-94); in the symbol loop the
 * range can drop below 0x100 at most once per decision */
FFGPU_HD void ffrac_enc_shift1(FFRacEnc *c)
{
    if (c->pending < 0) {
        c->pending = c->low >> 8;
    } else if (c->low <= 0xFF00) {
        ffrac_emit(c, c->pending);
        for (; c->run; c->run--)
            ffrac_emit(c, 0xFF);
        c->pending = c->low >> 8;
    } else if (c->low >= 0x10000) {
        ffrac_emit(c, c->pending + 1);
        for (; c->run; c->run--)
            ffrac_emit(c, 0x00);
        c->pending = (c->low >> 8) & 0xFF;
    } else {
        c->run++;
    }
    c->low = (c->low & 0xFF) << 8;
    c->range <<= 8;
}

/* state slot of decision `step` of a residual with exponent e (put_symbol_inline,
 * ffv1enc.c:185-231): 0 zero flag | 1..e+1 unary exponent | e+2..2e+1 mantissa | 2e+2 sign */
#define FF_STAB_STRIDE 40
#define FF_STAB_ROWS   19
FFGPU_HD int ff_slot_of(int e, int step)
{
    if (step == 0)
        return 0;
    if (step <= e + 1)
        return 1 + ff_min(step - 1, 9);
    if (step <= 2 * e + 1)
        return 22 + ff_min(2 * e + 1 - step, 9);
    return 11 + ff_min(e, 10);
}

#if defined(__CUDACC__)
static __shared__ uint8_t ff_s_stab[FF_STAB_ROWS * FF_STAB_STRIDE];
/* shared-memory byte access through an explicit 32-bit shared address: one LDS/STS each,
 * no generic-address arithmetic in the decision loop */
__device__ __forceinline__ uint32_t ff_lds8(uint32_t sa)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(sa));
    return v;
}
/* identity the assembler cannot see through (must be called by a converged warp) */
__device__ __forceinline__ uint32_t ff_opaque(uint32_t v)
{
    return __shfl_sync(__activemask(), v, (int)(threadIdx.x & 31));
}
__device__ __forceinline__ uint32_t ff_lds16u(uint32_t sa)
{
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(sa));
    return v;
}
/* range decoder refill (rangecoder.h:123-139), out of line: in the loop it would be turned
 * into a dozen predicated instructions that every decision issues and whose scoreboard wait
 * on the prefetched byte every decision pays.  in: low, pos, overread, nbyte; out likewise */
static __device__ __noinline__ uint4 ff_dec_refill(int low, uint32_t pos, int overread, uint32_t nbyte,
                                            const uint8_t *buf, uint32_t end)
{
    const int in = pos < end;
    uint4 r;
    r.x = (uint32_t)((low << 8) + (in ? (int)nbyte : 0));
    r.y = pos + (uint32_t)in;
    r.z = (uint32_t)(overread + !in);
    r.w = buf[r.y];                              /* the arena is padded past `end` */
    return r;
}
__device__ __forceinline__ int ff_lds16s(uint32_t sa)
{
    int v;
    asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(sa));
    return v;
}
__device__ __forceinline__ void ff_sts8(uint32_t sa, uint32_t v)
{
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(sa), "r"(v) : "memory");
}
#endif

#if defined(__CUDA_ARCH__)
#define FF_UNLIKELY(c) __builtin_expect(!!(c), 0)
#else
#define FF_UNLIKELY(c) (c)
#endif

/* Stage B, range coder.  The 32 lanes of a warp code 32 different slices, so the loop is
 * written as ONE BINARY DECISION PER ITERATION with a small per-lane state machine
 * (put_symbol_inline, ffv1enc.c:185-231, unrolled over `step`): a lane that is inside a
 * large residual (2e+3 decisions) does not hold up lanes that only code zero flags.
 *   step 0           zero flag                 state[0]
 *   1..e             unary exponent ones       state[1 + min(i,9)]
 *   e+1              unary terminator          state[1 + min(e,9)]
 *   e+2..2e+1        mantissa, MSB first       state[22 + min(i,9)]
 *   2e+2             sign                      state[11 + min(e,10)]
 * returns the slice's byte count; *overflow != 0 if the arena was too small */
/* first pass (AV_CODEC_FLAG_PASS1): the counters put_symbol_inline keeps, ffv1enc.c:193-199 */
typedef struct FFPassStats {
    unsigned long long *rc_stat;    /* [256][2]: decisions by adaptive state value and bit          */
    unsigned long long *rc_stat2;   /* [contexts of the table][32][2]: by context, state slot, bit  */
    int ctx_count;                  /* contexts of the encoder's quant table (sets share counters)  */
} FFPassStats;

FFGPU_HD void ff_pass_count(const FFPassStats *st, int state, int ctx, int slot, int bit)
{
    unsigned long long *a = st->rc_stat + 2 * state + bit;
    unsigned long long *b = st->rc_stat2 + ((size_t)(ctx % st->ctx_count) * 32 + slot) * 2 + bit;
#if defined(__CUDA_ARCH__)
    atomicAdd(a, 1ull);
    atomicAdd(b, 1ull);
#else
    (*a)++;
    (*b)++;
#endif
}

template <bool STATS>               /* STATS: a first pass, count every decision */
FFGPU_HD uint32_t ff_encode_slice_range(const FFDevSlice &sl, const uint32_t *tokens,
                                        uint8_t *state, const FFRacTables *tab_,
                                        const FFRacPrefix &pre, const uint8_t *pre_bytes,
                                        uint8_t *out, uint32_t *overflow, uint32_t *row_,
                                        const int *rct = 0,     /* version 4: {by, ry} of the slice */
                                        uint32_t v4_room = 0,   /* version 4: the reference's slice buffer */
                                        const FFPassStats *pass = 0)  /* first pass: count the decisions */
{
    FFRacEnc c;
    const uint32_t n = sl.ntok;
    /* Version 4 gives every slice (16384 + 12 * width * height) / slice_count bytes
     * (ffv1enc.c:1180-1181, :1221-1224) and encode_line gives up when fewer than 35 * w are
     * left at the start of a line (ffv1enc_template.c:33-37); the reference then rewrites the
     * slice as raw "PCM" bits (:1107-1117).  The byte count only grows, so the guard fires
     * iff it fires at the start of the LAST line: remember the count there.  This encoder
     * does not write PCM slices; it reports "encoded frame too large" where the reference
     * would have written one, instead of silently producing a different packet. */
    const uint32_t guard_tok = v4_room ? n - (uint32_t)sl.seg_w[sl.nseg - 1] : 0xFFFFFFFFu;
    uint32_t guard_pos = 0;
    uint32_t i = 0, nb;
    int cur_ctx = -1;
    int e = 0, step = 0, nsteps = 0;
    uint64_t seq = 0;                                /* bit k = value of decision k of the residual */
    (void)row_;

#if defined(__CUDA_ARCH__)
    uint32_t row_sa = (uint32_t)__cvta_generic_to_shared(&ff_s_rows[threadIdx.x * FF_ROW_WORDS]);
    uint32_t tab_sa = (uint32_t)__cvta_generic_to_shared(&ff_s_tab);
    uint32_t stab_sa = (uint32_t)__cvta_generic_to_shared(ff_s_stab);
    /* a shuffle from the own lane is opaque to ptxas: otherwise it rematerialises the
     * shared-window base (S2R CgaCtaId + shifts, a dozen instructions) in front of every
     * decision instead of keeping three registers */
    row_sa = ff_opaque(row_sa);
    tab_sa = ff_opaque(tab_sa);
    stab_sa = ff_opaque(stab_sa);
    uint32_t srow = stab_sa, slot = 0;
    const uint32_t nchunks = (n + 3) >> 2;
    for (uint32_t ch = 0; ch < FF_TOK_AHEAD; ch++) {
        if (ch < nchunks)
            ff_cp_async16(&ff_s_tok[ch * FF_CODE_THREADS + threadIdx.x], tokens + 4 * ch);
        ff_cp_async_commit();
    }
#endif
    ff_enc_resume(&c, pre, pre_bytes, out, sl.bs_cap);
    if (rct)
        ff_enc_v4_header_tail(&c, tab_, pre, rct[0], rct[1], 0);
    for (;;) {
        int bit, s, r1, rb;
        if (step == nsteps) {                        /* fetch the next residual */
            uint32_t tok, a, neg;
            int ctx, diff;
            if (i == n)
                break;
            if (i == guard_tok)
                guard_pos = c.pos;
#if defined(__CUDA_ARCH__)
            if ((i & 3) == 0) {                      /* entering chunk i/4: issue chunk i/4 + AHEAD */
                const uint32_t ch = (i >> 2) + FF_TOK_AHEAD;
                if (ch < nchunks)
                    ff_cp_async16(&ff_s_tok[(ch & (FF_TOK_CHUNKS - 1)) * FF_CODE_THREADS + threadIdx.x],
                                  tokens + 4 * ch);
                ff_cp_async_commit();
                ff_cp_async_wait<FF_TOK_AHEAD>();
            }
            tok = ((const uint32_t *)&ff_s_tok[((i >> 2) & (FF_TOK_CHUNKS - 1)) * FF_CODE_THREADS +
                                               threadIdx.x])[i & 3];
#else
            tok = tokens[i];
#endif
            i++;
            ctx = (int)(tok & FF_TOKEN_CTX_MASK);
            diff = (int32_t)tok >> FF_TOKEN_CTX_BITS;
            if (ctx != cur_ctx) {
                if (cur_ctx >= 0)
                    ff_row_store(FF_ROWW, state + (size_t)cur_ctx * FF_CONTEXT_SIZE);
                ff_row_load(FF_ROWW, state + (size_t)ctx * FF_CONTEXT_SIZE);
                cur_ctx = ctx;
            }
            a = (uint32_t)(diff < 0 ? -diff : diff);
            neg = diff < 0;
            e = ffrac_ilog2(a);
            /* the residual's decisions as a bit string, first decision in bit 0:
             * zero flag, e ones, a zero, the e mantissa bits MSB first, the sign */
            if (a) {
                uint32_t m = a & ((1u << e) - 1), rev = 0;
#if defined(__CUDA_ARCH__)
                rev = e ? __brev(m) >> (32 - e) : 0; /* mantissa reversed: MSB goes out first */
#else
                for (int k = 0; k < e; k++)
                    rev |= ((m >> k) & 1u) << (e - 1 - k);
#endif
                seq = ((uint64_t)((1u << e) - 1) << 1) | ((uint64_t)rev << (e + 2)) |
                      ((uint64_t)neg << (2 * e + 2));
                nsteps = 2 * e + 3;
            } else {
                seq = 1;
                nsteps = 1;
            }
            step = 0;
#if defined(__CUDA_ARCH__)
            srow = stab_sa + (uint32_t)e * FF_STAB_STRIDE;
            slot = 0;
#endif
        }
        bit = (int)(seq & 1);
        seq >>= 1;
#if defined(__CUDA_ARCH__)
        {
            const uint32_t sa = row_sa + slot;
            s = (int)ff_lds8(sa);
            if (STATS)
                ff_pass_count(pass, s, cur_ctx, (int)slot, bit);
            r1 = (c.range * s) >> 8;                 /* put_rac, rangecoder.h:104-121 */
            rb = c.range - r1;
            ff_sts8(sa, ff_lds8(tab_sa + (uint32_t)s + (bit ? 0u : 256u)));
            step++;
            slot = ff_lds8(srow + (uint32_t)step);   /* slot of the next decision */
        }
#else
        {
            const int slot = ff_slot_of(e, step);
            s = FF_ROWB(slot);
            if (STATS)
                ff_pass_count(pass, s, cur_ctx, slot, bit);
            r1 = (c.range * s) >> 8;
            rb = c.range - r1;
            FF_ROWB(slot) = FF_TAB(s + (bit ? 0 : 256));
            step++;
        }
#endif
        c.low += bit ? rb : 0;
        c.range = bit ? r1 : rb;
        if (c.range < 0x100)
            ffrac_enc_shift1(&c);
    }
    if (cur_ctx >= 0)
        ff_row_store(FF_ROWW, state + (size_t)cur_ctx * FF_CONTEXT_SIZE);
    nb = ffrac_enc_finish(&c, tab_, 1);              /* ffv1enc.c:1242 */
    *overflow = c.overflow;
    if (v4_room && (int64_t)v4_room - (int64_t)guard_pos < (int64_t)sl.seg_w[sl.nseg - 1] * 35)
        *overflow = 1;
    return nb;
}

#if defined(__CUDACC__)
/* renorm_encoder step out of line, values in and out through registers: inlined at every
 * decision site it makes the lone coder's loop ten times larger than the instruction caches
 * (L0 6 KB, L1.5 32 KB), and a lone warp waits out every instruction fetch */
typedef struct FFEncRenorm {
    int32_t low, range, pending, run;
    uint32_t pos, overflow;
} FFEncRenorm;
static __device__ __noinline__ FFEncRenorm ff_enc_renorm(int32_t low, int32_t range, int32_t pending, int32_t run,
                                                         uint32_t pos, uint32_t overflow, uint8_t *buf, uint32_t cap)
{
    FFRacEnc c;
    FFEncRenorm r;
    c.low = low; c.range = range; c.pending = pending; c.run = run;
    c.buf = buf; c.pos = pos; c.cap = cap; c.overflow = overflow;
    ffrac_enc_shift1(&c);
    r.low = c.low; r.range = c.range; r.pending = c.pending; r.run = c.run;
    r.pos = c.pos; r.overflow = c.overflow;
    return r;
}
#endif

/* Stage B for launches that give every slice a warp of its own (E.lane_stride == 32: streams
 * with a handful of large slices).  One live lane per warp and a few warps per SM: nothing
 * diverges and nothing hides latency, the launch lasts as long as the lone lane of the
 * heaviest slice needs.  So here put_symbol_inline (ffv1enc.c:185-231) is written the
 * straight way, as nested loops over the decisions of a residual with the bit values known at
 * compile time where the symbol layout fixes them, slot 0 of the current row (the zero flag)
 * in a register, the next token fetched while the current one is coded, and the
 * renormalisation out of line (the loop has to stay inside the instruction caches: a lone
 * warp waits out every instruction fetch).  Same packets as
 * ff_encode_slice_range, byte for byte; no first-pass counters (those launches take the
 * other form). */
FFGPU_HD uint32_t ff_encode_slice_range_lone(const FFDevSlice &sl, const uint32_t *tokens,
                                             uint8_t *state, const FFRacTables *tab_,
                                             const FFRacPrefix &pre, const uint8_t *pre_bytes,
                                             uint8_t *out, uint32_t *overflow, uint32_t *row_,
                                             const int *rct = 0, uint32_t v4_room = 0)
{
    FFRacEnc c;
    const uint32_t n = sl.ntok;
    const uint32_t guard_tok = v4_room ? n - (uint32_t)sl.seg_w[sl.nseg - 1] : 0xFFFFFFFFu;  /* see above */
    uint32_t guard_pos = 0, nb;
    int cur_ctx = -1;
    uint32_t s0 = 128;
    (void)row_;
#if defined(__CUDA_ARCH__)
    uint32_t row_sa = (uint32_t)__cvta_generic_to_shared(&ff_s_rows[threadIdx.x * FF_ROW_WORDS]);
    uint32_t tab_sa = (uint32_t)__cvta_generic_to_shared(&ff_s_tab);
    row_sa = ff_opaque(row_sa);
    tab_sa = ff_opaque(tab_sa);
#define FF_E_LD(slot) ff_lds8(row_sa + (uint32_t)(slot))
#define FF_E_ST(slot, v) ff_sts8(row_sa + (uint32_t)(slot), (v))
#define FF_E_NEXT(s, bit) ff_lds8(tab_sa + (uint32_t)(s) + ((bit) ? 0u : 256u))
#define FF_E_RENORM()                                                                         \
    do {                                                                                      \
        const FFEncRenorm rn_ = ff_enc_renorm(c.low, c.range, c.pending, c.run, c.pos,        \
                                              c.overflow, c.buf, c.cap);                      \
        c.low = rn_.low; c.range = rn_.range; c.pending = rn_.pending; c.run = rn_.run;       \
        c.pos = rn_.pos; c.overflow = rn_.overflow;                                           \
    } while (0)
#else
#define FF_E_RENORM() ffrac_enc_shift1(&c)
#define FF_E_LD(slot) ((uint32_t)FF_ROWB(slot))
#define FF_E_ST(slot, v) (FF_ROWB(slot) = (uint8_t)(v))
#define FF_E_NEXT(s, bit) ((uint32_t)FF_TAB((s) + ((bit) ? 0 : 256)))
#endif
/* put_rac + one renorm_encoder step, rangecoder.h:71-121 */
#define FF_E_PUT(s, bit)                                                          \
    do {                                                                          \
        const int r1_ = (c.range * (int)(s)) >> 8, rb_ = c.range - r1_;           \
        c.low += (bit) ? rb_ : 0;                                                 \
        c.range = (bit) ? r1_ : rb_;                                              \
        if (FF_UNLIKELY(c.range < 0x100))                                         \
            FF_E_RENORM();                                                        \
    } while (0)
/* a decision on a slot of the row other than 0 */
#define FF_E_SLOT(slot, bit)                                                      \
    do {                                                                          \
        const uint32_t sl_ = (uint32_t)(slot);                                    \
        const uint32_t st_ = FF_E_LD(sl_);                                        \
        FF_E_PUT(st_, bit);                                                       \
        FF_E_ST(sl_, FF_E_NEXT(st_, bit));                                        \
    } while (0)

    ff_enc_resume(&c, pre, pre_bytes, out, sl.bs_cap);
    if (rct)
        ff_enc_v4_header_tail(&c, tab_, pre, rct[0], rct[1], 0);
    {
        /* the two tokens after the current one, fetched ahead: a new line of the token
         * stream comes from L2 (one in 32 tokens), about two zero-flag tokens away */
        uint32_t nxt = n ? tokens[0] : 0u, nxt2 = n > 1 ? tokens[1] : 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t tok = nxt;
            const int ctx = (int)(tok & FF_TOKEN_CTX_MASK);
            const int diff = (int32_t)tok >> FF_TOKEN_CTX_BITS;
            nxt = nxt2;
            if (i + 2 < n)
                nxt2 = tokens[i + 2];
            if (i == guard_tok)
                guard_pos = c.pos;
            if (ctx != cur_ctx) {
                if (cur_ctx >= 0) {
                    FF_E_ST(0, s0);
                    ff_row_store(FF_ROWW, state + (size_t)cur_ctx * FF_CONTEXT_SIZE);
                }
                ff_row_load(FF_ROWW, state + (size_t)ctx * FF_CONTEXT_SIZE);
                s0 = FF_E_LD(0);
                cur_ctx = ctx;
            }
            if (diff == 0) {
                FF_E_PUT(s0, 1);
                s0 = FF_E_NEXT(s0, 1);
            } else {
                const uint32_t a = (uint32_t)(diff < 0 ? -diff : diff);
                const int e = ffrac_ilog2(a);
                FF_E_PUT(s0, 0);
                s0 = FF_E_NEXT(s0, 0);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
                for (int j = 0; j < e; j++)                     /* unary exponent */
                    FF_E_SLOT(1 + ff_min(j, 9), 1);
                FF_E_SLOT(1 + ff_min(e, 9), 0);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
                for (int j = e - 1; j >= 0; j--) {              /* mantissa, MSB first */
                    const int bit = (int)((a >> j) & 1u);
                    FF_E_SLOT(22 + ff_min(j, 9), bit);
                }
                {
                    const int bit = diff < 0;
                    FF_E_SLOT(11 + ff_min(e, 10), bit);
                }
            }
        }
    }
    if (cur_ctx >= 0) {
        FF_E_ST(0, s0);
        ff_row_store(FF_ROWW, state + (size_t)cur_ctx * FF_CONTEXT_SIZE);
    }
#undef FF_E_LD
#undef FF_E_ST
#undef FF_E_NEXT
#undef FF_E_PUT
#undef FF_E_SLOT
#undef FF_E_RENORM
    nb = ffrac_enc_finish(&c, tab_, 1);              /* ffv1enc.c:1242 */
    *overflow = c.overflow;
    if (v4_room && (int64_t)v4_room - (int64_t)guard_pos < (int64_t)sl.seg_w[sl.nseg - 1] * 35)
        *overflow = 1;
    return nb;
}


/* ------------------------------------------------------------------ */
/* stage B in two halves, for streams with few, large slices            */
/* ------------------------------------------------------------------ */
/* In the ENCODER every context and every bit is known once stage A has run.  What is serial
 * is (1) the sequence of adaptive states of each (context, slot) and (2) the arithmetic
 * coder's low / range / carry.  (1) only orders decisions that share a context, (2) is a
 * handful of instructions per decision once the probabilities are known.  A launch with few
 * slices (BASELINE C3/C5: 4 or 9 per picture) leaves the GPU empty and lasts as long as one
 * lane needs for a whole slice, so there the two are separated:
 *   ff_chain_token          (k_chain_states, one WARP per slice) 32 tokens at a time: tokens
 *                           of different contexts run their state chains in parallel, tokens
 *                           of one context in coding order; every decision leaves a 16-bit
 *                           record (state | bit << 8) at its place in coding order
 *   ff_encode_slice_records (k_code_records, one lane per slice) put_rac + renorm_encoder
 *                           (rangecoder.h:71-121) over the records
 * The packets are the same bytes as ff_encode_slice_range writes. */
FFGPU_HD uint32_t ff_chain_token(uint32_t tok, uint8_t *rows, const FFRacTables *tab, uint16_t *rec)
{
    const int ctx = (int)(tok & FF_TOKEN_CTX_MASK);
    const int diff = (int32_t)tok >> FF_TOKEN_CTX_BITS;
    const uint32_t a = (uint32_t)(diff < 0 ? -diff : diff);
    uint8_t *row = rows + (size_t)ctx * FF_CONTEXT_SIZE;
    if (!a) {
        const int s = row[0];
        rec[0] = (uint16_t)(s | 0x100);
        row[0] = tab->one[s];
        return 1;
    } else {
        const int e = ffrac_ilog2(a), n = 2 * e + 3;
        for (int step = 0; step < n; step++) {
            const int slot = ff_slot_of(e, step);
            int bit;
            if (step == 0)
                bit = 0;
            else if (step <= e)
                bit = 1;
            else if (step == e + 1)
                bit = 0;
            else if (step <= 2 * e + 1)
                bit = (int)((a >> (2 * e + 1 - step)) & 1u);
            else
                bit = diff < 0;
            const int s = row[slot];
            rec[step] = (uint16_t)(s | (bit << 8));
            row[slot] = bit ? tab->one[s] : tab->zero[s];
        }
        return (uint32_t)n;
    }
}

FFGPU_HD uint32_t ff_encode_slice_records(const FFDevSlice &sl, const uint16_t *rec, uint32_t nrec,
                                          uint32_t guard_rec, const FFRacTables *tab,
                                          const FFRacPrefix &pre, const uint8_t *pre_bytes,
                                          uint8_t *out, uint32_t *overflow, const int *rct, uint32_t v4_room)
{
    FFRacEnc c;
    uint32_t guard_pos = 0, nb;
    ff_enc_resume(&c, pre, pre_bytes, out, sl.bs_cap);
    if (rct)
        ff_enc_v4_header_tail(&c, tab, pre, rct[0], rct[1], 0);
    for (uint32_t i = 0; i < nrec; i++) {
        const uint32_t r = rec[i];
        const int s = (int)(r & 0xFFu), bit = (int)(r >> 8);
        const int r1 = (c.range * s) >> 8, rb = c.range - r1;
        if (i == guard_rec)                          /* see ff_encode_slice_range: version 4 */
            guard_pos = c.pos;
        c.low += bit ? rb : 0;
        c.range = bit ? r1 : rb;
        if (c.range < 0x100)
            ffrac_enc_shift1(&c);
    }
    nb = ffrac_enc_finish(&c, tab, 1);
    *overflow = c.overflow;
    if (v4_room && (int64_t)v4_room - (int64_t)guard_pos < (int64_t)sl.seg_w[sl.nseg - 1] * 35)
        *overflow = 1;
    return nb;
}

/* number of binary decisions the range coder spends on a token: the slice's total is its
 * serial cost and drives the longest-first scheduling of stage B */
FFGPU_HD uint32_t ff_token_weight(uint32_t tok)
{
    const int diff = (int32_t)tok >> FF_TOKEN_CTX_BITS;
    const uint32_t a = (uint32_t)(diff < 0 ? -diff : diff);
    return a ? 2u * (uint32_t)ffrac_ilog2(a) + 3u : 1u;
}

/* ------------------------------------------------------------------ */
/* stage B: Golomb-Rice coder                                            */
/* ------------------------------------------------------------------ */

/* VlcState (ffv1.h:61-66) packed in 8 bytes: lo = drift | error_sum << 16,
 * hi = (uint8)bias | count << 8 */
typedef struct FFVlc {
    int drift;
    int error_sum;
    int bias;
    int count;
} FFVlc;

#define FF_VLC_INIT_LO (4u << 16)
#define FF_VLC_INIT_HI (1u << 8)

FFGPU_HD FFVlc ff_vlc_load(const uint2 *p)
{
    const uint2 v = *p;
    FFVlc s;
    s.drift = (int16_t)(v.x & 0xFFFF);
    s.error_sum = (int)(v.x >> 16);
    s.bias = (int8_t)(v.y & 0xFF);
    s.count = (int)((v.y >> 8) & 0xFF);
    return s;
}

FFGPU_HD void ff_vlc_store(uint2 *p, const FFVlc &s)
{
    uint2 v;
    v.x = ((uint32_t)s.drift & 0xFFFFu) | ((uint32_t)(s.error_sum & 0xFFFF) << 16);
    v.y = ((uint32_t)s.bias & 0xFFu) | ((uint32_t)s.count << 8);
    *p = v;
}

/* update_vlc_state, ffv1.h:162-188 */
FFGPU_HD void ff_vlc_adapt(FFVlc *s, int v)
{
    int drift = s->drift, count = s->count;
    /* modulo 2^32 like the reference's wrapping ints: a damaged stream delivers any v */
    s->error_sum = (int)(((uint32_t)s->error_sum + (v < 0 ? 0u - (uint32_t)v : (uint32_t)v)) & 0xFFFF);
    drift = (int)((uint32_t)drift + (uint32_t)v);
    if (count == 128) {
        count >>= 1;
        drift >>= 1;
        s->error_sum >>= 1;
    }
    count++;
    if (drift <= -count) {
        s->bias = ff_max(s->bias - 1, -128);
        drift = ff_max(drift + count, -count + 1);
    } else if (drift > 0) {
        s->bias = ff_min(s->bias + 1, 127);
        drift = ff_min(drift - count, 0);
    }
    s->drift = (int16_t)drift;
    s->count = count;
}

FFGPU_HD int ff_vlc_k(const FFVlc &s)
{
    int i = s.count, k = 0;
    while (i < s.error_sum) {
        k++;
        i += i;
    }
    return k;
}

/* MSB-first bit writer (put_bits.h) with a 64-bit accumulator */
typedef struct FFBitW {
    uint8_t *buf;
    uint32_t pos, cap;     /* bytes written / capacity */
    uint64_t acc;          /* pending bits, right aligned */
    int      nacc;
    uint32_t overflow;
} FFBitW;

FFGPU_HD void ff_bw_put(FFBitW *b, int n, uint32_t v)
{
    b->acc = (b->acc << n) | (uint64_t)v;
    b->nacc += n;
    while (b->nacc >= 8) {
        b->nacc -= 8;
        if (b->pos < b->cap)
            b->buf[b->pos] = (uint8_t)(b->acc >> b->nacc);
        else
            b->overflow = 1;
        b->pos++;
    }
}

/* flush_put_bits: zero padding to a byte boundary */
FFGPU_HD void ff_bw_flush(FFBitW *b)
{
    if (b->nacc > 0)
        ff_bw_put(b, 8 - b->nacc, 0);
}

/* put_vlc_symbol ffv1enc.c:240-262 + set_sr_golomb / set_ur_golomb golomb.h:676-731 */
FFGPU_HD void ff_vlc_put(FFBitW *b, uint2 *sp, int v, int bits)
{
    FFVlc s = ff_vlc_load(sp);
    int k, code, u, e;
    v = ff_fold(v - s.bias, bits);
    k = ff_vlc_k(s);
    code = v ^ ((2 * s.drift + s.count) >> 31);
    u = -2 * code - 1;
    u ^= u >> 31;
    e = u >> k;
    if (e < 12)
        ff_bw_put(b, e + k + 1, (1u << k) + ((uint32_t)u & ((1u << k) - 1)));
    else
        ff_bw_put(b, 12 + bits, (uint32_t)(u - 11));
    ff_vlc_adapt(&s, v);
    ff_vlc_store(sp, s);
}

/* encode_line's Golomb branch incl. run mode, ffv1enc_template.c:58-119, over all lines */
FFGPU_HD uint32_t ff_encode_slice_golomb(const FFDevParams &P, const FFDevSlice &sl,
                                         const uint32_t *tokens, uint2 *vstate,
                                         const FFRacPrefix &pre, const uint8_t *pre_bytes,
                                         uint8_t *out, uint32_t *overflow,
                                         const FFRacTables *tab = 0, const int *rct = 0)
{
    FFBitW bw;
    uint32_t i, t = 0, start = pre.golomb_start;
    int seg;
    if (rct) {
        /* version 4: the header is finished and its coder closed here (see FFRacPrefix) */
        FFRacEnc c;
        ff_enc_resume(&c, pre, pre_bytes, out, sl.bs_cap);
        start = ff_enc_v4_header_tail(&c, tab, pre, rct[0], rct[1], 1);
    } else {
        for (i = 0; i < pre.nbytes && i < sl.bs_cap; i++)
            out[i] = pre_bytes[pre.byte_off + i];
    }
    bw.buf = out;
    bw.pos = start;
    bw.cap = sl.bs_cap;
    bw.acc = 0;
    bw.nacc = 0;
    bw.overflow = 0;
    for (seg = 0; seg < sl.nseg; seg++) {
        const int w = sl.seg_w[seg];
        int run_index = 0, line;
        for (line = 0; line < sl.seg_lines[seg]; line++) {
            /* context 0 of the line's plane-context set */
            const int set = P.colorspace == 0 ? P.cp[seg].set : P.cp[line % P.ncoded].set;
            const uint32_t ctx0 = (uint32_t)P.set_base[set];
            int run_count = 0, run_mode = 0, x;
            for (x = 0; x < w; x++) {
                const uint32_t tok = tokens[t++];
                const uint32_t ctx = tok & FF_TOKEN_CTX_MASK;
                int diff = (int32_t)tok >> FF_TOKEN_CTX_BITS;
                if (ctx == ctx0)
                    run_mode = 1;
                if (run_mode) {
                    if (diff) {
                        while (run_count >= 1 << ff_run_log2(run_index)) {
                            run_count -= 1 << ff_run_log2(run_index);
                            run_index++;
                            ff_bw_put(&bw, 1, 1);
                        }
                        ff_bw_put(&bw, 1 + ff_run_log2(run_index), (uint32_t)run_count);
                        if (run_index)
                            run_index--;
                        run_count = 0;
                        run_mode = 0;
                        if (diff > 0)
                            diff--;
                    } else {
                        run_count++;
                    }
                }
                if (!run_mode)
                    ff_vlc_put(&bw, vstate + ctx, diff, P.cbits);
            }
            if (run_mode) {
                while (run_count >= 1 << ff_run_log2(run_index)) {
                    run_count -= 1 << ff_run_log2(run_index);
                    run_index++;
                    ff_bw_put(&bw, 1, 1);
                }
                if (run_count)
                    ff_bw_put(&bw, 1, 1);
            }
        }
    }
    ff_bw_flush(&bw);                           /* ffv1enc.c:1244-1245 */
    *overflow = bw.overflow;
    return bw.pos;
}

/* ------------------------------------------------------------------ */
/* packet assembly                                                      */
/* ------------------------------------------------------------------ */

/* bytes slice i occupies in the packet: payload, 24-bit size, optional 0x00 + CRC
 * (ffv1enc.c:1248-1261) */
FFGPU_HD uint32_t ff_slice_packed_size(const FFDevParams &P, int i, uint32_t payload)
{
    uint32_t n = payload;
    if (i > 0 || P.version > 2)
        n += 3;
    if (P.ec)
        n += 5;
    return n;
}

/* CRC-32 IEEE (0x04C11DB7) MSB-first, init 0, no final xor: libavutil/crc.c:336
 * AV_CRC_32_IEEE.  The register is kept in textbook form; av_crc()'s byte-swapped value
 * stored with AV_WL32 (ffv1enc.c:1258) puts exactly these four bytes, MSB first. */
FFGPU_HD uint32_t ff_crc_table_entry(int i)
{
    uint32_t c = (uint32_t)i << 24;
    int k;
    for (k = 0; k < 8; k++)
        c = (c << 1) ^ ((c >> 31) ? 0x04C11DB7u : 0u);
    return c;
}

/* copy slice i's payload to its place in the packet and append the trailer:
 * 24-bit big-endian size, then (ec) a zero error-status byte and the CRC over
 * payload+size+status, ffv1enc.c:1248-1261.  One thread per slice. */
FFGPU_HD void ff_pack_slice(const FFDevParams &P, int i, const uint8_t *src, uint32_t payload,
                            uint8_t *dst, const uint32_t *crc_tab)
{
    uint32_t n, j, r = 0;
    for (n = 0; n < payload; n++)
        dst[n] = src[n];
    if (i > 0 || P.version > 2) {
        dst[n++] = (uint8_t)(payload >> 16);
        dst[n++] = (uint8_t)(payload >> 8);
        dst[n++] = (uint8_t)payload;
    }
    if (P.ec) {
        dst[n++] = 0;
        for (j = 0; j < n; j++)
            r = (r << 8) ^ crc_tab[(r >> 24) ^ dst[j]];
        dst[n++] = (uint8_t)(r >> 24);
        dst[n++] = (uint8_t)(r >> 16);
        dst[n++] = (uint8_t)(r >> 8);
        dst[n++] = (uint8_t)r;
    }
}

/* ------------------------------------------------------------------ */
/* decoder                                                               */
/* ------------------------------------------------------------------ */

/* MSB-first bit reader (get_bits.h, checked reader) */
typedef struct FFBitR {
    const uint8_t *buf;
    int64_t pos, size_bits;
} FFBitR;

FFGPU_HD uint32_t ff_br_peek32(const FFBitR *r)
{
    const uint8_t *p = r->buf + (r->pos >> 3);
    const uint64_t w = ((uint64_t)p[0] << 32) | ((uint64_t)p[1] << 24) | ((uint64_t)p[2] << 16) |
                       ((uint64_t)p[3] << 8) | (uint64_t)p[4];
    return (uint32_t)(w >> (8 - (int)(r->pos & 7)));
}

FFGPU_HD void ff_br_skip(FFBitR *r, int n)
{
    r->pos += n;
    if (r->pos > r->size_bits + 8)
        r->pos = r->size_bits + 8;
}

FFGPU_HD uint32_t ff_br_get(FFBitR *r, int n)
{
    uint32_t v;
    if (!n)
        return 0;
    v = ff_br_peek32(r) >> (32 - n);
    ff_br_skip(r, n);
    return v;
}

/* get_vlc_symbol ffv1dec.c:71-94 + get_sr_golomb / get_ur_golomb golomb.h:373-413,529-534 */
FFGPU_HD int ff_vlc_get(FFBitR *r, uint2 *sp, int bits)
{
    FFVlc s = ff_vlc_load(sp);
    const int k = ff_vlc_k(s);
    uint32_t buf = ff_br_peek32(r);
    const int log = ffrac_ilog2(buf);
    uint32_t u;
    int v, ret;
    if (log > 31 - 12) {
        buf >>= log - k;
        buf += (uint32_t)(30 - log) << k;
        ff_br_skip(r, 32 + k - log);
        u = buf;
    } else {
        ff_br_skip(r, 12);
        u = ff_br_get(r, bits) + 11;
    }
    v = (int)(u >> 1) ^ -(int)(u & 1);
    v ^= (2 * s.drift + s.count) >> 31;
    ret = ff_fold((int)((uint32_t)v + (uint32_t)s.bias), bits);
    ff_vlc_adapt(&s, v);
    ff_vlc_store(sp, s);
    return ret;
}

/* everything a slice decode needs besides the stream parameters */
typedef struct FFDecCtx {
    const int16_t *qt_all;      /* quant table sets, FF_QT_STRIDE apart               */
    const FFRacTables *tab;
    uint8_t *rstate;            /* range: [total_ctx][32]                             */
    uint2   *vstate;            /* golomb: [total_ctx]                                */
    int32_t *lines;             /* [ncoded][2][line_stride] scratch                   */
    int      line_stride;
    uint8_t *frame;             /* output picture                                     */
    int gate_wait;              /* sample set-up gating: longest idle wait, in iterations */
    uint32_t *touched;          /* one bit per context of the slice: its state row exists.   */
                                /* NULL: every row was initialised before the slice started  */
    int any_five;               /* some quant table of the stream uses 5 context inputs      */
    int lone;                   /* the slice has a warp of its own: ff_decode_slice_range_planar_lone */
} FFDecCtx;

FFGPU_HD int ff_wrap_sample(const FFDevParams &P, int v)
{
    return P.use32 ? v : (int)(int16_t)v;
}

/* decode_line (Golomb branch incl. run mode), ffv1dec_template.c:70-113 */
FFGPU_HD int ff_decode_line_golomb(const FFDevParams &P, FFBitR *br, const FFDecCtx &D,
                                   const int16_t *qt, uint2 *vstate, int w,
                                   const int32_t *prev, int32_t *cur, int *run_index_io)
{
    const int five = qt[FF_MAX_CTX_INPUTS * 256];
    const uint32_t mask = (1u << P.cbits) - 1;
    int run_index = *run_index_io, run_count = 0, run_mode = 0;
    int x;
    int T = prev[0], LT = cur[0], L = prev[0], LL = 0;
    if (br->size_bits - br->pos < 1)              /* is_input_end */
        return -1;
    for (x = 0; x < w; x++) {
        int RT = prev[ff_min(x + 1, w - 1)];
        int ctx, sign = 0, diff, v;
        if (!(x & 1023) && br->size_bits - br->pos < 1)
            return -1;
        ctx = qt[(L - LT) & 0xFF] + qt[256 + ((LT - T) & 0xFF)] + qt[512 + ((T - RT) & 0xFF)];
        if (five) {
            const int TT = cur[x];
            ctx += qt[768 + ((LL - L) & 0xFF)] + qt[1024 + ((TT - T) & 0xFF)];
        }
        if (ctx < 0) {
            ctx = -ctx;
            sign = 1;
        }
        if (ctx == 0 && run_mode == 0)
            run_mode = 1;
        if (run_mode) {
            if (run_count == 0 && run_mode == 1) {
                if (ff_br_get(br, 1)) {
                    run_count = 1 << ff_run_log2(run_index);
                    if (x + run_count <= w)
                        run_index++;
                } else {
                    run_count = ff_run_log2(run_index) ? (int)ff_br_get(br, ff_run_log2(run_index)) : 0;
                    if (run_index)
                        run_index--;
                    run_mode = 2;
                }
            }
            /* zero residuals: each sample equals its prediction */
            while (run_count > 1 && w - x > 1) {
                v = ff_wrap_sample(P, ff_median3(L, L + T - LT, T));
                cur[x] = v;
                LL = L;
                L = v;
                LT = T;
                T = RT;
                x++;
                RT = prev[ff_min(x + 1, w - 1)];
                run_count--;
            }
            run_count--;
            if (run_count < 0) {
                run_mode = 0;
                run_count = 0;
                diff = ff_vlc_get(br, vstate + ctx, P.cbits);
                if (diff >= 0)
                    diff++;
            } else {
                diff = 0;
            }
        } else {
            diff = ff_vlc_get(br, vstate + ctx, P.cbits);
        }
        if (sign)
            diff = -diff;
        v = ff_wrap_sample(P, (int)(((uint32_t)ff_median3(L, L + T - LT, T) + (uint32_t)diff) & mask));
        cur[x] = v;
        LL = L;
        L = v;
        LT = T;
        T = RT;
    }
    *run_index_io = run_index;
    return 0;
}

/* decode_plane's store, ffv1dec.c:142-161 */
FFGPU_HD void ff_store_line_ycc(const FFDevParams &P, uint8_t *frame, int k, int X0, int Y,
                                int w, const int32_t *line)
{
    const FFDevPlane cp = P.cp[k];
    uint8_t *p = frame + P.plane_off[cp.mem] + (size_t)Y * P.pitch[cp.mem] + (size_t)X0 * cp.step + cp.off;
    int x;
    for (x = 0; x < w; x++) {
        const int v = line[x];
        if (P.sbits <= 8) {
            p[(size_t)x * cp.step] = (uint8_t)v;
        } else {
            uint32_t o;
            if (P.packed_lsb)
                o = (uint32_t)v & 0xFFFF;
            else
                o = (uint32_t)((v << (16 - P.sbits)) | ((v & 0xFFFF) >> (2 * P.sbits - 16))) & 0xFFFF;
            p[2 * x] = (uint8_t)o;
            p[2 * x + 1] = (uint8_t)(o >> 8);
        }
    }
}

/* decode_rgb_frame's inverse RCT + store, ffv1dec_template.c:160-190 */
FFGPU_HD void ff_store_line_rgb(const FFDevParams &P, uint8_t *frame, int X0, int Y, int w,
                                const int32_t *lg, const int32_t *lb, const int32_t *lr,
                                const int32_t *la, int rct_by = 1, int rct_ry = 1, int pcm = 0)
{
    const int offset = 1 << P.sbits;
    int x;
    for (x = 0; x < w; x++) {
        int g = lg[x], b = lb[x], r = lr[x];
        const int a = la ? la[x] : 0;
        if (!pcm) {                                  /* ffv1dec_template.c:169-176 */
            b -= offset;
            r -= offset;
            g -= (b * rct_by + r * rct_ry) >> 2;
            b += g;
            r += g;
        }
        if (P.layout == FF_LAY_BGR32) {
            uint8_t *p = frame + P.plane_off[0] + (size_t)Y * P.pitch[0] + (size_t)(X0 + x) * 4;
            const uint32_t v = (uint32_t)b + ((uint32_t)g << 8) + ((uint32_t)r << 16) + ((uint32_t)a << 24);
            p[0] = (uint8_t)v;
            p[1] = (uint8_t)(v >> 8);
            p[2] = (uint8_t)(v >> 16);
            p[3] = (uint8_t)(v >> 24);
        } else {
            int v0, v1;
            uint8_t *p0 = frame + P.plane_off[0] + (size_t)Y * P.pitch[0] + (size_t)(X0 + x) * 2;
            uint8_t *p1 = frame + P.plane_off[1] + (size_t)Y * P.pitch[1] + (size_t)(X0 + x) * 2;
            uint8_t *p2 = frame + P.plane_off[2] + (size_t)Y * P.pitch[2] + (size_t)(X0 + x) * 2;
            if (P.use32 || P.transparency) {
                v0 = g; v1 = b;
            } else {
                v0 = b; v1 = g;
            }
            p0[0] = (uint8_t)v0; p0[1] = (uint8_t)(v0 >> 8);
            p1[0] = (uint8_t)v1; p1[1] = (uint8_t)(v1 >> 8);
            p2[0] = (uint8_t)r;  p2[1] = (uint8_t)(r >> 8);
            if (P.transparency) {
                uint8_t *p3 = frame + P.plane_off[3] + (size_t)Y * P.pitch[3] + (size_t)(X0 + x) * 2;
                p3[0] = (uint8_t)a; p3[1] = (uint8_t)(a >> 8);
            }
        }
    }
}

/* ff_init_range_decoder + decode_slice_header (ffv1dec.c:167-244) + the Golomb hand-over
 * (ffv1dec.c:312-319) + the slice CRC check (ffv1dec.c:905-922) for slices the host left to
 * the device.  Fills the coder state / rectangle / quant table indices of the work item. */
FFGPU_HD_COLD void ff_dec_slice_header(const FFDevParams &P, const FFDecHdr &H, FFDecSlice *w,
                                  const uint8_t *pkt, const FFRacTables *tab,
                                  const uint32_t *crc_tab, FFDecResult *res)
{
    const uint8_t *base = pkt + w->pkt_off;
    FFRacDec c;
    uint8_t st[FF_CONTEXT_SIZE];
    int i, bad = 0;
    if (P.ec) {
        uint32_t r = 0;
        for (uint32_t j = 0; j < w->size; j++)
            r = (r << 8) ^ crc_tab[(r >> 24) ^ base[j]];
        if (r)
            res->flags |= FF_RES_CRC_BAD;
    }
    if (w->size < 2) {
        w->skip = 1;
        res->flags |= FF_RES_HDR_BAD;
        return;
    }
    ffrac_dec_init(&c, base, w->size);
    for (i = 0; i < FF_CONTEXT_SIZE; i++)
        st[i] = 128;
    {
        const unsigned sx = (unsigned)ffrac_get_symbol(&c, tab, st, 0) * (unsigned)P.width;
        const unsigned sy = (unsigned)ffrac_get_symbol(&c, tab, st, 0) * (unsigned)P.height;
        const unsigned sw = ((unsigned)ffrac_get_symbol(&c, tab, st, 0) + 1U) * (unsigned)P.width + sx;
        const unsigned sh = ((unsigned)ffrac_get_symbol(&c, tab, st, 0) + 1U) * (unsigned)P.height + sy;
        w->x = (int)sx / P.nh;
        w->y = (int)sy / P.nv;
        w->w = (int)sw / P.nh - w->x;
        w->h = (int)sh / P.nv - w->y;
        if ((unsigned)w->w > (unsigned)P.width || (unsigned)w->h > (unsigned)P.height)
            bad = 1;
        else if ((unsigned)w->x + (uint64_t)w->w > (unsigned)P.width ||
                 (unsigned)w->y + (uint64_t)w->h > (unsigned)P.height)
            bad = 1;
    }
    for (i = 0; i < P.nsets && !bad; i++) {
        const int idx = ffrac_get_symbol(&c, tab, st, 0);
        if ((unsigned)idx >= (unsigned)H.qt_count || H.ctx_count[idx & 7] > H.ctx_cap)
            bad = 1;                       /* "quant_table_index out of range" / arena too small */
        else
            w->qidx[i] = idx;
    }
    if (bad) {
        w->x = w->y = w->w = w->h = 0;
        w->skip = 1;
        res->flags |= FF_RES_HDR_BAD;
        return;
    }
    ffrac_get_symbol(&c, tab, st, 0);      /* picture structure: the frame takes slice 0's */
    ffrac_get_symbol(&c, tab, st, 0);      /* sample aspect ratio                          */
    ffrac_get_symbol(&c, tab, st, 0);
    if (P.ac == FF_AC_GOLOMB) {
        if ((P.version == 3 && H.micro_version > 1) || P.version > 3) {
            uint8_t term = 129;
            ffrac_get(&c, tab, &term);
        }
        w->golomb_start = c.pos - 1;       /* version > 2 here */
    }
    w->low = c.low;
    w->range = c.range;
    w->pos = c.pos;
    w->overread = c.overread;
    if (c.end < w->size)
        w->size = c.end;
    if (!w->w || !w->h)
        w->skip = 1;
}

/* line iterator over the coded lines of a slice, in coding order (YCbCr: plane after plane,
 * RGB: G,B,R[,A] interleaved per picture line) */
typedef struct FFLineIt {
    int k, y;          /* current coded plane / line */
    int w, h;          /* geometry of plane k        */
} FFLineIt;

FFGPU_HD int ff_line_first(const FFDevParams &P, const FFDecSlice &d, FFLineIt *it)
{
    it->k = 0;
    it->y = 0;
    it->w = ff_crshift(d.w, P.cp[0].hs);
    it->h = ff_crshift(d.h, P.cp[0].vs);
    return it->w > 0 && it->h > 0;
}

FFGPU_HD int ff_line_next(const FFDevParams &P, const FFDecSlice &d, FFLineIt *it)
{
    if (P.colorspace == 0) {
        if (++it->y < it->h)
            return 1;
        if (++it->k >= P.ncoded)
            return 0;
        it->y = 0;
        it->w = ff_crshift(d.w, P.cp[it->k].hs);
        it->h = ff_crshift(d.h, P.cp[it->k].vs);
        return 1;
    }
    if (++it->k < P.ncoded)
        return 1;
    it->k = 0;
    return ++it->y < it->h;
}

/* decode_slice after the header for the range coder, ffv1dec.c:304-359 with decode_line
 * (ffv1dec_template.c:23-126) and get_symbol_inline (ffv1dec.c:42-64) flattened to one
 * binary decision per loop iteration, for the same reason as in the encoder: the lanes of
 * a warp decode different slices and must not wait for each other's long residuals. */
FFGPU_HD void ff_decode_slice_range(const FFDevParams &P, const FFDecSlice &d, const uint8_t *pkt,
                                    const FFDecCtx &D, FFDecResult *res, uint32_t *row_)
{
#if !defined(__CUDA_ARCH__)                          /* host build: tables through pointers */
    const FFRacTables *tab_ = D.tab;
    const int16_t *qt_all_ = D.qt_all;
#endif
    const uint32_t mask = (1u << P.cbits) - 1;
    const int use32 = P.use32;
    /* how a finished sample reaches the picture: 0 at the end of the line (RGB), 1 one byte,
     * 2 one little-endian u16, 3 u16 with the MSB-aligned replication of ffv1dec.c:158 */
    const int smode = P.colorspace ? 0 : P.sbits <= 8 ? 1 : P.packed_lsb ? 2 : 3;
    const int shl = 16 - P.sbits, shr = 2 * P.sbits - 16;
    FFRacDec c;
    FFLineIt it;
    int x, err = 0, cur_ctx = -1, need_new = 1, waited = 0;
    int w = 0, five = 0, sign = 0, e = 0, mi = 0, slot = 0;
    uint32_t a = 0;
    int T = 0, LT = 0, L = 0, LL = 0, RT = 0;
    int q0 = 0, q1 = 0, q2 = 0, q3 = 0;
    int qo = 0;                                      /* offset of the line's quant table set */
    int32_t *cur = D.lines;
    const int32_t *prev = D.lines;
    uint8_t *outp = D.frame;
    const uint8_t *prow = D.frame, *pprow = D.frame;   /* picture rows y-1 and y-2 (planar modes) */
    int ostep = 0, havep = 0, havepp = 0, usepic = 0;
    int sbase = 0;
    (void)row_; (void)waited;

    /* planar YCbCr, full-resolution planes: the previous lines are read back from the output
     * picture itself (what decode_plane just stored, ffv1dec.c:142-161), so no separate line
     * buffer is written.  Subsampled chroma planes keep private line buffers: with an odd
     * luma offset the chroma rectangles of neighbouring slices share a column/row
     * (ffv1dec.c:324-327), and another thread may be writing it. */
#define FF_PIC(rowp, xx) (smode == 1 ? (int)(rowp)[(size_t)(xx) * ostep]                                     \
                          : smode == 2 ? (int)(int16_t) * (const uint16_t *)((rowp) + (size_t)(xx) * ostep)   \
                                       : (int)(int16_t)(*(const uint16_t *)((rowp) + (size_t)(xx) * ostep) >> shl))
#define FF_PREV(xx) (usepic ? (havep ? FF_PIC(prow, xx) : 0) : prev[xx])
#define FF_PREV2(xx) (usepic ? (havepp ? FF_PIC(pprow, xx) : 0) : cur[xx])
#if defined(__CUDA_ARCH__)
    uint32_t row_sa = (uint32_t)__cvta_generic_to_shared(&ff_s_rows[threadIdx.x * FF_ROW_WORDS]);
    uint32_t tab_sa = (uint32_t)__cvta_generic_to_shared(&ff_s_tab);
    row_sa = ff_opaque(row_sa);                      /* see ff_encode_slice_range */
    tab_sa = ff_opaque(tab_sa);
#endif
    c.buf = pkt + d.pkt_off;
    c.low = d.low;
    c.range = d.range;
    c.pos = d.pos;
    c.end = d.size;
    c.overread = d.overread;

    for (int k = 0; k < P.ncoded; k++)
        if (!smode || P.cp[k].hs || P.cp[k].vs)
            for (x = 0; x < 2 * D.line_stride; x++)
                D.lines[(size_t)k * 2 * D.line_stride + x] = 0;

    if (!ff_line_first(P, d, &it))
        goto finish;
    x = -1;                                          /* -1: the line has to be set up */
    for (;;) {
        int s, r1, bit, done, diff;
#if defined(__CUDA_ARCH__)
        /* The sample set-up below is the expensive, divergent part of the loop.  The lanes
         * of a warp finish their symbols at different iterations, so running it whenever ANY
         * lane needs a new sample executes it almost every iteration for a handful of lanes.
         * Instead a lane that needs a sample idles until every active lane needs one, or
         * until the longest-waiting lane has idled D.gate_wait iterations (`waited` counts
         * the iterations since the last set-up in which some lane was waiting; it is the
         * same in every lane). */
        {
            const unsigned act = __activemask();
            const unsigned need = __ballot_sync(act, need_new);
            if (need) {
                if (need == act || waited >= D.gate_wait)
                    waited = 0;
                else {
                    waited++;
                    if (need_new)
                        continue;
                }
            }
        }
#endif
        if (need_new) {
            int ctx;
            if (x < 0 || x == w) {
                if (x == w) {                        /* a line is complete */
                    if (smode == 0 && it.k == P.ncoded - 1) {
                        const int o = (it.y & 1) ? D.line_stride : 0;
                        ff_store_line_rgb(P, D.frame, d.x, d.y + it.y, d.w,
                                          D.lines + o, D.lines + 2 * D.line_stride + o,
                                          D.lines + 4 * D.line_stride + o,
                                          P.ncoded > 3 ? D.lines + 6 * D.line_stride + o : (const int32_t *)0,
                                          d.rct_by, d.rct_ry, 0);
                    }
                    if (!ff_line_next(P, d, &it))
                        break;
                }
                {
                    /* everything that is constant along the line is computed here once */
                    const FFDevPlane cp = P.cp[it.k];
                    int32_t *l0 = D.lines + (size_t)it.k * 2 * D.line_stride;
                    cur = (it.y & 1) ? l0 + D.line_stride : l0;
                    prev = (it.y & 1) ? l0 : l0 + D.line_stride;
                    w = smode ? it.w : d.w;
                    qo = d.qidx[cp.set] * FF_QT_STRIDE;
                    five = FF_QT(qo, FF_MAX_CTX_INPUTS * 256);
                    sbase = P.set_base[cp.set];
                    if (smode) {
                        outp = D.frame + P.plane_off[cp.mem] +
                               (size_t)((d.y >> cp.vs) + it.y) * P.pitch[cp.mem] +
                               (size_t)(d.x >> cp.hs) * cp.step + cp.off;
                        ostep = cp.step;
                        prow = outp - P.pitch[cp.mem];
                        pprow = prow - P.pitch[cp.mem];
                        havep = it.y >= 1;
                        havepp = it.y >= 2;
                        usepic = !cp.hs && !cp.vs;
                    }
                }
                x = 0;
                T = FF_PREV(0);
                LT = FF_PREV2(0);
                L = T;
                LL = 0;
                /* look-ahead on the previous line: q0..q3 = prev[min(x+1..x+4, w-1)] */
                q0 = FF_PREV(ff_min(1, w - 1));
                q1 = FF_PREV(ff_min(2, w - 1));
                q2 = FF_PREV(ff_min(3, w - 1));
                q3 = FF_PREV(ff_min(4, w - 1));
                if (c.overread > 2) {                /* is_input_end at line start */
                    err = 1;
                    break;
                }
            } else if (!(x & 1023) && c.overread > 2) {
                err = 1;
                break;
            }
            RT = q0;
            q0 = q1;
            q1 = q2;
            q2 = q3;
            q3 = FF_PREV(ff_min(x + 5, w - 1));
            ctx = FF_QT(qo, (L - LT) & 0xFF) + FF_QT(qo, 256 + ((LT - T) & 0xFF)) +
                  FF_QT(qo, 512 + ((T - RT) & 0xFF));
            if (five)
                ctx += FF_QT(qo, 768 + ((LL - L) & 0xFF)) + FF_QT(qo, 1024 + ((FF_PREV2(x) - T) & 0xFF));
            sign = ctx < 0;
            ctx = sbase + (sign ? -ctx : ctx);
            if (ctx != cur_ctx) {
                if (cur_ctx >= 0)
                    ff_row_store(FF_ROWW, D.rstate + (size_t)cur_ctx * FF_CONTEXT_SIZE);
                ff_row_load(FF_ROWW, D.rstate + (size_t)ctx * FF_CONTEXT_SIZE);
                cur_ctx = ctx;
            }
            need_new = 0;
            slot = 0;
        }
        /* one binary decision: get_rac + refill, rangecoder.h:123-152 */
#if defined(__CUDA_ARCH__)
        s = (int)ff_lds8(row_sa + (uint32_t)slot);
        r1 = (c.range * s) >> 8;
        c.range -= r1;
        bit = c.low >= c.range;
        ff_sts8(row_sa + (uint32_t)slot, ff_lds8(tab_sa + (uint32_t)s + (bit ? 0u : 256u)));
#else
        s = FF_ROWB(slot);
        r1 = (c.range * s) >> 8;
        c.range -= r1;
        bit = c.low >= c.range;
        FF_ROWB(slot) = FF_TAB(s + (bit ? 0 : 256));
#endif
        c.low -= bit ? c.range : 0;
        c.range = bit ? r1 : c.range;
        if (c.range < 0x100) {
            c.range <<= 8;
            c.low <<= 8;
            if (c.pos < c.end)
                c.low += c.buf[c.pos++];
            else
                c.overread++;
        }
        /* get_symbol_inline (ffv1dec.c:42-64) as a walk over the state slots:
         * 0 zero flag | 1..10 unary exponent | 22..31 mantissa | 11..21 sign */
        done = 0;
        diff = 0;
        if (slot == 0) {
            if (bit)
                done = 1;
            else {
                slot = 1;
                e = 0;
            }
        } else if (slot < 11) {
            if (bit) {
                if (++e > 31) {                      /* get_symbol returns AVERROR_INVALIDDATA */
                    diff = FFRAC_SYMBOL_ERROR;
                    done = 1;
                }
                slot = 1 + ff_min(e, 9);
            } else {
                a = 1;
                mi = e - 1;
                slot = e ? 22 + ff_min(mi, 9) : 11;
            }
        } else if (slot >= 22) {
            a += a + (uint32_t)bit;
            mi--;
            slot = mi < 0 ? 11 + ff_min(e, 10) : 22 + ff_min(mi, 9);
        } else {
            diff = bit ? (int)(0u - a) : (int)a;
            done = 1;
        }
        if (done) {
            int v;
            diff = sign ? FF_NEG32(diff) : diff;
            v = (int)(((uint32_t)ff_median3(L, L + T - LT, T) + (uint32_t)diff) & mask);
            v = use32 ? v : (int)(int16_t)v;
            if (!usepic)
                cur[x] = v;
            if (smode == 1) {                 /* decode_plane's store, ffv1dec.c:142-161 */
                *outp = (uint8_t)v;
            } else if (smode == 2) {
                *(uint16_t *)outp = (uint16_t)v;
            } else if (smode == 3) {
                *(uint16_t *)outp = (uint16_t)((v << shl) | ((v & 0xFFFF) >> shr));
            }
            outp += ostep;
            LL = L;
            L = v;
            LT = T;
            T = RT;
            x++;
            need_new = 1;
        }
    }
finish:
    if (cur_ctx >= 0)
        ff_row_store(FF_ROWW, D.rstate + (size_t)cur_ctx * FF_CONTEXT_SIZE);
    /* end-of-slice check, ffv1dec.c:351-359 */
    if (P.version > 2) {
        uint8_t term = 129;
        ffrac_get(&c, D.tab, &term);
    }
    res->end_pos = c.pos;
    res->overread = c.overread;
    res->error = err;
#undef FF_PIC
#undef FF_PREV
#undef FF_PREV2
}

/* decode_slice after the header, Golomb-Rice streams (ffv1dec.c:304-350) */
FFGPU_HD void ff_decode_slice_golomb(const FFDevParams &P, const FFDecSlice &d, const uint8_t *pkt,
                                     const FFDecCtx &D, FFDecResult *res)
{
    const uint8_t *base = pkt + d.pkt_off;
    FFBitR br;
    int k, y, x, err = 0;

    br.buf = base + d.golomb_start;
    br.size_bits = (int64_t)(d.size - d.golomb_start) * 8;
    br.pos = 0;

    /* zero the line buffers: memset of sample_buffer, ffv1dec.c:131 / ffv1dec_template.c:146 */
    for (k = 0; k < P.ncoded; k++)
        for (x = 0; x < 2 * D.line_stride; x++)
            D.lines[(size_t)k * 2 * D.line_stride + x] = 0;

    if (P.colorspace == 0) {
        for (k = 0; k < P.ncoded; k++) {
            const FFDevPlane cp = P.cp[k];
            const int w = ff_crshift(d.w, cp.hs), h = ff_crshift(d.h, cp.vs);
            const int16_t *qt = D.qt_all + (size_t)d.qidx[cp.set] * FF_QT_STRIDE;
            const size_t sbase = (size_t)P.set_base[cp.set];
            int32_t *l0 = D.lines + (size_t)k * 2 * D.line_stride;
            int32_t *l1 = l0 + D.line_stride;
            int run_index = 0;
            for (y = 0; y < h; y++) {
                int32_t *cur = (y & 1) ? l1 : l0;
                const int32_t *prev = (y & 1) ? l0 : l1;
                if (ff_decode_line_golomb(P, &br, D, qt, D.vstate + sbase, w, prev, cur, &run_index) < 0) {
                    err = 1;
                    break;                       /* decode_plane returns, next plane still runs */
                }
                ff_store_line_ycc(P, D.frame, k, d.x >> cp.hs, (d.y >> cp.vs) + y, w, cur);
            }
        }
    } else {
        int run_index = 0;
        for (y = 0; y < d.h && !err; y++) {
            for (k = 0; k < P.ncoded; k++) {
                const FFDevPlane cp = P.cp[k];
                const int16_t *qt = D.qt_all + (size_t)d.qidx[cp.set] * FF_QT_STRIDE;
                const size_t sbase = (size_t)P.set_base[cp.set];
                int32_t *l0 = D.lines + (size_t)k * 2 * D.line_stride;
                int32_t *l1 = l0 + D.line_stride;
                int32_t *cur = (y & 1) ? l1 : l0;
                const int32_t *prev = (y & 1) ? l0 : l1;
                if (ff_decode_line_golomb(P, &br, D, qt, D.vstate + sbase, d.w, prev, cur, &run_index) < 0) {
                    err = 1;
                    break;
                }
            }
            if (!err) {
                const int o = (y & 1) ? D.line_stride : 0;
                ff_store_line_rgb(P, D.frame, d.x, d.y + y, d.w,
                                  D.lines + o, D.lines + 2 * D.line_stride + o,
                                  D.lines + 4 * D.line_stride + o,
                                  P.ncoded > 3 ? D.lines + 6 * D.line_stride + o : (const int32_t *)0,
                                  d.rct_by, d.rct_ry, 0);
            }
        }
    }
    res->end_pos = d.pos;
    res->overread = d.overread;
    res->error = err;

}

/* ------------------------------------------------------------------ */
/* decoder, planar YCbCr / gray, range coder: the specialised hot path   */
/* ------------------------------------------------------------------ */
/* Same algorithm and loop shape as ff_decode_slice_range (one binary decision per iteration,
 * gated per-sample set-up), specialised at compile time for the sample container
 * (SMODE 1: one byte, 2: little-endian u16 packed at the LSB) and for 3- or 5-input contexts,
 * and trimmed for instruction count, the quantity that bounds a launch with tens of
 * thousands of resident slice decoders:
 *   - every line is described by three byte pointers (previous line, line before that,
 *     output row) that are set up once per line; the first lines point at a zeroed row, so
 *     the sample loop has no "is there a line above" selects;
 *   - subsampled planes keep their private two-line ring (see ff_decode_slice_range) but in
 *     the picture's own sample container, so one load path serves both cases;
 *   - the next bitstream byte is fetched right after the previous one was consumed, so the
 *     renormalisation never waits for a global load;
 *   - adaptive state rows are created on first touch (D.touched, one bit per context of the
 *     slice) instead of by a kernel that fills the whole arena: ff_ffv1_clear_slice_state
 *     (ffv1.c:182-207) without the memset traffic.  Only for streams whose every frame is a
 *     key frame and that carry no initial-state tables. */
/* samples of the previous line fetched ahead of their use (2 or 4) */
#ifndef FF_DEC_AHEAD
#define FF_DEC_AHEAD 4
#endif
template <int SMODE> struct FFPix;
template <> struct FFPix<1> { typedef uint8_t type; };
template <> struct FFPix<2> { typedef uint16_t type; };

/* A sample of a previous line, wrapped to the coder's int16.  On the device the load is an
 * explicit ld.global with the sign extension done by the load unit: written in C the
 * extension becomes a separate instruction that the compiler schedules right behind the
 * load, and the warp then waits out the whole memory latency of a value it only needs four
 * samples later (21 % of all stall samples of the kernel in the first profile). */
template <int SMODE>
FFGPU_HD int ff_pix_load(const uint8_t *p)
{
#if defined(__CUDA_ARCH__)
    int v;
    if (SMODE == 1)
        asm volatile("ld.global.u8 %0, [%1];" : "=r"(v) : "l"(p));
    else
        asm volatile("ld.global.s16 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#else
    if (SMODE == 1)
        return (int)*p;
    return (int)(int16_t)*(const uint16_t *)p;
#endif
}

template <int SMODE, bool FIVE>
FFGPU_HD void ff_decode_slice_range_planar(const FFDevParams &P, const FFDecSlice &d, const uint8_t *pkt,
                                           const FFDecCtx &D, FFDecResult *res, uint32_t *row_, int live)
{
#if !defined(__CUDA_ARCH__)
    const FFRacTables *tab_ = D.tab;
    const int16_t *qt_all_ = D.qt_all;
#endif
    typedef typename FFPix<SMODE>::type pix_t;
    const uint32_t mask = (1u << P.cbits) - 1;
    const uint8_t *buf = pkt + d.pkt_off;
    int low = d.low, range = d.range, overread = d.overread;
    uint32_t pos = d.pos;
    const uint32_t end = d.size;
    uint32_t nbyte;                                   /* buf[pos], fetched ahead */
    FFLineIt it;
    /* state: 1 the lane needs the next sample set up, 2 it is inside a symbol, 0 its slice is
     * finished (or it never had one) */
    unsigned state = 1;
    int x = -1, w = 0, err = 0, cur_ctx = -1;
    int five = 0, sign = 0, e = 0, mi = 0, slot = 0;
    /* the adaptive state of the coming decision, and the current row's slot 0 (the zero
     * flag, four of five decisions in flat pictures): both live in registers, so the
     * decision does not start with a shared-memory round trip.  The row's byte 0 in shared
     * memory is only brought up to date when the row is written back. */
    uint32_t sreg = 128, s0 = 128;
    uint32_t a = 0;
    int T = 0, LT = 0, L = 0, LL = 0, RT = 0, q0 = 0, q1 = 0, q2 = 0, q3 = 0;
    int qo = 0, sbase = 0, step = (int)sizeof(pix_t);
    (void)q2; (void)q3;
    /* byte pointers of the current line: previous line, the line before it, output row, and
     * (subsampled planes) the private copy of the line being decoded */
    const uint8_t *pl = D.frame, *ppl = D.frame;
    uint8_t *ol = D.frame, *ll = 0;
    uint8_t *const scratch = (uint8_t *)D.lines;
    const size_t lbytes = (size_t)D.line_stride * sizeof(int32_t);    /* one scratch line */
    (void)row_;
#if defined(__CUDA_ARCH__)
    const int gate_wait = D.gate_wait;
    int waited = 0;
    uint32_t row_sa = (uint32_t)__cvta_generic_to_shared(&ff_s_rows[threadIdx.x * FF_ROW_WORDS]);
    uint32_t tab_sa = (uint32_t)__cvta_generic_to_shared(ff_s_tab16);
    uint32_t qt_sa = (uint32_t)__cvta_generic_to_shared(ff_s_qt);
    row_sa = ff_opaque(row_sa);                      /* see ff_encode_slice_range */
    tab_sa = ff_opaque(tab_sa);
    qt_sa = ff_opaque(qt_sa);
    uint32_t q_sa = qt_sa;                           /* quant table set of the current line */
#define FF_QTL(i) ff_lds16s(q_sa + 2u * (uint32_t)(i))
#define FF_ST_LD(sl) ff_lds8(row_sa + (uint32_t)(sl))
#define FF_ST_ST(sl, v) ff_sts8(row_sa + (uint32_t)(sl), (v))
#define FF_TAB16(st) ff_lds16u(tab_sa + 2u * (st))
#else
#define FF_QTL(i) FF_QT(qo, i)
#define FF_ST_LD(sl) ((uint32_t)FF_ROWB(sl))
#define FF_ST_ST(sl, v) (FF_ROWB(sl) = (uint8_t)(v))
#define FF_TAB16(st) ((uint32_t)FF_TAB(st) | ((uint32_t)FF_TAB(256 + (st)) << 8))
#endif

    /* zero rows: line 0 of plane 0's scratch is the row "above the slice" of every
     * full-resolution plane; subsampled planes start from two zeroed private lines */
    for (int k = 0; k < (live ? P.ncoded : 0); k++) {
        const int sub = P.cp[k].hs || P.cp[k].vs;
        if (k && !sub)
            continue;
        const int pw = ff_crshift(d.w, P.cp[k].hs);
        const int nw = (pw * 2 + 3) / 4 + 1;
        for (int l = 0; l < (sub ? 2 : 1); l++) {
            uint32_t *z = (uint32_t *)(scratch + ((size_t)k * 2 + l) * lbytes);
            for (int i = 0; i < nw; i++)
                z[i] = 0;
        }
    }
    nbyte = live ? buf[pos] : 0;
    if (!live || !ff_line_first(P, d, &it))
        state = 0;

    for (;;) {
        int s, r1, bit, done, diff;
#if defined(__CUDA_ARCH__)
        {
            /* Set-up gating, see ff_decode_slice_range.  Two warp-wide votes tell whether
             * some lane needs a sample and whether some lane is inside a symbol; lanes whose
             * slice is finished stay in the loop and vote for neither, so no active mask has
             * to be tracked and the whole warp leaves together.  While both kinds exist the
             * lanes that need a sample idle, at most gate_wait iterations.  `waited` is the
             * same in every lane. */
            const bool any_need = __any_sync(0xffffffffu, state == 1);
            const bool any_busy = __any_sync(0xffffffffu, state == 2);
            if (!any_need && !any_busy)
                break;
            const bool hold = any_need && any_busy && waited < gate_wait;
            waited = hold ? waited + 1 : 0;
            if (state == 0 || (hold && state == 1))
                continue;
        }
#else
        if (state == 0)
            break;
#endif
        if (state == 1) {
            int ctx;
            if (x < 0 || x == w) {
                if (x == w && !ff_line_next(P, d, &it)) {
                    state = 0;
                    continue;
                }
                {
                    const FFDevPlane cp = P.cp[it.k];
                    const int sub = cp.hs || cp.vs;
                    uint8_t *row = D.frame + P.plane_off[cp.mem] +
                                   (size_t)((d.y >> cp.vs) + it.y) * P.pitch[cp.mem] +
                                   (size_t)(d.x >> cp.hs) * cp.step + cp.off;
                    w = it.w;
                    qo = d.qidx[cp.set] * FF_QT_STRIDE;
#if defined(__CUDA_ARCH__)
                    q_sa = qt_sa + 2u * (uint32_t)qo;
#endif
                    five = FIVE ? FF_QT(qo, FF_MAX_CTX_INPUTS * 256) : 0;
                    sbase = P.set_base[cp.set];
                    ol = row;
                    /* bytes between samples: 2 only for the interleaved gray+alpha layout,
                     * whose planes are never subsampled; the zero row is sized for it */
                    step = SMODE == 1 ? cp.step : 2;
                    if (sub) {
                        uint8_t *l0 = scratch + (size_t)it.k * 2 * lbytes;
                        ll = (it.y & 1) ? l0 + lbytes : l0;
                        pl = (it.y & 1) ? l0 : l0 + lbytes;
                        ppl = ll;                     /* still holds line y-2 (or zeros) */
                    } else {
                        ll = 0;
                        pl = it.y >= 1 ? row - P.pitch[cp.mem] : scratch;
                        ppl = it.y >= 2 ? row - 2 * (size_t)P.pitch[cp.mem] : scratch;
                    }
                }
                x = 0;
                T = ff_pix_load<SMODE>(pl);
                LT = ff_pix_load<SMODE>(ppl);
                L = T;
                LL = 0;
                /* look-ahead on the previous line: q0.. = prev[min(x+1.., w-1)] */
                q0 = ff_pix_load<SMODE>(pl + (size_t)ff_min(1, w - 1) * step);
                q1 = ff_pix_load<SMODE>(pl + (size_t)ff_min(2, w - 1) * step);
#if FF_DEC_AHEAD == 4
                q2 = ff_pix_load<SMODE>(pl + (size_t)ff_min(3, w - 1) * step);
                q3 = ff_pix_load<SMODE>(pl + (size_t)ff_min(4, w - 1) * step);
#endif
                if (overread > 2) {                  /* is_input_end at line start */
                    err = 1;
                    state = 0;
                    continue;
                }
            } else if (!(x & 1023) && overread > 2) {
                err = 1;
                state = 0;
                continue;
            }
            RT = q0;
            q0 = q1;
#if FF_DEC_AHEAD == 4
            q1 = q2;
            q2 = q3;
            q3 = ff_pix_load<SMODE>(pl + (size_t)ff_min(x + 5, w - 1) * step);
#else
            q1 = ff_pix_load<SMODE>(pl + (size_t)ff_min(x + 3, w - 1) * step);
#endif
            ctx = FF_QTL((L - LT) & 0xFF) + FF_QTL(256 + ((LT - T) & 0xFF)) +
                  FF_QTL(512 + ((T - RT) & 0xFF));
            if (FIVE && five) {
                const int TT = ff_pix_load<SMODE>(ppl + (size_t)x * step);
                ctx += FF_QTL(768 + ((LL - L) & 0xFF)) + FF_QTL(1024 + ((TT - T) & 0xFF));
            }
            sign = ctx < 0;
            ctx = sbase + (sign ? -ctx : ctx);
            if (ctx != cur_ctx) {
                if (cur_ctx >= 0) {
                    FF_ST_ST(0, s0);
                    ff_row_store(FF_ROWW, D.rstate + (size_t)cur_ctx * FF_CONTEXT_SIZE);
                }
                if (D.touched) {
                    const uint32_t tw = D.touched[ctx >> 5], tb = 1u << (ctx & 31);
                    if (tw & tb) {
                        ff_row_load(FF_ROWW, D.rstate + (size_t)ctx * FF_CONTEXT_SIZE);
                    } else {
                        uint32_t *rw = FF_ROWW;
                        D.touched[ctx >> 5] = tw | tb;
                        rw[0] = rw[1] = rw[2] = rw[3] = rw[4] = rw[5] = rw[6] = rw[7] = 0x80808080u;
                    }
                } else {
                    ff_row_load(FF_ROWW, D.rstate + (size_t)ctx * FF_CONTEXT_SIZE);
                }
                s0 = FF_ST_LD(0);
                cur_ctx = ctx;
            }
            state = 2;
            slot = 0;
            sreg = s0;
        }
        /* one binary decision: get_rac + refill, rangecoder.h:123-152 */
        {
            const uint32_t t16 = FF_TAB16(sreg);     /* both successor states, fetched early */
            uint32_t ns;
            s = (int)sreg;
            r1 = (range * s) >> 8;
            range -= r1;
            bit = low >= range;
            ns = bit ? (t16 & 0xFFu) : (t16 >> 8);
            if (slot == 0)
                s0 = ns;
            else
                FF_ST_ST(slot, ns);
            sreg = ns;                               /* if the next decision uses the same slot */
        }
        low -= bit ? range : 0;
        range = bit ? r1 : range;
        if (range < 0x100) {
#if defined(__CUDA_ARCH__)
            const uint4 rf = ff_dec_refill(low, pos, overread, nbyte, buf, end);
            range <<= 8;
            low = (int)rf.x;
            pos = rf.y;
            overread = (int)rf.z;
            nbyte = rf.w;
#else
            const int in = pos < end;
            range <<= 8;
            low = (low << 8) + (in ? (int)nbyte : 0);
            pos += (uint32_t)in;
            overread += !in;
            nbyte = buf[pos];                        /* the arena is padded past `end` */
#endif
        }
        const int slot_was = slot;
        /* get_symbol_inline (ffv1dec.c:42-64) as a walk over the state slots:
         * 0 zero flag | 1..10 unary exponent | 22..31 mantissa | 11..21 sign */
        done = 0;
        diff = 0;
        if (slot == 0) {
            if (bit)
                done = 1;
            else {
                slot = 1;
                e = 0;
            }
        } else if (slot < 11) {
            if (bit) {
                if (++e > 31) {                      /* get_symbol returns AVERROR_INVALIDDATA */
                    diff = FFRAC_SYMBOL_ERROR;
                    done = 1;
                }
                slot = 1 + ff_min(e, 9);
            } else {
                a = 1;
                mi = e - 1;
                slot = e ? 22 + ff_min(mi, 9) : 11;
            }
        } else if (slot >= 22) {
            a += a + (uint32_t)bit;
            mi--;
            slot = mi < 0 ? 11 + ff_min(e, 10) : 22 + ff_min(mi, 9);
        } else {
            diff = bit ? (int)(0u - a) : (int)a;
            done = 1;
        }
        if (!done && slot != slot_was)               /* state of the next decision, fetched ahead */
            sreg = FF_ST_LD(slot);
        if (done) {
            int v;
            diff = sign ? FF_NEG32(diff) : diff;
            v = (int)(((uint32_t)ff_median3(L, L + T - LT, T) + (uint32_t)diff) & mask);
            v = (int)(int16_t)v;
            *(pix_t *)(ol + (size_t)x * step) = (pix_t)v;     /* decode_plane's store, ffv1dec.c:142-161 */
            if (ll)
                *(pix_t *)(ll + (size_t)x * step) = (pix_t)v;
            LL = L;
            L = v;
            LT = T;
            T = RT;
            x++;
            state = 1;
        }
    }
    if (!live)
        return;
    if (cur_ctx >= 0 && !D.touched) {                /* lazily created states die with the launch */
        FF_ST_ST(0, s0);
        ff_row_store(FF_ROWW, D.rstate + (size_t)cur_ctx * FF_CONTEXT_SIZE);
    }
#undef FF_QTL
#undef FF_ST_LD
#undef FF_ST_ST
#undef FF_TAB16
    {
        FFRacDec c;
        c.buf = buf; c.low = low; c.range = range; c.pos = pos; c.end = end; c.overread = overread;
        /* end-of-slice check, ffv1dec.c:351-359 */
        if (P.version > 2) {
            uint8_t term = 129;
            ffrac_get(&c, D.tab, &term);
        }
        res->end_pos = c.pos;
        res->overread = c.overread;
    }
    res->error = err;
}

/* The same slice decoder for launches that give every slice a warp of its own (streams with a
 * handful of large slices: D.lane_stride == 32, one live lane per warp and at most a few warps
 * per SM).  Nothing diverges there and nothing hides latency either: the launch lasts as long
 * as the lone lane of the heaviest slice needs, i.e. instructions per sample times the
 * dependent-issue latency.  So this form is written for instruction count on one lane:
 * get_symbol_inline (ffv1dec.c:42-64) as straight nested loops instead of the
 * one-decision-per-iteration state machine above (no votes, no per-decision bookkeeping of
 * where the lane is inside a symbol).  Same results, byte for byte. */
template <int SMODE, bool FIVE>
FFGPU_HD void ff_decode_slice_range_planar_lone(const FFDevParams &P, const FFDecSlice &d, const uint8_t *pkt,
                                                const FFDecCtx &D, FFDecResult *res, uint32_t *row_)
{
#if !defined(__CUDA_ARCH__)
    const FFRacTables *tab_ = D.tab;
    const int16_t *qt_all_ = D.qt_all;
#endif
    typedef typename FFPix<SMODE>::type pix_t;
    const uint32_t mask = (1u << P.cbits) - 1;
    const uint8_t *buf = pkt + d.pkt_off;
    int low = d.low, range = d.range, overread = d.overread;
    uint32_t pos = d.pos;
    const uint32_t end = d.size;
    uint32_t nbyte = buf[pos];                        /* buf[pos], fetched ahead */
    FFLineIt it;
    int err = 0, cur_ctx = -1;
    uint32_t s0 = 128;                                /* slot 0 of the current row, see above */
    uint8_t *const scratch = (uint8_t *)D.lines;
    const size_t lbytes = (size_t)D.line_stride * sizeof(int32_t);
    (void)row_;
#if defined(__CUDA_ARCH__)
    uint32_t row_sa = (uint32_t)__cvta_generic_to_shared(&ff_s_rows[threadIdx.x * FF_ROW_WORDS]);
    uint32_t tab_sa = (uint32_t)__cvta_generic_to_shared(ff_s_tab16);
    uint32_t qt_sa = (uint32_t)__cvta_generic_to_shared(ff_s_qt);
    row_sa = ff_opaque(row_sa);                      /* see ff_encode_slice_range */
    tab_sa = ff_opaque(tab_sa);
    qt_sa = ff_opaque(qt_sa);
    uint32_t q_sa = qt_sa;
#define FF_QTL(i) ff_lds16s(q_sa + 2u * (uint32_t)(i))
#define FF_ST_LD(sl) ff_lds8(row_sa + (uint32_t)(sl))
#define FF_ST_ST(sl, v) ff_sts8(row_sa + (uint32_t)(sl), (v))
#define FF_TAB16(st) ff_lds16u(tab_sa + 2u * (st))
#define FF_LONE_REFILL()                                                          \
    do {                                                                          \
        const uint4 rf = ff_dec_refill(low, pos, overread, nbyte, buf, end);      \
        range <<= 8;                                                              \
        low = (int)rf.x;                                                          \
        pos = rf.y;                                                               \
        overread = (int)rf.z;                                                     \
        nbyte = rf.w;                                                             \
    } while (0)
#else
#define FF_QTL(i) FF_QT(qo, i)
#define FF_ST_LD(sl) ((uint32_t)FF_ROWB(sl))
#define FF_ST_ST(sl, v) (FF_ROWB(sl) = (uint8_t)(v))
#define FF_TAB16(st) ((uint32_t)FF_TAB(st) | ((uint32_t)FF_TAB(256 + (st)) << 8))
#define FF_LONE_REFILL()                                                          \
    do {                                                                          \
        const int in = pos < end;                                                 \
        range <<= 8;                                                              \
        low = (low << 8) + (in ? (int)nbyte : 0);                                 \
        pos += (uint32_t)in;                                                      \
        overread += !in;                                                          \
        nbyte = buf[pos];                                                         \
    } while (0)
#endif
/* one binary decision (get_rac + refill, rangecoder.h:123-152) on the state value `st`:
 * the bit in `bit_`, the successor state in `ns_` */
#define FF_LONE_GET(st, bit_, ns_)                                                \
    do {                                                                          \
        const uint32_t st_ = (st);                                                \
        const uint32_t t16_ = FF_TAB16(st_);                                      \
        const int r1_ = (range * (int)st_) >> 8;                                  \
        range -= r1_;                                                             \
        bit_ = low >= range;                                                      \
        ns_ = bit_ ? (t16_ & 0xFFu) : (t16_ >> 8);                                \
        low -= bit_ ? range : 0;                                                  \
        range = bit_ ? r1_ : range;                                               \
        if (FF_UNLIKELY(range < 0x100))                                           \
            FF_LONE_REFILL();                                                     \
    } while (0)

    for (int k = 0; k < P.ncoded; k++) {              /* zero rows, as above */
        const int sub = P.cp[k].hs || P.cp[k].vs;
        if (k && !sub)
            continue;
        const int pw = ff_crshift(d.w, P.cp[k].hs);
        const int nw = (pw * 2 + 3) / 4 + 1;
        for (int l = 0; l < (sub ? 2 : 1); l++) {
            uint32_t *z = (uint32_t *)(scratch + ((size_t)k * 2 + l) * lbytes);
            for (int i = 0; i < nw; i++)
                z[i] = 0;
        }
    }
    if (ff_line_first(P, d, &it)) {
        do {
            const FFDevPlane cp = P.cp[it.k];
            const int sub = cp.hs || cp.vs;
            uint8_t *const ol = D.frame + P.plane_off[cp.mem] +
                                (size_t)((d.y >> cp.vs) + it.y) * P.pitch[cp.mem] +
                                (size_t)(d.x >> cp.hs) * cp.step + cp.off;
            const int w = it.w;
            const int qo = d.qidx[cp.set] * FF_QT_STRIDE;
            const int five = FIVE ? FF_QT(qo, FF_MAX_CTX_INPUTS * 256) : 0;
            const int sbase = P.set_base[cp.set];
            const int step = SMODE == 1 ? cp.step : 2;
            const uint8_t *pl, *ppl;
            uint8_t *ll;
            int T, LT, L, LL = 0, RT, R2, q0, qsum;
#if defined(__CUDA_ARCH__)
            q_sa = qt_sa + 2u * (uint32_t)qo;
#else
            (void)qo;
#endif
            if (sub) {
                uint8_t *l0 = scratch + (size_t)it.k * 2 * lbytes;
                ll = (it.y & 1) ? l0 + lbytes : l0;
                pl = (it.y & 1) ? l0 : l0 + lbytes;
                ppl = ll;
            } else {
                ll = 0;
                pl = it.y >= 1 ? ol - P.pitch[cp.mem] : scratch;
                ppl = it.y >= 2 ? ol - 2 * (size_t)P.pitch[cp.mem] : scratch;
            }
            T = ff_pix_load<SMODE>(pl);
            LT = ff_pix_load<SMODE>(ppl);
            L = T;
            /* the previous line is read three samples ahead of its use */
            RT = ff_pix_load<SMODE>(pl + (size_t)ff_min(1, w - 1) * step);
            R2 = ff_pix_load<SMODE>(pl + (size_t)ff_min(2, w - 1) * step);
            q0 = ff_pix_load<SMODE>(pl + (size_t)ff_min(3, w - 1) * step);
            if (overread > 2) {                       /* is_input_end at line start */
                err = 1;
                break;
            }
            /* qsum: the context of the coming sample before sign and set offset
             * (get_context, ffv1.h:173-186) */
            qsum = FF_QTL((L - LT) & 0xFF) + FF_QTL(256 + ((LT - T) & 0xFF)) +
                   FF_QTL(512 + ((T - RT) & 0xFF));
            if (FIVE && five) {
                const int TT = ff_pix_load<SMODE>(ppl);
                qsum += FF_QTL(768 + ((LL - L) & 0xFF)) + FF_QTL(1024 + ((TT - T) & 0xFF));
            }
            for (int x = 0; x < w; x++) {
                int bit, diff, v, part, qn;
                uint32_t ns;
                const int sign = qsum < 0;
                const int ctx = sbase + (sign ? -qsum : qsum);
                const int pred = ff_median3(L, L + T - LT, T);
                /* The context of sample x+1 needs this sample's value only for one of its
                 * inputs (two with five inputs).  The rest (`part`) is summed up here, and
                 * so is the v-dependent input under the assumption that the residual is
                 * zero (`qn`, four of five samples in flat pictures): both overlap with the
                 * decision below instead of following it. */
                const int vs = (int)(int16_t)((uint32_t)pred & mask);
                const int nq = ff_pix_load<SMODE>(pl + (size_t)ff_min(x + 4, w - 1) * step);
                if (FF_UNLIKELY(x && !(x & 1023) && overread > 2)) {
                    err = 1;
                    break;
                }
                part = FF_QTL(256 + ((T - RT) & 0xFF)) + FF_QTL(512 + ((RT - R2) & 0xFF));
                qn = FF_QTL((vs - T) & 0xFF);
                if (FIVE && five) {
                    const int TTn = ff_pix_load<SMODE>(ppl + (size_t)ff_min(x + 1, w - 1) * step);
                    part += FF_QTL(1024 + ((TTn - RT) & 0xFF));
                    qn += FF_QTL(768 + ((L - vs) & 0xFF));
                }
                if (ctx != cur_ctx) {
                    if (cur_ctx >= 0) {
                        FF_ST_ST(0, s0);
                        ff_row_store(FF_ROWW, D.rstate + (size_t)cur_ctx * FF_CONTEXT_SIZE);
                    }
                    if (D.touched) {
                        const uint32_t tw = D.touched[ctx >> 5], tb = 1u << (ctx & 31);
                        if (tw & tb) {
                            ff_row_load(FF_ROWW, D.rstate + (size_t)ctx * FF_CONTEXT_SIZE);
                        } else {
                            uint32_t *rw = FF_ROWW;
                            D.touched[ctx >> 5] = tw | tb;
                            rw[0] = rw[1] = rw[2] = rw[3] = rw[4] = rw[5] = rw[6] = rw[7] = 0x80808080u;
                        }
                    } else {
                        ff_row_load(FF_ROWW, D.rstate + (size_t)ctx * FF_CONTEXT_SIZE);
                    }
                    s0 = FF_ST_LD(0);
                    cur_ctx = ctx;
                }
                FF_LONE_GET(s0, bit, ns);             /* zero flag */
                s0 = ns;
                v = vs;
                if (!bit) {
                    int e = 0;
                    diff = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1                                      /* the loop has to stay inside the instruction caches */
#endif
                    for (;;) {                        /* unary exponent */
                        const uint32_t sl = 1u + (uint32_t)ff_min(e, 9);
                        FF_LONE_GET(FF_ST_LD(sl), bit, ns);
                        FF_ST_ST(sl, ns);
                        if (!bit)
                            break;
                        if (++e > 31) {               /* get_symbol returns AVERROR_INVALIDDATA */
                            diff = FFRAC_SYMBOL_ERROR;
                            break;
                        }
                    }
                    if (diff == 0) {
                        uint32_t a = 1;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
                        for (int i = e - 1; i >= 0; i--) {        /* mantissa */
                            const uint32_t sl = 22u + (uint32_t)ff_min(i, 9);
                            FF_LONE_GET(FF_ST_LD(sl), bit, ns);
                            FF_ST_ST(sl, ns);
                            a += a + (uint32_t)bit;
                        }
                        {
                            const uint32_t sl = 11u + (uint32_t)ff_min(e, 10);
                            FF_LONE_GET(FF_ST_LD(sl), bit, ns);
                            FF_ST_ST(sl, ns);
                        }
                        diff = bit ? (int)(0u - a) : (int)a;
                    }
                    diff = sign ? FF_NEG32(diff) : diff;
                    v = (int)(((uint32_t)pred + (uint32_t)diff) & mask);
                    v = (int)(int16_t)v;
                    qn = FF_QTL((v - T) & 0xFF);      /* the assumption did not hold */
                    if (FIVE && five)
                        qn += FF_QTL(768 + ((L - v) & 0xFF));
                }
                *(pix_t *)(ol + (size_t)x * step) = (pix_t)v;
                if (ll)
                    *(pix_t *)(ll + (size_t)x * step) = (pix_t)v;
                qsum = part + qn;
                LL = L;
                L = v;
                LT = T;
                T = RT;
                RT = R2;
                R2 = q0;
                q0 = nq;
            }
            if (err)
                break;
        } while (ff_line_next(P, d, &it));
    }
    if (cur_ctx >= 0 && !D.touched) {                /* lazily created states die with the launch */
        FF_ST_ST(0, s0);
        ff_row_store(FF_ROWW, D.rstate + (size_t)cur_ctx * FF_CONTEXT_SIZE);
    }
#undef FF_QTL
#undef FF_ST_LD
#undef FF_ST_ST
#undef FF_TAB16
#undef FF_LONE_GET
#undef FF_LONE_REFILL
    {
        FFRacDec c;
        c.buf = buf; c.low = low; c.range = range; c.pos = pos; c.end = end; c.overread = overread;
        if (P.version > 2) {                         /* end-of-slice check, ffv1dec.c:351-359 */
            uint8_t term = 129;
            ffrac_get(&c, D.tab, &term);
        }
        res->end_pos = c.pos;
        res->overread = c.overread;
    }
    res->error = err;
}

/* The same for RGB slices (decode_rgb_frame, ffv1dec_template.c:131-200): the planes of a pixel
 * row are coded line by line into int32 line buffers, and the row goes through the inverse
 * colour transform once its last plane is done. */
template <bool FIVE>
FFGPU_HD void ff_decode_slice_range_rgb_lone(const FFDevParams &P, const FFDecSlice &d, const uint8_t *pkt,
                                                const FFDecCtx &D, FFDecResult *res, uint32_t *row_)
{
#if !defined(__CUDA_ARCH__)
    const FFRacTables *tab_ = D.tab;
    const int16_t *qt_all_ = D.qt_all;
#endif
    const int use32 = P.use32;
    const uint32_t mask = (1u << P.cbits) - 1;
    const uint8_t *buf = pkt + d.pkt_off;
    int low = d.low, range = d.range, overread = d.overread;
    uint32_t pos = d.pos;
    const uint32_t end = d.size;
    uint32_t nbyte = buf[pos];                        /* buf[pos], fetched ahead */
    FFLineIt it;
    int err = 0, cur_ctx = -1;
    uint32_t s0 = 128;                                /* slot 0 of the current row, see above */
    (void)row_;
#if defined(__CUDA_ARCH__)
    uint32_t row_sa = (uint32_t)__cvta_generic_to_shared(&ff_s_rows[threadIdx.x * FF_ROW_WORDS]);
    uint32_t tab_sa = (uint32_t)__cvta_generic_to_shared(ff_s_tab16);
    uint32_t qt_sa = (uint32_t)__cvta_generic_to_shared(ff_s_qt);
    row_sa = ff_opaque(row_sa);                      /* see ff_encode_slice_range */
    tab_sa = ff_opaque(tab_sa);
    qt_sa = ff_opaque(qt_sa);
    uint32_t q_sa = qt_sa;
#define FF_QTL(i) ff_lds16s(q_sa + 2u * (uint32_t)(i))
#define FF_ST_LD(sl) ff_lds8(row_sa + (uint32_t)(sl))
#define FF_ST_ST(sl, v) ff_sts8(row_sa + (uint32_t)(sl), (v))
#define FF_TAB16(st) ff_lds16u(tab_sa + 2u * (st))
#define FF_LONE_REFILL()                                                          \
    do {                                                                          \
        const uint4 rf = ff_dec_refill(low, pos, overread, nbyte, buf, end);      \
        range <<= 8;                                                              \
        low = (int)rf.x;                                                          \
        pos = rf.y;                                                               \
        overread = (int)rf.z;                                                     \
        nbyte = rf.w;                                                             \
    } while (0)
#else
#define FF_QTL(i) FF_QT(qo, i)
#define FF_ST_LD(sl) ((uint32_t)FF_ROWB(sl))
#define FF_ST_ST(sl, v) (FF_ROWB(sl) = (uint8_t)(v))
#define FF_TAB16(st) ((uint32_t)FF_TAB(st) | ((uint32_t)FF_TAB(256 + (st)) << 8))
#define FF_LONE_REFILL()                                                          \
    do {                                                                          \
        const int in = pos < end;                                                 \
        range <<= 8;                                                              \
        low = (low << 8) + (in ? (int)nbyte : 0);                                 \
        pos += (uint32_t)in;                                                      \
        overread += !in;                                                          \
        nbyte = buf[pos];                                                         \
    } while (0)
#endif
/* one binary decision (get_rac + refill, rangecoder.h:123-152) on the state value `st`:
 * the bit in `bit_`, the successor state in `ns_` */
#define FF_LONE_GET(st, bit_, ns_)                                                \
    do {                                                                          \
        const uint32_t st_ = (st);                                                \
        const uint32_t t16_ = FF_TAB16(st_);                                      \
        const int r1_ = (range * (int)st_) >> 8;                                  \
        range -= r1_;                                                             \
        bit_ = low >= range;                                                      \
        ns_ = bit_ ? (t16_ & 0xFFu) : (t16_ >> 8);                                \
        low -= bit_ ? range : 0;                                                  \
        range = bit_ ? r1_ : range;                                               \
        if (FF_UNLIKELY(range < 0x100))                                           \
            FF_LONE_REFILL();                                                     \
    } while (0)

    for (int k = 0; k < P.ncoded; k++)                /* every plane starts from two zero lines */
        for (int i = 0; i < 2 * D.line_stride; i++)
            D.lines[(size_t)k * 2 * D.line_stride + i] = 0;
    if (ff_line_first(P, d, &it)) {
        do {
            const FFDevPlane cp = P.cp[it.k];
            int32_t *const l0 = D.lines + (size_t)it.k * 2 * D.line_stride;
            int32_t *const cur = (it.y & 1) ? l0 + D.line_stride : l0;      /* still holds line y-2 */
            const int32_t *const prev = (it.y & 1) ? l0 : l0 + D.line_stride;
            const int w = d.w;
            const int qo = d.qidx[cp.set] * FF_QT_STRIDE;
            const int five = FIVE ? FF_QT(qo, FF_MAX_CTX_INPUTS * 256) : 0;
            const int sbase = P.set_base[cp.set];
            int T, LT, L, LL = 0, RT, R2, q0, qsum;
#if defined(__CUDA_ARCH__)
            q_sa = qt_sa + 2u * (uint32_t)qo;
#else
            (void)qo;
#endif
            T = prev[0];
            LT = cur[0];
            L = T;
            RT = prev[ff_min(1, w - 1)];
            R2 = prev[ff_min(2, w - 1)];
            q0 = prev[ff_min(3, w - 1)];
            if (overread > 2) {                       /* is_input_end at line start */
                err = 1;
                break;
            }
            qsum = FF_QTL((L - LT) & 0xFF) + FF_QTL(256 + ((LT - T) & 0xFF)) +
                   FF_QTL(512 + ((T - RT) & 0xFF));
            if (FIVE && five)
                qsum += FF_QTL(768 + ((LL - L) & 0xFF)) + FF_QTL(1024 + ((cur[0] - T) & 0xFF));
            for (int x = 0; x < w; x++) {
                int bit, diff, v, part, qn;
                uint32_t ns;
                const int sign = qsum < 0;
                const int ctx = sbase + (sign ? -qsum : qsum);
                const int pred = ff_median3(L, L + T - LT, T);
                /* see ff_decode_slice_range_planar_lone: the context of sample x+1, but for
                 * the inputs that need this sample's value, is summed up ahead of the
                 * decision, and those under the assumption of a zero residual */
                const int vm = (int)((uint32_t)pred & mask);
                const int vs = use32 ? vm : (int)(int16_t)vm;
                const int nq = prev[ff_min(x + 4, w - 1)];
                if (FF_UNLIKELY(x && !(x & 1023) && overread > 2)) {
                    err = 1;
                    break;
                }
                part = FF_QTL(256 + ((T - RT) & 0xFF)) + FF_QTL(512 + ((RT - R2) & 0xFF));
                qn = FF_QTL((vs - T) & 0xFF);
                if (FIVE && five) {
                    part += FF_QTL(1024 + ((cur[ff_min(x + 1, w - 1)] - RT) & 0xFF));
                    qn += FF_QTL(768 + ((L - vs) & 0xFF));
                }
                if (ctx != cur_ctx) {
                    if (cur_ctx >= 0) {
                        FF_ST_ST(0, s0);
                        ff_row_store(FF_ROWW, D.rstate + (size_t)cur_ctx * FF_CONTEXT_SIZE);
                    }
                    ff_row_load(FF_ROWW, D.rstate + (size_t)ctx * FF_CONTEXT_SIZE);
                    s0 = FF_ST_LD(0);
                    cur_ctx = ctx;
                }
                FF_LONE_GET(s0, bit, ns);             /* zero flag */
                s0 = ns;
                v = vs;
                if (!bit) {
                    int e = 0;
                    diff = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
                    for (;;) {                        /* unary exponent */
                        const uint32_t sl = 1u + (uint32_t)ff_min(e, 9);
                        FF_LONE_GET(FF_ST_LD(sl), bit, ns);
                        FF_ST_ST(sl, ns);
                        if (!bit)
                            break;
                        if (++e > 31) {               /* get_symbol returns AVERROR_INVALIDDATA */
                            diff = FFRAC_SYMBOL_ERROR;
                            break;
                        }
                    }
                    if (diff == 0) {
                        uint32_t a = 1;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
                        for (int i = e - 1; i >= 0; i--) {        /* mantissa */
                            const uint32_t sl = 22u + (uint32_t)ff_min(i, 9);
                            FF_LONE_GET(FF_ST_LD(sl), bit, ns);
                            FF_ST_ST(sl, ns);
                            a += a + (uint32_t)bit;
                        }
                        {
                            const uint32_t sl = 11u + (uint32_t)ff_min(e, 10);
                            FF_LONE_GET(FF_ST_LD(sl), bit, ns);
                            FF_ST_ST(sl, ns);
                        }
                        diff = bit ? (int)(0u - a) : (int)a;
                    }
                    diff = sign ? FF_NEG32(diff) : diff;
                    v = (int)(((uint32_t)pred + (uint32_t)diff) & mask);
                    v = use32 ? v : (int)(int16_t)v;
                    qn = FF_QTL((v - T) & 0xFF);      /* the assumption did not hold */
                    if (FIVE && five)
                        qn += FF_QTL(768 + ((L - v) & 0xFF));
                }
                cur[x] = v;
                qsum = part + qn;
                LL = L;
                L = v;
                LT = T;
                T = RT;
                RT = R2;
                R2 = q0;
                q0 = nq;
            }
            if (err)
                break;
            if (it.k == P.ncoded - 1) {              /* the pixel row is complete: inverse RCT, store */
                const int o = (it.y & 1) ? D.line_stride : 0;
                ff_store_line_rgb(P, D.frame, d.x, d.y + it.y, d.w,
                                  D.lines + o, D.lines + 2 * D.line_stride + o,
                                  D.lines + 4 * D.line_stride + o,
                                  P.ncoded > 3 ? D.lines + 6 * D.line_stride + o : (const int32_t *)0,
                                  d.rct_by, d.rct_ry, 0);
            }
        } while (ff_line_next(P, d, &it));
    }
    if (cur_ctx >= 0) {
        FF_ST_ST(0, s0);
        ff_row_store(FF_ROWW, D.rstate + (size_t)cur_ctx * FF_CONTEXT_SIZE);
    }
#undef FF_QTL
#undef FF_ST_LD
#undef FF_ST_ST
#undef FF_TAB16
#undef FF_LONE_GET
#undef FF_LONE_REFILL
    {
        FFRacDec c;
        c.buf = buf; c.low = low; c.range = range; c.pos = pos; c.end = end; c.overread = overread;
        if (P.version > 2) {                         /* end-of-slice check, ffv1dec.c:351-359 */
            uint8_t term = 129;
            ffrac_get(&c, D.tab, &term);
        }
        res->end_pos = c.pos;
        res->overread = c.overread;
    }
    res->error = err;
}

/* A version 4 slice with slice_coding_mode == 1 ("PCM", decode_line ffv1dec_template.c:37-47):
 * every sample is `bits` raw decisions, each with a fresh state of 128; no prediction, no
 * contexts, no colour transform.  The reference encoder falls back to it when a slice
 * outgrows its buffer (ffv1enc.c:1107-1117); such slices are rare and run the plain way. */
FFGPU_HD void ff_decode_slice_pcm(const FFDevParams &P, const FFDecSlice &d, const uint8_t *pkt,
                                  const FFDecCtx &D, FFDecResult *res)
{
    FFRacDec c;
    FFLineIt it;
    /* decode_rgb_frame: bits + (slice_coding_mode != 1) -> the raw depth; decode_plane: bits */
    const int bits = P.colorspace ? P.sbits : P.cbits;
    int err = 0;
    c.buf = pkt + d.pkt_off;
    c.low = d.low;
    c.range = d.range;
    c.pos = d.pos;
    c.end = d.size;
    c.overread = d.overread;
    if (ff_line_first(P, d, &it)) {
        do {
            const int w = P.colorspace ? d.w : it.w;
            int32_t *line = D.lines + (size_t)it.k * 2 * D.line_stride;
            if (c.overread > 2) {                    /* is_input_end */
                err = 1;
                break;
            }
            for (int x = 0; x < w; x++) {
                int v = 0;
                for (int i = 0; i < bits; i++) {
                    uint8_t st = 128;
                    v += v + ffrac_get(&c, D.tab, &st);
                }
                line[x] = ff_wrap_sample(P, v);
            }
            if (P.colorspace == 0) {
                const FFDevPlane cp = P.cp[it.k];
                ff_store_line_ycc(P, D.frame, it.k, d.x >> cp.hs, (d.y >> cp.vs) + it.y, w, line);
            } else if (it.k == P.ncoded - 1) {
                ff_store_line_rgb(P, D.frame, d.x, d.y + it.y, d.w, D.lines, D.lines + 2 * D.line_stride,
                                  D.lines + 4 * D.line_stride,
                                  P.ncoded > 3 ? D.lines + 6 * D.line_stride : (const int32_t *)0, 1, 1, 1);
            }
        } while (ff_line_next(P, d, &it));
    }
    if (P.version > 2) {                             /* end-of-slice check, ffv1dec.c:351-359 */
        uint8_t term = 129;
        ffrac_get(&c, D.tab, &term);
    }
    res->end_pos = c.pos;
    res->overread = c.overread;
    res->error = err;
}

/* which specialisation of ff_decode_slice_range_planar serves the stream: 0 none (generic
 * path), 1 one byte per sample, 2 LSB-packed little-endian u16 */
FFGPU_HD int ff_decode_planar_mode(const FFDevParams *P)
{
    if (P->ac == FF_AC_GOLOMB || P->colorspace != 0 || P->use32)
        return 0;
    if (P->sbits <= 8)
        return 1;
    return (P->packed_lsb || P->sbits == 16) ? 2 : 0;
}

FFGPU_HD void ff_decode_slice(const FFDevParams &P, const FFDecSlice &d, const uint8_t *pkt,
                              const FFDecCtx &D, FFDecResult *res, uint32_t *row)
{
    if (d.pcm) {
        ff_decode_slice_pcm(P, d, pkt, D, res);
    } else if (P.ac == FF_AC_GOLOMB) {
        ff_decode_slice_golomb(P, d, pkt, D, res);
    } else if (ff_decode_planar_mode(&P)) {
        /* planar YCbCr / gray (+alpha), 8-bit or LSB-packed 16-bit containers */
        if (D.lone) {
            if (P.sbits <= 8) {
                if (D.any_five) ff_decode_slice_range_planar_lone<1, true>(P, d, pkt, D, res, row);
                else            ff_decode_slice_range_planar_lone<1, false>(P, d, pkt, D, res, row);
            } else {
                if (D.any_five) ff_decode_slice_range_planar_lone<2, true>(P, d, pkt, D, res, row);
                else            ff_decode_slice_range_planar_lone<2, false>(P, d, pkt, D, res, row);
            }
        } else if (P.sbits <= 8) {
            if (D.any_five) ff_decode_slice_range_planar<1, true>(P, d, pkt, D, res, row, 1);
            else            ff_decode_slice_range_planar<1, false>(P, d, pkt, D, res, row, 1);
        } else {
            if (D.any_five) ff_decode_slice_range_planar<2, true>(P, d, pkt, D, res, row, 1);
            else            ff_decode_slice_range_planar<2, false>(P, d, pkt, D, res, row, 1);
        }
#if !defined(__CUDA_ARCH__)          /* on the device k_decode<0, FIVE, true> calls it, see there */
    } else if (D.lone && P.colorspace != 0) {
        if (D.any_five) ff_decode_slice_range_rgb_lone<true>(P, d, pkt, D, res, row);
        else            ff_decode_slice_range_rgb_lone<false>(P, d, pkt, D, res, row);
#endif
    } else {
        ff_decode_slice_range(P, d, pkt, D, res, row);
    }
}

#endif /* FFGPU_FFV1_SLICE_CUH */
