/*
 * tests/emul/fake_kernels.cpp -- TEST INFRASTRUCTURE: the launchers of ffv1_launch.h executed
 * on the CPU, one loop iteration per work item, with the product's own __host__ __device__
 * slice functions (csrc/ffv1_slice.cuh) doing the work -- the same functions the CUDA kernels
 * of ffv1_kernels.cu wrap.  Together with fake_cuda.cpp it stands in for the device under
 * the unmodified ffgpu_api.cu, so that the library's HOST side can be driven through the
 * public C ABI in a container without a GPU (tests/test_host_pipeline_cpu.py).  The form of
 * the slice coders follows the launch shape exactly like the real launchers (one slice per
 * warp -> straight-line coders, records configured -> stage B in two halves).
 * Never part of the product: libffgpu.so contains ffv1_kernels.cu and nothing of this.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../ffmpeg_ffv2_b200/csrc/ffv1_host.h"
#include "../../ffmpeg_ffv2_b200/csrc/ffv1_launch.h"
#include "../../ffmpeg_ffv2_b200/csrc/ffv1_slice.cuh"

static int lone_allowed(void)
{
    const char *v = getenv("FFGPU_LONE");
    return !(v && !atoi(v));
}

extern "C" size_t ffk_sort_tmp_bytes(int n) { return (size_t)n * 8 + 256; }

extern "C" int ffk_encode_group(const FFDevParams *Pp, const FFEncDev *E, int nframes, ffk_stream)
{
    const FFDevParams &P = *Pp;
    const int golomb = P.ac == FF_AC_GOLOMB;
    int launches = 0;
    if (nframes <= 0 || nframes > 1024 || P.nslices > 1024)
        return FFGPU_EINVAL;
    const int n = nframes * P.nslices;
    /* version 4: k_rct_stat + k_rct_pick */
    if (P.colorspace != 0 && E->rct) {
        for (int f = 0; f < nframes; f++)
            for (int s = 0; s < P.nslices; s++) {
                const FFDevSlice &sl = E->slices[s];
                const uint8_t *frame = E->frames + (size_t)f * P.frame_bytes;
                int32_t stat[FF_RCT_CANDIDATES] = { 0 };
                for (int y = 1; y < sl.h; y++)
                    for (int x = 1; x < sl.w; x++) {
                        int32_t v[FF_RCT_CANDIDATES];
                        ff_rct_pixel_stat(P, frame, sl.x, sl.y, x, y, v);
                        for (int k = 0; k < FF_RCT_CANDIDATES; k++)
                            stat[k] = (int32_t)((uint32_t)stat[k] + (uint32_t)v[k]);
                    }
                int *rct = E->rct + 2 * ((size_t)f * P.nslices + s);
                ff_rct_pick(stat, &rct[0], &rct[1]);
            }
        launches += 2;
    }
    /* stage A */
    for (int f = 0; f < nframes; f++)
        for (int s = 0; s < P.nslices; s++) {
            const FFDevSlice &sl = E->slices[s];
            const uint8_t *frame = E->frames + (size_t)f * P.frame_bytes;
            uint32_t *tok = E->tokens + (size_t)f * P.frame_tokens + sl.tok_off;
            const int *rct = E->rct ? E->rct + 2 * ((size_t)f * P.nslices + s) : nullptr;
            uint32_t w = 0;
            for (uint32_t t = 0; t < sl.ntok; t++) {
                tok[t] = ff_symbolize_index(P, sl, frame, E->qt, t, rct ? rct[0] : 1, rct ? rct[1] : 1);
                w += ff_token_weight(tok[t]);
            }
            if (E->weight)
                E->weight[(size_t)f * P.nslices + s] = w;
        }
    launches++;
    /* k_fill_state / k_fill_state_initial */
    {
        const size_t per_slice = (size_t)P.total_ctx * (golomb ? 8 : FF_CONTEXT_SIZE);
        for (int f = 0; f < nframes; f++) {
            if (!E->frame_key[f])
                continue;
            uint8_t *p = E->state + (E->state_per_frame ? (size_t)f * per_slice * P.nslices : 0);
            for (int s = 0; s < P.nslices; s++) {
                uint8_t *q = p + (size_t)s * per_slice;
                if (golomb) {
                    uint2 *v = (uint2 *)q;
                    for (int c = 0; c < P.total_ctx; c++) { v[c].x = FF_VLC_INIT_LO; v[c].y = FF_VLC_INIT_HI; }
                } else if (E->initial) {
                    memcpy(q, E->initial, per_slice);
                } else {
                    memset(q, 128, per_slice);
                }
            }
        }
        launches++;
    }
    /* the sort and k_sched only choose an execution order: items run in index order here */
    /* stage B */
    {
        const bool lone = !golomb && E->lane_stride == 32 && !E->rc_stat && lone_allowed();
        bool split = !golomb && !lone && E->rec && E->weight;
        if (split) {                                   /* k_rec_offsets */
            unsigned long long sum = 0;
            for (int i = 0; i < n; i++) {
                E->rec_off[i] = sum;
                sum += E->weight[i];
            }
            E->rec_off[n] = sum;
            *E->split_ok = sum <= E->rec_cap;
            split = *E->split_ok != 0;
            launches += 3;
        }
        FFPassStats pass;
        pass.rc_stat = E->rc_stat;
        pass.rc_stat2 = E->rc_stat2;
        pass.ctx_count = E->stat_ctx_count;
        for (int gid = 0; gid < n; gid++) {
            const int f = gid / P.nslices, s = gid - f * P.nslices;
            const FFDevSlice &sl = E->slices[s];
            const size_t st_slot = (size_t)(E->state_per_frame ? f : 0) * P.nslices + s;
            const uint32_t *tok = E->tokens + (size_t)f * P.frame_tokens + sl.tok_off;
            const FFRacPrefix &pre = E->prefix[(size_t)E->frame_prefix_set[f] * P.nslices + s];
            uint8_t *bs = E->bs + (size_t)f * P.frame_bs + sl.bs_off;
            const int *rct = E->rct ? E->rct + 2 * (size_t)gid : nullptr;
            const uint32_t v4cap = E->rct ? (uint32_t)((16384 + (int64_t)P.width * P.height * 12) / P.nslices) : 0u;
            uint32_t ovf = 0, bytes;
            alignas(16) uint32_t row[FF_ROW_WORDS];
            if (golomb) {
                bytes = ff_encode_slice_golomb(P, sl, tok, (uint2 *)E->state + st_slot * P.total_ctx, pre,
                                               E->prefix_bytes, bs, &ovf, E->tab, rct);
            } else {
                uint8_t *rows = E->state + st_slot * P.total_ctx * FF_CONTEXT_SIZE;
                if (lone) {
                    bytes = ff_encode_slice_range_lone(sl, tok, rows, E->tab, pre, E->prefix_bytes, bs, &ovf, row,
                                                       rct, v4cap);
                } else if (split) {                    /* k_chain_states + k_code_records */
                    uint16_t *rec = E->rec + E->rec_off[gid];
                    const uint32_t guard_tok = E->rct ? sl.ntok - (uint32_t)sl.seg_w[sl.nseg - 1] : 0xFFFFFFFFu;
                    uint32_t off = 0;
                    for (uint32_t t = 0; t < sl.ntok; t++) {
                        if (t == guard_tok)
                            E->guard_rec[gid] = off;
                        off += ff_chain_token(tok[t], rows, E->tab, rec + off);
                    }
                    if (off != (uint32_t)(E->rec_off[gid + 1] - E->rec_off[gid]))
                        return FFGPU_EXTERNAL;
                    bytes = ff_encode_slice_records(sl, rec, off, E->rct ? E->guard_rec[gid] : 0xFFFFFFFFu, E->tab, pre,
                                                    E->prefix_bytes, bs, &ovf, rct, v4cap);
                } else if (E->rc_stat) {
                    bytes = ff_encode_slice_range<true>(sl, tok, rows, E->tab, pre, E->prefix_bytes, bs, &ovf, row, rct,
                                                        v4cap, &pass);
                } else {
                    bytes = ff_encode_slice_range<false>(sl, tok, rows, E->tab, pre, E->prefix_bytes, bs, &ovf, row, rct,
                                                         v4cap, nullptr);
                }
            }
            E->slice_bytes[gid] = bytes;
            if (ovf)
                *E->overflow |= 1u;
        }
        launches++;
    }
    /* k_pack_slice_scan, k_pack_frame_scan: the packets may reuse the token array, so every
     * slice is coded before the first byte is packed */
    {
        uint32_t units = 0;
        for (int f = 0; f < nframes; f++) {
            uint32_t off = 0;
            for (int s = 0; s < P.nslices; s++) {
                E->slice_off[(size_t)f * P.nslices + s] = off;
                off += ff_slice_packed_size(P, s, E->slice_bytes[(size_t)f * P.nslices + s]);
            }
            E->pkt_size[f] = off;
            E->pkt_off[f] = units;
            units += (off + 15u) >> 4;
        }
        E->pkt_off[nframes] = units;
    }
    /* k_pack_gather */
    {
        uint32_t crc[256];
        for (int i = 0; i < 256; i++)
            crc[i] = ff_crc_table_entry(i);
        if (!*E->overflow)                             /* overflowed slices have no valid size */
            for (int f = 0; f < nframes; f++)
                for (int s = 0; s < P.nslices; s++)
                    ff_pack_slice(P, s, E->bs + (size_t)f * P.frame_bs + E->slices[s].bs_off,
                                  E->slice_bytes[(size_t)f * P.nslices + s],
                                  E->pkt + ((size_t)E->pkt_off[f] << 4) + E->slice_off[(size_t)f * P.nslices + s], crc);
    }
    launches += 3;
    return launches;
}

extern "C" int ffk_decode_group(const FFDevParams *Pp, const FFDecDev *D, int nframes, ffk_stream)
{
    const FFDevParams &P = *Pp;
    const int golomb = P.ac == FF_AC_GOLOMB;
    if (nframes <= 0)
        return FFGPU_EINVAL;
    const int total = nframes * D->max_slices;
    const size_t per_slice = (size_t)P.total_ctx * (golomb ? 8 : FF_CONTEXT_SIZE);
    if (D->touched) {
        memset(D->touched, 0, (size_t)total * D->touched_words * sizeof(uint32_t));
    } else if (!D->initial && P.version < 4 && ((size_t)D->max_slices * per_slice) % 16 == 0) {
        /* k_dec_fill_state: the whole arena of a key frame */
        for (int f = 0; f < nframes; f++) {
            if (!D->work[(size_t)f * D->max_slices].key_frame)
                continue;
            uint8_t *p = D->state + (D->state_per_frame ? (size_t)f * D->max_slices * per_slice : 0);
            if (golomb) {
                uint2 *v = (uint2 *)p;
                for (size_t c = 0; c < (size_t)D->max_slices * P.total_ctx; c++) { v[c].x = FF_VLC_INIT_LO; v[c].y = FF_VLC_INIT_HI; }
            } else {
                memset(p, 128, (size_t)D->max_slices * per_slice);
            }
        }
    } else {
        /* k_dec_init_state: per slice, by quant table index */
        for (int f = 0; f < nframes; f++)
            for (int s = 0; s < D->nslices[f]; s++) {
                const FFDecSlice &w = D->work[(size_t)f * D->max_slices + s];
                if (!w.key_frame || w.skip)
                    continue;
                const size_t slot = (size_t)(D->state_per_frame ? f : 0) * D->max_slices + s;
                if (golomb) {
                    uint2 *v = (uint2 *)D->state + slot * P.total_ctx;
                    for (int c = 0; c < P.total_ctx; c++) { v[c].x = FF_VLC_INIT_LO; v[c].y = FF_VLC_INIT_HI; }
                } else {
                    uint8_t *p = D->state + slot * P.total_ctx * FF_CONTEXT_SIZE;
                    const size_t per_set = (size_t)D->max_ctx * FF_CONTEXT_SIZE;
                    for (int set = 0; set < P.nsets; set++) {
                        if (D->initial)
                            memcpy(p + set * per_set, D->initial + (size_t)w.qidx[set] * per_set, per_set);
                        else
                            memset(p + set * per_set, 128, per_set);
                    }
                }
            }
    }
    if (D->wide_used)
        *D->wide_used = 0;
    const int planar = D->generic ? 0 : ff_decode_planar_mode(&P);
    const bool lone = D->lane_stride == 32 && lone_allowed() &&
                      (planar || (!D->generic && P.colorspace && !golomb));
    uint32_t crc[256];
    for (int i = 0; i < 256; i++)
        crc[i] = ff_crc_table_entry(i);
    for (int gid = 0; gid < total; gid++) {
        const int f = gid / D->max_slices, s = gid - f * D->max_slices;
        FFDecResult r;
        memset(&r, 0, sizeof(r));
        r.flags = FF_RES_NOT_DECODED;
        if (s < D->nslices[f]) {
            FFDecSlice w = D->work[gid];
            r.x = w.x; r.y = w.y; r.w = w.w; r.h = w.h;
            if (w.parse && !w.skip) {
                ff_dec_slice_header(P, D->hdr, &w, D->pkt, D->tab, crc, &r);
                r.x = w.x; r.y = w.y; r.w = w.w; r.h = w.h;
            }
            r.size = w.size;
            if (!w.skip) {
                FFDecCtx C;
                memset(&C, 0, sizeof(C));
                r.flags &= ~FF_RES_NOT_DECODED;
                const size_t slot = (size_t)(D->state_per_frame ? f : 0) * D->max_slices + s;
                C.qt_all = D->qt;
                C.tab = D->tab;
                C.rstate = D->state + slot * P.total_ctx * FF_CONTEXT_SIZE;
                C.vstate = (uint2 *)D->state + slot * P.total_ctx;
                C.lines = D->lines + (size_t)gid * P.ncoded * 2 * D->line_stride;
                C.line_stride = D->line_stride;
                C.frame = D->frames + (size_t)f * P.frame_bytes;
                C.gate_wait = D->gate_wait;
                C.touched = D->touched ? D->touched + (size_t)gid * D->touched_words : nullptr;
                C.any_five = D->any_five;
                C.lone = lone;
                if (w.w + 8 > D->line_stride) {
                    const uint32_t slot_w = D->wide_used ? (*D->wide_used)++ : 0xFFFFFFFFu;
                    if (slot_w < (uint32_t)D->wide_count) {
                        C.lines = D->wide_lines + (size_t)slot_w * P.ncoded * 2 * D->wide_stride;
                        C.line_stride = D->wide_stride;
                    } else {
                        C.lines = nullptr;
                        r.flags |= FF_RES_HDR_BAD | FF_RES_NOT_DECODED;
                    }
                }
                if (C.lines) {
                    alignas(16) uint32_t row[FF_ROW_WORDS];
                    if (D->generic && !w.pcm && !golomb)   /* FFGPU_DEC_GENERIC: k_decode<0> for planar streams too */
                        ff_decode_slice_range(P, w, D->pkt, C, &r, row);
                    else
                        ff_decode_slice(P, w, D->pkt, C, &r, row);
                }
            }
        }
        D->result[gid] = r;
    }
    return 3;
}

extern "C" int ffk_conceal_rect(const FFDevParams *P, uint8_t *dst, const uint8_t *src, int x, int y, int w, int h,
                                int pixshift, ffk_stream)
{
    int nplanes = 0;
    for (int p = 0; p < FF_MAX_PLANES; p++)
        if (P->rows[p])
            nplanes = p + 1;
    const int planar_chroma = P->layout == FF_LAY_PLANAR && P->chroma_planes;
    for (int p = 0; p < nplanes; p++) {
        const int sh = (planar_chroma && (p == 1 || p == 2)) ? P->hs : 0;
        const int sv = (planar_chroma && (p == 1 || p == 2)) ? P->vs : 0;
        const int bpp = P->layout == FF_LAY_PLANAR ? (P->sbits > 8 ? 2 : 1) :
                        P->layout == FF_LAY_GBRP ? 2 : P->layout == FF_LAY_YA8 ? 2 : P->rgb_pixbytes;
        const int bw = ff_crshift(w, sh) * bpp, rows = ff_crshift(h, sv);
        const size_t xo = (size_t)((x >> sh) << pixshift);
        for (int r = 0; r < rows; r++) {
            const size_t o = P->plane_off[p] + (size_t)((y >> sv) + r) * P->pitch[p] + xo;
            memcpy(dst + o, src + o, bw);
        }
    }
    return 1;
}

extern "C" int ffk_copy_segments(const FFCopyArgs *a, ffk_stream)
{
    for (int s = 0; s < a->nseg; s++) {
        size_t bytes = a->seg[s].bytes;
        if (s == 0 && a->dyn_bytes) {
            bytes = (size_t)*a->dyn_bytes << a->dyn_shift;
            if (bytes > a->dyn_cap)
                bytes = 0;
        }
        memcpy(a->seg[s].dst, a->seg[s].src, bytes);
    }
    return 1;
}
