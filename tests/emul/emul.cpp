/*
 * tests/emul/emul.cpp -- TEST INFRASTRUCTURE: runs the product's __host__ __device__ slice
 * functions (ffmpeg_ffv2_b200/csrc/ffv1_slice.cuh) on the CPU, one loop iteration per CUDA
 * thread, so that the device logic can be checked against the oracle in a container that
 * has no GPU.  It is built into tests/emul/libffv1_emul.so, loaded only by the CPU tests,
 * and is NOT a fallback: libffgpu.so neither contains nor loads it.
 *
 * API shape = the oracle's (ffv1emul_ prefix) so tests/cpucodec.py can drive it.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../ffmpeg_ffv2_b200/csrc/ffv1_host.h"
#include "../../ffmpeg_ffv2_b200/csrc/ffv1_slice.cuh"

struct Params {
    int width, height; const char *pix_fmt;
    int slices, level, gop_size, coder, context, slicecrc, strict, threads, bits_per_raw_sample;
};

static std::vector<int16_t> make_qt(const FFStream &s)
{
    std::vector<int16_t> q((size_t)FF_MAX_QUANT_TABLES * FF_QT_STRIDE, 0);
    for (int i = 0; i < s.qt_count; i++) {
        memcpy(&q[(size_t)i * FF_QT_STRIDE], s.qt[i], sizeof(s.qt[i]));
        q[(size_t)i * FF_QT_STRIDE + FF_MAX_CTX_INPUTS * 256] = s.qt[i][3][127] || s.qt[i][4][127];
    }
    return q;
}

static void upload(const FFDevParams &P, const FFPixFmt *pf, int w, int h,
                   const uint8_t *const planes[4], const int ls[4], uint8_t *frame)
{
    for (int k = 0; k < pf->nplanes; k++) {
        int rb, rows;
        ff_plane_geometry(pf, w, h, k, &rb, &rows);
        for (int y = 0; y < rows; y++)
            memcpy(frame + P.plane_off[k] + (size_t)y * P.pitch[k], planes[k] + (ptrdiff_t)y * ls[k], rb);
    }
}

struct Enc {
    FFStream s;
    FFDevParams P;
    std::vector<FFDevSlice> sl;
    std::vector<int16_t> qt;
    std::vector<uint8_t> frame, rstate, bs, prebytes;
    std::vector<uint2> vstate;
    std::vector<uint32_t> tokens, crc;
    uint8_t *extradata = nullptr;
    int extradata_size = 0, gop = 12, pic = 0;
    int ps = 3, sar_num = 0, sar_den = 1;       /* per-picture slice header fields, ffv1enc.c:944-949 */
    /* two-pass coding */
    int pass1 = 0, gob_count = 0;
    std::vector<unsigned long long> rc_stat, rc_stat2;
};

static int g_pass1, g_pass2;
static const char *g_stats_in;
extern "C" void *ffv1emul_encoder_open(const Params *p, int *err);
extern "C" void *ffv1emul_encoder_open2(const Params *p, int pass1, int pass2, const char *stats_in, int *err)
{
    g_pass1 = pass1; g_pass2 = pass2; g_stats_in = stats_in;
    void *h = ffv1emul_encoder_open(p, err);
    g_pass1 = g_pass2 = 0; g_stats_in = nullptr;
    return h;
}

extern "C" void *ffv1emul_encoder_open(const Params *p, int *err)
{
    ffgpu_enc_options o;
    memset(&o, 0, sizeof(o));
    o.pass1 = g_pass1; o.pass2 = g_pass2; o.stats_in = g_stats_in;
    o.width = p->width; o.height = p->height; o.pix_fmt = p->pix_fmt; o.slices = p->slices;
    o.level = p->level; o.gop_size = p->gop_size; o.coder = p->coder; o.context = p->context;
    o.slicecrc = p->slicecrc; o.strict_std_compliance = p->strict;
    o.bits_per_raw_sample = p->bits_per_raw_sample;
    Enc *e = new Enc;
    *err = ff_stream_from_options(&e->s, &o);
    if (*err >= 0)
        *err = ff_write_extradata(&e->s, o.gop_size, &e->extradata, &e->extradata_size);
    if (*err < 0) { delete e; return nullptr; }
    e->gop = o.gop_size;
    e->pass1 = o.pass1;
    e->rc_stat.assign(512, 0);
    e->rc_stat2.assign((size_t)e->s.ctx_count[e->s.context_model] * 64, 0);
    e->sl.resize(e->s.nh * e->s.nv);
    ff_fill_dev_params(&e->s, 1, &e->P, e->sl.data());
    e->qt = make_qt(e->s);
    e->frame.resize(e->P.frame_bytes);
    e->tokens.resize(e->P.frame_tokens);
    e->bs.resize(e->P.frame_bs);
    e->rstate.assign((size_t)e->P.nslices * e->P.total_ctx * FF_CONTEXT_SIZE, 128);
    e->vstate.resize((size_t)e->P.nslices * e->P.total_ctx);
    e->crc.resize(256);
    for (int i = 0; i < 256; i++) e->crc[i] = ff_crc_table_entry(i);
    return e;
}

extern "C" int ffv1emul_encoder_extradata(void *h, const uint8_t **d)
{
    Enc *e = (Enc *)h; *d = e->extradata; return e->extradata_size;
}

extern "C" void ffv1emul_encoder_info(void *h, int info[8])
{
    Enc *e = (Enc *)h;
    info[0] = e->s.version; info[1] = e->s.micro_version; info[2] = e->s.ac; info[3] = e->s.nh;
    info[4] = e->s.nv; info[5] = e->s.ec; info[6] = e->s.bits; info[7] = e->s.colorspace;
}

extern "C" int ffv1emul_encode(void *h, const uint8_t *const planes[4], const int ls[4],
                               uint8_t *out, int cap, int *key)
{
    Enc *e = (Enc *)h;
    const FFDevParams &P = e->P;
    const int keyf = e->gop == 0 || e->pic % e->gop == 0;
    upload(P, e->s.pf, e->s.width, e->s.height, planes, ls, e->frame.data());
    std::vector<FFRacPrefix> pre(P.nslices);
    e->prebytes.assign((size_t)P.nslices * 2048, 0);
    std::vector<uint32_t> bytes(P.nslices);
    /* version 4: choose_rct_params per slice, one "thread" per pixel */
    std::vector<int> rct((size_t)P.nslices * 2, 1);
    if (P.version > 3)
        for (int i = 0; i < P.nslices; i++) {
            int32_t stat[FF_RCT_CANDIDATES] = { 0 };
            for (int y = 1; y < e->sl[i].h; y++)
                for (int x = 1; x < e->sl[i].w; x++) {
                    int32_t v[FF_RCT_CANDIDATES];
                    ff_rct_pixel_stat(P, e->frame.data(), e->sl[i].x, e->sl[i].y, x, y, v);
                    for (int k = 0; k < FF_RCT_CANDIDATES; k++)
                        stat[k] = (int32_t)((uint32_t)stat[k] + (uint32_t)v[k]);
                }
            ff_rct_pick(stat, &rct[2 * i], &rct[2 * i + 1]);
        }
    /* stage A: one "thread" per token */
    for (int i = 0; i < P.nslices; i++)
        for (uint32_t t = 0; t < e->sl[i].ntok; t++)
            e->tokens[e->sl[i].tok_off + t] = ff_symbolize_index(P, e->sl[i], e->frame.data(), e->qt.data(), t,
                                                                 rct[2 * i], rct[2 * i + 1]);
    /* state reset on key frames: ff_ffv1_clear_slice_state */
    e->gob_count += keyf;
    if (keyf) {
        memset(e->rstate.data(), 128, e->rstate.size());
        if (const uint8_t *init = e->s.initial[e->s.context_model]) {   /* second pass */
            const size_t per = (size_t)e->s.ctx_count[e->s.context_model] * FF_CONTEXT_SIZE;
            for (int i = 0; i < P.nslices; i++)
                for (int k = 0; k < P.nsets; k++)
                    memcpy(&e->rstate[((size_t)i * P.nsets + k) * per], init, per);
        }
        for (auto &v : e->vstate) { v.x = FF_VLC_INIT_LO; v.y = FF_VLC_INIT_HI; }
    }
    /* stage B: one "thread" per slice */
    for (int i = 0; i < P.nslices; i++) {
        FFSliceRect r = { e->sl[i].x, e->sl[i].y, e->sl[i].w, e->sl[i].h };
        uint32_t ovf = 0;
        alignas(16) uint32_t row[FF_ROW_WORDS];
        FFPassStats pass;
        pass.rc_stat = e->rc_stat.data();
        pass.rc_stat2 = e->rc_stat2.data();
        pass.ctx_count = e->s.ctx_count[e->s.context_model];
        int rc = ff_enc_slice_prefix(&e->s, i, &r, keyf, e->ps, e->sar_num, e->sar_den, &pre[i],
                                     &e->prebytes[(size_t)i * 2048], 2048);
        if (rc < 0) return rc;
        pre[i].byte_off = (uint32_t)i * 2048;
        if (P.ac == FF_AC_GOLOMB)
            bytes[i] = ff_encode_slice_golomb(P, e->sl[i], &e->tokens[e->sl[i].tok_off],
                                              &e->vstate[(size_t)i * P.total_ctx], pre[i], e->prebytes.data(),
                                              &e->bs[e->sl[i].bs_off], &ovf, &e->s.cur_tab,
                                              P.version > 3 ? &rct[2 * i] : nullptr);
        else if (getenv("FFV1_EMUL_SPLIT") && !e->pass1) {
            /* stage B in two halves: the state chains in coding order, then the bare coder */
            const FFDevSlice &sl = e->sl[i];
            std::vector<uint16_t> rec;
            uint32_t guard = 0xFFFFFFFFu;
            const uint32_t guard_tok = P.version > 3 ? sl.ntok - (uint32_t)sl.seg_w[sl.nseg - 1] : 0xFFFFFFFFu;
            uint16_t tmp[80];
            for (uint32_t t = 0; t < sl.ntok; t++) {
                if (t == guard_tok) guard = (uint32_t)rec.size();
                uint32_t n = ff_chain_token(e->tokens[sl.tok_off + t],
                                            &e->rstate[(size_t)i * P.total_ctx * FF_CONTEXT_SIZE], &e->s.cur_tab, tmp);
                if (n != ff_token_weight(e->tokens[sl.tok_off + t])) return FFGPU_EXTERNAL;
                rec.insert(rec.end(), tmp, tmp + n);
            }
            bytes[i] = ff_encode_slice_records(sl, rec.data(), (uint32_t)rec.size(), guard, &e->s.cur_tab, pre[i],
                                               e->prebytes.data(), &e->bs[sl.bs_off], &ovf,
                                               P.version > 3 ? &rct[2 * i] : nullptr,
                                               P.version > 3 ? (uint32_t)((16384 + (int64_t)P.width * P.height * 12) / P.nslices) : 0u);
        } else if (getenv("FFV1_EMUL_LONE") && !e->pass1)
            /* the one-slice-per-warp form of the coder */
            bytes[i] = ff_encode_slice_range_lone(e->sl[i], &e->tokens[e->sl[i].tok_off],
                                             &e->rstate[(size_t)i * P.total_ctx * FF_CONTEXT_SIZE], &e->s.cur_tab,
                                             pre[i], e->prebytes.data(), &e->bs[e->sl[i].bs_off], &ovf, row,
                                             P.version > 3 ? &rct[2 * i] : nullptr,
                                             P.version > 3 ? (uint32_t)((16384 + (int64_t)P.width * P.height * 12) / P.nslices) : 0u);
        else if (e->pass1)
            bytes[i] = ff_encode_slice_range<true>(e->sl[i], &e->tokens[e->sl[i].tok_off],
                                             &e->rstate[(size_t)i * P.total_ctx * FF_CONTEXT_SIZE], &e->s.cur_tab,
                                             pre[i], e->prebytes.data(), &e->bs[e->sl[i].bs_off], &ovf, row,
                                             P.version > 3 ? &rct[2 * i] : nullptr,
                                             P.version > 3 ? (uint32_t)((16384 + (int64_t)P.width * P.height * 12) / P.nslices) : 0u,
                                             e->pass1 ? &pass : nullptr);
        else
            bytes[i] = ff_encode_slice_range<false>(e->sl[i], &e->tokens[e->sl[i].tok_off],
                                             &e->rstate[(size_t)i * P.total_ctx * FF_CONTEXT_SIZE], &e->s.cur_tab,
                                             pre[i], e->prebytes.data(), &e->bs[e->sl[i].bs_off], &ovf, row,
                                             P.version > 3 ? &rct[2 * i] : nullptr,
                                             P.version > 3 ? (uint32_t)((16384 + (int64_t)P.width * P.height * 12) / P.nslices) : 0u,
                                             e->pass1 ? &pass : nullptr);
        if (ovf) return FFGPU_INVALIDDATA;
    }
    /* pack */
    size_t off = 0;
    for (int i = 0; i < P.nslices; i++) {
        uint32_t n = ff_slice_packed_size(P, i, bytes[i]);
        if (off + n > (size_t)cap) return FFGPU_ENOSPC;
        ff_pack_slice(P, i, &e->bs[e->sl[i].bs_off], bytes[i], out + off, e->crc.data());
        off += n;
    }
    if (key) *key = keyf;
    e->pic++;
    return (int)off;
}

/* picture structure (3 progressive, 1 / 2 interlaced top / bottom field first) and sample
 * aspect ratio of the pictures that follow */
extern "C" void ffv1emul_encoder_frame_props(void *h, int ps, int sar_num, int sar_den)
{
    Enc *e = (Enc *)h;
    e->ps = ps; e->sar_num = sar_num; e->sar_den = sar_den;
}

extern "C" int ffv1emul_encoder_stats_out(void *h, char *buf, int cap)
{
    Enc *e = (Enc *)h;
    const FFStream &s = e->s;
    size_t pos = 0;
    if (!e->pass1) return FFGPU_EINVAL;
#define OUT_(...) do { int w_ = snprintf(buf + pos, cap - pos, __VA_ARGS__); if (w_ < 0 || (size_t)w_ >= cap - pos) return FFGPU_ENOSPC; pos += w_; } while (0)
    for (int j = 0; j < 256; j++) OUT_("%llu %llu ", e->rc_stat[2 * j], e->rc_stat[2 * j + 1]);
    for (int i = 0; i < s.qt_count; i++)
        for (int j = 0; j < s.ctx_count[i]; j++)
            for (int m = 0; m < 32; m++) {
                const bool own = i == s.context_model;
                OUT_("%llu %llu ", own ? e->rc_stat2[((size_t)j * 32 + m) * 2] : 0ull,
                     own ? e->rc_stat2[((size_t)j * 32 + m) * 2 + 1] : 0ull);
            }
    OUT_("%d\n", e->gob_count);
#undef OUT_
    return (int)pos;
}

extern "C" void ffv1emul_encoder_close(void *h)
{
    Enc *e = (Enc *)h;
    if (!e) return;
    free(e->extradata);
    ff_stream_free(&e->s);
    delete e;
}

struct Dec {
    FFStream s;
    FFDecHostState hs;
    FFDevParams P;
    bool have_params = false;
    std::vector<FFDevSlice> sl;
    std::vector<int16_t> qt;
    std::vector<uint8_t> frame[2], rstate;
    std::vector<uint2> vstate;
    std::vector<int32_t> lines;
    int cur = 0, have_last = 0, damaged = 0;
    FFDecFrameInfo last_info;
};

extern "C" void *ffv1emul_decoder_open(int w, int h, const uint8_t *ex, int exsize, int threads, int *err)
{
    (void)threads;
    Dec *d = new Dec;
    memset(&d->s, 0, sizeof(d->s));
    memset(&d->hs, 0, sizeof(d->hs));
    d->s.width = w; d->s.height = h; d->s.nh = d->s.nv = 1;
    ff_default_tables(&d->s.def_tab);
    d->s.cur_tab = d->s.def_tab;
    *err = 0;
    if (!w || !h) *err = FFGPU_INVALIDDATA;
    if (*err >= 0 && exsize > 0) *err = ff_parse_extradata(&d->s, ex, exsize);
    if (*err < 0) { ff_stream_free(&d->s); delete d; return nullptr; }
    d->hs.max_slices = d->s.nh * d->s.nv;
    return d;
}

extern "C" int ffv1emul_decoder_damaged(void *h) { return ((Dec *)h)->damaged; }

/* what the C ABI reports beside the pictures: output format (NULL until known), stream
 * parameters, the frame header of a version 0/1 stream, the last picture's properties */
extern "C" const char *ffv1emul_decoder_pix_fmt(void *h)
{
    Dec *d = (Dec *)h;
    return d->s.pf ? d->s.pf->name : nullptr;
}

extern "C" void ffv1emul_decoder_info(void *h, int info[8])
{
    Dec *d = (Dec *)h;
    info[0] = d->s.version; info[1] = d->s.micro_version; info[2] = d->s.ac; info[3] = d->s.nh;
    info[4] = d->s.nv; info[5] = d->s.ec; info[6] = d->s.bits; info[7] = d->s.colorspace;
}

extern "C" int ffv1emul_decoder_probe(void *h, const uint8_t *pkt, int size)
{
    Dec *d = (Dec *)h;
    if (d->s.pf)
        return 0;
    if (d->s.version >= 2)                      /* parameters came with the extradata */
        return ff_pick_decoder_format(&d->s);
    std::vector<uint8_t> padded((size_t)size + 64, 0);
    memcpy(padded.data(), pkt, size);
    FFDecSlice tmp[1];
    FFDecFrameInfo info;
    FFDecHostState hs = d->hs;
    hs.max_slices = 1;
    const int r = ff_dec_parse_packet(&d->s, &hs, padded.data(), size, 0, tmp, &info);
    return r < 0 ? r : 0;
}

extern "C" void ffv1emul_decoder_frame_props(void *h, int props[5])
{
    const FFDecFrameInfo &i = ((Dec *)h)->last_info;
    props[0] = i.key_frame; props[1] = i.interlaced_frame; props[2] = i.top_field_first;
    props[3] = i.sar_num; props[4] = i.sar_den;
}

extern "C" int ffv1emul_decode(void *h, const uint8_t *pkt_in, int size, uint8_t *planes[4], int ls[4],
                               const char **fmt, int *key)
{
    Dec *d = (Dec *)h;
    std::vector<uint8_t> pkt((size_t)size + 64, 0);
    memcpy(pkt.data(), pkt_in, size);
    std::vector<FFDecSlice> work(FF_MAX_SLICES);
    FFDecFrameInfo info;
    d->hs.device_parse = 1;                 /* exercise the device-side header parser too */
    int n = ff_dec_parse_packet(&d->s, &d->hs, pkt.data(), size, 0, work.data(), &info);
    if (n < 0) return n;
    d->last_info = info;
    if (!d->have_params || info.key_frame) {
        d->sl.resize(d->s.nh * d->s.nv);
        ff_fill_dev_params(&d->s, 0, &d->P, d->sl.data());
        d->qt = make_qt(d->s);
        d->have_params = true;
        size_t st = (size_t)d->P.nslices * d->P.total_ctx;
        if (d->rstate.size() != st * FF_CONTEXT_SIZE) { d->rstate.assign(st * FF_CONTEXT_SIZE, 128); d->vstate.resize(st); }
        if (d->frame[0].size() != d->P.frame_bytes) { d->frame[0].assign(d->P.frame_bytes, 0); d->frame[1].assign(d->P.frame_bytes, 0); d->have_last = 0; }
    }
    const FFDevParams &P = d->P;
    d->cur ^= 1;
    uint8_t *frame = d->frame[d->cur].data();
    const int line_stride = P.width + 8;
    d->lines.resize((size_t)P.ncoded * 2 * line_stride);
    std::vector<FFDecResult> res(n);
    std::vector<uint32_t> crc(256);
    for (int i = 0; i < 256; i++) crc[i] = ff_crc_table_entry(i);
    FFDecHdr H;
    H.micro_version = d->s.micro_version; H.qt_count = d->s.qt_count; H.ctx_cap = P.total_ctx / P.nsets;
    for (int i = 0; i < FF_MAX_QUANT_TABLES; i++) H.ctx_count[i] = i < d->s.qt_count ? d->s.ctx_count[i] : 0;
    for (int i = 0; i < n; i++) {
        memset(&res[i], 0, sizeof(res[i]));
        if (work[i].parse && !work[i].skip) {
            ff_dec_slice_header(P, H, &work[i], pkt.data(), &d->s.cur_tab, crc.data(), &res[i]);
            if (res[i].flags & (FF_RES_CRC_BAD | FF_RES_HDR_BAD)) d->hs.damaged[i] = 1;
            FFSliceRect rc = { work[i].x, work[i].y, work[i].w, work[i].h };
            d->hs.rect[i] = rc;
        }
        if (work[i].skip) continue;
        uint8_t *rs = &d->rstate[(size_t)i * P.total_ctx * FF_CONTEXT_SIZE];
        uint2 *vs = &d->vstate[(size_t)i * P.total_ctx];
        if (work[i].key_frame) {
            for (int set = 0; set < P.nsets; set++) {
                const uint8_t *init = d->s.initial[work[i].qidx[set]];
                int cnt = d->s.ctx_count[work[i].qidx[set]];
                uint8_t *dst = rs + (size_t)P.set_base[set] * FF_CONTEXT_SIZE;
                if (init) memcpy(dst, init, (size_t)cnt * FF_CONTEXT_SIZE);
                else memset(dst, 128, (size_t)cnt * FF_CONTEXT_SIZE);
            }
            for (int c = 0; c < P.total_ctx; c++) { vs[c].x = FF_VLC_INIT_LO; vs[c].y = FF_VLC_INIT_HI; }
        }
        FFDecCtx D;
        D.qt_all = d->qt.data(); D.tab = &d->s.cur_tab; D.rstate = rs; D.vstate = vs;
        D.lines = d->lines.data(); D.line_stride = line_stride; D.frame = frame;
        D.gate_wait = FF_NEW_WAIT;
        D.any_five = 0;
        for (int q = 0; q < d->s.qt_count; q++)
            D.any_five |= d->qt[(size_t)q * FF_QT_STRIDE + FF_MAX_CTX_INPUTS * 256];
        /* lazily created states: exactly when the product uses them (every frame a key frame,
         * no initial-state tables) */
        std::vector<uint32_t> touched((size_t)(P.total_ctx + 31) / 32, 0);
        bool lazy = d->s.version > 2 && d->s.intra && work[i].key_frame && ff_decode_planar_mode(&P);
        for (int q = 0; q < d->s.qt_count; q++)
            if (d->s.initial[q]) lazy = false;
        D.touched = lazy ? touched.data() : nullptr;
        D.lone = getenv("FFV1_EMUL_LONE") != nullptr;   /* the one-slice-per-warp form of the decoder */
        if (lazy)                                   /* rows must come from the touched logic, not from the reset above */
            memset(rs, 0x55, (size_t)P.total_ctx * FF_CONTEXT_SIZE);
        alignas(16) uint32_t row[FF_ROW_WORDS];
        ff_decode_slice(P, work[i], pkt.data(), D, &res[i], row);
        if (P.ac != FF_AC_GOLOMB && P.version > 2) {
            int v = (int)work[i].size - (int)res[i].end_pos - 2 - 5 * P.ec;
            if (v) d->hs.damaged[i] = 1;          /* "bytestream end mismatching by %d" */
        }
    }
    /* concealment, ffv1dec.c:940-969 */
    d->damaged = 0;
    for (int i = n - 1; i >= 0; i--) {
        if (!d->hs.damaged[i]) continue;
        d->damaged++;
        if (!d->have_last) continue;
        const FFPixFmt *pf = d->s.pf;
        const FFSliceRect r = d->hs.rect[i];
        for (int p = 0; p < pf->nplanes; p++) {
            int sh = (pf->layout == FF_LAY_PLANAR && pf->chroma && (p == 1 || p == 2)) ? d->s.hs : 0;
            int sv = (pf->layout == FF_LAY_PLANAR && pf->chroma && (p == 1 || p == 2)) ? d->s.vs : 0;
            int pixshift = pf->depth > 8;
            int bw = ff_crshift(r.w, sh) * ff_bytes_per_pixel(pf), rows = ff_crshift(r.h, sv);
            size_t xo = (size_t)((r.x >> sh) << pixshift);
            for (int y = 0; y < rows; y++) {
                size_t o = P.plane_off[p] + (size_t)((r.y >> sv) + y) * P.pitch[p] + xo;
                memcpy(frame + o, d->frame[d->cur ^ 1].data() + o, bw);
            }
        }
    }
    d->have_last = 1;
    for (int p = 0; p < 4; p++) {
        planes[p] = p < d->s.pf->nplanes ? frame + P.plane_off[p] : nullptr;
        ls[p] = p < d->s.pf->nplanes ? P.pitch[p] : 0;
    }
    if (fmt) *fmt = d->s.pf->name;
    if (key) *key = info.key_frame;
    return size;
}

extern "C" void ffv1emul_decoder_close(void *h)
{
    Dec *d = (Dec *)h;
    if (!d) return;
    ff_stream_free(&d->s);
    delete d;
}

extern "C" int ffv1emul_plane_geometry(const char *fmt, int w, int h, int plane, int *bw, int *rows)
{
    return ff_plane_geometry(ff_find_pixfmt(fmt), w, h, plane, bw, rows);
}
