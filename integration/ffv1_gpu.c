/*
 * integration/ffv1_gpu.c -- the libavcodec side of the drop-in: two AVCodec objects that
 * keep FFV1's init/encode2/decode/close callbacks and AVOptions and forward the slice
 * pixel path to libffgpu.so (include/ffgpu.h).  A maintainer adds this one file to
 * libavcodec/ (see INTEGRATION.md); nothing else in FFmpeg changes.
 *
 * It replaces, callback for callback:
 *   ff_ffv1_encoder  libavcodec/ffv1enc.c:1323-1360  (encode_init :517, encode_frame :1122,
 *                                                     encode_close :1283, options :1291-1307)
 *   ff_ffv1_decoder  libavcodec/ffv1dec.c:1087-1101  (decode_init :818, decode_frame :837)
 *
 * The codecs are registered under the names "ffv1_gpu" (and can be renamed "ffv1" for a
 * literal drop-in); the bitstream is FFV1, so any FFV1 decoder reads the encoder's output
 * and the decoder reads any FFV1 v0/v1/v3 stream.
 */
#include "libavutil/avassert.h"
#include "libavutil/imgutils.h"
#include "libavutil/opt.h"
#include "libavutil/pixdesc.h"
#include "avcodec.h"
#include "decode.h"
#include "internal.h"

#include "ffgpu.h"

/* Pictures that already live on the GPU: the encoder also takes AV_PIX_FMT_CUDA
 * frames of an AVHWFramesContext whose sw_format is one of its formats (hwupload_cuda,
 * scale_cuda, a CUDA decoder in front of it), the pattern of libavcodec/nvenc.c:525-545,
 * :1654-1680.  The planes are device pointers; the library copies device to device.  Needs a
 * build with the CUDA hw context (CONFIG_CUDA). */
#if CONFIG_CUDA
#include "libavutil/hwcontext.h"
#include "libavutil/hwcontext_cuda.h"
#include "hwaccel.h"
#define CUDA_PUSH(s) do { if ((s)->cu_ctx) ffgpu_cuda_push_context((s)->cu_ctx); } while (0)
#define CUDA_POP(s)  do { if ((s)->cu_ctx) ffgpu_cuda_pop_context(); } while (0)
#else
#define CUDA_PUSH(s) do { } while (0)
#define CUDA_POP(s)  do { } while (0)
#endif

#define GPU_STATS_OUT_SIZE (1024 * 1024 * 6)   /* STATS_OUT_SIZE, ffv1enc.c:911 */

typedef struct FFV1GpuContext {
    AVClass *class;            /* first member: avcodec_open2 applies the AVOptions (utils.c:630-639) */
    /* AVOptions with the names, ranges and defaults of ffv1enc.c:1291-1307 */
    int ec;
    int ac;
    int context_model;
    /* new: placement */
    int gpu;
    int gpus;                  /* > 1: spread the stream over GPUs gpu .. gpu+gpus-1 */
    int max_batch;
    int depth;
    ffgpu_encoder *enc;
    ffgpu_decoder *dec;
    /* pictures handed to the GPU that have not come back yet, oldest first.  Encoder: a
     * reference to every input frame (the library reads the planes asynchronously until the
     * packet is out, like ffv1enc.c:1194-1196 keeps its av_frame_ref).  Decoder: the output
     * frames the pictures are downloaded into. */
    AVFrame **fifo;            /* ring that grows with what the launch groups hold in flight
                                * (up to gpu_batch x gpu_depth pictures per GPU) */
    int fifo_head, fifo_count, fifo_cap;
    AVPacket *pending;         /* decoder: a packet the library could not take yet */
    int draining, eof;
    void *cu_ctx;              /* AV_PIX_FMT_CUDA frames in or out: the hw device's CUcontext, else NULL */
    int negotiated;            /* decoder: the output format has been offered to the caller */
    int cuda_out;              /* decoder: pictures are written into AV_PIX_FMT_CUDA frames */
} FFV1GpuContext;

/* room for one more entry; the only way fifo_push can fail, so callers reserve BEFORE they
 * hand a picture to the library */
static int fifo_reserve(FFV1GpuContext *s)
{
    AVFrame **ring;
    int cap, i;
    if (s->fifo_count < s->fifo_cap)
        return 0;
    cap = s->fifo_cap ? 2 * s->fifo_cap : 256;
    if (!(ring = av_mallocz_array(cap, sizeof(*ring))))
        return AVERROR(ENOMEM);
    for (i = 0; i < s->fifo_count; i++)
        ring[i] = s->fifo[(s->fifo_head + i) % s->fifo_cap];
    av_free(s->fifo);
    s->fifo = ring;
    s->fifo_cap = cap;
    s->fifo_head = 0;
    return 0;
}

static int fifo_push(FFV1GpuContext *s, AVFrame *f)
{
    int ret = fifo_reserve(s);
    if (ret < 0)
        return ret;
    s->fifo[(s->fifo_head + s->fifo_count++) % s->fifo_cap] = f;
    return 0;
}

static AVFrame *fifo_pop(FFV1GpuContext *s)
{
    AVFrame *f;
    if (!s->fifo_count)
        return NULL;
    f = s->fifo[s->fifo_head];
    s->fifo[s->fifo_head] = NULL;
    s->fifo_head = (s->fifo_head + 1) % s->fifo_cap;
    s->fifo_count--;
    return f;
}

static void set_devices(const FFV1GpuContext *s, int *ndevices, int devices[FFGPU_MAX_DEVICES])
{
    int i;
    *ndevices = s->gpus > 1 ? FFMIN(s->gpus, FFGPU_MAX_DEVICES) : 0;
    for (i = 0; i < *ndevices; i++)
        devices[i] = s->gpu + i;
}

/* exported so that a test harness without libavutil/opt.c can set the private options */
void ff_ffv1_gpu_set_options(void *priv, int slicecrc, int coder, int context);
void ff_ffv1_gpu_set_options(void *priv, int slicecrc, int coder, int context)
{
    FFV1GpuContext *s = priv;
    s->ec = slicecrc;
    s->ac = coder;
    s->context_model = context;
}

static av_cold int gpu_encode_init(AVCodecContext *avctx)
{
    FFV1GpuContext *s = avctx->priv_data;
    ffgpu_enc_options o = { 0 };
    const uint8_t *ex;
    int ret, n;

    o.width  = avctx->width;
    o.height = avctx->height;
    o.pix_fmt = av_get_pix_fmt_name(avctx->pix_fmt);
#if CONFIG_CUDA
    if (avctx->pix_fmt == AV_PIX_FMT_CUDA) {
        AVHWFramesContext *frames;
        AVCUDADeviceContext *cuda;
        if (!avctx->hw_frames_ctx) {
            av_log(avctx, AV_LOG_ERROR, "AV_PIX_FMT_CUDA input needs hw_frames_ctx\n");
            return AVERROR(EINVAL);
        }
        frames = (AVHWFramesContext *)avctx->hw_frames_ctx->data;
        cuda   = frames->device_ctx->hwctx;
        s->cu_ctx = cuda->cuda_ctx;
        o.pix_fmt = av_get_pix_fmt_name(frames->sw_format);
    }
#endif
    o.slices = avctx->slices;
    o.level  = avctx->level;
    o.gop_size = avctx->gop_size;
    o.coder  = s->ac;
    o.context = s->context_model;
    o.slicecrc = s->ec;
    o.strict_std_compliance = avctx->strict_std_compliance;
    o.bits_per_raw_sample = avctx->bits_per_raw_sample;
    o.device = s->gpu;
    o.max_batch = s->max_batch;
    o.pipeline_depth = s->depth;
    set_devices(s, &o.ndevices, o.devices);
    /* two-pass coding, ffv1enc.c:528, :785-873, :912-916 */
    o.pass1 = !!(avctx->flags & AV_CODEC_FLAG_PASS1);
    o.pass2 = !!(avctx->flags & AV_CODEC_FLAG_PASS2);
    o.stats_in = avctx->stats_in;
    if (o.pass1 && !(avctx->stats_out = av_mallocz(GPU_STATS_OUT_SIZE)))
        return AVERROR(ENOMEM);
    CUDA_PUSH(s);
    ret = ffgpu_ffv1_encode_init(&s->enc, &o);
    CUDA_POP(s);
    if (ret < 0) {
        av_log(avctx, AV_LOG_ERROR, "%s\n", ffgpu_last_error());
        return ret;                            /* same AVERROR values as encode_init */
    }
    n = ffgpu_ffv1_encoder_extradata(s->enc, &ex);
    if (n > 0) {
        avctx->extradata = av_mallocz(n + AV_INPUT_BUFFER_PADDING_SIZE);
        if (!avctx->extradata)
            return AVERROR(ENOMEM);
        memcpy(avctx->extradata, ex, n);
        avctx->extradata_size = n;
    }
    {
        int info[8];
        ffgpu_ffv1_encoder_info(s->enc, info);
        avctx->bits_per_raw_sample = info[6];
    }
#if FF_API_CODED_FRAME
FF_DISABLE_DEPRECATION_WARNINGS
    avctx->coded_frame->pict_type = AV_PICTURE_TYPE_I;
FF_ENABLE_DEPRECATION_WARNINGS
#endif
    return 0;
}

/* AVCodec.send_frame / receive_packet (avcodec.h:3654-3662): the encoder sets
 * AV_CODEC_CAP_DELAY like ffv1enc.c:1332, pictures go into the launch-group pipeline and
 * packets come back in presentation order as groups finish; a NULL frame drains it.
 * AVERROR(EAGAIN) / AVERROR_EOF have the meaning libavcodec/encode.c expects. */
static int gpu_send_frame(AVCodecContext *avctx, const AVFrame *pict)
{
    FFV1GpuContext *s = avctx->priv_data;
    ffgpu_picture p = { { 0 } };
    AVFrame *ref;
    int ret, i;

    if (!pict) {
        CUDA_PUSH(s);
        ret = ffgpu_ffv1_encode_send_frame(s->enc, NULL);
        CUDA_POP(s);
        return ret;
    }
    if ((ret = fifo_reserve(s)) < 0)
        return ret;
    /* the planes are read asynchronously: hold a reference until the packet is out, and give
     * the library the planes of THAT reference (for a frame that is not refcounted
     * av_frame_clone copies the picture, and the caller may reuse its own right away) */
    if (!(ref = av_frame_clone(pict)))
        return AVERROR(ENOMEM);
    for (i = 0; i < 4; i++) {
        p.data[i] = ref->data[i];
        p.linesize[i] = ref->linesize[i];
    }
    p.interlaced_frame = pict->interlaced_frame;
    p.top_field_first = pict->top_field_first;
    p.sar_num = pict->sample_aspect_ratio.num;
    p.sar_den = pict->sample_aspect_ratio.den;
    p.pts = pict->pts;
    CUDA_PUSH(s);
    ret = ffgpu_ffv1_encode_send_frame(s->enc, &p);   /* FFGPU_EAGAIN == AVERROR(EAGAIN) */
    CUDA_POP(s);
    if (ret < 0) {
        if (ret != AVERROR(EAGAIN))
            av_log(avctx, AV_LOG_ERROR, "%s\n", ffgpu_last_error());
        av_frame_free(&ref);
        return ret;
    }
    return fifo_push(s, ref);                  /* cannot fail: room was reserved above */
}

static int gpu_receive_packet(AVCodecContext *avctx, AVPacket *pkt)
{
    FFV1GpuContext *s = avctx->priv_data;
    size_t size = 0;
    int key = 0, ret;
    int64_t pts = AV_NOPTS_VALUE;

    AVFrame *done;

    /* is a packet ready, and how large is it?  Nothing is allocated on EAGAIN / EOF, and the
     * packet gets its exact size instead of the worst case of ffv1enc.c:1131-1132 */
    CUDA_PUSH(s);
    ret = ffgpu_ffv1_encode_packet_ready(s->enc, &size);
    CUDA_POP(s);
    if (ret < 0) {                             /* EAGAIN, EOF, "encoded frame too large", ... */
        if (ret != AVERROR(EAGAIN) && ret != AVERROR_EOF)
            av_log(avctx, AV_LOG_ERROR, "%s\n", ffgpu_last_error());
        if (ret == AVERROR_EOF && avctx->stats_out) {
            /* first pass: the statistics appear with the end of the flush (ffv1enc.c:1134-1177) */
            CUDA_PUSH(s);
            if (ffgpu_ffv1_encoder_stats_out(s->enc, avctx->stats_out, GPU_STATS_OUT_SIZE) < 0)
                av_log(avctx, AV_LOG_ERROR, "%s\n", ffgpu_last_error());
            CUDA_POP(s);
        }
        return ret;
    }
    if (avctx->stats_out)
        avctx->stats_out[0] = 0;               /* ffv1enc.c:1264-1265 */
    /* a refcounted packet of exactly that size: avcodec_receive_packet hands it on as it is
     * (libavcodec/encode.c:434-438), there is no encode2 wrapper that would copy it out of
     * the shared byte_buffer ff_alloc_packet2 may return */
    if ((ret = av_new_packet(pkt, size)) < 0)
        return ret;
    CUDA_PUSH(s);
    ret = ffgpu_ffv1_encode_receive_packet(s->enc, pkt->data, pkt->size, &size, &key, &pts);
    CUDA_POP(s);
    if (ret < 0) {
        av_log(avctx, AV_LOG_ERROR, "%s\n", ffgpu_last_error());
        av_packet_unref(pkt);
        return ret;
    }
    done = fifo_pop(s);                        /* packets come back in input order */
    av_frame_free(&done);
    pkt->size = size;
    pkt->pts = pkt->dts = pts;
    if (key)
        pkt->flags |= AV_PKT_FLAG_KEY;
    return 0;
}

static av_cold int gpu_close(AVCodecContext *avctx)
{
    FFV1GpuContext *s = avctx->priv_data;
    AVFrame *f;
    /* the library first: it waits for the GPU, which may still read the frames' planes */
    CUDA_PUSH(s);
    ffgpu_ffv1_encode_close(s->enc);
    ffgpu_ffv1_decode_close(s->dec);
    CUDA_POP(s);
    s->enc = NULL;
    s->dec = NULL;
    while ((f = fifo_pop(s)))
        av_frame_free(&f);
    av_freep(&s->fifo);
    s->fifo_cap = s->fifo_head = 0;
    av_packet_free(&s->pending);
    av_freep(&avctx->stats_out);               /* like ff_ffv1_close, ffv1.c:235 */
    return 0;
}

/* Output format of the decoder.  Ordinarily the stream's own format.  In a build with the
 * CUDA hw context, and when the caller supplied a CUDA device or frames context, the decoder
 * also offers AV_PIX_FMT_CUDA through get_format (the way libavcodec/cuviddec.c:845-915
 * does): the pictures then STAY on the GPU -- ff_get_buffer hands out frames of an
 * AVHWFramesContext whose planes are device pointers, and the library downloads nothing
 * (hwupload-free chains such as ffv1_gpu -> scale_cuda -> an NVENC encoder).
 * hwcontext_cuda.c knows the layouts of yuv420p, yuv444p, yuv444p16 and 0rgb32/0bgr32; any
 * other stream format falls back to frames in host memory. */
static int gpu_output_format(AVCodecContext *avctx, enum AVPixelFormat sw_fmt)
{
    FFV1GpuContext *s = avctx->priv_data;
#if CONFIG_CUDA
    if (!s->negotiated && (avctx->hw_device_ctx || avctx->hw_frames_ctx)) {
        const enum AVPixelFormat fmts[3] = { AV_PIX_FMT_CUDA, sw_fmt, AV_PIX_FMT_NONE };
        int ret;
        avctx->sw_pix_fmt = sw_fmt;
        if ((ret = ff_get_format(avctx, fmts)) < 0)
            return ret;
        if (ret == AV_PIX_FMT_CUDA) {
            AVHWFramesContext *frames;
            if (!avctx->hw_frames_ctx) {
                if (!(avctx->hw_frames_ctx = av_hwframe_ctx_alloc(avctx->hw_device_ctx)))
                    return AVERROR(ENOMEM);
                frames = (AVHWFramesContext *)avctx->hw_frames_ctx->data;
                frames->format    = AV_PIX_FMT_CUDA;
                frames->sw_format = sw_fmt;
                frames->width     = avctx->coded_width  ? avctx->coded_width  : avctx->width;
                frames->height    = avctx->coded_height ? avctx->coded_height : avctx->height;
                if ((ret = av_hwframe_ctx_init(avctx->hw_frames_ctx)) < 0) {
                    av_log(avctx, AV_LOG_WARNING, "no CUDA frames of format %s: pictures go to host memory\n",
                           av_get_pix_fmt_name(sw_fmt));
                    av_buffer_unref(&avctx->hw_frames_ctx);
                }
            }
            if (avctx->hw_frames_ctx) {
                frames = (AVHWFramesContext *)avctx->hw_frames_ctx->data;
                if (frames->format != AV_PIX_FMT_CUDA || frames->sw_format != sw_fmt) {
                    av_log(avctx, AV_LOG_ERROR, "hw_frames_ctx does not hold CUDA frames of format %s\n",
                           av_get_pix_fmt_name(sw_fmt));
                    return AVERROR(EINVAL);
                }
                s->cu_ctx = ((AVCUDADeviceContext *)frames->device_ctx->hwctx)->cuda_ctx;
                s->cuda_out = 1;
            }
        }
    }
#endif
    s->negotiated = 1;
    avctx->pix_fmt = s->cuda_out ? AV_PIX_FMT_CUDA : sw_fmt;
    return 0;
}

static av_cold int gpu_decode_init(AVCodecContext *avctx)
{
    FFV1GpuContext *s = avctx->priv_data;
    ffgpu_dec_options o = { 0 };
    const char *name;
    int ret;

    o.width = avctx->width;
    o.height = avctx->height;
    o.extradata = avctx->extradata;
    o.extradata_size = avctx->extradata_size;
    o.device = s->gpu;
    o.max_batch = s->max_batch;
    o.pipeline_depth = s->depth;
    set_devices(s, &o.ndevices, o.devices);
    if (!(s->pending = av_packet_alloc()))
        return AVERROR(ENOMEM);
    if ((ret = ffgpu_ffv1_decode_init(&s->dec, &o)) < 0) {
        av_log(avctx, AV_LOG_ERROR, "%s\n", ffgpu_last_error());
        return ret;
    }
    if ((name = ffgpu_ffv1_decoder_pix_fmt(s->dec))) {
        int info[8];
        avctx->pix_fmt = av_get_pix_fmt(name);
        ffgpu_ffv1_decoder_info(s->dec, info);
        avctx->bits_per_raw_sample = info[6];
    }
    return 0;
}

static void fill_frame_props(AVFrame *frame, const ffgpu_picture_out *out)
{
    frame->pict_type = AV_PICTURE_TYPE_I;
    frame->key_frame = out->key_frame;
    frame->interlaced_frame = out->interlaced_frame;
    frame->top_field_first = out->top_field_first;
    frame->sample_aspect_ratio = (AVRational){ out->sar_num, out->sar_den };
}

/* AVCodec.decode: synchronous form (one packet in, one picture out); the pipelined form is
 * gpu_receive_frame below. */
static int gpu_decode_frame(AVCodecContext *avctx, void *data, int *got_frame, AVPacket *avpkt)
{
    FFV1GpuContext *s = avctx->priv_data;
    AVFrame *frame = data;
    ffgpu_picture_out out = { { 0 } };
    const char *name;
    int ret, i;

    /* v0/v1 streams announce their format in the first key frame */
    if (!(name = ffgpu_ffv1_decoder_pix_fmt(s->dec))) {
        if ((ret = ffgpu_ffv1_decoder_probe(s->dec, avpkt->data, avpkt->size)) < 0) {
            av_log(avctx, AV_LOG_ERROR, "%s\n", ffgpu_last_error());
            return ret;
        }
        name = ffgpu_ffv1_decoder_pix_fmt(s->dec);
    }
    if ((ret = gpu_output_format(avctx, av_get_pix_fmt(name))) < 0)
        return ret;
    if ((ret = ff_get_buffer(avctx, frame, AV_GET_BUFFER_FLAG_REF)) < 0)
        return ret;
    for (i = 0; i < 4; i++) {
        out.data[i] = frame->data[i];
        out.linesize[i] = frame->linesize[i];
    }
    CUDA_PUSH(s);
    ret = ffgpu_ffv1_decode_frame(s->dec, avpkt->data, avpkt->size, &out, got_frame);
    CUDA_POP(s);
    if (ret < 0) {
        av_log(avctx, AV_LOG_ERROR, "%s\n", ffgpu_last_error());
        av_frame_unref(frame);
        return ret;
    }
    fill_frame_props(frame, &out);
    if (out.damaged_slices)
        av_log(avctx, AV_LOG_ERROR, "%d damaged slice(s) concealed\n", out.damaged_slices);
    return ret;                                /* bytes consumed, like ffv1dec.c:982 */
}

/* AVCodec.receive_frame (avcodec.h:3654-3662): the pipelined decoder.  It replaces what
 * frame threading does for the CPU codec (ffv1dec.c:1042-1099, pthread_frame.c): packets are
 * pulled with ff_decode_get_packet() and queued on the GPU together with the AVFrame they
 * will be downloaded into; pictures come back in packet order as launch groups finish. */
static int gpu_receive_frame(AVCodecContext *avctx, AVFrame *frame)
{
    FFV1GpuContext *s = avctx->priv_data;
    ffgpu_picture_out out;
    const char *name;
    int ret, i;

    if (s->eof)                                /* drained: libavcodec may still ask again */
        return AVERROR_EOF;
    for (;;) {
        memset(&out, 0, sizeof(out));
        CUDA_PUSH(s);
        ret = ffgpu_ffv1_decode_receive_frame(s->dec, &out);
        CUDA_POP(s);
        if (ret == 0) {
            AVFrame *f = fifo_pop(s);
            av_assert0(f);
            av_frame_move_ref(frame, f);
            av_frame_free(&f);
            fill_frame_props(frame, &out);
            if (out.damaged_slices)
                av_log(avctx, AV_LOG_ERROR, "%d damaged slice(s) concealed\n", out.damaged_slices);
            return 0;
        }
        if (ret != AVERROR(EAGAIN)) {          /* AVERROR_EOF after the drain, or an error */
            if (ret != AVERROR_EOF)
                av_log(avctx, AV_LOG_ERROR, "%s\n", ffgpu_last_error());
            else
                s->eof = 1;
            return ret;
        }
        /* nothing ready: feed the GPU */
        if (s->draining)
            continue;                          /* flushed: the library blocks until pictures come */
        if (!s->pending->data) {
            ret = ff_decode_get_packet(avctx, s->pending);
            if (ret == AVERROR_EOF) {
                s->draining = 1;
                CUDA_PUSH(s);
                ret = ffgpu_ffv1_decode_send_packet(s->dec, NULL, 0, 0, NULL);
                CUDA_POP(s);
                if (ret < 0)
                    return ret;
                continue;
            }
            if (ret < 0)
                return ret;                    /* AVERROR(EAGAIN): the caller sends more packets */
        }
        if (!(name = ffgpu_ffv1_decoder_pix_fmt(s->dec))) {
            if ((ret = ffgpu_ffv1_decoder_probe(s->dec, s->pending->data, s->pending->size)) < 0) {
                av_log(avctx, AV_LOG_ERROR, "%s\n", ffgpu_last_error());
                av_packet_unref(s->pending);
                return ret;
            }
            name = ffgpu_ffv1_decoder_pix_fmt(s->dec);
        }
        if ((ret = gpu_output_format(avctx, av_get_pix_fmt(name))) < 0) {
            av_packet_unref(s->pending);
            return ret;
        }
        {
            AVFrame *f = av_frame_alloc();
            ffgpu_picture_out dst = { { 0 } };
            if (!f)
                return AVERROR(ENOMEM);
            if ((ret = ff_get_buffer(avctx, f, AV_GET_BUFFER_FLAG_REF)) < 0) {
                av_frame_free(&f);
                return ret;
            }
            /* ff_get_buffer took pts and the other properties from the packet
             * (ff_decode_frame_props); what decode_simple_receive_frame adds for .decode
             * codecs (decode.c:452-470) is done here for this receive_frame codec */
            f->pkt_dts = s->pending->dts;
            f->best_effort_timestamp = f->pts != AV_NOPTS_VALUE ? f->pts : f->pkt_dts;
            for (i = 0; i < 4; i++) {
                dst.data[i] = f->data[i];
                dst.linesize[i] = f->linesize[i];
            }
            CUDA_PUSH(s);
            ret = ffgpu_ffv1_decode_send_packet(s->dec, s->pending->data, s->pending->size,
                                                s->pending->pts, &dst);
            CUDA_POP(s);
            if (ret == AVERROR(EAGAIN)) {      /* every launch group is busy: receive blocks next */
                av_frame_free(&f);
                continue;
            }
            av_packet_unref(s->pending);
            if (ret < 0) {
                av_log(avctx, AV_LOG_ERROR, "%s\n", ffgpu_last_error());
                av_frame_free(&f);
                return ret;
            }
            if ((ret = fifo_push(s, f)) < 0) {
                av_frame_free(&f);
                return ret;
            }
        }
    }
}

static void gpu_flush(AVCodecContext *avctx)
{
    FFV1GpuContext *s = avctx->priv_data;
    ffgpu_picture_out out;
    AVFrame *f;
    int ret;
    if (!s->dec)
        return;
    /* avcodec_flush_buffers: drop what is in flight.  The library is drained to its EOF so
     * that it takes packets again afterwards (a flushed handle refuses them until then); the
     * pictures land in the frames queued for them, which are dropped. */
    CUDA_PUSH(s);
    if (!s->eof) {
        if (!s->draining)
            ffgpu_ffv1_decode_send_packet(s->dec, NULL, 0, 0, NULL);
        do {
            memset(&out, 0, sizeof(out));
            ret = ffgpu_ffv1_decode_receive_frame(s->dec, &out);
            if (ret != AVERROR_EOF && ret != AVERROR(EAGAIN) && (f = fifo_pop(s)))
                av_frame_free(&f);             /* a picture, or one that failed */
        } while (ret != AVERROR_EOF && ret != AVERROR(EAGAIN) && (ret >= 0 || s->fifo_count));
    }
    CUDA_POP(s);
    while ((f = fifo_pop(s)))
        av_frame_free(&f);
    av_packet_unref(s->pending);
    s->draining = s->eof = 0;
}

#define OFFSET(x) offsetof(FFV1GpuContext, x)
#define VE AV_OPT_FLAG_VIDEO_PARAM | AV_OPT_FLAG_ENCODING_PARAM
#define VD AV_OPT_FLAG_VIDEO_PARAM | AV_OPT_FLAG_DECODING_PARAM
static const AVOption enc_options[] = {
    /* unchanged from ffv1enc.c:1291-1307 */
    { "slicecrc", "Protect slices with CRCs", OFFSET(ec), AV_OPT_TYPE_BOOL, { .i64 = -1 }, -1, 1, VE },
    { "coder", "Coder type", OFFSET(ac), AV_OPT_TYPE_INT, { .i64 = 0 }, -2, 2, VE, "coder" },
        { "rice", "Golomb rice", 0, AV_OPT_TYPE_CONST, { .i64 = 0 }, INT_MIN, INT_MAX, VE, "coder" },
        { "range_def", "Range with default table", 0, AV_OPT_TYPE_CONST, { .i64 = -2 }, INT_MIN, INT_MAX, VE, "coder" },
        { "range_tab", "Range with custom table", 0, AV_OPT_TYPE_CONST, { .i64 = 2 }, INT_MIN, INT_MAX, VE, "coder" },
        { "ac", "Range with custom table (the ac option exists for compatibility and is deprecated)", 0,
          AV_OPT_TYPE_CONST, { .i64 = 1 }, INT_MIN, INT_MAX, VE, "coder" },
    { "context", "Context model", OFFSET(context_model), AV_OPT_TYPE_INT, { .i64 = 0 }, 0, 1, VE },
    /* new */
    { "gpu", "CUDA device ordinal", OFFSET(gpu), AV_OPT_TYPE_INT, { .i64 = 0 }, 0, 64, VE },
    { "gpus", "spread the stream over this many GPUs, starting at gpu", OFFSET(gpus), AV_OPT_TYPE_INT, { .i64 = 1 }, 1, FFGPU_MAX_DEVICES, VE },
    { "gpu_batch", "pictures per launch group (0 = auto)", OFFSET(max_batch), AV_OPT_TYPE_INT, { .i64 = 0 }, 0, 1024, VE },
    { "gpu_depth", "launch groups in flight (0 = auto)", OFFSET(depth), AV_OPT_TYPE_INT, { .i64 = 0 }, 0, 8, VE },
    { NULL }
};
static const AVOption dec_options[] = {
    { "gpu", "CUDA device ordinal", OFFSET(gpu), AV_OPT_TYPE_INT, { .i64 = 0 }, 0, 64, VD },
    { "gpus", "spread an intra-only stream over this many GPUs, starting at gpu", OFFSET(gpus), AV_OPT_TYPE_INT, { .i64 = 1 }, 1, FFGPU_MAX_DEVICES, VD },
    { "gpu_batch", "packets per launch group (0 = auto)", OFFSET(max_batch), AV_OPT_TYPE_INT, { .i64 = 0 }, 0, 1024, VD },
    { "gpu_depth", "launch groups in flight (0 = auto)", OFFSET(depth), AV_OPT_TYPE_INT, { .i64 = 0 }, 0, 8, VD },
    { NULL }
};

static const AVClass enc_class = {
    .class_name = "ffv1 gpu encoder", .item_name = av_default_item_name,
    .option = enc_options, .version = LIBAVUTIL_VERSION_INT,
};
static const AVClass dec_class = {
    .class_name = "ffv1 gpu decoder", .item_name = av_default_item_name,
    .option = dec_options, .version = LIBAVUTIL_VERSION_INT,
};

AVCodec ff_ffv1_gpu_encoder = {
    .name           = "ffv1_gpu",
    .long_name      = NULL_IF_CONFIG_SMALL("FFmpeg video codec #1 (B200 CUDA slice path)"),
    .type           = AVMEDIA_TYPE_VIDEO,
    .id             = AV_CODEC_ID_FFV1,
    .priv_data_size = sizeof(FFV1GpuContext),
    .init           = gpu_encode_init,
    .send_frame     = gpu_send_frame,
    .receive_packet = gpu_receive_packet,
    .close          = gpu_close,
    .capabilities   = AV_CODEC_CAP_DELAY,
    /* gpu_close copes with a half-made context: init needs no clean-up paths of its own
     * (the reference encoder sets no caps_internal and cleans up by hand, ffv1enc.c:517-928) */
    .caps_internal  = FF_CODEC_CAP_INIT_CLEANUP,
    .pix_fmts       = (const enum AVPixelFormat[]) {   /* ffv1enc.c:1333-1355 */
        AV_PIX_FMT_YUV420P,   AV_PIX_FMT_YUVA420P,  AV_PIX_FMT_YUVA422P,  AV_PIX_FMT_YUV444P,
        AV_PIX_FMT_YUVA444P,  AV_PIX_FMT_YUV440P,   AV_PIX_FMT_YUV422P,   AV_PIX_FMT_YUV411P,
        AV_PIX_FMT_YUV410P,   AV_PIX_FMT_0RGB32,    AV_PIX_FMT_RGB32,     AV_PIX_FMT_YUV420P16,
        AV_PIX_FMT_YUV422P16, AV_PIX_FMT_YUV444P16, AV_PIX_FMT_YUV444P9,  AV_PIX_FMT_YUV422P9,
        AV_PIX_FMT_YUV420P9,  AV_PIX_FMT_YUV420P10, AV_PIX_FMT_YUV422P10, AV_PIX_FMT_YUV444P10,
        AV_PIX_FMT_YUV420P12, AV_PIX_FMT_YUV422P12, AV_PIX_FMT_YUV444P12,
        AV_PIX_FMT_YUVA444P16, AV_PIX_FMT_YUVA422P16, AV_PIX_FMT_YUVA420P16,
        AV_PIX_FMT_YUVA444P10, AV_PIX_FMT_YUVA422P10, AV_PIX_FMT_YUVA420P10,
        AV_PIX_FMT_YUVA444P9, AV_PIX_FMT_YUVA422P9, AV_PIX_FMT_YUVA420P9,
        AV_PIX_FMT_GRAY16,    AV_PIX_FMT_GRAY8,     AV_PIX_FMT_GBRP9,     AV_PIX_FMT_GBRP10,
        AV_PIX_FMT_GBRP12,    AV_PIX_FMT_GBRP14,    AV_PIX_FMT_GBRAP10,   AV_PIX_FMT_GBRAP12,
        AV_PIX_FMT_YA8,       AV_PIX_FMT_GRAY10,    AV_PIX_FMT_GRAY12,    AV_PIX_FMT_GBRP16,
        AV_PIX_FMT_RGB48,     AV_PIX_FMT_GBRAP16,   AV_PIX_FMT_RGBA64,    AV_PIX_FMT_GRAY9,
        AV_PIX_FMT_YUV420P14, AV_PIX_FMT_YUV422P14, AV_PIX_FMT_YUV444P14,
        AV_PIX_FMT_YUV440P10, AV_PIX_FMT_YUV440P12,
#if CONFIG_CUDA
        AV_PIX_FMT_CUDA,                                /* new: frames resident on the GPU */
#endif
        AV_PIX_FMT_NONE
    },
    .priv_class     = &enc_class,
};

#if CONFIG_CUDA
/* pictures that stay on the GPU: offered like libavcodec/cuviddec.c:1132-1143 */
static const AVCodecHWConfigInternal *gpu_dec_hw_configs[] = {
    &(const AVCodecHWConfigInternal) {
        .public = {
            .pix_fmt     = AV_PIX_FMT_CUDA,
            .methods     = AV_CODEC_HW_CONFIG_METHOD_HW_DEVICE_CTX | AV_CODEC_HW_CONFIG_METHOD_HW_FRAMES_CTX |
                           AV_CODEC_HW_CONFIG_METHOD_INTERNAL,
            .device_type = AV_HWDEVICE_TYPE_CUDA
        },
        .hwaccel = NULL,
    },
    NULL
};
#endif

AVCodec ff_ffv1_gpu_decoder = {
    .name           = "ffv1_gpu",
    .long_name      = NULL_IF_CONFIG_SMALL("FFmpeg video codec #1 (B200 CUDA slice path)"),
    .type           = AVMEDIA_TYPE_VIDEO,
    .id             = AV_CODEC_ID_FFV1,
    .priv_data_size = sizeof(FFV1GpuContext),
    .init           = gpu_decode_init,
    .close          = gpu_close,
    .decode         = gpu_decode_frame,        /* synchronous form, used by the vtable harness  */
    .receive_frame  = gpu_receive_frame,       /* what libavcodec/decode.c:643 calls            */
    .flush          = gpu_flush,
    .capabilities   = AV_CODEC_CAP_DR1 | AV_CODEC_CAP_DELAY,
    .caps_internal  = FF_CODEC_CAP_INIT_CLEANUP,
    .priv_class     = &dec_class,
#if CONFIG_CUDA
    .hw_configs     = gpu_dec_hw_configs,
#endif
};
