/*
 * ffgpu.h -- C ABI of libffgpu.so: the B200-native FFV1 slice pixel path.
 *
 * This is the drop-in boundary for the reference's codec callbacks.  Every entry
 * point names the reference interface it replaces (paths relative to the
 * reference tree).  Plain C: pointers, sizes and ints only; no CUDA or C++ types.
 * All device memory, streams and pinned staging are owned by the handle.
 * There is NO CPU fallback: if no CUDA device is usable the pixel-path calls
 * return FFGPU_EXTERNAL.
 *
 * Return convention (same as libavcodec, avcodec.h:3629-3645): 0 or a positive
 * byte count on success, a negative AVERROR-compatible code on failure.
 */
#ifndef FFGPU_H
#define FFGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FFGPU_ABI_VERSION 2
#define FFGPU_MAX_DEVICES 16

/* error codes: numerically identical to libavutil/error.h so the libavcodec glue
 *  can return them unchanged */
#define FFGPU_EAGAIN       (-11)           /* AVERROR(EAGAIN): output not ready / input full      */
#define FFGPU_ENOMEM       (-12)
#define FFGPU_EINVAL       (-22)
#define FFGPU_ENOSPC       (-28)
#define FFGPU_ENOSYS       (-38)
#define FFGPU_EOF          (-541478725)    /* AVERROR_EOF                                          */
#define FFGPU_INVALIDDATA  (-1094995529)   /* AVERROR_INVALIDDATA                                  */
#define FFGPU_EXTERNAL     (-542398533)    /* AVERROR_EXTERNAL: CUDA runtime/driver failure        */

/* coder option values: the "coder" AVOption of ffv1enc.c:1291-1303 */
#define FFGPU_CODER_RICE        0
#define FFGPU_CODER_AC          1
#define FFGPU_CODER_RANGE_TAB   2
#define FFGPU_CODER_RANGE_DEF (-2)

#define FFGPU_LEVEL_UNKNOWN   (-99)        /* FF_LEVEL_UNKNOWN */

/* ------------------------------------------------------------------------
 * Options read by encode_init (ffv1enc.c:517-928): the codec-private AVOptions
 * (slicecrc, coder, context; ffv1enc.c:1291-1307) and the AVCodecContext fields
 * it consults (width, height, pix_fmt, slices, level, gop_size,
 * strict_std_compliance, bits_per_raw_sample).  The last block is new: where
 * and how wide to run on the GPU.
 * ------------------------------------------------------------------------ */
typedef struct ffgpu_enc_options {
    int width, height;
    const char *pix_fmt;          /* libavutil pixdesc name, little-endian spelling ("yuv420p10le") */
    int slices;                   /* AVCodecContext.slices        0 = automatic                     */
    int level;                    /* AVCodecContext.level         FFGPU_LEVEL_UNKNOWN = automatic   */
    int gop_size;                 /* AVCodecContext.gop_size      (libavcodec default 12)           */
    int coder;                    /* AVOption "coder"                                               */
    int context;                  /* AVOption "context" 0/1                                         */
    int slicecrc;                 /* AVOption "slicecrc" -1 auto, 0, 1                              */
    int strict_std_compliance;    /* AVCodecContext.strict_std_compliance                           */
    int bits_per_raw_sample;      /* AVCodecContext.bits_per_raw_sample, 0 = from pix_fmt           */
    /* GPU placement */
    int device;                   /* CUDA ordinal                                                   */
    int max_batch;                /* frames coded per launch group when every frame is a key frame  */
                                  /* (gop_size <= 1); 0 = default                                   */
    int pipeline_depth;           /* launch groups in flight for send/receive; 0 = default          */
    /* More than one GPU: with ndevices > 1 the handle spreads the stream over
     * devices[0..ndevices-1] -- picture i goes to GPU i mod N when every frame is a key frame
     * (gop_size <= 1), whole GOPs round-robin otherwise (adaptive states carry inside a GOP,
     * ffv1enc.c:1071) -- and receive_packet hands the packets back in presentation order, the
     * way the reference's frame threads do (pthread_frame.c; AV_CODEC_CAP_DELAY
     * ffv1enc.c:1332).  ndevices <= 1: the single GPU `device`. */
    int ndevices;
    int devices[FFGPU_MAX_DEVICES];
    /* Two-pass coding (AV_CODEC_FLAG_PASS1 / _PASS2 and AVCodecContext.stats_in,
     * ffv1enc.c:528, :785-873, :1134-1177).  pass1: the range coder's decisions are counted
     * (rc_stat / rc_stat2) and ffgpu_ffv1_encoder_stats_out() returns the text the reference
     * leaves in AVCodecContext.stats_out after the flush.  stats_in: that text, possibly of
     * several first passes concatenated; encode_init derives the sorted state transition
     * table and the initial states per context from it and writes them to the extradata. */
    int pass1, pass2;
    const char *stats_in;
} ffgpu_enc_options;

/* one picture: AVFrame.data/linesize plus the per-frame fields the slice header carries
 * (ffv1enc.c:944-949).  data[] may point to pageable or pinned HOST memory, or -- frames
 * that already live on the GPU, AV_PIX_FMT_CUDA -- to CUDA DEVICE memory; the
 * copy kind is taken from the pointer (unified addressing).  The planes must stay valid and
 * unchanged until the picture's packet has been received (the glue keeps an av_frame_ref). */
typedef struct ffgpu_picture {
    const uint8_t *data[4];
    int linesize[4];
    int interlaced_frame, top_field_first;
    int sar_num, sar_den;         /* AVFrame.sample_aspect_ratio */
    int64_t pts;
} ffgpu_picture;

typedef struct ffgpu_encoder ffgpu_encoder;
typedef struct ffgpu_decoder ffgpu_decoder;

/* ---- encoder: replaces ff_ffv1_encoder (ffv1enc.c:1323-1360) ---- */

/* AVCodec.init = encode_init, ffv1enc.c:517.  Host only: resolves version, coder, slice
 * grid, quant tables and writes the extradata; device resources are created lazily by the
 * first pixel-path call.  Same error codes as the reference for the same bad options. */
int ffgpu_ffv1_encode_init(ffgpu_encoder **enc, const ffgpu_enc_options *opt);

/* AVCodecContext.extradata written by write_extradata, ffv1enc.c:396-467.
 * Returns the size (0 for version < 2 streams). */
int ffgpu_ffv1_encoder_extradata(const ffgpu_encoder *enc, const uint8_t **data);

/* info[8]: version, micro_version, ac, num_h_slices, num_v_slices, ec,
 * bits_per_raw_sample, colorspace (what -debug pict prints, ffv1dec.c:512-526) */
void ffgpu_ffv1_encoder_info(const ffgpu_encoder *enc, int info[8]);

/* upper bound of one packet for this stream (what the caller must provide) */
size_t ffgpu_ffv1_encoder_max_packet(const ffgpu_encoder *enc);

/* AVCodec.encode2 = encode_frame, ffv1enc.c:1122-1281, synchronous: one picture in host
 * memory in, one packet in host memory out.  *key_frame mirrors AV_PKT_FLAG_KEY.
 * Returns 0, FFGPU_ENOSPC if pkt_cap is too small, FFGPU_INVALIDDATA
 * ("encoded frame too large", ffv1enc_template.c:34-44) if a slice outgrows its arena. */
int ffgpu_ffv1_encode_frame(ffgpu_encoder *enc, const ffgpu_picture *pic,
                            uint8_t *pkt, size_t pkt_cap, size_t *pkt_size, int *key_frame);

/* AVCodec.send_frame / receive_packet (avcodec.h:3654-3662) for an encoder that sets
 * AV_CODEC_CAP_DELAY (ffv1enc.c:1332): pictures are queued, uploaded and coded in launch
 * groups on CUDA streams; packets come back in presentation order.
 * send: pic == NULL flushes.  FFGPU_EAGAIN = queue full, call receive first.
 * receive: FFGPU_EAGAIN = nothing ready yet (send more or flush), FFGPU_EOF after a flush
 * once every packet was returned; the EOF also ends the flush, so the handle accepts
 * pictures again afterwards (avcodec_flush_buffers semantics). */
int ffgpu_ffv1_encode_send_frame(ffgpu_encoder *enc, const ffgpu_picture *pic);
int ffgpu_ffv1_encode_receive_packet(ffgpu_encoder *enc, uint8_t *pkt, size_t pkt_cap,
                                     size_t *pkt_size, int *key_frame, int64_t *pts);
/* receive_packet without taking the packet: 0 and its exact size when the next packet is
 * ready (so that the caller can allocate the AVPacket, ff_alloc_packet2, only then and only
 * that large), else FFGPU_EAGAIN / FFGPU_EOF / an error exactly as receive_packet would.
 * FFGPU_EOF is reported once: it ends the flush like receive_packet's. */
int ffgpu_ffv1_encode_packet_ready(ffgpu_encoder *enc, size_t *pkt_size);

/* Pictures already resident in device memory ("frames stay on the GPU").
 * d_frames holds nframes pictures back to back in the layout ffgpu_ffv1_frame_layout()
 * describes; cuda_stream is a cudaStream_t (NULL = the encoder's own stream).  Packets are
 * left in device memory; ffgpu_ffv1_encode_device_result() exposes them once the stream
 * has been synchronised.  All nframes must be key frames (gop_size <= 1) when nframes > 1. */
int ffgpu_ffv1_encode_device(ffgpu_encoder *enc, const void *d_frames, int nframes,
                             void *cuda_stream);
int ffgpu_ffv1_encode_device_result(ffgpu_encoder *enc, int frame, const void **d_pkt,
                                    size_t *pkt_size);
/* copy the packets of the last ffgpu_ffv1_encode_device() call to host memory */
int ffgpu_ffv1_encode_device_fetch(ffgpu_encoder *enc, int frame, uint8_t *pkt, size_t pkt_cap,
                                   size_t *pkt_size);

/* AVCodecContext.stats_out of a first pass (ffv1enc.c:1134-1177): the decision statistics of
 * every picture coded so far, as text.  Call it after the flush (it waits for the pictures
 * in flight).  Returns the length, or FFGPU_ENOSPC / FFGPU_EINVAL (not a first pass). */
int ffgpu_ffv1_encoder_stats_out(ffgpu_encoder *enc, char *buf, size_t cap);

/* AVCodec.close = encode_close, ffv1enc.c:1283 */
int ffgpu_ffv1_encode_close(ffgpu_encoder *enc);

/* ---- decoder: replaces ff_ffv1_decoder (ffv1dec.c:1087-1101) ---- */

typedef struct ffgpu_dec_options {
    int width, height;            /* AVCodecContext.width/height (from the container)  */
    const uint8_t *extradata;     /* AVCodecContext.extradata, may be NULL (v0/v1)      */
    int extradata_size;
    int device;
    int max_batch;                /* packets decoded per launch group (intra-only streams) */
    int pipeline_depth;
    /* More than one GPU: packet i is decoded on GPU i mod N and receive_frame returns the
     * pictures in packet order.  Only streams whose every frame is a key frame (version 3,
     * intra flag in the extradata) can be spread; any other stream runs on devices[0]. */
    int ndevices;
    int devices[FFGPU_MAX_DEVICES];
} ffgpu_dec_options;

typedef struct ffgpu_picture_out {
    uint8_t *data[4];             /* caller-provided destination planes (host memory)  */
    int linesize[4];
    /* filled on return */
    int key_frame;
    int interlaced_frame, top_field_first;
    int sar_num, sar_den;
    int damaged_slices;           /* slices concealed / left untouched (ffv1dec.c:940) */
    int64_t pts;
} ffgpu_picture_out;

/* AVCodec.init = decode_init, ffv1dec.c:818-835 (+ read_extra_header :413-528) */
int ffgpu_ffv1_decode_init(ffgpu_decoder **dec, const ffgpu_dec_options *opt);

/* pixdesc name of the output format chosen by read_header (ffv1dec.c:597-739); valid after
 * init for streams with extradata, else after the first key frame.  NULL if unknown. */
const char *ffgpu_ffv1_decoder_pix_fmt(const ffgpu_decoder *dec);
void ffgpu_ffv1_decoder_info(const ffgpu_decoder *dec, int info[8]);
/* v0/v1 streams carry their parameters in the first key frame instead of extradata
 * (read_header, ffv1dec.c:538-590): parse that header without decoding anything, so that
 * the output format is known before the caller allocates the picture.  0 or an error. */
int ffgpu_ffv1_decoder_probe(ffgpu_decoder *dec, const uint8_t *pkt, size_t pkt_size);

/* AVCodec.decode = decode_frame, ffv1dec.c:837-983, synchronous.  Returns the number of
 * bytes consumed (pkt_size) like the reference, negative on error. */
int ffgpu_ffv1_decode_frame(ffgpu_decoder *dec, const uint8_t *pkt, size_t pkt_size,
                            ffgpu_picture_out *out, int *got_frame);

/* AVCodec.send_packet / receive_frame semantics, pipelined over CUDA streams.
 * send: pkt == NULL flushes.  dst (optional) names the destination planes up front, the
 * way ff_thread_get_buffer() hands decode_frame its AVFrame before the slices are decoded
 * (ffv1dec.c:881): the picture is then downloaded straight into them by the group's
 * stream and receive_frame only reports completion and the frame properties.  Without dst
 * the picture is copied out inside receive_frame, into `out`.
 * FFGPU_EAGAIN / FFGPU_EOF as for the encoder. */
int ffgpu_ffv1_decode_send_packet(ffgpu_decoder *dec, const uint8_t *pkt, size_t pkt_size,
                                  int64_t pts, const ffgpu_picture_out *dst);
int ffgpu_ffv1_decode_receive_frame(ffgpu_decoder *dec, ffgpu_picture_out *out);

/* Decode nframes packets (host memory) into device-resident pictures laid out per
 * ffgpu_ffv1_frame_layout(); nothing is copied back. */
int ffgpu_ffv1_decode_device(ffgpu_decoder *dec, const uint8_t *const *pkts,
                             const size_t *pkt_sizes, int nframes, void *d_frames,
                             void *cuda_stream);

/* Damage report of the last ffgpu_ffv1_decode_device() batch (the send/receive path reports
 * the same through ffgpu_picture_out.damaged_slices): waits for the batch, then
 * damaged_slices[i] = slices of picture i with a CRC mismatch, an invalid slice header or a
 * bytestream-end mismatch (ffv1dec.c:905-922, :296-300, :351-359).  Device batches are NOT
 * concealed (there is no previous picture the library owns); a caller that needs the
 * reference's concealment copies the damaged pictures' predecessor itself or uses
 * send_packet/receive_frame.  Returns the total number of damaged slices or a negative error. */
int ffgpu_ffv1_decode_device_status(ffgpu_decoder *dec, int *damaged_slices, int nframes);

int ffgpu_ffv1_decode_close(ffgpu_decoder *dec);

/* ---- shared helpers ---- */

/* device picture layout used by the *_device entry points: per memory plane the byte
 * offset inside one picture, the row pitch and the number of rows; returns the picture
 * stride in bytes (0 on error). */
size_t ffgpu_ffv1_frame_layout(const char *pix_fmt, int width, int height,
                               size_t plane_offset[4], int plane_pitch[4], int plane_rows[4],
                               int plane_rowbytes[4]);

/* launch statistics of the handle (kernels launched by this library so far) */
uint64_t ffgpu_ffv1_encoder_launches(const ffgpu_encoder *enc);
uint64_t ffgpu_ffv1_decoder_launches(const ffgpu_decoder *dec);

/* Profiling of the *_device entry points: when enabled, CUDA events are recorded around
 * every kernel of the launch group; *_kernel_ms() returns the elapsed times of the last
 * group (after the stream was synchronised), in launch order:
 *   encoder: symbolize, fill_state, sort, code, pack_slice_scan, pack_frame_scan, pack_gather
 *   decoder: init_state, sort, decode
 * and the number of kernels, or a negative error. */
int ffgpu_ffv1_encoder_profile(ffgpu_encoder *enc, int enable);
int ffgpu_ffv1_encoder_kernel_ms(ffgpu_encoder *enc, float *ms, int n);
int ffgpu_ffv1_decoder_profile(ffgpu_decoder *dec, int enable);
int ffgpu_ffv1_decoder_kernel_ms(ffgpu_decoder *dec, float *ms, int n);

/* Serial cost of the last ffgpu_ffv1_encode_device() batch as stage A counted it: the number
 * of binary range-coder decisions of all slices (put_rac calls of encode_line,
 * ffv1enc_template.c:58-77 / ffv1enc.c:185-231) and of the heaviest slice.  The decoder
 * takes exactly the same decisions for the same stream.  Waits for the batch.  Zero for
 * Golomb-Rice streams (no binary decisions). */
int ffgpu_ffv1_encoder_decisions(ffgpu_encoder *enc, uint64_t *total, uint32_t *heaviest_slice);

/* AV_PIX_FMT_CUDA frames belong to the CUcontext of libavutil's CUDA hw device
 * (AVCUDADeviceContext.cuda_ctx, libavutil/hwcontext_cuda.h).  The glue brackets EVERY call
 * into this library -- init included, so that the handle's memory lives there -- with
 * push(cuda_ctx) ... pop(), as libavcodec/nvenc.c:1329-1340 does for NVENC.  Host-memory
 * callers never need these. */
int ffgpu_cuda_push_context(void *cu_context);
int ffgpu_cuda_pop_context(void);

/* last error text of the calling thread ("" if none) */
const char *ffgpu_last_error(void);

int ffgpu_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FFGPU_H */
