/*
 * tests/emul/tsan_e2e.c -- TEST INFRASTRUCTURE: the C host loop of bench.py (tools/e2e_driver.c:
 * an encoder thread and a decoder thread on the public C ABI) over the host-pipeline build with
 * the thread sanitizer: two routing handles over two stand-in devices, pageable pictures through
 * the copy-thread pool.  Any data race between the two caller threads and the pool is reported.
 */
#include "../../tools/e2e_driver.c"
int main(void)
{
    const int w = 352, h = 288, N = 64, NS = 4;
    ffgpu_enc_options eo; memset(&eo, 0, sizeof(eo));
    eo.width = w; eo.height = h; eo.pix_fmt = "yuv420p10le"; eo.slices = 12; eo.level = FFGPU_LEVEL_UNKNOWN;
    eo.gop_size = 1; eo.slicecrc = -1; eo.max_batch = 4; eo.pipeline_depth = 3;
    eo.ndevices = 2; eo.devices[0] = 0; eo.devices[1] = 1;
    ffgpu_encoder *enc = NULL; ffgpu_decoder *dec = NULL;
    if (ffgpu_ffv1_encode_init(&enc, &eo) < 0) { printf("enc init: %s\n", ffgpu_last_error()); return 1; }
    const uint8_t *ex; int exn = ffgpu_ffv1_encoder_extradata(enc, &ex);
    ffgpu_dec_options dopt; memset(&dopt, 0, sizeof(dopt));
    dopt.width = w; dopt.height = h; dopt.extradata = ex; dopt.extradata_size = exn; dopt.max_batch = 4; dopt.pipeline_depth = 3;
    dopt.ndevices = 2; dopt.devices[0] = 0; dopt.devices[1] = 1;
    if (ffgpu_ffv1_decode_init(&dec, &dopt) < 0) { printf("dec init: %s\n", ffgpu_last_error()); return 1; }
    ffgpu_picture src[4]; ffgpu_picture_out dst[4];
    memset(src, 0, sizeof(src)); memset(dst, 0, sizeof(dst));
    for (int i = 0; i < NS; i++) {
        int pw[3] = { w * 2, w, w }, ph[3] = { h, h / 2, h / 2 };
        for (int k = 0; k < 3; k++) {
            uint16_t *p = malloc((size_t)pw[k] * ph[k]);
            for (int j = 0; j < pw[k] / 2 * ph[k]; j++) p[j] = (uint16_t)((j * 7 + i * 13 + (j >> 6)) & 1023);
            src[i].data[k] = (const uint8_t *)p; src[i].linesize[k] = pw[k];
            dst[i].data[k] = malloc((size_t)pw[k] * ph[k]); dst[i].linesize[k] = pw[k];
        }
        src[i].sar_den = 1; src[i].pts = i;
    }
    E2ERun r; memset(&r, 0, sizeof(r));
    r.enc = enc; r.dec = dec; r.nframes = N; r.nsrc = NS; r.src = src; r.ndst = NS; r.dst = dst; r.timeout_s = 120;
    r.pkt = calloc(N, sizeof(*r.pkt)); r.pkt_size = calloc(N, sizeof(*r.pkt_size));
    int rc = ffgpu_e2e_run(&r);
    printf("rc=%d decoded=%d damaged=%d %s\n", rc, r.decoded, r.damaged, r.message);
    int bad = 0;
    for (int i = 0; i < NS; i++) for (int k = 0; k < 3; k++) {
        int pw[3] = { w * 2, w, w }, ph[3] = { h, h / 2, h / 2 };
        /* dst[i] holds the last picture decoded into it: picture N - NS + i -> src[(N-NS+i) % NS] = src[i] */
        if (memcmp(dst[i].data[k], src[i].data[k], (size_t)pw[k] * ph[k])) bad++;
    }
    printf("mismatching planes: %d\n", bad);
    ffgpu_ffv1_encode_close(enc); ffgpu_ffv1_decode_close(dec);
    return rc || bad;
}
