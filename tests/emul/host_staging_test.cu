/*
 * tests/emul/host_staging_test.cu -- TEST INFRASTRUCTURE: the host-side staging copies of
 * ffgpu_api.cu (stage_picture, the copy-thread pool, needs_staging) run without a GPU.  The
 * file includes the library source so that its static functions can be called directly;
 * nothing here touches the device.  Top-down and bottom-up (negative linesize) pictures, padded
 * caller pitches, every plane layout, 1 and several copy threads (FFGPU_COPY_THREADS).
 */
#include "../../ffmpeg_ffv2_b200/csrc/ffgpu_api.cu"
#include <vector>
int main()
{
    const char *fmts[] = { "yuv420p10le", "bgr0", "yuv420p", "gbrp16le", "ya8" };
    int bad = 0;
    for (const char *fmt : fmts)
        for (int w : { 33, 176, 1920 })
            for (int h : { 7, 98, 1081 }) {
                ffgpu_enc_options o; memset(&o, 0, sizeof(o));
                o.width = w; o.height = h; o.pix_fmt = fmt; o.level = FFGPU_LEVEL_UNKNOWN; o.gop_size = 1; o.slicecrc = -1;
                FFStream s; memset(&s, 0, sizeof(s));
                if (ff_stream_from_options(&s, &o) < 0) { printf("opt %s\n", fmt); return 1; }
                std::vector<FFDevSlice> sl(s.nh * s.nv);
                FFDevParams P;
                ff_fill_dev_params(&s, 1, &P, sl.data());
                std::vector<uint8_t> stage(P.frame_bytes, 0xEE), back(P.frame_bytes, 0);
                std::vector<std::vector<uint8_t>> planes(4), planes2(4);
                uint8_t *data[4] = {0,0,0,0}, *data2[4] = {0,0,0,0}; int ls[4] = {0,0,0,0}, ls2[4] = {0,0,0,0};
                for (int k = 0; k < s.pf->nplanes; k++) {
                    int rb, rows; ff_plane_geometry(s.pf, w, h, k, &rb, &rows);
                    const int pitch = rb + 13;
                    planes[k].resize((size_t)pitch * rows); planes2[k].assign((size_t)pitch * rows, 0);
                    for (size_t i = 0; i < planes[k].size(); i++) planes[k][i] = (uint8_t)(i * 7 + k * 31 + (i >> 8));
                    data[k] = planes[k].data() + (size_t)pitch * (rows - 1); ls[k] = -pitch;      // bottom-up
                    data2[k] = planes2[k].data() + (size_t)pitch * (rows - 1); ls2[k] = -pitch;
                }
                if (stage_picture(&P, s.pf, w, h, data, ls, stage.data(), 1) < 0) { printf("stage failed\n"); return 1; }
                // expected: row y of the picture (top-down) = memory row rows-1-y
                for (int k = 0; k < s.pf->nplanes; k++) {
                    int rb, rows; ff_plane_geometry(s.pf, w, h, k, &rb, &rows);
                    const int pitch = rb + 13;
                    for (int y = 0; y < rows; y++)
                        if (memcmp(stage.data() + P.plane_off[k] + (size_t)y * P.pitch[k],
                                   planes[k].data() + (size_t)(rows - 1 - y) * pitch, rb)) { bad++; break; }
                }
                if (stage_picture(&P, s.pf, w, h, data2, ls2, stage.data(), 0) < 0) { printf("unstage failed\n"); return 1; }
                for (int k = 0; k < s.pf->nplanes; k++) {
                    int rb, rows; ff_plane_geometry(s.pf, w, h, k, &rb, &rows);
                    const int pitch = rb + 13;
                    for (int y = 0; y < rows; y++)
                        if (memcmp(planes2[k].data() + (size_t)y * pitch, planes[k].data() + (size_t)y * pitch, rb)) { bad++; break; }
                }
                // and the ordinary top-down picture still round-trips
                for (int k = 0; k < s.pf->nplanes; k++) { int rb, rows; ff_plane_geometry(s.pf, w, h, k, &rb, &rows);
                    data[k] = planes[k].data(); ls[k] = rb + 13; data2[k] = planes2[k].data(); ls2[k] = rb + 13; planes2[k].assign(planes2[k].size(), 0); }
                stage_picture(&P, s.pf, w, h, data, ls, stage.data(), 1);
                stage_picture(&P, s.pf, w, h, data2, ls2, stage.data(), 0);
                for (int k = 0; k < s.pf->nplanes; k++) { int rb, rows; ff_plane_geometry(s.pf, w, h, k, &rb, &rows);
                    for (int y = 0; y < rows; y++) if (memcmp(planes2[k].data() + (size_t)y * (rb + 13), planes[k].data() + (size_t)y * (rb + 13), rb)) { bad++; break; } }
                int neg[4] = { -1, 0, 0, 0 };
                if (needs_staging(planes[0].data(), neg) != 1) { printf("needs_staging\n"); bad++; }
                ff_stream_free(&s);
            }
    printf(bad ? "FAILED %d\n" : "host copy ok\n", bad);
    return bad != 0;
}
