#!/bin/bash
# TEST INFRASTRUCTURE.  Builds the reference's own `ffmpeg` program -- from a scratch COPY of
# /root/reference, which is never modified -- with integration/ffv1_gpu.c added to libavcodec
# exactly as INTEGRATION.md describes (one file, two Makefile lines, two allcodecs.c lines)
# and linked against libffgpu.so.  Output: oracle/_ref/ffmpeg (git-ignored, travels to the
# GPU box).  It serves three purposes:
#   - the drop-in proof: `ffmpeg -c:v ffv1_gpu` against `ffmpeg -c:v ffv1` (framemd5 / packet md5)
#   - the lavfi sources BASELINE.json names (testsrc2, mandelbrot, noise) for bench.py
#   - a second CPU baseline (the reference codec inside its own program)
# configure flags: SURVEY.md 8c (C only, no asm: binutils here rejects x86/mathops.h:125).
set -e
REPO="$(cd "$(dirname "$0")/.." && pwd)"
SRC=${FFV1_REFERENCE:-/root/reference}
WORK=${FFV1_FFBUILD:-/tmp/ffv1gpu_ffbuild}
OUT="$REPO/oracle/_ref"
[ -d "$SRC/libavcodec" ] || { echo "reference tree not found at $SRC" >&2; exit 1; }
[ -f "$REPO/ffmpeg_ffv2_b200/libffgpu.so" ] || { echo "build libffgpu.so first" >&2; exit 1; }
mkdir -p "$OUT" "$WORK"
if [ ! -d "$WORK/src/libavcodec" ]; then
    mkdir -p "$WORK/src"
    (cd "$SRC" && tar cf - --exclude=.git .) | (cd "$WORK/src" && tar xf -)
    chmod -R u+w "$WORK/src"
    # the three edits of INTEGRATION.md
    sed -i 's|^OBJS-$(CONFIG_FFV1_ENCODER) .*|&\nOBJS-$(CONFIG_FFV1_GPU_ENCODER)        += ffv1_gpu.o\nOBJS-$(CONFIG_FFV1_GPU_DECODER)        += ffv1_gpu.o|' "$WORK/src/libavcodec/Makefile"
    sed -i 's|^extern AVCodec ff_ffv1_decoder;|&\nextern AVCodec ff_ffv1_gpu_encoder;\nextern AVCodec ff_ffv1_gpu_decoder;|' "$WORK/src/libavcodec/allcodecs.c"
fi
cp "$REPO/integration/ffv1_gpu.c" "$WORK/src/libavcodec/ffv1_gpu.c"
cp "$REPO/include/ffgpu.h" "$WORK/src/libavcodec/ffgpu.h"
mkdir -p "$WORK/build"
cd "$WORK/build"
FILTERS=testsrc2,testsrc,mandelbrot,scale,format,null,noise,nullsrc,geq,color,trim,select,vflip
# configure again when the feature list above changed (the scratch tree outlives a run)
if [ ! -f ffbuild/config.mak ] || [ "$(cat .filters 2>/dev/null)" != "$FILTERS" ]; then
    "$WORK/src/configure" --disable-asm --disable-doc --disable-everything --disable-autodetect \
        --enable-encoder=ffv1,ffv1_gpu,rawvideo,wrapped_avframe \
        --enable-decoder=ffv1,ffv1_gpu,rawvideo,wrapped_avframe \
        --enable-muxer=nut,avi,matroska,framemd5,framecrc,md5,null,rawvideo \
        --enable-demuxer=nut,avi,matroska,rawvideo \
        --enable-protocol=file,pipe,md5 --enable-indev=lavfi \
        --enable-filter=$FILTERS \
        --disable-ffplay --disable-ffprobe \
        --extra-ldflags="-L$REPO/ffmpeg_ffv2_b200" --extra-libs="-lffgpu -lpthread -ldl -lrt" \
        > configure.log 2>&1 || { tail -20 configure.log; tail -30 ffbuild/config.log; exit 1; }
    echo "$FILTERS" > .filters
fi
make -j"$(nproc)" ffmpeg > make.log 2>&1 || { tail -30 make.log; exit 1; }
cp ffmpeg "$OUT/ffmpeg"
echo "built $OUT/ffmpeg"
