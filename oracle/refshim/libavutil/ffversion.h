/* hand-written stand-in for the generated libavutil/ffversion.h */
#ifndef AVUTIL_FFVERSION_H
#define AVUTIL_FFVERSION_H
#define FFMPEG_VERSION "oracle-refshim"
#endif
