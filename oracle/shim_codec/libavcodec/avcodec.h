/* TEST INFRASTRUCTURE — stand-in for <libavcodec/avcodec.h> so that the reference's
 * file.c compiles UNMODIFIED into oracle/_ref.  Only saveImage()'s pixel-format
 * conversion (file.c:197-260) and saveImageDirect() (file.c:134-176) ever run; the
 * codec calls below are declared here and defined in oracle/ref_harness.c as
 * "unavailable" (they are reached only for formats saveImageDirect() refuses). */
#pragma once
#include <stdint.h>
#include "libavutil/frame.h"
enum AVCodecID { AV_CODEC_ID_NONE = 0, AV_CODEC_ID_PPM, AV_CODEC_ID_PGM, AV_CODEC_ID_PBM };
typedef struct AVRational { int num, den; } AVRational;
typedef struct AVCodec { int id; } AVCodec;
typedef struct AVCodecParameters { enum AVCodecID codec_id; int width, height, format; } AVCodecParameters;
typedef struct AVCodecContext { int width, height; int pix_fmt; AVRational time_base; } AVCodecContext;
typedef struct AVPacket { int stream_index; uint8_t *data; int size; } AVPacket;
const AVCodec *avcodec_find_decoder(enum AVCodecID id);
const AVCodec *avcodec_find_encoder(enum AVCodecID id);
AVCodecContext *avcodec_alloc_context3(const AVCodec *codec);
int avcodec_parameters_to_context(AVCodecContext *c, const AVCodecParameters *p);
int avcodec_open2(AVCodecContext *c, const AVCodec *codec, void *opts);
int avcodec_send_packet(AVCodecContext *c, const AVPacket *pkt);
int avcodec_receive_frame(AVCodecContext *c, AVFrame *f);
int avcodec_send_frame(AVCodecContext *c, const AVFrame *f);
int avcodec_receive_packet(AVCodecContext *c, AVPacket *pkt);
void avcodec_free_context(AVCodecContext **c);
AVPacket *av_packet_alloc(void);
void av_packet_free(AVPacket **pkt);
