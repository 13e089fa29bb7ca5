/* TEST INFRASTRUCTURE — stand-in for <libavformat/avformat.h>; see libavcodec/avcodec.h here. */
#pragma once
#include "libavcodec/avcodec.h"
#define AVIO_FLAG_WRITE 2
typedef struct AVIOContext AVIOContext;
typedef struct AVStream { AVCodecParameters *codecpar; AVRational time_base; } AVStream;
typedef struct AVFormatContext { unsigned nb_streams; AVStream **streams; void *priv_data; AVIOContext *pb; } AVFormatContext;
int avformat_open_input(AVFormatContext **s, const char *url, void *fmt, void *opts);
int avformat_find_stream_info(AVFormatContext *s, void *opts);
void av_dump_format(AVFormatContext *s, int index, const char *url, int is_output);
int av_read_frame(AVFormatContext *s, AVPacket *pkt);
void avformat_close_input(AVFormatContext **s);
int avformat_alloc_output_context2(AVFormatContext **ctx, void *oformat, const char *format_name, const char *filename);
AVStream *avformat_new_stream(AVFormatContext *s, const AVCodec *c);
int avio_open(AVIOContext **s, const char *url, int flags);
int avformat_write_header(AVFormatContext *s, void *opts);
int av_write_frame(AVFormatContext *s, AVPacket *pkt);
int av_write_trailer(AVFormatContext *s);
void avformat_free_context(AVFormatContext *s);
