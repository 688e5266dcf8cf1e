/* oracle/libav_stub.c -- TEST INFRASTRUCTURE ONLY (never linked or loaded by the product).
 *
 * A stand-in for the 46 libavcodec / libavformat / libavutil / libswscale entry points that the
 * reference's video_io sources call (h264.cpp; `nm -u` of the objects built by build_ref.sh),
 * compiled against the reference's own ffmpeg 7.1 headers
 * (/root/reference/extra/ffmpeg/ffmpeg-7.1-msvc/include).  ffmpeg 7.1 + libx264 / kvazaar are
 * fetched from the network by the reference's build and are absent here, so the bitstream stage
 * is replaced by an IDENTITY "codec" and a trivial container:
 *
 *   - the "encoder" (avcodec_send_frame / avcodec_receive_packet) copies the planes of the
 *     AVFrame it is handed -- exactly as H264Capture::AddFrame (h264.cpp:1022-1238) laid them
 *     out, rows un-padded -- plus the frame's pict_type into a packet, no delay;
 *   - the "muxer" writes a fixed header and one fixed-size record per packet;
 *   - the "demuxer" / "decoder" hand the same planes back to VideoGrabber::toArray
 *     (h264.cpp:3016-3051) in an AVFrame whose rows are padded again.
 *
 * With it the UNMODIFIED reference writer (H264_Saver, incl. the lossy pre-conditioner and the
 * attribute trailer) and reader (H264_Loader -> IRFileLoader::readImage with bad-pixel removal
 * and motion correction) run end to end, and the file a writer leaves behind IS the recording of
 * the planes and key-frame decisions AddFrame produced (parse with oracle/refvio.py).  The file starts like an mp4 ("ftyp" at byte 4) because that is how
 * IRFileLoader::findFileType (IRFileLoader.cpp:117-122) recognises the writer's files.
 *
 * File layout (little endian):
 *   header  : "RIR1ftyp" | i32 codec_id | i32 width | i32 height | i32 pix_fmt | i32 fps
 *             | i64 nb_frames (patched by av_write_trailer) | i32 comment_len | comment bytes
 *   record  : u32 0x52454346 ("FCER") | i32 flags | i32 pict_type | i64 pts | i64 dts
 *             | i64 duration | i64 pos | u32 size | payload
 *   end     : u32 0x444E4546 ("FEND")
 * Anything after the end marker (the reference's attribute trailer) is ignored.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <errno.h>

#include <libavcodec/avcodec.h>
#include <libavformat/avformat.h>
#include <libavformat/avio.h>
#include <libavutil/opt.h>
#include <libavutil/dict.h>
#include <libavdevice/avdevice.h>
#include <libswscale/swscale.h>

#define REC_MAGIC 0x52454346u
#define END_MAGIC 0x444E4546u

/* ------------------------------------------------------------------ small utilities */
void *av_malloc(size_t size) { return malloc(size ? size : 1); }
void av_free(void *p) { free(p); }
void av_log_set_level(int level) { (void)level; }
void avdevice_register_all(void) {}
void sws_freeContext(struct SwsContext *c) { (void)c; }
int av_opt_set(void *obj, const char *name, const char *val, int flags)
{
    (void)obj; (void)name; (void)val; (void)flags;
    return 0;
}
void av_dump_format(AVFormatContext *ic, int index, const char *url, int is_output)
{
    (void)ic; (void)index; (void)url; (void)is_output;
}

/* ------------------------------------------------------------------ dictionary */
struct AVDictionary {
    int count;
    AVDictionaryEntry *elems;
};

int av_dict_set(AVDictionary **pm, const char *key, const char *value, int flags)
{
    (void)flags;
    AVDictionary *m = *pm;
    if (!m) {
        m = (AVDictionary *)calloc(1, sizeof(*m));
        *pm = m;
    }
    for (int i = 0; i < m->count; ++i)
        if (strcmp(m->elems[i].key, key) == 0) {
            free(m->elems[i].value);
            m->elems[i].value = strdup(value ? value : "");
            return 0;
        }
    m->elems = (AVDictionaryEntry *)realloc(m->elems, (size_t)(m->count + 1) * sizeof(AVDictionaryEntry));
    m->elems[m->count].key = strdup(key);
    m->elems[m->count].value = strdup(value ? value : "");
    m->count++;
    return 0;
}

AVDictionaryEntry *av_dict_get(const AVDictionary *m, const char *key, const AVDictionaryEntry *prev, int flags)
{
    (void)flags;
    if (!m)
        return NULL;
    int start = prev ? (int)(prev - m->elems) + 1 : 0;
    for (int i = start; i < m->count; ++i)
        if (strcmp(m->elems[i].key, key) == 0)
            return &m->elems[i];
    return NULL;
}

static void dict_free(AVDictionary **pm)
{
    AVDictionary *m = *pm;
    if (!m)
        return;
    for (int i = 0; i < m->count; ++i) {
        free(m->elems[i].key);
        free(m->elems[i].value);
    }
    free(m->elems);
    free(m);
    *pm = NULL;
}

/* ------------------------------------------------------------------ codecs */
static const AVCodec codec_x264 = {.name = "libx264", .long_name = "identity stub (yuv444p)", .type = AVMEDIA_TYPE_VIDEO, .id = AV_CODEC_ID_H264};
static const AVCodec codec_kvazaar = {.name = "libkvazaar", .long_name = "identity stub (yuv420p)", .type = AVMEDIA_TYPE_VIDEO, .id = AV_CODEC_ID_HEVC};
static const AVCodec codec_hevc = {.name = "hevc", .long_name = "identity stub", .type = AVMEDIA_TYPE_VIDEO, .id = AV_CODEC_ID_HEVC};
static const AVCodec codec_h264 = {.name = "h264", .long_name = "identity stub", .type = AVMEDIA_TYPE_VIDEO, .id = AV_CODEC_ID_H264};

const AVCodec *avcodec_find_encoder_by_name(const char *name)
{
    if (!name)
        return NULL;
    if (strcmp(name, "libx264") == 0)
        return &codec_x264;
    if (strcmp(name, "libkvazaar") == 0)
        return &codec_kvazaar;
    return NULL; /* nvenc, ffv1, vp8/9, av1, dirac: not provided */
}
const AVCodec *avcodec_find_encoder(enum AVCodecID id)
{
    if (id == AV_CODEC_ID_H264)
        return &codec_x264;
    if (id == AV_CODEC_ID_HEVC)
        return &codec_kvazaar;
    return NULL;
}
const AVCodec *avcodec_find_decoder(enum AVCodecID id)
{
    if (id == AV_CODEC_ID_H264)
        return &codec_h264;
    if (id == AV_CODEC_ID_HEVC)
        return &codec_hevc;
    return NULL;
}

/* per-context state of the identity codec */
typedef struct StubCodec {
    /* encoder: one pending packet */
    uint8_t *pending;
    int pending_size;
    int pending_pict;
    int64_t pending_pts;
    /* decoder: the last frame's padded planes */
    uint8_t *planes;
    size_t planes_size;
    int have_frame;
    int64_t frame_dts, frame_pts;
    int frame_flags;
} StubCodec;

AVCodecContext *avcodec_alloc_context3(const AVCodec *codec)
{
    AVCodecContext *c = (AVCodecContext *)calloc(1, sizeof(AVCodecContext));
    c->codec = codec;
    c->codec_type = AVMEDIA_TYPE_VIDEO;
    c->codec_id = codec ? codec->id : AV_CODEC_ID_NONE;
    c->pix_fmt = AV_PIX_FMT_NONE;
    c->priv_data = calloc(1, sizeof(StubCodec));
    return c;
}
void avcodec_free_context(AVCodecContext **pc)
{
    if (!pc || !*pc)
        return;
    StubCodec *s = (StubCodec *)(*pc)->priv_data;
    if (s) {
        free(s->pending);
        free(s->planes);
        free(s);
    }
    free(*pc);
    *pc = NULL;
}
int avcodec_close(AVCodecContext *c) { (void)c; return 0; }
int avcodec_open2(AVCodecContext *c, const AVCodec *codec, AVDictionary **options)
{
    (void)options;
    if (!c || !codec)
        return AVERROR(EINVAL);
    c->codec = codec;
    return 0;
}
void avcodec_flush_buffers(AVCodecContext *c)
{
    StubCodec *s = (StubCodec *)c->priv_data;
    s->have_frame = 0;
}
int avcodec_parameters_to_context(AVCodecContext *c, const AVCodecParameters *par)
{
    c->codec_type = par->codec_type;
    c->codec_id = par->codec_id;
    c->width = par->width;
    c->height = par->height;
    c->pix_fmt = (enum AVPixelFormat)par->format;
    return 0;
}
int avcodec_parameters_from_context(AVCodecParameters *par, const AVCodecContext *c)
{
    par->codec_type = c->codec_type;
    par->codec_id = c->codec_id;
    par->width = c->width;
    par->height = c->height;
    par->format = c->pix_fmt;
    return 0;
}
int avcodec_parameters_copy(AVCodecParameters *dst, const AVCodecParameters *src)
{
    *dst = *src;
    dst->extradata = NULL;
    dst->extradata_size = 0;
    dst->coded_side_data = NULL;
    dst->nb_coded_side_data = 0;
    return 0;
}

/* plane geometry of the two pixel formats the reference uses */
static void plane_dims(int fmt, int w, int h, int pw[3], int ph[3])
{
    pw[0] = w;
    ph[0] = h;
    if (fmt == AV_PIX_FMT_YUV420P) {
        pw[1] = pw[2] = (w + 1) / 2;
        ph[1] = ph[2] = (h + 1) / 2;
    } else { /* AV_PIX_FMT_YUV444P */
        pw[1] = pw[2] = w;
        ph[1] = ph[2] = h;
    }
}
static int align_up(int v, int a) { return (v + a - 1) / a * a; }

/* ------------------------------------------------------------------ frames */
AVFrame *av_frame_alloc(void)
{
    AVFrame *f = (AVFrame *)calloc(1, sizeof(AVFrame));
    f->format = -1;
    f->pts = AV_NOPTS_VALUE;
    f->pkt_dts = AV_NOPTS_VALUE;
    return f;
}
static void buffer_unref(AVBufferRef **pb)
{
    if (pb && *pb) {
        free((*pb)->data);
        free(*pb);
        *pb = NULL;
    }
}
void av_frame_free(AVFrame **pf)
{
    if (!pf || !*pf)
        return;
    buffer_unref(&(*pf)->buf[0]);
    free(*pf);
    *pf = NULL;
}
/* Like libavutil: every plane in ONE buffer (h264.cpp:1090 clears buf[0] to clear the frame), rows
 * padded to `align` bytes -- plus one extra alignment unit so that linesize != width even for
 * widths that are already multiples of 32 and the row-stride handling of AddFrame is exercised. */
int av_frame_get_buffer(AVFrame *f, int align)
{
    if (align <= 0)
        align = 32;
    int pw[3], ph[3];
    plane_dims(f->format, f->width, f->height, pw, ph);
    size_t total = 0, off[3];
    for (int i = 0; i < 3; ++i) {
        f->linesize[i] = align_up(pw[i], align) + align;
        off[i] = total;
        total += (size_t)f->linesize[i] * (size_t)align_up(ph[i], 32) + 64;
    }
    AVBufferRef *b = (AVBufferRef *)calloc(1, sizeof(AVBufferRef));
    b->data = (uint8_t *)malloc(total);
    memset(b->data, 0xA5, total); /* padding is poisoned: nothing may depend on it */
    b->size = total;
    f->buf[0] = b;
    for (int i = 0; i < 3; ++i)
        f->data[i] = b->data + off[i];
    return 0;
}

/* ------------------------------------------------------------------ packets */
void av_init_packet(AVPacket *pkt)
{
    pkt->pts = AV_NOPTS_VALUE;
    pkt->dts = AV_NOPTS_VALUE;
    pkt->pos = -1;
    pkt->duration = 0;
    pkt->flags = 0;
    pkt->stream_index = 0;
    pkt->buf = NULL;
    pkt->side_data = NULL;
    pkt->side_data_elems = 0;
    pkt->opaque = NULL;
    pkt->opaque_ref = NULL;
    pkt->time_base.num = 0;
    pkt->time_base.den = 1;
}
void av_packet_unref(AVPacket *pkt)
{
    buffer_unref(&pkt->buf);
    av_init_packet(pkt);
    pkt->data = NULL;
    pkt->size = 0;
}
static void packet_set_payload(AVPacket *pkt, uint8_t *data, int size)
{
    AVBufferRef *b = (AVBufferRef *)calloc(1, sizeof(AVBufferRef));
    b->data = data;
    b->size = (size_t)size;
    pkt->buf = b;
    pkt->data = data;
    pkt->size = size;
}

/* ------------------------------------------------------------------ identity encoder */
int avcodec_send_frame(AVCodecContext *c, const AVFrame *f)
{
    StubCodec *s = (StubCodec *)c->priv_data;
    if (!f)
        return 0; /* flush: nothing is ever delayed */
    int pw[3], ph[3];
    plane_dims(c->pix_fmt, c->width, c->height, pw, ph);
    size_t total = 0;
    for (int i = 0; i < 3; ++i)
        total += (size_t)pw[i] * ph[i];
    free(s->pending);
    s->pending = (uint8_t *)malloc(total);
    uint8_t *o = s->pending;
    for (int i = 0; i < 3; ++i)
        for (int y = 0; y < ph[i]; ++y) {
            memcpy(o, f->data[i] + (size_t)y * f->linesize[i], (size_t)pw[i]);
            o += pw[i];
        }
    s->pending_size = (int)total;
    s->pending_pict = (int)f->pict_type;
    s->pending_pts = f->pts;
    return 0;
}
int avcodec_receive_packet(AVCodecContext *c, AVPacket *pkt)
{
    StubCodec *s = (StubCodec *)c->priv_data;
    if (!s->pending)
        return AVERROR_EOF;
    packet_set_payload(pkt, s->pending, s->pending_size);
    s->pending = NULL;
    pkt->pts = pkt->dts = s->pending_pts;
    pkt->flags = (s->pending_pict == AV_PICTURE_TYPE_I) ? AV_PKT_FLAG_KEY : 0;
    /* the frame's pict_type travels in the packet so that the muxer can record the key-frame
     * decision of AddFrame (h264.cpp:1050-1064) */
    pkt->opaque = (void *)(intptr_t)s->pending_pict;
    return 0;
}

/* ------------------------------------------------------------------ identity decoder */
int avcodec_send_packet(AVCodecContext *c, const AVPacket *pkt)
{
    StubCodec *s = (StubCodec *)c->priv_data;
    s->have_frame = 0;
    if (!pkt || !pkt->data || pkt->size <= 0)
        return 0; /* draining */
    int pw[3], ph[3];
    plane_dims(c->pix_fmt, c->width, c->height, pw, ph);
    size_t need = 0, total = 0;
    for (int i = 0; i < 3; ++i)
        need += (size_t)pw[i] * ph[i];
    if ((size_t)pkt->size != need)
        return AVERROR_INVALIDDATA;
    int ls[3];
    size_t off[3];
    for (int i = 0; i < 3; ++i) {
        ls[i] = align_up(pw[i], 64) + 64;
        off[i] = total;
        total += (size_t)ls[i] * ph[i];
    }
    if (s->planes_size != total) {
        free(s->planes);
        s->planes = (uint8_t *)malloc(total);
        s->planes_size = total;
    }
    memset(s->planes, 0x5A, total);
    const uint8_t *in = pkt->data;
    for (int i = 0; i < 3; ++i)
        for (int y = 0; y < ph[i]; ++y) {
            memcpy(s->planes + off[i] + (size_t)y * ls[i], in, (size_t)pw[i]);
            in += pw[i];
        }
    s->have_frame = 1;
    s->frame_dts = pkt->dts;
    s->frame_pts = pkt->pts;
    s->frame_flags = pkt->flags;
    return 0;
}
int avcodec_receive_frame(AVCodecContext *c, AVFrame *f)
{
    StubCodec *s = (StubCodec *)c->priv_data;
    if (!s->have_frame)
        return AVERROR_EOF;
    s->have_frame = 0;
    int pw[3], ph[3];
    plane_dims(c->pix_fmt, c->width, c->height, pw, ph);
    size_t total = 0;
    for (int i = 0; i < 3; ++i) {
        f->linesize[i] = align_up(pw[i], 64) + 64;
        f->data[i] = s->planes + total;
        total += (size_t)f->linesize[i] * ph[i];
    }
    f->width = c->width;
    f->height = c->height;
    f->format = c->pix_fmt;
    f->pts = s->frame_pts;
    f->pkt_dts = s->frame_dts;
    f->pict_type = (s->frame_flags & AV_PKT_FLAG_KEY) ? AV_PICTURE_TYPE_I : AV_PICTURE_TYPE_P;
#if FF_API_FRAME_KEY
    f->key_frame = (s->frame_flags & AV_PKT_FLAG_KEY) ? 1 : 0;
#endif
    return 0;
}

/* ------------------------------------------------------------------ byte IO */
/* an AVIOContext is either one of ours over a FILE* (read_packet == NULL, opaque = FILE*) or the
 * caller's callbacks (avio_alloc_context; VideoGrabber::Open, h264.cpp:2795-2801) */
AVIOContext *avio_alloc_context(unsigned char *buffer, int buffer_size, int write_flag, void *opaque,
                                int (*read_packet)(void *, uint8_t *, int),
                                int (*write_packet)(void *, const uint8_t *, int),
                                int64_t (*seek)(void *, int64_t, int))
{
    AVIOContext *io = (AVIOContext *)calloc(1, sizeof(AVIOContext));
    io->buffer = buffer;
    io->buffer_size = buffer_size;
    io->write_flag = write_flag;
    io->opaque = opaque;
    io->read_packet = read_packet;
    io->write_packet = write_packet;
    io->seek = seek;
    return io;
}
int avio_open(AVIOContext **pio, const char *url, int flags)
{
    FILE *fp = fopen(url, (flags & AVIO_FLAG_WRITE) ? "w+b" : "rb");
    if (!fp)
        return AVERROR(errno ? errno : EIO);
    AVIOContext *io = (AVIOContext *)calloc(1, sizeof(AVIOContext));
    io->opaque = fp;
    io->write_flag = (flags & AVIO_FLAG_WRITE) ? 1 : 0;
    *pio = io;
    return 0;
}
int avio_close(AVIOContext *io)
{
    if (!io)
        return 0;
    int r = 0;
    if (!io->read_packet && !io->write_packet && io->opaque)
        r = fclose((FILE *)io->opaque);
    free(io);
    return r ? AVERROR(EIO) : 0;
}
int avio_closep(AVIOContext **pio)
{
    int r = avio_close(*pio);
    *pio = NULL;
    return r;
}

/* demuxer / muxer private state, hung off AVFormatContext.priv_data */
typedef struct StubFormat {
    int custom_io;     /* pb belongs to the caller */
    int64_t rpos;      /* read position (the caller's reader is shared with others) */
    int64_t data_start;
    int64_t nb_frames; /* muxer: records written */
    int64_t record_size;
    int header_written;
} StubFormat;

static int io_read(AVFormatContext *s, void *dst, int n)
{
    StubFormat *st = (StubFormat *)s->priv_data;
    AVIOContext *io = s->pb;
    int got = 0;
    if (io->read_packet) {
        if (io->seek && io->seek(io->opaque, st->rpos, SEEK_SET) < 0)
            return -1;
        while (got < n) {
            int r = io->read_packet(io->opaque, (uint8_t *)dst + got, n - got);
            if (r <= 0)
                break;
            got += r;
        }
    } else {
        FILE *fp = (FILE *)io->opaque;
        if (fseeko(fp, (off_t)st->rpos, SEEK_SET) != 0)
            return -1;
        got = (int)fread(dst, 1, (size_t)n, fp);
    }
    st->rpos += got;
    return got == n ? 0 : -1;
}

/* ------------------------------------------------------------------ muxer */
static const AVOutputFormat stub_oformat = {.name = "rirstub", .long_name = "identity stub container", .extensions = "h264,h265,mp4", .video_codec = AV_CODEC_ID_H264, .flags = 0};

const AVOutputFormat *av_guess_format(const char *short_name, const char *filename, const char *mime_type)
{
    (void)short_name; (void)filename; (void)mime_type;
    return &stub_oformat;
}
AVFormatContext *avformat_alloc_context(void)
{
    AVFormatContext *s = (AVFormatContext *)calloc(1, sizeof(AVFormatContext));
    s->priv_data = calloc(1, sizeof(StubFormat));
    s->duration = AV_NOPTS_VALUE;
    return s;
}
int avformat_alloc_output_context2(AVFormatContext **ctx, const AVOutputFormat *oformat, const char *format_name, const char *filename)
{
    (void)format_name; (void)filename;
    AVFormatContext *s = avformat_alloc_context();
    s->oformat = oformat ? oformat : &stub_oformat;
    *ctx = s;
    return 0;
}
AVStream *avformat_new_stream(AVFormatContext *s, const AVCodec *c)
{
    (void)c;
    AVStream *st = (AVStream *)calloc(1, sizeof(AVStream));
    st->codecpar = (AVCodecParameters *)calloc(1, sizeof(AVCodecParameters));
    st->codecpar->format = -1;
    st->index = (int)s->nb_streams;
    s->streams = (AVStream **)realloc(s->streams, (s->nb_streams + 1) * sizeof(AVStream *));
    s->streams[s->nb_streams++] = st;
    return st;
}
void avformat_free_context(AVFormatContext *s)
{
    if (!s)
        return;
    for (unsigned i = 0; i < s->nb_streams; ++i) {
        free(s->streams[i]->codecpar);
        free(s->streams[i]);
    }
    free(s->streams);
    dict_free(&s->metadata);
    free(s->priv_data);
    free(s);
}

static void put_i32(FILE *fp, int32_t v) { fwrite(&v, 4, 1, fp); }
static void put_i64(FILE *fp, int64_t v) { fwrite(&v, 8, 1, fp); }

int avformat_write_header(AVFormatContext *s, AVDictionary **options)
{
    (void)options;
    if (!s->pb || !s->nb_streams)
        return AVERROR(EINVAL);
    FILE *fp = (FILE *)s->pb->opaque;
    AVStream *st = s->streams[0];
    AVDictionaryEntry *e = av_dict_get(s->metadata, "comment", NULL, 0);
    const char *comment = e ? e->value : "";
    fwrite("RIR1ftyp", 1, 8, fp);
    put_i32(fp, (int32_t)st->codecpar->codec_id);
    put_i32(fp, st->codecpar->width);
    put_i32(fp, st->codecpar->height);
    put_i32(fp, st->codecpar->format);
    put_i32(fp, st->time_base.num ? st->time_base.den / st->time_base.num : 0);
    put_i64(fp, 0);
    put_i32(fp, (int32_t)strlen(comment));
    fwrite(comment, 1, strlen(comment), fp);
    ((StubFormat *)s->priv_data)->header_written = 1;
    return 0;
}
int av_interleaved_write_frame(AVFormatContext *s, AVPacket *pkt)
{
    StubFormat *st = (StubFormat *)s->priv_data;
    if (!pkt)
        return 0;
    if (!st->header_written || !s->pb)
        return AVERROR(EINVAL);
    FILE *fp = (FILE *)s->pb->opaque;
    uint32_t magic = REC_MAGIC, size = (uint32_t)pkt->size;
    fwrite(&magic, 4, 1, fp);
    put_i32(fp, pkt->flags);
    /* encoder packets carry the frame's pict_type in `opaque`; re-muxed packets carry what
     * av_read_frame found in the record */
    put_i32(fp, (int32_t)(intptr_t)pkt->opaque);
    put_i64(fp, pkt->pts);
    put_i64(fp, pkt->dts);
    put_i64(fp, pkt->duration);
    put_i64(fp, pkt->pos);
    fwrite(&size, 4, 1, fp);
    fwrite(pkt->data, 1, size, fp);
    st->nb_frames++;
    return ferror(fp) ? AVERROR(EIO) : 0;
}
int av_write_trailer(AVFormatContext *s)
{
    StubFormat *st = (StubFormat *)s->priv_data;
    if (!st->header_written || !s->pb)
        return AVERROR(EINVAL);
    FILE *fp = (FILE *)s->pb->opaque;
    uint32_t magic = END_MAGIC;
    fwrite(&magic, 4, 1, fp);
    off_t end = ftello(fp);
    fseeko(fp, 8 + 5 * 4, SEEK_SET);
    put_i64(fp, st->nb_frames);
    fseeko(fp, end, SEEK_SET);
    fflush(fp);
    return 0;
}

/* ------------------------------------------------------------------ demuxer */
int avformat_open_input(AVFormatContext **ps, const char *url, const AVInputFormat *fmt, AVDictionary **options)
{
    (void)fmt; (void)options;
    AVFormatContext *s = *ps;
    int own = 0;
    if (!s) {
        s = avformat_alloc_context();
        own = 1;
    }
    StubFormat *st = (StubFormat *)s->priv_data;
    if (s->pb) {
        st->custom_io = 1;
    } else {
        if (!url || !*url || avio_open(&s->pb, url, AVIO_FLAG_READ) < 0)
            goto fail;
    }
    {
        char magic[8];
        int32_t h[5], clen;
        int64_t nb;
        st->rpos = 0;
        if (io_read(s, magic, 8) || memcmp(magic, "RIR1ftyp", 8) != 0)
            goto fail;
        if (io_read(s, h, 20) || io_read(s, &nb, 8) || io_read(s, &clen, 4) || clen < 0 || clen > 4096)
            goto fail;
        char *comment = (char *)calloc(1, (size_t)clen + 1);
        if (clen && io_read(s, comment, clen)) {
            free(comment);
            goto fail;
        }
        if (clen)
            av_dict_set(&s->metadata, "comment", comment, 0);
        free(comment);
        st->data_start = st->rpos;
        AVStream *vs = avformat_new_stream(s, NULL);
        vs->codecpar->codec_type = AVMEDIA_TYPE_VIDEO;
        vs->codecpar->codec_id = (enum AVCodecID)h[0];
        vs->codecpar->width = h[1];
        vs->codecpar->height = h[2];
        vs->codecpar->format = h[3];
        int fps = h[4] > 0 ? h[4] : 25;
        vs->time_base.num = 1;
        vs->time_base.den = fps;
        vs->r_frame_rate.num = fps;
        vs->r_frame_rate.den = 1;
        vs->avg_frame_rate = vs->r_frame_rate;
        vs->nb_frames = nb;
        s->duration = nb * (int64_t)AV_TIME_BASE / fps;
        int pw[3], ph[3];
        plane_dims(h[3], h[1], h[2], pw, ph);
        int64_t payload = 0;
        for (int i = 0; i < 3; ++i)
            payload += (int64_t)pw[i] * ph[i];
        st->record_size = 4 + 4 + 4 + 8 * 4 + 4 + payload;
        st->nb_frames = nb;
    }
    *ps = s;
    return 0;
fail:
    if (s->pb && !st->custom_io)
        avio_closep(&s->pb);
    if (own || 1)
        avformat_free_context(s);
    *ps = NULL;
    return AVERROR_INVALIDDATA;
}
int avformat_find_stream_info(AVFormatContext *s, AVDictionary **options)
{
    (void)s; (void)options;
    return 0;
}
void avformat_close_input(AVFormatContext **ps)
{
    if (!ps || !*ps)
        return;
    AVFormatContext *s = *ps;
    StubFormat *st = (StubFormat *)s->priv_data;
    if (s->pb) {
        if (st->custom_io) {
            free(s->pb->buffer);
            free(s->pb);
        } else
            avio_close(s->pb);
        s->pb = NULL;
    }
    avformat_free_context(s);
    *ps = NULL;
}
int av_read_frame(AVFormatContext *s, AVPacket *pkt)
{
    StubFormat *st = (StubFormat *)s->priv_data;
    uint32_t magic = 0, size = 0;
    int32_t flags, pict;
    int64_t q[4];
    int64_t at = st->rpos;
    if (io_read(s, &magic, 4) || magic != REC_MAGIC) {
        st->rpos = at; /* stay on the end marker */
        return AVERROR_EOF;
    }
    if (io_read(s, &flags, 4) || io_read(s, &pict, 4) || io_read(s, q, 32) || io_read(s, &size, 4))
        return AVERROR_EOF;
    uint8_t *data = (uint8_t *)malloc(size ? size : 1);
    if (io_read(s, data, (int)size)) {
        free(data);
        return AVERROR_EOF;
    }
    av_init_packet(pkt);
    packet_set_payload(pkt, data, (int)size);
    pkt->flags = flags;
    pkt->opaque = (void *)(intptr_t)pict;
    pkt->pts = q[0];
    pkt->dts = q[1];
    pkt->duration = q[2];
    pkt->pos = q[3];
    pkt->stream_index = 0;
    return 0;
}
/* every record is a key frame of the identity codec and records have one size, so a seek lands
 * exactly on the record whose dts is the largest one <= timestamp (dts = index * 12800 after
 * H264Capture::Remux, h264.cpp:1396-1401; = index before) */
int av_seek_frame(AVFormatContext *s, int stream_index, int64_t timestamp, int flags)
{
    (void)stream_index; (void)flags;
    StubFormat *st = (StubFormat *)s->priv_data;
    if (st->nb_frames <= 0) {
        st->rpos = st->data_start;
        return 0;
    }
    /* read the dts step from the second record if there is one */
    int64_t step = 1;
    if (st->nb_frames > 1) {
        int64_t q[2];
        st->rpos = st->data_start + st->record_size + 12;
        if (io_read(s, q, 16) == 0 && q[1] > 0)
            step = q[1];
    }
    int64_t idx = timestamp <= 0 ? 0 : timestamp / step;
    if (idx >= st->nb_frames)
        idx = st->nb_frames - 1;
    st->rpos = st->data_start + idx * st->record_size;
    return 0;
}
