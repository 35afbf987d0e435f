// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product.
//
// A thin extern "C" door onto the UNMODIFIED reference classes, compiled by
// oracle/Makefile together with the reference's own translation units straight
// from /root/reference (nothing is copied into this repo).  It lets tests and
// the bench's cpu_baseline leg call
//   ImageScaler::Create/Scale                  (src/image-scaler.h:33-39)
//   Framebuffer::AlphaComposeBackground        (src/framebuffer.h:103-106)
//   UnicodeBlockCanvas::Send                   (src/unicode-block-canvas.h:45-46)
//   ImageSource::CalcScaleToFitDisplay         (src/image-source.h:84-87)
//   rgba_t::As256TermColor / ParseColor        (src/framebuffer.h:37-58)
// with plain pointers.  Bytes written by the canvas are captured through the
// reference's own BufferedWriteSequencer into a memfd (or /dev/null for timing).
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <csignal>
#include <cstdint>
#include <cstring>
#include <memory>

#include "buffered-write-sequencer.h"
#include "display-options.h"
#include "framebuffer.h"
#include "image-scaler.h"
#include "image-source.h"
#include "terminal-canvas.h"
#include "unicode-block-canvas.h"

namespace {
static volatile sig_atomic_t g_never_interrupted = 0;

// CalcScaleToFitDisplay is protected static; a derived struct re-exports it.
struct FitDoor : public timg::ImageSource {
    using timg::ImageSource::CalcScaleToFitDisplay;
};

inline timg::rgba_t unpack(uint32_t v) {
    timg::rgba_t c;
    memcpy(&c, &v, 4);
    return c;
}

struct BlockCanvasDoor {
    int fd;
    off_t consumed = 0;
    timg::BufferedWriteSequencer *seq;
    timg::UnicodeBlockCanvas *canvas;
};
}  // namespace

extern "C" {

int ref_calc_fit(int img_w, int img_h, int width, int height, int cell_x_px,
                 int cell_y_px, float width_stretch, int upscale,
                 int upscale_integer, int fill_width, int fill_height,
                 int fit_in_rotated, int *tw, int *th) {
    timg::DisplayOptions o;
    o.width           = width;
    o.height          = height;
    o.cell_x_px       = cell_x_px;
    o.cell_y_px       = cell_y_px;
    o.width_stretch   = width_stretch;
    o.upscale         = upscale;
    o.upscale_integer = upscale_integer;
    o.fill_width      = fill_width;
    o.fill_height     = fill_height;
    return FitDoor::CalcScaleToFitDisplay(img_w, img_h, o, fit_in_rotated, tw,
                                          th)
               ? 1
               : 0;
}

// fmt: 0 = kRGBA, 1 = kRGB32 (BGRA in memory).
int ref_scale(const uint8_t *in, int iw, int ih, int fmt, uint8_t *out, int ow,
              int oh) {
    timg::Framebuffer src(iw, ih);
    memcpy((void *)src.begin(), in, (size_t)iw * ih * 4);
    timg::Framebuffer dst(ow, oh);
    auto scaler = timg::ImageScaler::Create(
        iw, ih,
        fmt == 0 ? timg::ImageScaler::ColorFmt::kRGBA
                 : timg::ImageScaler::ColorFmt::kRGB32,
        ow, oh);
    if (!scaler) return -1;
    scaler->Scale(src, &dst);
    memcpy(out, (const void *)dst.begin(), (size_t)ow * oh * 4);
    return 0;
}

// Persistent-framebuffer variant for timing (no per-call alloc/copy).
void *ref_fb_new(int w, int h) { return new timg::Framebuffer(w, h); }
void ref_fb_free(void *fb) { delete (timg::Framebuffer *)fb; }
uint8_t *ref_fb_data(void *fb) {
    return (uint8_t *)((timg::Framebuffer *)fb)->begin();
}
int ref_scale_fb(void *in_fb, void *out_fb) {
    timg::Framebuffer *in  = (timg::Framebuffer *)in_fb;
    timg::Framebuffer *out = (timg::Framebuffer *)out_fb;
    auto scaler            = timg::ImageScaler::Create(
        in->width(), in->height(), timg::ImageScaler::ColorFmt::kRGBA,
        out->width(), out->height());
    if (!scaler) return -1;
    scaler->Scale(*in, out);
    return 0;
}

// bg / pattern are the 4 rgba_t bytes in memory order (r | g<<8 | b<<16 | a<<24).
// has_bg == 0 models "-b none" (a null bgcolor_getter).
void ref_compose(uint8_t *fb, int w, int h, int has_bg, uint32_t bg,
                 uint32_t pattern, int pw, int ph, int start_row) {
    timg::Framebuffer f(w, h);
    memcpy((void *)f.begin(), fb, (size_t)w * h * 4);
    const timg::rgba_t bgc = unpack(bg);
    timg::Framebuffer::bgcolor_query q;
    if (has_bg) q = [bgc]() { return bgc; };
    f.AlphaComposeBackground(q, unpack(pattern), pw, ph, start_row);
    memcpy(fb, (const void *)f.begin(), (size_t)w * h * 4);
}
void ref_compose_fb(void *fbp, int has_bg, uint32_t bg, uint32_t pattern,
                    int pw, int ph, int start_row) {
    timg::Framebuffer *f   = (timg::Framebuffer *)fbp;
    const timg::rgba_t bgc = unpack(bg);
    timg::Framebuffer::bgcolor_query q;
    if (has_bg) q = [bgc]() { return bgc; };
    f->AlphaComposeBackground(q, unpack(pattern), pw, ph, start_row);
}

int ref_as256(uint32_t rgba) { return unpack(rgba).As256TermColor(); }

uint32_t ref_parse_color(const char *s) {
    const timg::rgba_t c = timg::rgba_t::ParseColor(s);
    uint32_t v;
    memcpy(&v, &c, 4);
    return v;
}

// capture != 0: bytes go to a memfd and are returned by ref_blocks_send;
// capture == 0: bytes go to /dev/null (timing runs).
void *ref_blocks_new(int quarter, int upper, int color8, int capture) {
    BlockCanvasDoor *d = new BlockCanvasDoor;
    d->fd = capture ? memfd_create("timg_ref_out", 0) : open("/dev/null", O_WRONLY);
    d->seq = new timg::BufferedWriteSequencer(d->fd, false, 4, true,
                                              g_never_interrupted);
    d->canvas = new timg::UnicodeBlockCanvas(d->seq, quarter, upper, color8);
    return d;
}

void ref_blocks_prefix(void *h, const char *data, int len) {
    ((BlockCanvasDoor *)h)->canvas->AddPrefixNextSend(data, len);
}

// Returns number of bytes the canvas produced for this Send (copied to out if
// capturing and they fit), or -1 if they did not fit.
long ref_blocks_send(void *h, int x, int dy, const uint8_t *fb, int w, int hgt,
                     int seq_type, char *out, long cap) {
    BlockCanvasDoor *d = (BlockCanvasDoor *)h;
    timg::Framebuffer f(w, hgt);
    memcpy((void *)f.begin(), fb, (size_t)w * hgt * 4);
    d->canvas->Send(x, dy, f, (timg::SeqType)seq_type, timg::Duration());
    d->seq->Flush();
    if (!out) return 0;
    const off_t end = lseek(d->fd, 0, SEEK_END);
    const long n    = (long)(end - d->consumed);
    if (n > cap) return -1;
    if (n > 0 && pread(d->fd, out, n, d->consumed) != n) return -2;
    d->consumed = end;
    return n;
}
// Timing variant on a persistent framebuffer; no flush, no copy.
void ref_blocks_send_fb(void *h, int x, int dy, void *fbp) {
    BlockCanvasDoor *d = (BlockCanvasDoor *)h;
    d->canvas->Send(x, dy, *(timg::Framebuffer *)fbp,
                    timg::SeqType::FrameImmediate, timg::Duration());
}
void ref_blocks_flush(void *h) { ((BlockCanvasDoor *)h)->seq->Flush(); }

void ref_blocks_free(void *h) {
    BlockCanvasDoor *d = (BlockCanvasDoor *)h;
    delete d->canvas;
    delete d->seq;
    close(d->fd);
    delete d;
}

}  // extern "C"
