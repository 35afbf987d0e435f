/* TEST INFRASTRUCTURE ONLY (part of oracle/liboracle.so) -- the CPU baseline driver of bench.py.
 *
 * Runs the reference's per-frame CPU path on N native threads, one independent scaler + canvas per
 * thread over disjoint frames, the way timg's own loader pool works (src/timg.cc:913-968) -- no Python
 * in the timed region.  The stages are passed in as plain C function pointers so the same driver
 * runs on the reference's own translation units (oracle/_ref/libtimg_ref.so: ref_scale, ref_compose,
 * ref_blocks_*) or, where that library is absent, on the restatements in this directory:
 *   scale    ImageScaler::Scale                          src/image-scaler.cc:75-97
 *   compose  Framebuffer::AlphaComposeBackground         src/framebuffer.cc:108-150
 *   sixel    SixelCanvas::Send's pad + pad-strip compose src/sixel-canvas.cc:109-120, then the libsixel
 *            restatement orc_sixel_encode (libsixel itself is not vendored in the reference)
 *   blocks   UnicodeBlockCanvas::Send                    src/unicode-block-canvas.cc:323-403
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef int (*orc_scale_fn)(const uint8_t *in, int iw, int ih, int fmt, uint8_t *out, int ow, int oh);
typedef void (*orc_compose_fn)(uint8_t *fb, int w, int h, int has_bg, uint32_t bg, uint32_t pattern, int pw, int ph, int start_row);
typedef void *(*orc_blocks_new_fn)(int quarter, int upper, int color8, int capture);
typedef long (*orc_blocks_send_fn)(void *h, int x, int dy, const uint8_t *fb, int w, int hgt, int seq_type, char *out, long cap);
typedef void (*orc_blocks_free_fn)(void *h);

long orc_sixel_encode(const uint8_t *rgba, int w, int h, int mode, char *out, long cap, uint8_t *palette_out,
                      int *ncolors_out, int *origcolors_out, uint8_t *index_out);

typedef struct {
    const uint8_t *frames;
    int n_distinct, iw, ih, n_jobs, ow, oh, has_bg, mode, flags, animation, threads;
    uint32_t bg;
    orc_scale_fn sf;
    orc_compose_fn cf;
    orc_blocks_new_fn bnew;
    orc_blocks_send_fn bsend;
    orc_blocks_free_fn bfree;
    long *sizes;
    int next;                       /* sixel: shared job counter */
    pthread_mutex_t mu;
} plan_t;

typedef struct { plan_t *p; int tid; } arg_t;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void *sixel_worker(void *vp) {
    arg_t *a = (arg_t *)vp;
    plan_t *p = a->p;
    const int hp = (p->oh + 5) / 6 * 6;                    /* round_to_sixel, src/sixel-canvas.cc:91-94 */
    const size_t fbytes = (size_t)p->ow * hp * 4;
    uint8_t *fb = (uint8_t *)malloc(fbytes);
    const long cap = 1024 + (long)p->ow * hp * 5 + 256 * 24;   /* the reference's own bound, :123 (+ palette slack) */
    char *out = (char *)malloc((size_t)cap);
    for (;;) {
        pthread_mutex_lock(&p->mu);
        const int j = p->next++;
        pthread_mutex_unlock(&p->mu);
        if (j >= p->n_jobs) break;
        const uint8_t *src = p->frames + (size_t)(j % p->n_distinct) * p->iw * p->ih * 4;
        p->sf(src, p->iw, p->ih, 0, fb, p->ow, p->oh);
        p->cf(fb, p->ow, p->oh, p->has_bg, p->bg, 0, 0, 0, 0);             /* the source's compose, e.g. src/stb-image-source.cc:56-60 */
        memset(fb + (size_t)p->ow * p->oh * 4, 0, fbytes - (size_t)p->ow * p->oh * 4);
        p->cf(fb, p->ow, hp, p->has_bg, p->bg, 0, 0, 0, p->oh);            /* the canvas' pad strip, src/sixel-canvas.cc:115-118 */
        const long n = orc_sixel_encode(fb, p->ow, hp, p->mode, out, cap, NULL, NULL, NULL, NULL);
        if (p->sizes) p->sizes[j] = n;
    }
    free(out); free(fb);
    return NULL;
}

/* n_jobs frames (job j uses distinct frame j % n_distinct) on `threads` native threads; returns wall seconds */
double orc_cpu_sixel_jobs(const uint8_t *frames, int n_distinct, int iw, int ih, int n_jobs, int ow, int oh, int has_bg,
                          uint32_t bg, int mode, int threads, orc_scale_fn sf, orc_compose_fn cf, long *sizes) {
    plan_t p;
    memset(&p, 0, sizeof p);
    p.frames = frames; p.n_distinct = n_distinct; p.iw = iw; p.ih = ih; p.n_jobs = n_jobs; p.ow = ow; p.oh = oh;
    p.has_bg = has_bg; p.bg = bg; p.mode = mode; p.threads = threads; p.sf = sf; p.cf = cf; p.sizes = sizes;
    pthread_mutex_init(&p.mu, NULL);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    arg_t *args = (arg_t *)malloc(sizeof(arg_t) * (size_t)threads);
    const double t0 = now_s();
    for (int t = 0; t < threads; ++t) { args[t].p = &p; args[t].tid = t; pthread_create(&th[t], NULL, sixel_worker, &args[t]); }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    const double dt = now_s() - t0;
    free(args); free(th);
    pthread_mutex_destroy(&p.mu);
    return dt;
}

static void *blocks_worker(void *vp) {
    arg_t *a = (arg_t *)vp;
    plan_t *p = a->p;
    /* contiguous chunk per thread: an animation's delta frames chain inside it (the first frame of a chunk is full) */
    const int lo = (int)((long)p->n_jobs * a->tid / p->threads), hi = (int)((long)p->n_jobs * (a->tid + 1) / p->threads);
    if (lo >= hi) return NULL;
    uint8_t *fb = (uint8_t *)malloc((size_t)p->ow * p->oh * 4);
    void *canvas = p->bnew(p->flags & 1, (p->flags >> 1) & 1, (p->flags >> 2) & 1, 0 /* bytes to /dev/null */);
    const int rows = (p->oh + 1) / 2;                      /* cell rows of one frame: the cursor goes back up by this */
    for (int j = lo; j < hi; ++j) {
        const uint8_t *src = p->frames + (size_t)(j % p->n_distinct) * p->iw * p->ih * 4;
        p->sf(src, p->iw, p->ih, 0, fb, p->ow, p->oh);
        p->cf(fb, p->ow, p->oh, p->has_bg, p->bg, 0, 0, 0, 0);
        const int delta = p->animation && j > lo;
        p->bsend(canvas, 0, delta ? -rows : 0, fb, p->ow, p->oh, delta ? 2 : 1, NULL, 0);
    }
    p->bfree(canvas);
    free(fb);
    return NULL;
}

double orc_cpu_blocks_jobs(const uint8_t *frames, int n_distinct, int iw, int ih, int n_jobs, int ow, int oh, int has_bg,
                           uint32_t bg, int flags, int animation, int threads, orc_scale_fn sf, orc_compose_fn cf,
                           orc_blocks_new_fn bnew, orc_blocks_send_fn bsend, orc_blocks_free_fn bfree) {
    plan_t p;
    memset(&p, 0, sizeof p);
    p.frames = frames; p.n_distinct = n_distinct; p.iw = iw; p.ih = ih; p.n_jobs = n_jobs; p.ow = ow; p.oh = oh;
    p.has_bg = has_bg; p.bg = bg; p.flags = flags; p.animation = animation; p.threads = threads; p.sf = sf; p.cf = cf;
    p.bnew = bnew; p.bsend = bsend; p.bfree = bfree;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    arg_t *args = (arg_t *)malloc(sizeof(arg_t) * (size_t)threads);
    const double t0 = now_s();
    for (int t = 0; t < threads; ++t) { args[t].p = &p; args[t].tid = t; pthread_create(&th[t], NULL, blocks_worker, &args[t]); }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    const double dt = now_s() - t0;
    free(args); free(th);
    return dt;
}

/* the scaler restatement behind the 7-argument stage signature (used only where oracle/_ref is absent) */
int orc_stb_resize(const uint8_t *in, int iw, int ih, int fmt, uint8_t *out, int ow, int oh, int *info);
int orc_stb_resize7(const uint8_t *in, int iw, int ih, int fmt, uint8_t *out, int ow, int oh) {
    return orc_stb_resize(in, iw, ih, fmt, out, ow, oh, NULL);
}
