/* TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
 *
 * CPU restatement (plain C, strict IEEE: build with -ffp-contract=off, no fast-math)
 * of the in-tree arithmetic of hzeller/timg's block-mode hot path.  Every function
 * cites the reference file:line it follows (paths relative to /root/reference).
 * Pinned bit-exactly against the reference itself (oracle/_ref/libtimg_ref.so) by
 * tests/test_oracle_vs_reference.py and against tests/golden/ fixtures.
 *
 * Pixel layout: 4 bytes R,G,B,A per pixel, row-major, tightly packed
 * (src/framebuffer.h:26-61).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint8_t r, g, b, a; } px_t;

static inline uint32_t px_u32(px_t p) {
    uint32_t v; memcpy(&v, &p, 4); return v;
}

/* ---- src/framebuffer.h:37-52  rgba_t::As256TermColor ------------------- */
static int cube_index(int v) {          /* cut points are the midpoints of the xterm cube levels */
    if (v < 0x5f / 2) return 0;
    if (v < (0x5f + 0x87) / 2) return 1;
    if (v < (0x87 + 0xaf) / 2) return 2;
    if (v < (0xaf + 0xd7) / 2) return 3;
    if (v < (0xd7 + 0xff) / 2) return 4;
    return 5;
}
int orc_as256(uint32_t rgba) {
    px_t c; memcpy(&c, &rgba, 4);
    if (c.r == c.g && c.g == c.b) return (232 + (c.r * 23 / 255)) & 0xff;
    return 16 + 36 * cube_index(c.r) + 6 * cube_index(c.g) + cube_index(c.b);
}

/* ---- src/framebuffer.h:138-174  LinearColor ------------------------------ */
typedef struct { float r, g, b, a; } lin_t;

static inline lin_t lin_of(px_t c) {
    lin_t l;
    l.r = (float)(c.r * c.r); l.g = (float)(c.g * c.g); l.b = (float)(c.b * c.b);
    l.a = (float)c.a;
    return l;
}
static inline uint8_t ungamma(float v) {                /* framebuffer.h:169-172 */
    const float s = sqrtf(v);
    return (s > 255) ? 255 : (uint8_t)s;
}
static inline px_t lin_repack(lin_t l) {                /* framebuffer.h:150-152 */
    px_t p = { ungamma(l.r), ungamma(l.g), ungamma(l.b), (uint8_t)l.a };
    return p;
}
static inline float lin_dist(lin_t a, lin_t b) {        /* framebuffer.h:145-148 */
    const float dr = b.r - a.r, dg = b.g - a.g, db = b.b - a.b;
    return dr * dr + dg * dg + db * db;
}

/* ---- src/framebuffer.cc:108-150  AlphaComposeBackground ------------------- */
static inline px_t blend_onto(px_t p, lin_t bg) {       /* framebuffer.h:155-161 */
    lin_t c = lin_of(p);
    const float a = c.a, ia = 0xff - c.a;
    c.r = (c.r * a + bg.r * ia) / 0xff;
    c.g = (c.g * a + bg.g * ia) / 0xff;
    c.b = (c.b * a + bg.b * ia) / 0xff;
    c.a = 0xff;
    return lin_repack(c);
}

void orc_compose_bg(uint8_t *fb, int w, int h, int has_bg, uint32_t bg_u32,
                    uint32_t pat_u32, int pw, int ph, int start_row) {
    if (!has_bg) return;                                 /* :111 "-b none" */
    px_t *px = (px_t *)fb;
    const long n = (long)w * h;
    long pos = (long)start_row * w;
    for (; pos < n; ++pos) if (px[pos].a < 0xff) break;  /* :113-116 */
    if (pos >= n) return;                                /* :117 */
    px_t bg, pat; memcpy(&bg, &bg_u32, 4); memcpy(&pat, &pat_u32, 4);
    if (bg.a == 0x00) return;                            /* :121 */
    if (pat.a == 0x00 || pat_u32 == bg_u32 || pw <= 0 || ph <= 0) {  /* :124-132 */
        const lin_t lbg = lin_of(bg);
        for (; pos < n; ++pos) {
            if (px[pos].a == 0xff) continue;
            px[pos] = blend_onto(px[pos], lbg);
        }
        return;
    }
    const lin_t choice[2] = { lin_of(bg), lin_of(pat) }; /* :135 */
    const int sx = (int)(pos % w), sy = (int)(pos / w);
    for (int y = sy; y < h; ++y) {
        const int yp = y / ph;
        for (int x = (y == sy ? sx : 0); x < w; ++x, ++pos) {
            if (px[pos].a == 0xff) continue;
            px[pos] = blend_onto(px[pos], choice[((x / pw) + yp) % 2]);
        }
    }
}

/* ---- src/image-source.cc:47-153  CalcScaleToFitDisplay -------------------- */
int orc_calc_fit(int img_w, int img_h, int width, int height, int cell_x_px,
                 int cell_y_px, float width_stretch_in, int upscale,
                 int upscale_integer, int fill_width, int fill_height,
                 int fit_in_rotated, int *tw, int *th) {
    float ws = width_stretch_in;
    if (fit_in_rotated) {                                /* :52-56 */
        int t = width; width = height; height = t;
        t = fill_width; fill_width = fill_height; fill_height = t;
        ws = 1.0f / width_stretch_in;
    }
    const float kMax = 5.0;                              /* :59-63 */
    if (ws > kMax) ws = kMax;
    if (ws < 1 / kMax) ws = 1 / kMax;
    if (ws > 1.0f) width = (int)(width / ws);            /* :65-70: int op= float */
    else height = (int)(height * ws);
    const float wf = (float)width / img_w;
    const float hf = (float)height / img_h;
    /* the reference compares float with the double literal 1.0 -> promote */
    if (!upscale && (fill_height || (double)wf > 1.0) &&
        (fill_width || (double)hf > 1.0)) {              /* :75-86 */
        *tw = img_w; *th = img_h;
        if (cell_x_px == 2) { *tw *= 2; return 1; }
        return 0;
    }
    *tw = width; *th = height;
    if (fill_width && fill_height) {
        const float f = (wf > hf) ? wf : hf;
        *tw = (int)roundf(f * img_w); *th = (int)roundf(f * img_h);
    } else if (fill_height) {
        *tw = (int)roundf(hf * img_w);
    } else if (fill_width) {
        *th = (int)roundf(wf * img_h);
    } else {
        const float f = (wf < hf) ? wf : hf;
        *tw = (int)roundf(f * img_w); *th = (int)roundf(f * img_h);
    }
    if (ws > 1.0f) *tw = (int)(*tw * ws);                /* :120-125 */
    else *th = (int)(*th / ws);
    if (cell_x_px > 0 && cell_x_px <= 2 && cell_y_px > 0 && cell_y_px <= 2) {
        *tw = *tw / cell_x_px * cell_x_px;               /* :129-133 */
        *th = *th / cell_y_px * cell_y_px;
    }
    if (*tw <= 0) *tw = 1;
    if (*th <= 0) *th = 1;
    if (upscale_integer && *tw > img_w && *th > img_h) { /* :139-150 */
        const float ac = cell_x_px == 2 ? 2 : 1;
        const float f1 = 1.0f * *tw / ac / img_w;
        const float f2 = 1.0f * *th / img_h;
        const float sm = f1 < f2 ? f1 : f2;
        if (sm > 1.0f) {
            /* floor() is the double overload in the reference: float->double */
            *tw = (int)((double)ac * floor((double)sm) * img_w);
            *th = (int)(floor((double)sm) * img_h);
        }
    }
    return (*tw != img_w || *th != img_h) ? 1 : 0;
}

/* ---- src/unicode-block-canvas.cc ----------------------------------------- */
enum { kBackground, kTopLeft, kTopRight, kBotLeft, kBotRight, kLeftBar,
       kTopLeftBotRight, kLowerBlock, kUpperBlock };
/* UTF-8 of U+2598,259D,2596,2597,258C,259A,2584,2580 (:78-88); [0] is ' ' */
static const unsigned char kGlyphLast[9] = { 0, 0x98, 0x9D, 0x96, 0x97, 0x8C, 0x9A, 0x84, 0x80 };

typedef struct { px_t fg, bg; int block; } pick_t;

static inline int transparent(px_t c) { return c.a < 0x60; }   /* :154 */

/* avd(): mean in list order, then sum of distances in list order
 * (src/framebuffer.h:177-194).  *res must be zero-initialised by caller. */
static float avd(lin_t *res, const lin_t *v, int n) {
    for (int i = 0; i < n; ++i) {
        res->r += v[i].r; res->g += v[i].g; res->b += v[i].b; res->a += v[i].a;
    }
    res->r /= n; res->g /= n; res->b /= n; res->a /= n;
    float sum = 0;
    for (int i = 0; i < n; ++i) sum += lin_dist(*res, v[i]);
    return sum;
}

static pick_t best_glyph_half(px_t top, px_t bot, int upper) {   /* :164-172 */
    pick_t p;
    if (px_u32(top) == px_u32(bot) || (transparent(top) && transparent(bot))) {
        p.fg = top; p.bg = bot; p.block = kBackground; return p;
    }
    if (upper) { p.fg = top; p.bg = bot; p.block = kUpperBlock; return p; }
    p.fg = bot; p.bg = top; p.block = kLowerBlock; return p;
}

static pick_t best_glyph_quarter(const px_t *top, const px_t *bot, int upper) { /* :174-227 */
    const lin_t tl = lin_of(top[0]), tr = lin_of(top[1]);
    const lin_t bl = lin_of(bot[0]), br = lin_of(bot[1]);
    const lin_t zero = {0, 0, 0, 0};
    pick_t p;
    const int tt = transparent(top[0]) && transparent(top[1]);
    const int bt = transparent(bot[0]) && transparent(bot[1]);
    if (tt && bt) { p.fg = bot[0]; p.bg = top[0]; p.block = kBackground; return p; }
    if (tt) {
        lin_t m = zero; lin_t v[2] = { bl, br }; avd(&m, v, 2);
        p.fg = lin_repack(m); p.bg = top[0]; p.block = kLowerBlock; return p;
    }
    if (bt) {
        lin_t m = zero; lin_t v[2] = { tl, tr }; avd(&m, v, 2);
        p.fg = lin_repack(m); p.bg = bot[0]; p.block = kUpperBlock; return p;
    }
    lin_t best_fg = zero, best_bg = zero; int best_block = kBackground;
    float best_d = 1e12f;
    for (int b = 0; b < 8; ++b) {
        float d; lin_t fg = zero, bg = zero;
        const int block = b < 7 ? b : (upper ? kUpperBlock : kLowerBlock);
        switch (block) {
        case kBackground: { lin_t v[4] = {tl, tr, bl, br}; d = avd(&bg, v, 4); fg = bg; } break;
        case kTopLeft:    { lin_t v[3] = {tr, bl, br};     d = avd(&bg, v, 3); fg = tl; } break;
        case kTopRight:   { lin_t v[3] = {tl, bl, br};     d = avd(&bg, v, 3); fg = tr; } break;
        case kBotLeft:    { lin_t v[3] = {tl, tr, br};     d = avd(&bg, v, 3); fg = bl; } break;
        case kBotRight:   { lin_t v[3] = {tl, tr, bl};     d = avd(&bg, v, 3); fg = br; } break;
        case kLeftBar:    { lin_t v[2] = {tr, br}, u[2] = {tl, bl};
                            d = avd(&bg, v, 2); d = d + avd(&fg, u, 2); } break;
        case kTopLeftBotRight: { lin_t v[2] = {tr, bl}, u[2] = {tl, br};
                            d = avd(&bg, v, 2); d = d + avd(&fg, u, 2); } break;
        case kLowerBlock: { lin_t v[2] = {tl, tr}, u[2] = {bl, br};
                            d = avd(&bg, v, 2); d = d + avd(&fg, u, 2); } break;
        default:          { lin_t v[2] = {bl, br}, u[2] = {tl, tr};
                            d = avd(&bg, v, 2); d = d + avd(&fg, u, 2); } break;
        }
        if (d < best_d) {
            best_fg = fg; best_bg = bg; best_block = block;
            if (d < 1) break;
            best_d = d;
        }
    }
    p.fg = lin_repack(best_fg); p.bg = lin_repack(best_bg); p.block = best_block;
    return p;
}

static char *put_u8_semi(char *o, unsigned v) {          /* :474-491 (no LUT needed) */
    if (v >= 100) { *o++ = (char)('0' + v / 100); v %= 100; *o++ = (char)('0' + v / 10); *o++ = (char)('0' + v % 10); }
    else if (v >= 10) { *o++ = (char)('0' + v / 10); *o++ = (char)('0' + v % 10); }
    else *o++ = (char)('0' + v);
    *o++ = ';';
    return o;
}
static char *put_color(char *o, px_t c, int color8) {    /* :113-122 */
    if (color8) return put_u8_semi(o, (unsigned)orc_as256(px_u32(c)));
    o = put_u8_semi(o, c.r); o = put_u8_semi(o, c.g); return put_u8_semi(o, c.b);
}
static char *put_str(char *o, const char *s) { size_t n = strlen(s); memcpy(o, s, n); return o + n; }

/* Canvas state (src/unicode-block-canvas.h:70-79). */
typedef struct {
    int quarter, upper, color8;
    px_t *backing; size_t backing_px;
    int last_h, last_x;
} orc_canvas;

void *orc_blocks_new(int quarter, int upper, int color8) {
    orc_canvas *c = (orc_canvas *)calloc(1, sizeof *c);
    c->quarter = quarter; c->upper = upper; c->color8 = color8;
    return c;
}
void orc_blocks_free(void *h) { orc_canvas *c = (orc_canvas *)h; free(c->backing); free(c); }

/* Worst-case size like RequestBuffers (:405-424) plus room for a prefix. */
long orc_blocks_bound(int w, int h) { return 64 + (long)((h + 1) / 2) * (16 + (long)w * 39 + 5); }

/* One Send() (:323-403) without the TerminalCanvas prefix machinery: the caller
 * passes the already-built prefix bytes (cursor-up etc).  x is the indent in
 * PIXELS as the reference passes it; dy as in the reference.
 * Returns the number of bytes written to out (0 == "nothing changed"). */
long orc_blocks_send(void *h, int x, int dy, const uint8_t *fbp, int width, int height,
                     const char *prefix, int prefix_len, char *out) {
    orc_canvas *cv = (orc_canvas *)h;
    const int N = cv->quarter ? 2 : 1;
    const px_t *fb = (const px_t *)fbp;
    char *pos = out;
    /* dy<0 cursor-up is part of prefix in the reference (MoveCursorDY, :329); caller builds it. */
    if (prefix_len > 0) { memcpy(pos, prefix, (size_t)prefix_len); pos += prefix_len; }
    if (cv->quarter) x /= 2;                              /* :334 */
    const char *before = pos;

    const size_t need = (size_t)(width + 1) * (height + 1);  /* :430-434 */
    if (need > cv->backing_px) {
        cv->backing = (px_t *)realloc(cv->backing, need * sizeof(px_t));
        cv->backing_px = need;
    }
    px_t *prev = cv->backing;
    const int emit_diff = (x == cv->last_x) && (cv->last_h > 0) && (abs(dy) == cv->last_h); /* :344-346 */

    px_t *empty = (px_t *)calloc((size_t)width + 1, sizeof(px_t));   /* :436-440 */
    const int row_offset = ((height % 2 != 0) && !cv->upper) ? -1 : 0;  /* :356-358 */
    int y_skip = 0;
    for (int y = 0; y < height; y += 2) {
        const int row = y + row_offset;
        const px_t *t = row < 0 ? empty : fb + (size_t)width * row;
        const px_t *b = (row + 1) >= height ? empty : fb + (size_t)width * (row + 1);
        /* ---- AppendDoubleRow (:230-321) ---- */
        pick_t last; memset(&last, 0, sizeof last);
        px_t last_fg; memset(&last_fg, 0, sizeof last_fg);
        int fg_unknown = 1, bg_unknown = 1, x_skip = x;
        const char *start = pos;
        for (int cx = 0; cx < width; cx += N, prev += 2 * N, t += N, b += N) {
            if (emit_diff) {
                int same = (N == 1)
                    ? (px_u32(t[0]) == px_u32(prev[0]) && px_u32(b[0]) == px_u32(prev[1]))
                    : (px_u32(t[0]) == px_u32(prev[0]) && px_u32(t[1]) == px_u32(prev[1]) &&
                       px_u32(b[0]) == px_u32(prev[2]) && px_u32(b[1]) == px_u32(prev[3]));
                if (same) { ++x_skip; continue; }
            }
            if (y_skip) {                                 /* :249-258 */
                if (y_skip <= 4) { memset(pos, '\n', (size_t)y_skip); pos += y_skip; }
                else pos += sprintf(pos, "\033[%dB", y_skip);
                y_skip = 0;
            }
            if (x_skip > 0) { pos += sprintf(pos, "\033[%dC", x_skip); x_skip = 0; }
            const pick_t pick = (N == 1) ? best_glyph_half(t[0], b[0], cv->upper)
                                         : best_glyph_quarter(t, b, cv->upper);
            int emitted = 0;
            if (pick.block != kBackground && (fg_unknown || px_u32(pick.fg) != px_u32(last_fg))) {
                pos = put_str(pos, "\033[");
                pos = put_str(pos, cv->color8 ? "38;5;" : "38;2;");
                pos = put_color(pos, pick.fg, cv->color8);
                emitted = 1; last_fg = pick.fg; fg_unknown = 0;
            }
            if (bg_unknown || px_u32(pick.bg) != px_u32(last.bg)) {
                if (!emitted) pos = put_str(pos, "\033[");
                if (transparent(pick.bg)) pos = put_str(pos, "49;");
                else {
                    pos = put_str(pos, cv->color8 ? "48;5;" : "48;2;");
                    pos = put_color(pos, pick.bg, cv->color8);
                }
                emitted = 1; bg_unknown = 0;
            }
            if (emitted) pos[-1] = 'm';
            if (pick.block == kBackground) *pos++ = ' ';
            else { *pos++ = (char)0xE2; *pos++ = (char)0x96; *pos++ = (char)kGlyphLast[pick.block]; }
            last = pick;
            if (N == 1) { prev[0] = t[0]; prev[1] = b[0]; }
            else { prev[0] = t[0]; prev[1] = t[1]; prev[2] = b[0]; prev[3] = b[1]; }
        }
        if (pos == start) y_skip++;
        else pos = put_str(pos, "\033[0m\n");
    }
    free(empty);
    cv->last_h = height; cv->last_x = x;
    if (before == pos) return 0;                          /* :390-395: size stays 0 */
    if (y_skip) pos += sprintf(pos, "\033[%dB", y_skip);  /* :397-399 */
    return (long)(pos - out);
}
