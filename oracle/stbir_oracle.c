/* TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
 *
 * CPU restatement of what ImageScaler::Scale computes in the reference's STB build
 * (src/image-scaler.cc:75-97 -> third_party/stb/stb_image_resize2.h), for RGBA8 /
 * BGRA8 input, RGBA8 output, edge clamp, default filters (Mitchell when shrinking, BOX =
 * "trapezoid" when enlarging because of src/image-scaler.cc:32, point sample at scale 1),
 * non-premultiplied alpha with STB's "fancy" alpha weighting (7 float channels).
 * Line numbers below are into third_party/stb/stb_image_resize2.h.
 *
 * The arithmetic is restated operation by operation (float unless STB uses double) so
 * that the result is BIT-IDENTICAL to the reference; tests/test_scale_oracle.py pins it
 * against oracle/_ref/libtimg_ref.so (the real STB code) over many geometries.
 * Build with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SMALL_FLOAT ((float)1 / (1 << 20) / (1 << 20) / (1 << 20) / (1 << 20) / (1 << 20) / (1 << 20)) /* :1104 */

enum { F_POINT = 0, F_BOX = 1, F_MITCHELL = 2 };

typedef struct {
    int in_size, out_size;
    float scale, inv_scale;
    int rational; uint32_t num, den;
    int filter;
    int is_gather;              /* 1 upsample gather, 2 downsample gather, 0 vertical scatter */
    int fpw, margin, cw;        /* filter_pixel_width, filter_pixel_margin, coefficient_width */
    int *n0, *n1;               /* per output pixel, input range after edge folding */
    float *c;                   /* [out_size][cw] */
    int widest;
    int *lead;                  /* horizontal only: leading zero taps added by the pack step (:3803-3858) */
} axis_t;

/* ---- filter kernels :2845-2937 ---- */
static float k_trapezoid(float x, float scale) {
    float halfscale = scale / 2;
    float t = 0.5f + halfscale;
    if (x < 0.0f) x = -x;
    if (x >= t) return 0.0f;
    float r = 0.5f - halfscale;
    if (x <= r) return 1.0f;
    return (t - x) / scale;
}
static float k_mitchell(float x) {
    if (x < 0.0f) x = -x;
    if (x < 1.0f) return (16.0f + x * x * (21.0f * x - 36.0f)) / 18.0f;
    else if (x < 2.0f) return (32.0f + x * (-60.0f + x * (36.0f - 7.0f * x))) / 18.0f;
    return 0.0f;
}
static float kernel(int f, float x, float s) {
    return f == F_POINT ? 1.0f : f == F_BOX ? k_trapezoid(x, s) : k_mitchell(x);
}
static float support(int f, float s) {
    return f == F_POINT ? 0.5f : f == F_BOX ? 0.5f + s / 2.0f : 2.0f;
}

/* ---- :7473-7549 continued-fraction rational estimate ---- */
static int to_rational(double f, uint32_t limit, uint32_t *numer, uint32_t *denom, int limit_denom) {
    double err;
    uint64_t top, bot, numer_last = 0, denom_last = 1, numer_est = 1, denom_est = 0;
    top = (uint64_t)(f * (double)(1 << 25));
    bot = 1 << 25;
    for (;;) {
        uint64_t est, temp;
        if ((limit_denom ? denom_est : numer_est) >= limit) break;
        if (denom_est) {
            err = ((double)numer_est / (double)denom_est) - f;
            if (err < 0.0) err = -err;
            if (err < (1.0 / (double)(1 << 24))) { *numer = (uint32_t)numer_est; *denom = (uint32_t)denom_est; return 1; }
        }
        if (bot == 0) break;
        est = top / bot; temp = top % bot; top = bot; bot = temp;
        temp = est * denom_est + denom_last; denom_last = denom_est; denom_est = temp;
        temp = est * numer_est + numer_last; numer_last = numer_est; numer_est = temp;
    }
    if (limit_denom) { numer_est = (uint64_t)(f * (double)limit + 0.5); denom_est = limit; }
    else { numer_est = limit; denom_est = (uint64_t)(((double)limit / f) + 0.5); }
    *numer = (uint32_t)numer_est; *denom = (uint32_t)denom_est;
    err = denom_est ? (((double)(uint32_t)numer_est / (double)(uint32_t)denom_est) - f) : 1.0;
    if (err < 0.0) err = -err;
    return (err < (1.0 / (double)(1 << 24))) ? 1 : 0;
}

static void in_pixel_range(int *first, int *last, float out_center, float out_radius, float inv_scale) { /* :3242 */
    float lo = out_center - out_radius, hi = out_center + out_radius;
    float ilo = (lo + 0.0f) * inv_scale, ihi = (hi + 0.0f) * inv_scale;
    int f = (int)floorf(ilo + 0.5f), l = (int)floorf(ihi - 0.5f);
    if (l < f) l = f;
    *first = f; *last = l;
}
static void out_pixel_range(int *first, int *last, float in_center, float in_radius, float scale, int out_size) { /* :3364 */
    float lo = in_center - in_radius, hi = in_center + in_radius;
    float olo = lo * scale - 0.0f, ohi = hi * scale - 0.0f;
    int f = (int)floorf(olo + 0.5f), l = (int)floorf(ohi - 0.5f);
    if (f < 0) f = 0;
    if (l >= out_size) l = out_size - 1;
    *first = f; *last = l;
}

static void axis_free(axis_t *A) { free(A->n0); free(A->n1); free(A->c); free(A->lead); memset(A, 0, sizeof *A); }

static int build_axis(axis_t *A, int in_size, int out_size, int always_gather) {
    memset(A, 0, sizeof *A);
    A->in_size = in_size; A->out_size = out_size;
    const double scale_d = ((double)out_size / (double)in_size) * (((double)out_size / (double)out_size) / 1.0); /* :7573-7581 */
    A->scale = (float)scale_d; A->inv_scale = (float)(1.0 / scale_d);
    A->rational = to_rational(scale_d, (scale_d <= 1.0) ? (uint32_t)out_size : (uint32_t)in_size, &A->num, &A->den,
                              scale_d >= 1.0);                                           /* :7596 */
    const float scale = A->scale, inv_scale = A->inv_scale;
    /* filter choice :6499-6509 (pixel_shift == 0 here) */
    A->filter = F_MITCHELL;
    if (scale >= (1.0f - SMALL_FLOAT)) A->filter = (scale <= (1.0f + SMALL_FLOAT)) ? F_POINT : F_BOX;
    /* :2962-2970 */
    if (scale >= (1.0f - SMALL_FLOAT)) A->fpw = (int)ceilf(support(A->filter, 1.0f / scale) * 2.0f);
    else A->fpw = (int)ceilf(support(A->filter, scale) * 2.0f / scale);
    A->is_gather = 0;                                                                   /* :6530-6534 */
    if (scale >= (1.0f - SMALL_FLOAT)) A->is_gather = 1;
    else if (always_gather || A->fpw <= 32) A->is_gather = 2;
    A->cw = (A->is_gather == 1) ? (int)ceilf(support(A->filter, 1.0f / scale) * 2.0f)
                                : (int)ceilf(support(A->filter, scale) * 2.0f / scale);  /* :2974-2990; prescatter width == fpw */
    A->margin = A->fpw / 2;

    const int n_out = out_size, cw = A->cw;
    A->n0 = (int *)calloc((size_t)n_out, sizeof(int));
    A->n1 = (int *)calloc((size_t)n_out, sizeof(int));
    A->lead = (int *)calloc((size_t)n_out, sizeof(int));
    A->c = (float *)calloc((size_t)n_out * cw + 16, sizeof(float));
    if (!A->n0 || !A->n1 || !A->c || !A->lead) return -1;

    const int polyphase = A->rational && ((int)A->num < n_out);
    const int numerator = (int)A->num, denominator = (int)A->den;

    if (A->is_gather == 1) {                      /* :3267-3327 gather upsample */
        const float out_radius = support(A->filter, inv_scale) * scale;
        const int end = polyphase ? numerator : n_out;
        for (int n = 0; n < end; n++) {
            float *cg = A->c + (size_t)n * cw;
            float out_center = (float)n + 0.5f;
            float in_center_of_out = (out_center + 0.0f) * inv_scale;
            int first, last;
            in_pixel_range(&first, &last, out_center, out_radius, inv_scale);
            if ((last - first + 1) > cw) last = first + cw - 1;
            int last_non_zero = -1;
            for (int i = 0; i <= last - first; i++) {
                float in_center = (float)(i + first) + 0.5f;
                float coeff = kernel(A->filter, in_center_of_out - in_center, inv_scale);
                if ((coeff < SMALL_FLOAT) && (coeff > -SMALL_FLOAT)) {
                    if (i == 0) { ++first; i--; continue; }
                    coeff = 0;
                } else last_non_zero = i;
                cg[i] = coeff;
            }
            last = last_non_zero + first;
            A->n0[n] = first; A->n1[n] = last;
        }
    } else {                                      /* :3382-3458 gather downsample */
        const float in_radius = support(A->filter, scale) * inv_scale;
        int first_out_inited = -1;
        for (int in_pixel = -A->margin; in_pixel < in_size + A->margin; in_pixel++) {
            float in_center = (float)in_pixel + 0.5f;
            float out_center_of_in = in_center * scale - 0.0f;
            int ofirst, olast;
            out_pixel_range(&ofirst, &olast, in_center, in_radius, scale, n_out);
            if (ofirst > olast) continue;
            if (polyphase) {
                if (ofirst == numerator) break;
                if (olast >= numerator) olast = numerator - 1;
            }
            for (int i = 0; i <= olast - ofirst; i++) {
                float out_center = (float)(i + ofirst) + 0.5f;
                float x = out_center - out_center_of_in;
                float coeff = kernel(A->filter, x, scale) * scale;
                if ((coeff < SMALL_FLOAT) && (coeff > -SMALL_FLOAT)) coeff = 0.0f;
                int o = i + ofirst;
                float *cg = A->c + (size_t)o * cw;
                if (o > first_out_inited) {
                    first_out_inited = o;
                    A->n0[o] = in_pixel; A->n1[o] = in_pixel; cg[0] = coeff;
                } else {
                    if (cg[0] == 0.0f) A->n0[o] = in_pixel;   /* zap a leading zero (:3444-3448) */
                    A->n1[o] = in_pixel;
                    if (in_pixel - A->n0[o] >= cw) return -2;
                    cg[in_pixel - A->n0[o]] = coeff;
                }
            }
        }
    }

    /* ---- :3466-3635 cleanup: renormalise in double, polyphase copy, fold clamped edges ---- */
    {
        const int end = polyphase ? numerator : n_out;
        for (int n = 0; n < end; n++) {
            float *cg = A->c + (size_t)n * cw;
            double total = 0;
            const int e = A->n1[n] - A->n0[n];
            for (int i = 0; i <= e; i++) total += (double)cg[i];
            if ((total < SMALL_FLOAT) && (total > -SMALL_FLOAT)) { A->n1[n] = A->n0[n]; cg[0] = 0.0f; }
            else if ((total < (1.0f - SMALL_FLOAT)) || (total > (1.0f + SMALL_FLOAT))) {
                const double fs = ((double)1.0) / total;
                for (int i = 0; i <= e; i++) cg[i] = (float)(cg[i] * fs);
            }
        }
        if (polyphase) {
            for (int n = numerator; n < n_out; n++) {
                A->n0[n] = A->n0[n - numerator] + denominator;
                A->n1[n] = A->n1[n - numerator] + denominator;
            }
            memmove(A->c + (size_t)numerator * cw, A->c, (size_t)(n_out - numerator) * cw * sizeof(float));
            /* stbir_overlapping_memcpy copies forward: phase n takes phase n-numerator */
            for (int n = numerator; n < n_out; n++)
                memcpy(A->c + (size_t)n * cw, A->c + (size_t)(n - numerator) * cw, (size_t)cw * sizeof(float));
        }
        const int last_in = in_size - 1;
        int widest = -1;
        for (int n = 0; n < n_out; n++) {
            float *cg = A->c + (size_t)n * cw;
            if (A->n1[n] > last_in) {                         /* right edge first (:3563-3571) */
                const int start = A->n0[n], endi = A->n1[n];
                A->n1[n] = last_in;
                for (int i = in_size; i <= endi; i++) {
                    if (last_in < A->n0[n]) return -3;
                    cg[last_in - A->n0[n]] += cg[i - start];
                }
            }
            if (A->n0[n] < 0) {                               /* left edge (:3574-3593) */
                const int old_n0 = A->n0[n];
                if (A->n1[n] < 0) return -4;
                for (int i = -1; i > old_n0; i--) cg[0 - old_n0] += cg[i - old_n0];
                const float save = cg[0];
                for (int i = 0; i <= A->n1[n]; i++) cg[i] = cg[i - old_n0];
                A->n0[n] = 0;
                cg[0] += save;
            }
            if (A->n0[n] <= A->n1[n]) {
                int diff = A->n1[n] - A->n0[n] + 1;
                while (diff && (cg[diff - 1] == 0.0f)) --diff;
                A->n1[n] = A->n0[n] + diff - 1;
                if (A->n0[n] <= A->n1[n] && diff > widest) widest = diff;
                for (int i = diff; i < cw; i++) cg[i] = 0.0f;
            }
        }
        A->widest = widest;
    }
    return 0;
}

/* The pack step (:3803-3858) moves contributors that would read past the decoded row
 * back into it and pads with LEADING zero taps; that only matters because horizontal
 * taps alternate between two accumulators, so the tap parity shifts.  row_end is
 * conservative.n1 + 1 == in_size for edge clamp (:6659-6664). */
static void horizontal_pack_parity(axis_t *A) {
    const int widest = A->widest, row_end = A->in_size;
    for (int n = A->out_size - 1; n >= 0 && (A->n0[n] + widest * 2) >= row_end; n--) {
        if ((A->n0[n] + widest) > row_end) {
            int stop = widest;
            if (widest > 12) {
                const int mod = widest & 3;
                stop = (((A->n1[n] - A->n0[n] + 1) - mod + 3) & ~3) + mod;
                if (stop < (8 + mod)) stop = 8 + mod;
            }
            if ((A->n0[n] + stop) > row_end) A->lead[n] = A->n0[n] - (row_end - stop);
        }
    }
}

/* vertical-first cost heuristic: weights trained by the STB author for 4 and 7 float
 * channels (:6770-6822, rows = classification, cols = the four weights) */
static const float kWeights4[8][4] = {
    {0.00000f, 0.50000f, 0.00000f, 0.71875f}, {0.06250f, 0.84375f, 0.00000f, 0.87500f},
    {1.00000f, 0.50000f, 0.50000f, 0.96875f}, {1.00000f, 0.09375f, 0.31250f, 0.50000f},
    {1.00000f, 1.00000f, 1.00000f, 1.00000f}, {1.00000f, 0.03125f, 0.03125f, 0.53125f},
    {0.18750f, 0.12500f, 0.00000f, 1.00000f}, {0.00000f, 1.00000f, 0.03125f, 0.18750f}};
static const float kWeights7[8][4] = {
    {0.00000f, 0.59375f, 0.00000f, 0.96875f}, {0.06250f, 0.81250f, 0.06250f, 0.59375f},
    {0.75000f, 0.43750f, 0.12500f, 0.96875f}, {0.87500f, 0.06250f, 0.18750f, 0.43750f},
    {1.00000f, 1.00000f, 1.00000f, 1.00000f}, {0.15625f, 0.12500f, 1.00000f, 1.00000f},
    {0.06250f, 0.12500f, 0.00000f, 1.00000f}, {0.00000f, 1.00000f, 0.03125f, 0.34375f}};

static int vertical_first(const float (*wt)[4], const axis_t *H, const axis_t *V) {   /* :6859-6905 */
    int cls;
    if ((V->out_size <= 4) || (H->out_size <= 4)) cls = (V->out_size < H->out_size) ? 6 : 7;
    else if (V->scale <= 1.0f) cls = V->is_gather ? 1 : 0;
    else if (V->scale <= 2.0f) cls = 2;
    else if (V->scale <= 3.0f) cls = 3;
    else if (V->scale <= 4.0f) cls = 5;
    else cls = 6;
    const float *w = wt[cls];
    double h_cost = (float)H->fpw * w[0] + H->scale * (float)V->fpw * w[1];
    double v_cost = (float)V->fpw * w[2] + V->scale * (float)H->fpw * w[3];
    return (v_cost <= h_cost) ? 1 : 0;
}

/* horizontal gather of `ch` interleaved float channels, :5801-6008 / :10290-10470:
 * <=3 taps sequential, otherwise even taps -> x, odd taps -> y, result x + y. */
static void h_gather(const axis_t *H, const float *src, float *dst, int ch) {
    for (int o = 0; o < H->out_size; o++) {
        const float *cg = H->c + (size_t)o * H->cw;
        const int n0 = H->n0[o], cnt = H->n1[o] - n0 + 1, lead = H->lead[o];
        for (int k = 0; k < ch; k++) {
            if (H->widest <= 3) {
                float t = src[(size_t)n0 * ch + k] * cg[0];
                for (int i = 1; i < cnt; i++) t += src[(size_t)(n0 + i) * ch + k] * cg[i];
                dst[(size_t)o * ch + k] = t;
            } else {
                float acc[2] = {0.0f, 0.0f};
                int started[2] = {0, 0};
                for (int i = 0; i < cnt; i++) {
                    const int p = (i + lead) & 1;
                    const float v = src[(size_t)(n0 + i) * ch + k] * cg[i];
                    if (!started[p]) { acc[p] = v; started[p] = 1; } else acc[p] += v;
                }
                dst[(size_t)o * ch + k] = acc[0] + acc[1];
            }
        }
    }
}

/* public: returns 0, fills out (ow*oh*4).  fmt 0 = RGBA, 1 = BGRA in. */
int orc_stb_resize(const uint8_t *in, int iw, int ih, int fmt, uint8_t *out, int ow, int oh, int *info /*[4] or NULL*/) {
    axis_t H, V;
    if (build_axis(&H, iw, ow, 1)) return -1;
    if (build_axis(&V, ih, oh, 0)) return -1;
    const int both_point = (H.filter == F_POINT) && (V.filter == F_POINT);
    const int ch = both_point ? 4 : 7;                                  /* :6938-6952 */
    horizontal_pack_parity(&H);
    const int vfirst = vertical_first(both_point ? kWeights4 : kWeights7, &H, &V);
    if (info) { info[0] = vfirst; info[1] = H.widest; info[2] = V.widest; info[3] = ch; }
    const float inv255 = 1.0f / 255.0f;
    const int ro = fmt ? 2 : 0, bo = fmt ? 0 : 2;

    /* decode every input row: byte * (1/255) (:8300) then fancy alpha weight (:4160-4172) */
    float *D = (float *)malloc((size_t)iw * ih * ch * sizeof(float));
    if (!D) return -1;
    for (size_t p = 0; p < (size_t)iw * ih; p++) {
        const float r = (float)in[4 * p + ro] * inv255, g = (float)in[4 * p + 1] * inv255;
        const float b = (float)in[4 * p + bo] * inv255, a = (float)in[4 * p + 3] * inv255;
        float *d = D + p * ch;
        d[0] = r; d[1] = g; d[2] = b; d[3] = a;
        if (ch == 7) { d[4] = r * a; d[5] = g * a; d[6] = b * a; }
    }
    float *E = (float *)malloc((size_t)ow * oh * ch * sizeof(float));   /* resampled, pre-encode */
    if (!E) { free(D); return -1; }

    if (!vfirst) {
        /* horizontal first: every input row -> ow, then vertical gather (sequential sum in
         * input-row order, :10080-10170; scatter accumulates in the same order, :9864-10012) */
        float *T = (float *)malloc((size_t)ow * ih * ch * sizeof(float));
        if (!T) { free(D); free(E); return -1; }
        for (int y = 0; y < ih; y++) {
            if (H.filter == F_POINT && H.scale == 1.0f) memcpy(T + (size_t)y * ow * ch, D + (size_t)y * iw * ch, (size_t)ow * ch * sizeof(float));
            else h_gather(&H, D + (size_t)y * iw * ch, T + (size_t)y * ow * ch, ch);
        }
        for (int oy = 0; oy < oh; oy++) {
            const float *cg = V.c + (size_t)oy * V.cw;
            const int n0 = V.n0[oy], cnt = V.n1[oy] - n0 + 1;
            for (size_t i = 0; i < (size_t)ow * ch; i++) {
                float t = T[(size_t)n0 * ow * ch + i] * cg[0];
                for (int k = 1; k < cnt; k++) t += T[(size_t)(n0 + k) * ow * ch + i] * cg[k];
                E[(size_t)oy * ow * ch + i] = t;
            }
        }
        free(T);
    } else {
        float *T = (float *)malloc((size_t)iw * ch * sizeof(float));
        if (!T) { free(D); free(E); return -1; }
        for (int oy = 0; oy < oh; oy++) {
            const float *cg = V.c + (size_t)oy * V.cw;
            const int n0 = V.n0[oy], cnt = V.n1[oy] - n0 + 1;
            for (size_t i = 0; i < (size_t)iw * ch; i++) {
                float t = D[(size_t)n0 * iw * ch + i] * cg[0];
                for (int k = 1; k < cnt; k++) t += D[(size_t)(n0 + k) * iw * ch + i] * cg[k];
                T[i] = t;
            }
            if (H.filter == F_POINT && H.scale == 1.0f) memcpy(E + (size_t)oy * ow * ch, T, (size_t)ow * ch * sizeof(float));
            else h_gather(&H, T, E + (size_t)oy * ow * ch, ch);
        }
        free(T);
    }
    /* un-weight (:4247-4294) and encode v*255+0.5, clamp, truncate (:8329-8437) */
    for (size_t p = 0; p < (size_t)ow * oh; p++) {
        const float *e = E + p * ch;
        float r = e[0], g = e[1], b = e[2];
        const float a = e[3];
        if (ch == 7 && !(a < SMALL_FLOAT)) { const float ia = 1.0f / a; r = e[4] * ia; g = e[5] * ia; b = e[6] * ia; }
        const float v[4] = {r, g, b, a};
        for (int k = 0; k < 4; k++) {
            float f = v[k] * 255.0f + 0.5f;
            if (f < 0) f = 0; else if (f > 255) f = 255;
            out[4 * p + k] = (uint8_t)f;
        }
    }
    free(D); free(E); axis_free(&H); axis_free(&V);
    return 0;
}

/* Tables of one axis in the same shape the product exports (b200timg_resample_plan), for the
 * host-logic parity test.  axis 0 = horizontal, 1 = vertical. */
int orc_stb_plan(int iw, int ih, int ow, int oh, int axis, int *widest, int *flags, int32_t *first,
                 int32_t *count, int32_t *lead, float *coeff) {
    axis_t H, V;
    if (build_axis(&H, iw, ow, 1)) return -1;
    if (build_axis(&V, ih, oh, 0)) return -1;
    horizontal_pack_parity(&H);
    const int both_point = (H.filter == F_POINT) && (V.filter == F_POINT);
    const int vf = vertical_first(both_point ? kWeights4 : kWeights7, &H, &V);
    const axis_t *A = axis == 0 ? &H : &V;
    const int wd = A->widest < 1 ? 1 : A->widest;
    if (widest) *widest = wd;
    if (flags) *flags = (vf ? 1 : 0) | (both_point ? 2 : 0) | (H.widest <= 3 ? 4 : 0);
    for (int o = 0; o < A->out_size; o++) {
        int cnt = A->n1[o] - A->n0[o] + 1;
        if (cnt < 1) cnt = 1;
        if (first) first[o] = A->n0[o];
        if (count) count[o] = cnt;
        if (lead) lead[o] = axis == 0 ? A->lead[o] : 0;
        if (coeff) for (int i = 0; i < wd; i++) coeff[(size_t)o * wd + i] = i < cnt ? A->c[(size_t)o * A->cw + i] : 0.0f;
    }
    axis_free(&H); axis_free(&V);
    return 0;
}
