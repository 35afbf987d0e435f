// See resample_tables.h.  Pure host code (compiled by nvcc only so the library has one
// toolchain); no device functions here.  Floating-point contraction must stay off for this
// file: every expression below rounds exactly like the reference's scalar C.
#include "resample_tables.h"

#include <cmath>
#include <cstring>

namespace b200timg {
namespace {

// 2^-120, the library's "effectively zero" threshold (:1104)
const float kTiny = std::ldexp(1.0f, -120);

struct Filter {
    AxisFilter kind;
    // kernel value at distance x; s is the scale argument the reference passes (:2845-2937)
    float eval(float x, float s) const {
        if (x < 0.0f) x = -x;
        switch (kind) {
        case AxisFilter::kPoint: return 1.0f;
        case AxisFilter::kBox: {                          // "trapezoid": box that handles partial coverage
            const float half = s / 2;
            const float t = 0.5f + half;
            if (x >= t) return 0.0f;
            const float r = 0.5f - half;
            if (x <= r) return 1.0f;
            return (t - x) / s;
        }
        case AxisFilter::kMitchell:
            if (x < 1.0f) return (16.0f + x * x * (21.0f * x - 36.0f)) / 18.0f;
            if (x < 2.0f) return (32.0f + x * (-60.0f + x * (36.0f - 7.0f * x))) / 18.0f;
            return 0.0f;
        }
        return 0.0f;
    }
    float support(float s) const {
        switch (kind) {
        case AxisFilter::kPoint: return 0.5f;
        case AxisFilter::kBox: return 0.5f + s / 2.0f;
        default: return 2.0f;
        }
    }
};

// Best rational approximation by continued fractions, accepted when within one float ulp
// (:7473-7549).  Used to detect polyphase-periodic coefficient sets.
bool rational_within_float(double f, uint32_t limit, bool limit_is_denominator, uint32_t *num, uint32_t *den) {
    uint64_t top = (uint64_t)(f * (double)(1 << 25)), bot = 1u << 25;
    uint64_t n_prev = 0, d_prev = 1, n_cur = 1, d_cur = 0;
    const double tol = 1.0 / (double)(1 << 24);
    for (;;) {
        if ((limit_is_denominator ? d_cur : n_cur) >= limit) break;
        if (d_cur) {
            double err = ((double)n_cur / (double)d_cur) - f;
            if (err < 0.0) err = -err;
            if (err < tol) { *num = (uint32_t)n_cur; *den = (uint32_t)d_cur; return true; }
        }
        if (bot == 0) break;
        const uint64_t q = top / bot, rem = top % bot;
        top = bot; bot = rem;
        uint64_t t = q * d_cur + d_prev; d_prev = d_cur; d_cur = t;
        t = q * n_cur + n_prev; n_prev = n_cur; n_cur = t;
    }
    if (limit_is_denominator) { n_cur = (uint64_t)(f * (double)limit + 0.5); d_cur = limit; }
    else { n_cur = limit; d_cur = (uint64_t)(((double)limit / f) + 0.5); }
    *num = (uint32_t)n_cur; *den = (uint32_t)d_cur;
    double err = d_cur ? (((double)(uint32_t)n_cur / (double)(uint32_t)d_cur) - f) : 1.0;
    if (err < 0.0) err = -err;
    return err < tol;
}

// Working form while building: ranges may still hang over the image edges.
struct Work {
    int width;                       // taps reserved per output
    std::vector<int> lo, hi;
    std::vector<float> c;
    float *row(int o) { return c.data() + (size_t)o * width; }
};

bool build_axis(int in_size, int out_size, bool is_horizontal, AxisTable *T) {
    if (in_size <= 0 || out_size <= 0) return false;
    T->in_size = in_size; T->out_size = out_size;
    const double scale_d = ((double)out_size / (double)in_size) * (((double)out_size / (double)out_size) / 1.0);
    const float scale = (float)scale_d;
    const float inv_scale = (float)(1.0 / scale_d);
    T->scale = scale; T->inv_scale = inv_scale;
    uint32_t num = 0, den = 0;
    const bool rational = rational_within_float(scale_d, scale_d <= 1.0 ? (uint32_t)out_size : (uint32_t)in_size,
                                                scale_d >= 1.0, &num, &den);
    const bool enlarging = scale >= (1.0f - kTiny);
    Filter flt{AxisFilter::kMitchell};                                    // :6499-6509
    if (enlarging) flt.kind = (scale <= (1.0f + kTiny)) ? AxisFilter::kPoint : AxisFilter::kBox;
    T->filter = flt.kind;
    T->filter_pixel_width = enlarging ? (int)std::ceil(flt.support(1.0f / scale) * 2.0f)
                                      : (int)std::ceil(flt.support(scale) * 2.0f / scale);   // :2962-2970
    T->gather_mode = enlarging ? 1 : ((is_horizontal || T->filter_pixel_width <= 32) ? 2 : 0); // :6530-6534
    const int margin = T->filter_pixel_width / 2;

    Work W;
    W.width = enlarging ? (int)std::ceil(flt.support(1.0f / scale) * 2.0f)
                        : (int)std::ceil(flt.support(scale) * 2.0f / scale);               // :2974-2990
    W.lo.assign(out_size, 0); W.hi.assign(out_size, -1);
    W.c.assign((size_t)out_size * W.width + 8, 0.0f);

    const bool periodic = rational && ((int)num < out_size);
    const int period_out = (int)num, period_in = (int)den;

    if (enlarging) {
        // one output at a time: which inputs fall under the (scaled) kernel (:3267-3327)
        const float radius = flt.support(inv_scale) * scale;
        const int n_calc = periodic ? period_out : out_size;
        for (int o = 0; o < n_calc; ++o) {
            const float centre = (float)o + 0.5f;
            const float centre_in = (centre + 0.0f) * inv_scale;
            const float lo_f = ((centre - radius) + 0.0f) * inv_scale;
            const float hi_f = ((centre + radius) + 0.0f) * inv_scale;
            int first = (int)std::floor(lo_f + 0.5f), last = (int)std::floor(hi_f - 0.5f);
            if (last < first) last = first;
            if (last - first + 1 > W.width) last = first + W.width - 1;
            float *c = W.row(o);
            int last_nz = -1;
            for (int i = 0; i <= last - first; ++i) {
                const float pc = (float)(i + first) + 0.5f;
                float v = flt.eval(centre_in - pc, inv_scale);
                if (v < kTiny && v > -kTiny) {
                    if (i == 0) { ++first; --i; continue; }    // drop leading zeros
                    v = 0;
                } else last_nz = i;
                c[i] = v;
            }
            W.lo[o] = first; W.hi[o] = last_nz + first;
        }
    } else {
        // one input at a time: which outputs does it reach (:3382-3458)
        const float radius = flt.support(scale) * inv_scale;
        int inited = -1;
        for (int p = -margin; p < in_size + margin; ++p) {
            const float pc = (float)p + 0.5f;
            const float pc_out = pc * scale - 0.0f;
            const float lo_f = (pc - radius) * scale - 0.0f, hi_f = (pc + radius) * scale - 0.0f;
            int ofirst = (int)std::floor(lo_f + 0.5f), olast = (int)std::floor(hi_f - 0.5f);
            if (ofirst < 0) ofirst = 0;
            if (olast >= out_size) olast = out_size - 1;
            if (ofirst > olast) continue;
            if (periodic) {
                if (ofirst == period_out) break;
                if (olast >= period_out) olast = period_out - 1;
            }
            for (int o = ofirst; o <= olast; ++o) {
                const float oc = (float)o + 0.5f;
                float v = flt.eval(oc - pc_out, scale) * scale;
                if (v < kTiny && v > -kTiny) v = 0.0f;
                float *c = W.row(o);
                if (o > inited) { inited = o; W.lo[o] = p; W.hi[o] = p; c[0] = v; }
                else {
                    if (c[0] == 0.0f) W.lo[o] = p;
                    W.hi[o] = p;
                    if (p - W.lo[o] >= W.width) return false;
                    c[p - W.lo[o]] = v;
                }
            }
        }
    }

    // normalise each tap set to sum 1 in double (:3482-3514)
    {
        const int n_calc = periodic ? period_out : out_size;
        for (int o = 0; o < n_calc; ++o) {
            float *c = W.row(o);
            const int e = W.hi[o] - W.lo[o];
            double total = 0;
            for (int i = 0; i <= e; ++i) total += (double)c[i];
            if (total < kTiny && total > -kTiny) { W.hi[o] = W.lo[o]; c[0] = 0.0f; }
            else if (total < (1.0f - kTiny) || total > (1.0f + kTiny)) {
                const double k = 1.0 / total;
                for (int i = 0; i <= e; ++i) c[i] = (float)(c[i] * k);
            }
        }
    }
    // replicate the first period across the axis (:3518-3536)
    if (periodic)
        for (int o = period_out; o < out_size; ++o) {
            W.lo[o] = W.lo[o - period_out] + period_in;
            W.hi[o] = W.hi[o - period_out] + period_in;
            std::memcpy(W.row(o), W.row(o - period_out), (size_t)W.width * sizeof(float));
        }
    // fold taps hanging over a clamped edge into the edge pixel, right then left (:3560-3594)
    const int last_in = in_size - 1;
    int widest = 1;
    for (int o = 0; o < out_size; ++o) {
        float *c = W.row(o);
        if (W.hi[o] > last_in) {
            const int lo = W.lo[o], hi = W.hi[o];
            if (last_in < lo) return false;
            W.hi[o] = last_in;
            for (int p = in_size; p <= hi; ++p) c[last_in - lo] += c[p - lo];
        }
        if (W.lo[o] < 0) {
            const int lo = W.lo[o];
            if (W.hi[o] < 0) return false;
            for (int p = -1; p > lo; --p) c[-lo] += c[p - lo];
            const float outermost = c[0];
            for (int i = 0; i <= W.hi[o]; ++i) c[i] = c[i - lo];
            W.lo[o] = 0;
            c[0] += outermost;
        }
        int n = W.hi[o] - W.lo[o] + 1;
        while (n > 0 && c[n - 1] == 0.0f) --n;           // trailing zeros are dropped (:3598-3602)
        if (n < 1) { n = 1; c[0] = 0.0f; }               // a fully-zero set still reads one tap
        W.hi[o] = W.lo[o] + n - 1;
        if (n > widest) widest = n;
    }

    T->widest = widest;
    T->first.assign(out_size, 0); T->count.assign(out_size, 0); T->lead.assign(out_size, 0);
    T->coeff.assign((size_t)out_size * widest, 0.0f);
    for (int o = 0; o < out_size; ++o) {
        T->first[o] = W.lo[o];
        T->count[o] = W.hi[o] - W.lo[o] + 1;
        std::memcpy(T->coeff.data() + (size_t)o * widest, W.row(o), (size_t)T->count[o] * sizeof(float));
    }
    if (is_horizontal) {
        // The reference's packed horizontal loops always read `widest` taps (or a rounded
        // count when widest > 12); near the right edge it slides the window back and pads
        // with leading zeros (:3803-3858).  Values are unchanged but taps alternate between
        // two accumulators, so the number of leading zeros shifts the parity.
        const int row_end = in_size;
        for (int o = out_size - 1; o >= 0 && (T->first[o] + widest * 2) >= row_end; --o) {
            if (T->first[o] + widest > row_end) {
                int span = widest;
                if (widest > 12) {
                    const int mod = widest & 3;
                    span = ((T->count[o] - mod + 3) & ~3) + mod;
                    if (span < 8 + mod) span = 8 + mod;
                }
                if (T->first[o] + span > row_end) T->lead[o] = T->first[o] - (row_end - span);
            }
        }
    }
    return true;
}

// Cost model deciding which axis goes first (:6859-6905); weights are the library's trained
// constants for 4-channel and 7-channel float pixels (:6770-6822).
const float kCost4[8][4] = {{0.00000f, 0.50000f, 0.00000f, 0.71875f}, {0.06250f, 0.84375f, 0.00000f, 0.87500f},
                            {1.00000f, 0.50000f, 0.50000f, 0.96875f}, {1.00000f, 0.09375f, 0.31250f, 0.50000f},
                            {1.00000f, 1.00000f, 1.00000f, 1.00000f}, {1.00000f, 0.03125f, 0.03125f, 0.53125f},
                            {0.18750f, 0.12500f, 0.00000f, 1.00000f}, {0.00000f, 1.00000f, 0.03125f, 0.18750f}};
const float kCost7[8][4] = {{0.00000f, 0.59375f, 0.00000f, 0.96875f}, {0.06250f, 0.81250f, 0.06250f, 0.59375f},
                            {0.75000f, 0.43750f, 0.12500f, 0.96875f}, {0.87500f, 0.06250f, 0.18750f, 0.43750f},
                            {1.00000f, 1.00000f, 1.00000f, 1.00000f}, {0.15625f, 0.12500f, 1.00000f, 1.00000f},
                            {0.06250f, 0.12500f, 0.00000f, 1.00000f}, {0.00000f, 1.00000f, 0.03125f, 0.34375f}};

bool decide_vertical_first(const AxisTable &h, const AxisTable &v, bool seven_channels) {
    int cls;
    if (v.out_size <= 4 || h.out_size <= 4) cls = (v.out_size < h.out_size) ? 6 : 7;
    else if (v.scale <= 1.0f) cls = v.gather_mode ? 1 : 0;
    else if (v.scale <= 2.0f) cls = 2;
    else if (v.scale <= 3.0f) cls = 3;
    else if (v.scale <= 4.0f) cls = 5;
    else cls = 6;
    const float *w = seven_channels ? kCost7[cls] : kCost4[cls];
    const double h_cost = (float)h.filter_pixel_width * w[0] + h.scale * (float)v.filter_pixel_width * w[1];
    const double v_cost = (float)v.filter_pixel_width * w[2] + v.scale * (float)h.filter_pixel_width * w[3];
    return v_cost <= h_cost;
}

}  // namespace

void free_plan(ResamplePlan *p) { delete p; }

bool build_resample_plan(int in_w, int in_h, int out_w, int out_h, ResamplePlan *plan) {
    if (!build_axis(in_w, out_w, true, &plan->h)) return false;
    if (!build_axis(in_h, out_h, false, &plan->v)) return false;
    plan->copy_only = plan->h.filter == AxisFilter::kPoint && plan->v.filter == AxisFilter::kPoint;
    plan->vertical_first = decide_vertical_first(plan->h, plan->v, !plan->copy_only);
    plan->h_sequential = plan->h.widest <= 3;
    return true;
}

}  // namespace b200timg
