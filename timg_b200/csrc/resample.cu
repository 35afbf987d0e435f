// K1: ImageScaler::Scale on the device -- separable polyphase resampling of RGBA8 frames
// with the reference's (STB build) exact arithmetic: decode byte*(1/255), alpha-weighted
// 7-channel float pixels [R G B A R*A G*A B*A], horizontal taps alternating between two
// accumulators (or one when <=3 taps), vertical taps summed in input-row order, un-weight by
// 1/A unless A < 2^-120, encode trunc(clamp(v*255+0.5)).  Which axis runs first follows the
// reference's cost model (resample_tables.cu).  Every * and + is a separate IEEE rounding
// (__fmul_rn/__fadd_rn), like the reference's SSE2 code.
//
// v1 layout: one thread per output pixel, taps read straight from global memory (the
// working set of a tile stays in L1/L2).  Algorithmic bytes: 4*iw*ih read + 4*ow*oh written.
#include "common.cuh"
#include "resample_tables.h"

namespace b200timg {

struct ResampleParams {
    int iw, ih, ow, oh, out_frame_rows, n_frames;
    int bgra;
    int h_widest, v_widest, h_sequential;
    const int32_t *h_first, *h_count, *h_lead, *v_first, *v_count;
    const float *h_coeff, *v_coeff;
};

struct Px7 { float c[7]; };

__device__ __forceinline__ void decode7(uint32_t p, int bgra, float *d) {
    const float k = 1.0f / 255.0f;
    const float c0 = fmul((float)(p & 0xff), k), c1 = fmul((float)((p >> 8) & 0xff), k);
    const float c2 = fmul((float)((p >> 16) & 0xff), k), a = fmul((float)(p >> 24), k);
    const float r = bgra ? c2 : c0, b = bgra ? c0 : c2;
    d[0] = r; d[1] = c1; d[2] = b; d[3] = a;
    d[4] = fmul(r, a); d[5] = fmul(c1, a); d[6] = fmul(b, a);
}

__device__ __forceinline__ uint32_t encode_px(const float *e) {
    float r = e[0], g = e[1], b = e[2];
    const float a = e[3];
    const float tiny = 7.5231638452626401e-37f;          // 2^-120
    if (!(a < tiny)) {
        const float ia = fdiv(1.0f, a);
        r = fmul(e[4], ia); g = fmul(e[5], ia); b = fmul(e[6], ia);
    }
    auto enc = [](float v) -> uint32_t {
        float f = fadd(fmul(v, 255.0f), 0.5f);
        f = f < 0.0f ? 0.0f : (f > 255.0f ? 255.0f : f);
        return __float2uint_rz(f);
    };
    return pack_rgba(enc(r), enc(g), enc(b), enc(a));
}

template <bool VFIRST>
__global__ void __launch_bounds__(256)
resample_direct_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, ResampleParams P) {
    const int ox = blockIdx.x * 32 + (threadIdx.x & 31);
    const int oy = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int f = blockIdx.z;
    if (ox >= P.ow || oy >= P.oh) return;
    const uint32_t *src = in + (long long)f * P.iw * P.ih;
    const int hn0 = P.h_first[ox], hcnt = P.h_count[ox];
    const int hpar = P.h_sequential ? 0 : (P.h_lead[ox] & 1);
    const int hmask = P.h_sequential ? 0 : 1;
    const int vn0 = P.v_first[oy], vcnt = P.v_count[oy];
    const float *hc = P.h_coeff + (long long)ox * P.h_widest;
    const float *vc = P.v_coeff + (long long)oy * P.v_widest;
    float res[7];
    if (!VFIRST) {
        // rows: horizontal sum per input row, then accumulate rows in order
        for (int k = 0; k < vcnt; ++k) {
            const uint32_t *row = src + (long long)(vn0 + k) * P.iw + hn0;
            float acc[2][7];
#pragma unroll
            for (int c = 0; c < 7; ++c) { acc[0][c] = 0.0f; acc[1][c] = 0.0f; }
            for (int i = 0; i < hcnt; ++i) {
                float d[7];
                decode7(row[i], P.bgra, d);
                const float w = hc[i];
                const int p = (i + hpar) & hmask;
#pragma unroll
                for (int c = 0; c < 7; ++c) {
                    const float t = fmul(d[c], w);
                    if (p) acc[1][c] = fadd(acc[1][c], t); else acc[0][c] = fadd(acc[0][c], t);
                }
            }
            const float wv = vc[k];
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                const float hsum = fadd(acc[0][c], acc[1][c]);
                const float t = fmul(hsum, wv);
                res[c] = (k == 0) ? t : fadd(res[c], t);
            }
        }
    } else {
        float acc[2][7];
#pragma unroll
        for (int c = 0; c < 7; ++c) { acc[0][c] = 0.0f; acc[1][c] = 0.0f; }
        for (int i = 0; i < hcnt; ++i) {
            const uint32_t *col = src + (long long)vn0 * P.iw + hn0 + i;
            float vs[7];
            for (int k = 0; k < vcnt; ++k) {
                float d[7];
                decode7(col[(long long)k * P.iw], P.bgra, d);
                const float wv = vc[k];
#pragma unroll
                for (int c = 0; c < 7; ++c) {
                    const float t = fmul(d[c], wv);
                    vs[c] = (k == 0) ? t : fadd(vs[c], t);
                }
            }
            const float w = hc[i];
            const int p = (i + hpar) & hmask;
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                const float t = fmul(vs[c], w);
                if (p) acc[1][c] = fadd(acc[1][c], t); else acc[0][c] = fadd(acc[0][c], t);
            }
        }
#pragma unroll
        for (int c = 0; c < 7; ++c) res[c] = fadd(acc[0][c], acc[1][c]);
    }
    out[((long long)f * P.out_frame_rows + oy) * P.ow + ox] = encode_px(res);
}

// both axes point-sampled (scale 1): plain copy, with the BGRA swizzle if asked (:6938-6940).
__global__ void __launch_bounds__(256)
resample_copy_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, ResampleParams P) {
    const long long frame_px = (long long)P.ow * P.oh;
    const long long total = frame_px * P.n_frames;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
        const long long f = g / frame_px, i = g - f * frame_px;
        const int oy = (int)(i / P.ow), ox = (int)(i - (long long)oy * P.ow);
        uint32_t p = in[(long long)f * P.iw * P.ih + (long long)P.v_first[oy] * P.iw + P.h_first[ox]];
        if (P.bgra) p = (p & 0xff00ff00u) | ((p & 0xff) << 16) | ((p >> 16) & 0xff);
        out[((long long)f * P.out_frame_rows + oy) * P.ow + ox] = p;
    }
}

// ---- plan cache + upload ----------------------------------------------------------------
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int launch_scale(b200timg_ctx *ctx, const uint8_t *d_in, int iw, int ih, int fmt, uint8_t *d_out,
                 int ow, int oh, int out_frame_rows, int n_frames) {
    if (out_frame_rows < oh) return ctx->fail(B200TIMG_EINVAL, "scale: frame rows < out height");
    if ((reinterpret_cast<uintptr_t>(d_in) & 3) || (reinterpret_cast<uintptr_t>(d_out) & 3))
        return ctx->fail(B200TIMG_EINVAL, "scale: pixel buffers must be 4-byte aligned");
    if (!ctx->plan) ctx->plan = new ResamplePlan();
    int *key = ctx->plan_key;
    const bool hit = key[0] == iw && key[1] == ih && key[2] == ow && key[3] == oh;
    ResamplePlan *pl = ctx->plan;
    if (!hit) {
        key[0] = 0;
        if (!build_resample_plan(iw, ih, ow, oh, pl)) return ctx->fail(B200TIMG_EINVAL, "scale: degenerate geometry");
    }
    // table layout in ctx->tables: [h_first|h_count|h_lead|v_first|v_count|h_coeff|v_coeff]
    const size_t o_hf = 0, o_hc = o_hf + align_up(sizeof(int32_t) * ow, 16), o_hl = o_hc + align_up(sizeof(int32_t) * ow, 16);
    const size_t o_vf = o_hl + align_up(sizeof(int32_t) * ow, 16), o_vc = o_vf + align_up(sizeof(int32_t) * oh, 16);
    const size_t o_hk = o_vc + align_up(sizeof(int32_t) * oh, 16);
    const size_t o_vk = o_hk + align_up(sizeof(float) * pl->h.coeff.size(), 16);
    const size_t total = o_vk + align_up(sizeof(float) * pl->v.coeff.size(), 16);
    if (!hit) {
        B2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));          // tables may be in use by earlier launches
        B2_CUDA(ctx, ctx->tables.reserve(total));
        B2_CUDA(ctx, ctx->pinned_io.reserve(total));
        char *h = ctx->pinned_io.as<char>();
        memcpy(h + o_hf, pl->h.first.data(), sizeof(int32_t) * ow);
        memcpy(h + o_hc, pl->h.count.data(), sizeof(int32_t) * ow);
        memcpy(h + o_hl, pl->h.lead.data(), sizeof(int32_t) * ow);
        memcpy(h + o_vf, pl->v.first.data(), sizeof(int32_t) * oh);
        memcpy(h + o_vc, pl->v.count.data(), sizeof(int32_t) * oh);
        memcpy(h + o_hk, pl->h.coeff.data(), sizeof(float) * pl->h.coeff.size());
        memcpy(h + o_vk, pl->v.coeff.data(), sizeof(float) * pl->v.coeff.size());
        B2_CUDA(ctx, cudaMemcpyAsync(ctx->tables.p, h, total, cudaMemcpyHostToDevice, ctx->stream));
        B2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));          // pinned_io is reused by callers
        key[0] = iw; key[1] = ih; key[2] = ow; key[3] = oh;
    }
    const char *t = ctx->tables.as<char>();
    ResampleParams P;
    P.iw = iw; P.ih = ih; P.ow = ow; P.oh = oh; P.out_frame_rows = out_frame_rows; P.n_frames = n_frames;
    P.bgra = fmt == B200TIMG_FMT_RGB32;
    P.h_widest = pl->h.widest; P.v_widest = pl->v.widest; P.h_sequential = pl->h_sequential ? 1 : 0;
    P.h_first = reinterpret_cast<const int32_t *>(t + o_hf);
    P.h_count = reinterpret_cast<const int32_t *>(t + o_hc);
    P.h_lead = reinterpret_cast<const int32_t *>(t + o_hl);
    P.v_first = reinterpret_cast<const int32_t *>(t + o_vf);
    P.v_count = reinterpret_cast<const int32_t *>(t + o_vc);
    P.h_coeff = reinterpret_cast<const float *>(t + o_hk);
    P.v_coeff = reinterpret_cast<const float *>(t + o_vk);
    const uint32_t *in = reinterpret_cast<const uint32_t *>(d_in);
    uint32_t *out = reinterpret_cast<uint32_t *>(d_out);
    if (pl->copy_only) {
        long long blocks = ((long long)ow * oh * n_frames + 255) / 256;
        if (blocks > (long long)ctx->sm_count * 16) blocks = (long long)ctx->sm_count * 16;
        B2_KERNEL(ctx, "resample_copy_kernel");
        resample_copy_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(in, out, P);
    } else {
        if (n_frames > 65535) return ctx->fail(B200TIMG_EINVAL, "scale: too many frames for one launch");
        const dim3 grid((ow + 31) / 32, (oh + 7) / 8, n_frames);
        B2_KERNEL(ctx, "resample_direct_kernel");
        if (pl->vertical_first) resample_direct_kernel<true><<<grid, 256, 0, ctx->stream>>>(in, out, P);
        else resample_direct_kernel<false><<<grid, 256, 0, ctx->stream>>>(in, out, P);
    }
    B2_LAUNCH_CHECK(ctx);
    return B200TIMG_OK;
}

}  // namespace b200timg

// Host-only introspection of the resampling plan (tests pin it against the oracle on CPU).
extern "C" int b200timg_resample_plan(int in_w, int in_h, int out_w, int out_h, int axis, int *widest,
                                      int *flags, int32_t *first, int32_t *count, int32_t *lead,
                                      float *coeff, size_t coeff_cap) {
    b200timg::ResamplePlan plan;
    if (!b200timg::build_resample_plan(in_w, in_h, out_w, out_h, &plan)) return B200TIMG_EINVAL;
    const b200timg::AxisTable &T = axis == 0 ? plan.h : plan.v;
    if (widest) *widest = T.widest;
    if (flags) *flags = (plan.vertical_first ? 1 : 0) | (plan.copy_only ? 2 : 0) | (plan.h_sequential ? 4 : 0);
    if (coeff_cap < T.coeff.size()) return B200TIMG_ENOSPC;
    if (first) memcpy(first, T.first.data(), sizeof(int32_t) * T.out_size);
    if (count) memcpy(count, T.count.data(), sizeof(int32_t) * T.out_size);
    if (lead) memcpy(lead, T.lead.data(), sizeof(int32_t) * T.out_size);
    if (coeff) memcpy(coeff, T.coeff.data(), sizeof(float) * T.coeff.size());
    return B200TIMG_OK;
}
