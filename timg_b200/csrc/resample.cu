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
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "resample_tables.h"

namespace b200timg {

struct ResampleParams {
    int iw, ih, ow, oh, out_frame_rows, n_frames;
    int bgra;
    int h_widest, v_widest, h_sequential;
    const int32_t *h_first, *h_count, *h_lead, *v_first, *v_count;
    const float *h_coeff, *v_coeff;
    ComposeSpec cs;              // fused AlphaComposeBackground (cs.active == 0: none)
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
        const float f = fadd(fmul(v, 255.0f), 0.5f);
        return __float2uint_rz(fminf(fmaxf(f, 0.0f), 255.0f));     // NaN -> 0 either way (cvt.rzi of NaN is 0)
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
    out[((long long)f * P.out_frame_rows + oy) * P.ow + ox] = compose_at(P.cs, encode_px(res), ox, oy);
}

// ---- tiled separable kernel ---------------------------------------------------------------
// One CTA produces a TW x TH output tile.  The input window the tile needs is decoded ONCE into
// shared memory as float4 pixels, the first pass (vertical or horizontal, whichever the
// reference's cost model picks) writes a float4 intermediate tile to shared memory, the second
// pass reads it.  Arithmetic and summation order are exactly those of resample_direct_kernel;
// only the sharing of partial sums between neighbouring output pixels is new.
// Only 4 of the 7 channels are carried (A, R*A, G*A, B*A): the un-weighted R,G,B are needed only
// where the filtered alpha is < 2^-120 (fully transparent output), and tiles that contain such a
// pixel run a second pass for them.
struct TileGeom { int tw, th, nix_max, niy_max; };

// byte k of p as a float without the (quarter-rate) I2F unit: 0x4B0000bb is 2^23 + bb exactly.
__device__ __forceinline__ float byte_f(uint32_t p, int k) {
    return fsub(__uint_as_float(__byte_perm(p, 0x4B000000u, 0x7540u | (uint32_t)k)), 8388608.0f);
}
__device__ __forceinline__ float4 decode_pm(uint32_t p, int bgra) {       // (R*A, G*A, B*A, A)
    const float k = 1.0f / 255.0f;
    const float c0 = fmul(byte_f(p, 0), k), c1 = fmul(byte_f(p, 1), k), c2 = fmul(byte_f(p, 2), k), a = fmul(byte_f(p, 3), k);
    const float r = bgra ? c2 : c0, b = bgra ? c0 : c2;
    return make_float4(fmul(r, a), fmul(c1, a), fmul(b, a), a);
}
__device__ __forceinline__ float4 decode_plain(uint32_t p, int bgra) {    // (R, G, B, -)
    const float k = 1.0f / 255.0f;
    const float c0 = fmul(byte_f(p, 0), k), c1 = fmul(byte_f(p, 1), k), c2 = fmul(byte_f(p, 2), k);
    return make_float4(bgra ? c2 : c0, c1, bgra ? c0 : c2, 0.0f);
}
// pass 0: (R*A, G*A, B*A, A); pass 1: (R, G, B, 0) -- one code path: the weight is A or 1.0 (x*1.0f is exact)
__device__ __forceinline__ float4 decode_sel(uint32_t p, int bgra, bool plain) {
    const float k = 1.0f / 255.0f;
    const float c0 = fmul(byte_f(p, 0), k), c1 = fmul(byte_f(p, 1), k), c2 = fmul(byte_f(p, 2), k), a = fmul(byte_f(p, 3), k);
    const float r = bgra ? c2 : c0, b = bgra ? c0 : c2, m = plain ? 1.0f : a;
    return make_float4(fmul(r, m), fmul(c1, m), fmul(b, m), plain ? 0.0f : a);
}
__device__ __forceinline__ float4 mul4(float4 v, float w) { return make_float4(fmul(v.x, w), fmul(v.y, w), fmul(v.z, w), fmul(v.w, w)); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z), fadd(a.w, b.w)); }

constexpr int RT = 256;      // threads per tile CTA
constexpr int RPPT = 4;      // max output pixels per thread (TW*TH <= RT*RPPT)

// Tap loops.  HW / VW > 0: compile-time tap budget (taps beyond an output's own count are skipped
// by predicate, the loop is fully unrolled); 0: run-time loop.  Horizontal taps go alternately to
// two accumulators and the two are added at the end -- which accumulator gets the even taps does
// not matter because IEEE addition is commutative, so the reference's "leading zero taps" (which
// only flip that assignment) need no special handling here.
template <int VW>
__device__ __forceinline__ float4 vsum(const float4 *col, int stride, int cnt, const float *vc) {
    float4 a = mul4(col[0], vc[0]);
    if (VW > 0) {
#pragma unroll
        for (int k = 1; k < VW; ++k) if (k < cnt) a = add4(a, mul4(col[k * stride], vc[k]));
    } else {
        for (int k = 1; k < cnt; ++k) a = add4(a, mul4(col[k * stride], vc[k]));
    }
    return a;
}
template <int HW>
__device__ __forceinline__ float4 hsum(const float4 *row, int cnt, const float *hc, bool sequential) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 a0 = z, a1 = z;
    if (sequential) {
        a0 = mul4(row[0], hc[0]);
#pragma unroll
        for (int i = 1; i < 3; ++i) if (i < cnt) a0 = add4(a0, mul4(row[i], hc[i]));
        return a0;
    }
    if (HW > 0) {
#pragma unroll
        for (int i = 0; i < HW; ++i) if (i < cnt) { const float4 t = mul4(row[i], hc[i]); if (i & 1) a1 = add4(a1, t); else a0 = add4(a0, t); }
    } else {
        for (int i = 0; i < cnt; ++i) { const float4 t = mul4(row[i], hc[i]); if (i & 1) a1 = add4(a1, t); else a0 = add4(a0, t); }
    }
    return add4(a0, a1);
}

template <bool VFIRST, int HW, int VW>
__global__ void __launch_bounds__(RT)
resample_tiled_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, ResampleParams P, TileGeom G) {
    extern __shared__ float4 s_px[];                  // Din[niy][nix], T, then the tile's tap tables
    __shared__ int s_ext[4];                          // ix0, ix1, iy0, iy1
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, f = blockIdx.z;
    const int ox0 = blockIdx.x * G.tw, oy0 = blockIdx.y * G.th;
    const int tw = min(G.tw, P.ow - ox0), th = min(G.th, P.oh - oy0);
    const int dstride = G.nix_max;
    float4 *Din = s_px;
    float4 *T = s_px + (size_t)G.niy_max * dstride;
    const int t_elems = VFIRST ? G.th * dstride : G.niy_max * G.tw;
    int *s_hmeta = reinterpret_cast<int *>(T + t_elems);           // [tw][2] first, count
    int *s_vmeta = s_hmeta + 2 * G.tw;                             // [th][2]
    const int hstride = P.h_widest | 1, vstride = P.v_widest | 1;  // odd: conflict-free
    float *s_hc = reinterpret_cast<float *>(s_vmeta + 2 * G.th);   // [tw][hstride]
    float *s_vc = s_hc + G.tw * hstride;                           // [th][vstride]
    if (tid == 0) { s_ext[0] = 0x7fffffff; s_ext[1] = -1; s_ext[2] = 0x7fffffff; s_ext[3] = -1; }
    __syncthreads();
    if (tid < tw) {
        const int a = P.h_first[ox0 + tid], c = P.h_count[ox0 + tid];
        s_hmeta[2 * tid] = a; s_hmeta[2 * tid + 1] = c;
        atomicMin(&s_ext[0], a); atomicMax(&s_ext[1], a + c - 1);
    }
    if (tid >= 64 && tid - 64 < th) {
        const int t = tid - 64, a = P.v_first[oy0 + t], c = P.v_count[oy0 + t];
        s_vmeta[2 * t] = a; s_vmeta[2 * t + 1] = c;
        atomicMin(&s_ext[2], a); atomicMax(&s_ext[3], a + c - 1);
    }
    for (int e = tid; e < tw * P.h_widest; e += RT) { const int x = e / P.h_widest, i = e - x * P.h_widest; s_hc[x * hstride + i] = P.h_coeff[(long long)(ox0 + x) * P.h_widest + i]; }
    for (int e = tid; e < th * P.v_widest; e += RT) { const int y = e / P.v_widest, i = e - y * P.v_widest; s_vc[y * vstride + i] = P.v_coeff[(long long)(oy0 + y) * P.v_widest + i]; }
    __syncthreads();
    const int ix0 = s_ext[0], nix = s_ext[1] - ix0 + 1, iy0 = s_ext[2], niy = s_ext[3] - iy0 + 1;
    const uint32_t *src = in + (long long)f * P.iw * P.ih;
    const bool hseq = P.h_sequential != 0;
    const float tiny = 7.5231638452626401e-37f;       // 2^-120
    // this thread's output pixels: one column tx, rows ty0 + q*rstep (they share the horizontal taps)
    const int tx = tid & (G.tw - 1), ty0 = tid / G.tw, rstep = RT / G.tw;
    const bool col_ok = tx < tw;

    float4 res[RPPT], plain[RPPT];
    bool need_plain = false;
    for (int pass = 0; pass < 2; ++pass) {
        for (int ly = wid; ly < niy; ly += RT / 32) {            // stage + decode once; a warp per row
            const uint32_t *row = src + (long long)(iy0 + ly) * P.iw + ix0;
            float4 *drow = Din + ly * dstride;
            for (int lx = lane; lx < nix; lx += 32) drow[lx] = pass == 0 ? decode_pm(row[lx], P.bgra) : decode_plain(row[lx], P.bgra);
        }
        __syncthreads();
        float4 acc[RPPT];
        if (VFIRST) {
            for (int ty = wid; ty < th; ty += RT / 32) {         // T[ty][lx] = sum_k Din[v_first+k][lx]*cv[k], rows in order
                const int n0 = s_vmeta[2 * ty] - iy0, cnt = s_vmeta[2 * ty + 1];
                const float *vc = s_vc + ty * vstride;
                const float4 *dcol = Din + n0 * dstride;
                float4 *trow = T + ty * dstride;
                for (int lx = lane; lx < nix; lx += 32) trow[lx] = vsum<VW>(dcol + lx, dstride, cnt, vc);
            }
            __syncthreads();
            if (col_ok) {
                const int n0 = s_hmeta[2 * tx] - ix0, cnt = s_hmeta[2 * tx + 1];
                const float *hc = s_hc + tx * hstride;
#pragma unroll
                for (int q = 0; q < RPPT; ++q) {
                    const int ty = ty0 + q * rstep;
                    if (ty < th) acc[q] = hsum<HW>(T + ty * dstride + n0, cnt, hc, hseq);
                }
            }
        } else {
            if (col_ok) {                                        // T[ly][tx] = sum_i Din[ly][h_first+i]*ch[i]
                const int n0 = s_hmeta[2 * tx] - ix0, cnt = s_hmeta[2 * tx + 1];
                const float *hc = s_hc + tx * hstride;
                for (int ly = ty0; ly < niy; ly += rstep) T[ly * G.tw + tx] = hsum<HW>(Din + ly * dstride + n0, cnt, hc, hseq);
            }
            __syncthreads();
            if (col_ok) {
#pragma unroll
                for (int q = 0; q < RPPT; ++q) {
                    const int ty = ty0 + q * rstep;
                    if (ty < th) {
                        const int n0 = s_vmeta[2 * ty] - iy0, cnt = s_vmeta[2 * ty + 1];
                        acc[q] = vsum<VW>(T + n0 * G.tw + tx, G.tw, cnt, s_vc + ty * vstride);
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < RPPT; ++q) { if (pass == 0) res[q] = acc[q]; else plain[q] = acc[q]; }
        if (pass == 0) {
#pragma unroll
            for (int q = 0; q < RPPT; ++q) if (col_ok && ty0 + q * rstep < th && res[q].w < tiny) need_plain = true;
            if (!__syncthreads_or(need_plain)) break;          // also orders T/Din reuse for pass 1
        }
    }
    if (col_ok) {
#pragma unroll
        for (int q = 0; q < RPPT; ++q) {
            const int ty = ty0 + q * rstep;
            if (ty < th) {
                float v[7];
                v[3] = res[q].w; v[4] = res[q].x; v[5] = res[q].y; v[6] = res[q].z;
                const bool transparent = res[q].w < tiny;
                v[0] = transparent ? plain[q].x : 0.f; v[1] = transparent ? plain[q].y : 0.f; v[2] = transparent ? plain[q].z : 0.f;
                out[((long long)f * P.out_frame_rows + oy0 + ty) * P.ow + ox0 + tx] = compose_at(P.cs, encode_px(v), ox0 + tx, oy0 + ty);
            }
        }
    }
}

// ---- fixed-tap fast path ------------------------------------------------------------------
// Same algorithm as resample_tiled_kernel, specialised so the inner loops carry no predicates and
// no run-time trip counts: every output uses exactly HC horizontal and VC vertical taps (the
// per-axis widest count rounded up to 2/4/6/8); taps an output does not have are zero coefficients
// reading staged (finite) pixels, which adds +0 and changes nothing.  The staged window is padded
// accordingly (zeros outside the image), its origin per tile column/row comes from the host.
struct FixedGeom { int nix, niy; const int32_t *tile_ix0, *tile_iy0; };
constexpr int FTW = 64;

template <bool VFIRST, int HC, int VC, int TH, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
resample_fixed_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, ResampleParams P, FixedGeom G) {
    extern __shared__ float4 s_px[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, f = blockIdx.z;
    const int ox0 = blockIdx.x * FTW, oy0 = blockIdx.y * TH;
    const int ix0 = G.tile_ix0[blockIdx.x], iy0 = G.tile_iy0[blockIdx.y];
    const int nix = G.nix, niy = G.niy;
    float4 *Din = s_px;
    float4 *T = s_px + (size_t)niy * nix;
    int *s_hfirst = reinterpret_cast<int *>(T + (VFIRST ? TH * nix : niy * FTW));   // [FTW]
    int *s_vfirst = s_hfirst + FTW;                                                   // [TH]
    float *s_hc = reinterpret_cast<float *>(s_vfirst + TH);                          // [FTW][HC+1]
    float *s_vc = s_hc + FTW * (HC + 1);                                              // [TH][VC+1]
    if (tid < FTW) {
        const int ox = ox0 + tid;
        const bool ok = ox < P.ow;
        s_hfirst[tid] = ok ? P.h_first[ox] - ix0 : 0;
#pragma unroll
        for (int i = 0; i < HC; ++i) s_hc[tid * (HC + 1) + i] = (ok && i < P.h_widest) ? P.h_coeff[(long long)ox * P.h_widest + i] : 0.0f;
    } else if (tid < FTW + TH) {
        const int t = tid - FTW, oy = oy0 + t;
        const bool ok = oy < P.oh;
        s_vfirst[t] = ok ? P.v_first[oy] - iy0 : 0;
#pragma unroll
        for (int i = 0; i < VC; ++i) s_vc[t * (VC + 1) + i] = (ok && i < P.v_widest) ? P.v_coeff[(long long)oy * P.v_widest + i] : 0.0f;
    }
    const uint32_t *src = in + (long long)f * P.iw * P.ih;
    const bool hseq = P.h_sequential != 0;
    const float tiny = 7.5231638452626401e-37f;       // 2^-120
    constexpr int NW = NT / 32, RSTEP = NT / FTW, RPT = TH / RSTEP;   // warps, row stride, rows per thread
    const int tx = tid & (FTW - 1), tyb = tid >> 6;   // this thread's pixels: column tx, rows tyb + RSTEP*q

    float4 res[RPT], plain[RPT];
    bool need_plain = false;
    for (int pass = 0; pass < 2; ++pass) {
        // stage + decode the window once.  All of a thread's global loads (up to 4 rows x 4 column
        // groups) are issued before the first one is consumed, so their latencies overlap.
        for (int ly0 = 0; ly0 < niy; ly0 += 4 * NW)
            for (int lx0 = 0; lx0 < nix; lx0 += 128) {
                uint32_t pv[4][4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int ly = ly0 + wid + r * NW, y = iy0 + ly;
                    const uint32_t *row = src + (long long)y * P.iw + ix0;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int lx = lx0 + lane + 32 * c;
                        pv[r][c] = (ly < niy && lx < nix && y < P.ih && ix0 + lx < P.iw) ? row[lx] : 0u;
                    }
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int ly = ly0 + wid + r * NW;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int lx = lx0 + lane + 32 * c;
                        if (ly < niy && lx < nix) Din[ly * nix + lx] = decode_sel(pv[r][c], P.bgra, pass != 0);
                    }
                }
            }
        __syncthreads();
        float4 acc[RPT];
        if (VFIRST) {
#pragma unroll
            for (int r = 0; r < TH / NW; ++r) {
                const int ty = wid + r * NW;
                float vc[VC];
#pragma unroll
                for (int k = 0; k < VC; ++k) vc[k] = s_vc[ty * (VC + 1) + k];
                const float4 *rp[VC];
#pragma unroll
                for (int k = 0; k < VC; ++k) rp[k] = Din + (s_vfirst[ty] + k) * nix + lane;
                float4 *trow = T + ty * nix + lane;
                for (int lx0 = 0; lx0 < nix; lx0 += 128) {
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int o = lx0 + 32 * it;
                        if (o + lane < nix) {
                            float4 a = mul4(rp[0][o], vc[0]);
#pragma unroll
                            for (int k = 1; k < VC; ++k) a = add4(a, mul4(rp[k][o], vc[k]));
                            trow[o] = a;
                        }
                    }
                }
            }
            __syncthreads();
            float hc[HC];
#pragma unroll
            for (int i = 0; i < HC; ++i) hc[i] = s_hc[tx * (HC + 1) + i];
            const float4 *tbase = T + s_hfirst[tx];
#pragma unroll
            for (int q = 0; q < RPT; ++q) {
                const float4 *row = tbase + (tyb + RSTEP * q) * nix;
                if (hseq) {
                    float4 a = mul4(row[0], hc[0]);
#pragma unroll
                    for (int i = 1; i < (HC < 3 ? HC : 3); ++i) a = add4(a, mul4(row[i], hc[i]));
                    acc[q] = a;
                } else {
                    float4 a0 = mul4(row[0], hc[0]), a1 = mul4(row[1], hc[1]);
#pragma unroll
                    for (int i = 2; i < HC; ++i) { if (i & 1) a1 = add4(a1, mul4(row[i], hc[i])); else a0 = add4(a0, mul4(row[i], hc[i])); }
                    acc[q] = add4(a0, a1);
                }
            }
        } else {
            float hc[HC];
#pragma unroll
            for (int i = 0; i < HC; ++i) hc[i] = s_hc[tx * (HC + 1) + i];
            const int hn0 = s_hfirst[tx];
            for (int ly = tyb; ly < niy; ly += RSTEP) {
                const float4 *row = Din + ly * nix + hn0;
                float4 r;
                if (hseq) {
                    r = mul4(row[0], hc[0]);
#pragma unroll
                    for (int i = 1; i < (HC < 3 ? HC : 3); ++i) r = add4(r, mul4(row[i], hc[i]));
                } else {
                    float4 a0 = mul4(row[0], hc[0]), a1 = mul4(row[1], hc[1]);
#pragma unroll
                    for (int i = 2; i < HC; ++i) { if (i & 1) a1 = add4(a1, mul4(row[i], hc[i])); else a0 = add4(a0, mul4(row[i], hc[i])); }
                    r = add4(a0, a1);
                }
                T[ly * FTW + tx] = r;
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < RPT; ++q) {
                const int ty = tyb + RSTEP * q;
                const float4 *col = T + s_vfirst[ty] * FTW + tx;
                const float *vc = s_vc + ty * (VC + 1);
                float4 a = mul4(col[0], vc[0]);
#pragma unroll
                for (int k = 1; k < VC; ++k) a = add4(a, mul4(col[k * FTW], vc[k]));
                acc[q] = a;
            }
        }
#pragma unroll
        for (int q = 0; q < RPT; ++q) { if (pass == 0) res[q] = acc[q]; else plain[q] = acc[q]; }
        if (pass == 0) {
#pragma unroll
            for (int q = 0; q < RPT; ++q) if (ox0 + tx < P.ow && oy0 + tyb + RSTEP * q < P.oh && res[q].w < tiny) need_plain = true;
            if (!__syncthreads_or(need_plain)) break;
        }
    }
    if (ox0 + tx < P.ow) {
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
            const int oy = oy0 + tyb + RSTEP * q;
            if (oy < P.oh) {
                float v[7];
                v[3] = res[q].w; v[4] = res[q].x; v[5] = res[q].y; v[6] = res[q].z;
                const bool transparent = res[q].w < tiny;
                v[0] = transparent ? plain[q].x : 0.f; v[1] = transparent ? plain[q].y : 0.f; v[2] = transparent ? plain[q].z : 0.f;
                out[((long long)f * P.out_frame_rows + oy) * P.ow + ox0 + tx] = compose_at(P.cs, encode_px(v), ox0 + tx, oy);
            }
        }
    }
}

// ---- planar fast path (vertical pass first, <= 8 taps per axis) ---------------------------
// Same arithmetic again, reorganised around what limits resample_fixed_kernel on B200 (shared-memory
// wavefronts and issue slots, see profiles/): every filtered channel is an independent plane, so the
// window is staged as separate float planes and
//   * the vertical pass produces 4 neighbouring columns per thread (one LDS.128 per tap and plane),
//   * the horizontal pass maps lanes to output ROWS: a warp works on one output column at a time, its
//     taps and start index are warp-uniform and every tap is one conflict-free wavefront (odd pitch),
//   * the encoded pixels go through a small transpose buffer so global stores stay coalesced.
// Channel passes: a tile whose window is fully opaque (the common case: photos, video) needs only the
// three colour planes -- decoded alpha is exactly 1.0f (255 * fl(1/255) rounds to 1), so R*A == R, and
// the filtered alpha is the same tap sum over the constant 1.0, computed from the coefficients alone in
// the same order.  Other tiles run (R*A, G*A, B*A), then A, and -- only if some output alpha is
// < 2^-120 -- the un-weighted (R, G, B) planes, exactly like the float4 kernels above.
struct PlanarGeom { int nix, niy, sp, tp; unsigned grp_magic; const int32_t *tile_ix0, *tile_iy0; };   // grp_magic: floor(2^32/(sp/4))+1
constexpr int PTH = 32;          // output rows per tile (= lanes of the horizontal pass)
constexpr int PNT = 256;

template <int HC, int VC, int PTW, int MINB>
__global__ void __launch_bounds__(PNT, MINB)
resample_planar_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, ResampleParams P, PlanarGeom G) {
    extern __shared__ float4 s_px[];
    float *S = reinterpret_cast<float *>(s_px);                   // [3][niy][sp]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, f = blockIdx.z;
    const int ox0 = blockIdx.x * PTW, oy0 = blockIdx.y * PTH;
    const int ix0 = G.tile_ix0[blockIdx.x], iy0 = G.tile_iy0[blockIdx.y];
    const int niy = G.niy, sp = G.sp, tp = G.tp;
    const int splane = niy * sp, tplane = PTH * tp;
    float *T = S + 3 * splane;                                    // [3][PTH][tp]
    float *s_hc = T + 3 * tplane + ((4 - ((3 * tplane) & 3)) & 3); // [PTW][8], 16-byte aligned
    float *s_vc = s_hc + PTW * 8;                                 // [PTH][8]
    int *s_hfirst = reinterpret_cast<int *>(s_vc + PTH * 8);      // [PTW]
    int *s_vfirst = s_hfirst + PTW;                               // [PTH]
    uint32_t *O = reinterpret_cast<uint32_t *>(S);                // [PTH][PTW+1], reuses S after the last pass
    if (tid < PTW) {
        const int ox = ox0 + tid;
        const bool ok = ox < P.ow;
        s_hfirst[tid] = ok ? P.h_first[ox] - ix0 : 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s_hc[tid * 8 + i] = (ok && i < P.h_widest) ? P.h_coeff[(long long)ox * P.h_widest + i] : 0.0f;
    } else if (tid < PTW + PTH) {
        const int t = tid - PTW, oy = oy0 + t;
        const bool ok = oy < P.oh;
        s_vfirst[t] = ok ? P.v_first[oy] - iy0 : 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s_vc[t * 8 + i] = (ok && i < P.v_widest) ? P.v_coeff[(long long)oy * P.v_widest + i] : 0.0f;
    }
    const uint32_t *src = in + (long long)f * P.iw * P.ih;
    const bool hseq = P.h_sequential != 0;
    const int kr = P.bgra ? 2 : 0, kb = P.bgra ? 0 : 2;          // byte index of R and B in the source pixel
    const float tiny = 7.5231638452626401e-37f;       // 2^-120
    const float k255 = 1.0f / 255.0f;
    constexpr int NJ = PTW / 8;                       // output columns per warp: wid + 8*j, row = lane
    const int ngrp = sp >> 2;                         // 4-column groups per window row
    const int n_stage = niy * ngrp;

    // ---- staging: a unit is 4 neighbouring source pixels (one 16-byte load) -> one float4 per plane
    uint4 raw[4];
    int soff[4];                                      // S offset of the unit, -1: none
    auto load_chunk = [&](int u0) -> bool {           // all loads of a thread are issued before the first use
        bool ok255 = true;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int u = u0 + tid + r * PNT;
            const int ly = (int)__umulhi((unsigned)u, G.grp_magic), g = u - ly * ngrp;
            const int y = iy0 + ly, x = ix0 + 4 * g;
            raw[r] = make_uint4(0u, 0u, 0u, 0u);
            soff[r] = u < n_stage ? ly * sp + 4 * g : -1;
            if (u < n_stage && y < P.ih && x < P.iw) {
                raw[r] = __ldg(reinterpret_cast<const uint4 *>(src + (long long)y * P.iw + x));
                ok255 = ok255 && ((raw[r].x & raw[r].y & raw[r].z & raw[r].w) >= 0xff000000u);
            }
        }
        return ok255;
    };
    // mode 0: R*A,G*A,B*A   1: A   2: R,G,B
    auto decode_chunk = [&](int mode) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (soff[r] < 0) continue;
            const uint32_t pv[4] = {raw[r].x, raw[r].y, raw[r].z, raw[r].w};
            float *d = S + soff[r];
            if (mode == 1) {
                *reinterpret_cast<float4 *>(d) = make_float4(fmul(byte_f(pv[0], 3), k255), fmul(byte_f(pv[1], 3), k255),
                                                             fmul(byte_f(pv[2], 3), k255), fmul(byte_f(pv[3], 3), k255));
                continue;
            }
            float c[3][4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                c[0][e] = fmul(byte_f(pv[e], kr), k255); c[1][e] = fmul(byte_f(pv[e], 1), k255); c[2][e] = fmul(byte_f(pv[e], kb), k255);
                if (mode == 0) {
                    const float a = fmul(byte_f(pv[e], 3), k255);
                    c[0][e] = fmul(c[0][e], a); c[1][e] = fmul(c[1][e], a); c[2][e] = fmul(c[2][e], a);
                }
            }
            *reinterpret_cast<float4 *>(d) = make_float4(c[0][0], c[0][1], c[0][2], c[0][3]);
            *reinterpret_cast<float4 *>(d + splane) = make_float4(c[1][0], c[1][1], c[1][2], c[1][3]);
            *reinterpret_cast<float4 *>(d + 2 * splane) = make_float4(c[2][0], c[2][1], c[2][2], c[2][3]);
        }
    };
    auto stage = [&](int mode, bool have_first_chunk) {
        for (int u0 = 0; u0 < n_stage; u0 += 4 * PNT) {
            if (!(have_first_chunk && u0 == 0)) load_chunk(u0);
            if (mode == 0) decode_chunk(0); else if (mode == 1) decode_chunk(1); else decode_chunk(2);
        }
    };
    // ---- vertical: T[c][ty][4g..4g+3] = sum_k S[c][vfirst[ty]+k][4g..] * vc[ty][k], rows in order.
    // A warp covers 4 output rows x 8 column groups: with an odd T pitch its scalar stores hit 32
    // different banks, and each LDS.128 touches four 128-byte row segments (the minimum).
    auto vertical = [&](auto np_tag) {
        constexpr int NP = decltype(np_tag)::value;
        const int ty = 4 * wid + (lane >> 3);
        const float4 v0 = *reinterpret_cast<const float4 *>(s_vc + ty * 8), v1 = *reinterpret_cast<const float4 *>(s_vc + ty * 8 + 4);
        const float vc[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        const float *sb = S + s_vfirst[ty] * sp + 4 * (lane & 7);
        float *t = T + ty * tp + 4 * (lane & 7);
        for (int g = lane & 7; g < ngrp; g += 8, sb += 32, t += 32) {
#pragma unroll
            for (int c = 0; c < NP; ++c) {
                float4 a = mul4(*reinterpret_cast<const float4 *>(sb + c * splane), vc[0]);
#pragma unroll
                for (int k = 1; k < VC; ++k) a = add4(a, mul4(*reinterpret_cast<const float4 *>(sb + c * splane + k * sp), vc[k]));
                float *tc = t + c * tplane;
                tc[0] = a.x; tc[1] = a.y; tc[2] = a.z; tc[3] = a.w;
            }
        }
    };
    // tap sum of one output: HC taps alternating between two accumulators (or sequential when <= 3 taps)
    auto hsum1 = [&](auto tap, const float *hc) -> float {
        if (hseq) {
            float a = fmul(tap(0), hc[0]);
#pragma unroll
            for (int i = 1; i < (HC < 3 ? HC : 3); ++i) a = fadd(a, fmul(tap(i), hc[i]));
            return a;
        }
        float a0 = fmul(tap(0), hc[0]), a1 = fmul(tap(1), hc[1]);
#pragma unroll
        for (int i = 2; i < HC; ++i) { if (i & 1) a1 = fadd(a1, fmul(tap(i), hc[i])); else a0 = fadd(a0, fmul(tap(i), hc[i])); }
        return fadd(a0, a1);
    };
    // ---- horizontal: lane = output row, the warp's column changes with j
    auto horizontal = [&](auto np_tag, float (*r)[3]) {
        constexpr int NP = decltype(np_tag)::value;
        const float *trow = T + lane * tp;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int tx = wid + 8 * j;
            const float4 h0 = *reinterpret_cast<const float4 *>(s_hc + tx * 8), h1 = *reinterpret_cast<const float4 *>(s_hc + tx * 8 + 4);
            const float hc[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
            const float *row = trow + s_hfirst[tx];
#pragma unroll
            for (int c = 0; c < NP; ++c) {
                const float *rc = row + c * tplane;
                r[j][c] = hsum1([&](int i) { return rc[i]; }, hc);
            }
        }
    };
    using I1 = std::integral_constant<int, 1>;
    using I3 = std::integral_constant<int, 3>;

    float pm[NJ][3], al[NJ], pl[NJ][3];
    bool have_pl = false;
    // is every pixel of the window opaque?  (decoded alpha is then exactly 1.0f)
    bool ok255 = load_chunk(0);
    for (int u0 = 4 * PNT; u0 < n_stage; u0 += 4 * PNT) ok255 = load_chunk(u0) && ok255;
    const bool opaque = __syncthreads_and(ok255) != 0;           // also orders the table writes above
    const bool one_chunk = n_stage <= 4 * PNT;                   // then raw[] still holds the window
#pragma unroll 1
    for (int rep = 0; rep < 2; ++rep) {
        // rep 0: weighted colour planes (un-weighted == weighted when opaque); rep 1: un-weighted colour planes
        if (rep == 0 && !opaque) stage(0, one_chunk); else stage(2, rep == 0 && one_chunk);
        __syncthreads();
        vertical(I3());
        __syncthreads();
        if (rep == 1) { horizontal(I3(), pl); have_pl = true; break; }
        horizontal(I3(), pm);
        if (opaque) {
            // filtered alpha of an all-ones window: the same two tap sums over the constant 1.0f
            const float4 v0 = *reinterpret_cast<const float4 *>(s_vc + lane * 8), v1 = *reinterpret_cast<const float4 *>(s_vc + lane * 8 + 4);
            const float vc[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            float av = vc[0];
#pragma unroll
            for (int k = 1; k < VC; ++k) av = fadd(av, vc[k]);
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int tx = wid + 8 * j;
                const float4 h0 = *reinterpret_cast<const float4 *>(s_hc + tx * 8), h1 = *reinterpret_cast<const float4 *>(s_hc + tx * 8 + 4);
                const float hc[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
                al[j] = hsum1([&](int) { return av; }, hc);
            }
            break;
        }
        // alpha plane
        __syncthreads();                               // T of the colour pass has been consumed
        stage(1, false);
        __syncthreads();
        vertical(I1());
        __syncthreads();
        float a1[NJ][3];
        horizontal(I1(), a1);
        bool need_plain = false;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            al[j] = a1[j][0];
            if (ox0 + wid + 8 * j < P.ow && oy0 + lane < P.oh && al[j] < tiny) need_plain = true;
        }
        if (!__syncthreads_or(need_plain)) break;
    }
    __syncthreads();                                   // every warp is done with S and T
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int tx = wid + 8 * j;
        float v[7];
        v[3] = al[j]; v[4] = pm[j][0]; v[5] = pm[j][1]; v[6] = pm[j][2];
        if (have_pl) { v[0] = pl[j][0]; v[1] = pl[j][1]; v[2] = pl[j][2]; }      // only read when al < 2^-120
        else { v[0] = pm[j][0]; v[1] = pm[j][1]; v[2] = pm[j][2]; }
        O[lane * (PTW + 1) + tx] = compose_at(P.cs, encode_px(v), ox0 + tx, oy0 + lane);
    }
    __syncthreads();
    {
        static_assert(PTW == 32, "store loop maps a lane to a column");
        uint32_t *orow = out + ((long long)f * P.out_frame_rows + oy0 + wid) * P.ow + ox0 + lane;
        const bool col_ok = ox0 + lane < P.ow;
#pragma unroll
        for (int i = 0; i < PTH / 8; ++i)
            if (col_ok && oy0 + wid + 8 * i < P.oh) orow[(long long)(8 * i) * P.ow] = O[(wid + 8 * i) * (PTW + 1) + lane];
    }
}

typedef void (*PlanarFn)(const uint32_t *, uint32_t *, ResampleParams, PlanarGeom);
template <int HC, int PTW, int MINB>
static PlanarFn planar_v(int vc) {
    switch (vc) {
    case 2: return resample_planar_kernel<HC, 2, PTW, MINB>;
    case 4: return resample_planar_kernel<HC, 4, PTW, MINB>;
    case 6: return resample_planar_kernel<HC, 6, PTW, MINB>;
    default: return resample_planar_kernel<HC, 8, PTW, MINB>;
    }
}
template <int PTW, int MINB>
static PlanarFn planar_h(int hc, int vc) {
    switch (hc) {
    case 2: return planar_v<2, PTW, MINB>(vc);
    case 4: return planar_v<4, PTW, MINB>(vc);
    case 6: return planar_v<6, PTW, MINB>(vc);
    default: return planar_v<8, PTW, MINB>(vc);
    }
}

typedef void (*FixedFn)(const uint32_t *, uint32_t *, ResampleParams, FixedGeom);
struct FixedVariant { FixedFn fn; int th, nt; };
// Tile 64x16, 256 threads, 3 CTAs/SM (80 registers) measured best on B200 among {16x256x3, 16x256x4,
// 8x128x8, 8x256x4, 16x512x2, 32x512x2} (4.21 / 4.76 / 5.01 / 4.79 / 5.45 / 6.75 ms for 64 C2 frames).
template <bool VF, int HC>
static FixedVariant fixed_v(int vc, char v) {
    switch (vc) {
    case 2: return {resample_fixed_kernel<VF, HC, 2, 16, 256, 3>, 16, 256};
    case 4: return {resample_fixed_kernel<VF, HC, 4, 16, 256, 3>, 16, 256};
    case 6: return {resample_fixed_kernel<VF, HC, 6, 16, 256, 3>, 16, 256};
    default: return {resample_fixed_kernel<VF, HC, 8, 16, 256, 3>, 16, 256};
    }
}
template <bool VF>
static FixedVariant fixed_h(int hc, int vc, char v) {
    switch (hc) {
    case 2: return fixed_v<VF, 2>(vc, v);
    case 4: return fixed_v<VF, 4>(vc, v);
    case 6: return fixed_v<VF, 6>(vc, v);
    default: return fixed_v<VF, 8>(vc, v);
    }
}
static int fixed_class(int widest) { return widest <= 2 ? 2 : widest <= 4 ? 4 : widest <= 6 ? 6 : 8; }

typedef void (*TiledFn)(const uint32_t *, uint32_t *, ResampleParams, TileGeom);
template <bool VF, int HW>
static TiledFn pick_v(int vclass) {
    switch (vclass) {
    case 4: return resample_tiled_kernel<VF, HW, 4>;
    case 6: return resample_tiled_kernel<VF, HW, 6>;
    case 8: return resample_tiled_kernel<VF, HW, 8>;
    default: return resample_tiled_kernel<VF, HW, 0>;
    }
}
template <bool VF>
static TiledFn pick_h(int hclass, int vclass) {
    switch (hclass) {
    case 4: return pick_v<VF, 4>(vclass);
    case 6: return pick_v<VF, 6>(vclass);
    case 8: return pick_v<VF, 8>(vclass);
    default: return pick_v<VF, 0>(vclass);
    }
}
static int tap_class(int widest) { return widest <= 4 ? 4 : widest <= 6 ? 6 : widest <= 8 ? 8 : 0; }

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
        out[((long long)f * P.out_frame_rows + oy) * P.ow + ox] = compose_at(P.cs, p, ox, oy);
    }
}

// ---- plan cache + upload ----------------------------------------------------------------
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int launch_scale(b200timg_ctx *ctx, const uint8_t *d_in, int iw, int ih, int fmt, uint8_t *d_out,
                 int ow, int oh, int out_frame_rows, int n_frames, const ComposeSpec *cs) {
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
    if (cs) P.cs = *cs; else { memset(&P.cs, 0, sizeof P.cs); P.cs.pw = P.cs.ph = 1; }
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
        // fast path: both axes need <= 8 taps -> fixed-tap kernel on 64x16 tiles
        // fastest path: vertical pass first, <= 8 taps per axis, 16-byte aligned rows -> planar kernel
        if (pl->vertical_first && pl->h.widest <= 8 && pl->v.widest <= 8 && (iw & 3) == 0 &&
            (reinterpret_cast<uintptr_t>(d_in) & 15) == 0 && !getenv("B200TIMG_NO_PLANAR") && !getenv("B200TIMG_NO_FIXED")) {
            const int hc = fixed_class(pl->h.widest), vc = fixed_class(pl->v.widest);
            const int ptw = 32;
            const int ntx = (ow + ptw - 1) / ptw, nty = (oh + PTH - 1) / PTH;
            std::vector<int32_t> tix(ntx), tiy(nty);
            int nix = 1, niy = 1;
            for (int j = 0; j < ntx; ++j) {
                int lo = 0x7fffffff, hi = -1;
                for (int x = j * ptw; x < std::min(ow, (j + 1) * ptw); ++x) { lo = std::min(lo, pl->h.first[x]); hi = std::max(hi, pl->h.first[x] + hc - 1); }
                lo &= ~3;                                    // aligned 16-byte loads
                tix[j] = lo; nix = std::max(nix, hi - lo + 1);
            }
            for (int j = 0; j < nty; ++j) {
                int lo = 0x7fffffff, hi = -1;
                for (int y = j * PTH; y < std::min(oh, (j + 1) * PTH); ++y) { lo = std::min(lo, pl->v.first[y]); hi = std::max(hi, pl->v.first[y] + vc - 1); }
                tiy[j] = lo; niy = std::max(niy, hi - lo + 1);
            }
            const int sp = (nix + 3) & ~3, tp = sp | 1;   // odd T pitch >= the 4-column groups written per row
            const size_t psmem = sizeof(float) * (3 * ((size_t)niy * sp + (size_t)PTH * tp) + 4 + (size_t)ptw * 8 + PTH * 8) + sizeof(int) * (ptw + PTH);
            const size_t smem_cap = 75 * 1024;           // 3 CTAs/SM (80 registers): 7.96 ms vs 10.2 ms at 2 CTAs/SM, 148 C2 frames
            if (psmem <= smem_cap && (size_t)PTH * (ptw + 1) <= 3 * (size_t)niy * sp) {
                B2_CUDA(ctx, ctx->misc.reserve(4096 + sizeof(int32_t) * (size_t)(ntx + nty)));
                int32_t *d_t = reinterpret_cast<int32_t *>(ctx->misc.as<char>() + 4096);
                B2_CUDA(ctx, cudaMemcpyAsync(d_t, tix.data(), sizeof(int32_t) * ntx, cudaMemcpyHostToDevice, ctx->stream));
                B2_CUDA(ctx, cudaMemcpyAsync(d_t + ntx, tiy.data(), sizeof(int32_t) * nty, cudaMemcpyHostToDevice, ctx->stream));
                PlanarGeom PG{nix, niy, sp, tp, (unsigned)(0x100000000ull / (unsigned)(sp / 4)) + 1u, d_t, d_t + ntx};
                PlanarFn fn = planar_h<32, 3>(hc, vc);
                B2_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
                const dim3 grid(ntx, nty, n_frames);
                B2_KERNEL(ctx, "resample_planar_kernel");
                fn<<<grid, PNT, psmem, ctx->stream>>>(in, out, P, PG);
                B2_LAUNCH_CHECK(ctx);
                return B200TIMG_OK;
            }
        }
        if (pl->h.widest <= 8 && pl->v.widest <= 8 && !getenv("B200TIMG_NO_FIXED")) {
            const int hc = fixed_class(pl->h.widest), vc = fixed_class(pl->v.widest);
            const FixedVariant fv = pl->vertical_first ? fixed_h<true>(hc, vc, 'A') : fixed_h<false>(hc, vc, 'A');
            const int FTHv = fv.th;
            const int ntx = (ow + FTW - 1) / FTW, nty = (oh + FTHv - 1) / FTHv;
            std::vector<int32_t> tix(ntx), tiy(nty);
            int nix = 1, niy = 1;
            for (int j = 0; j < ntx; ++j) {
                int lo = 0x7fffffff, hi = -1;
                for (int x = j * FTW; x < std::min(ow, (j + 1) * FTW); ++x) { lo = std::min(lo, pl->h.first[x]); hi = std::max(hi, pl->h.first[x] + hc - 1); }
                tix[j] = lo; nix = std::max(nix, hi - lo + 1);
            }
            for (int j = 0; j < nty; ++j) {
                int lo = 0x7fffffff, hi = -1;
                for (int y = j * FTHv; y < std::min(oh, (j + 1) * FTHv); ++y) { lo = std::min(lo, pl->v.first[y]); hi = std::max(hi, pl->v.first[y] + vc - 1); }
                tiy[j] = lo; niy = std::max(niy, hi - lo + 1);
            }
            const size_t fsmem = sizeof(float4) * ((size_t)nix * niy + (pl->vertical_first ? (size_t)FTHv * nix : (size_t)niy * FTW))
                               + sizeof(int) * (FTW + FTHv) + sizeof(float) * ((size_t)FTW * (hc + 1) + (size_t)FTHv * (vc + 1));
            if (fsmem <= 100 * 1024) {
                // tile origins live behind the flag word in ctx->misc (re-uploaded per call: a few hundred bytes)
                B2_CUDA(ctx, ctx->misc.reserve(4096 + sizeof(int32_t) * (size_t)(ntx + nty)));
                int32_t *d_t = reinterpret_cast<int32_t *>(ctx->misc.as<char>() + 4096);
                B2_CUDA(ctx, cudaMemcpyAsync(d_t, tix.data(), sizeof(int32_t) * ntx, cudaMemcpyHostToDevice, ctx->stream));
                B2_CUDA(ctx, cudaMemcpyAsync(d_t + ntx, tiy.data(), sizeof(int32_t) * nty, cudaMemcpyHostToDevice, ctx->stream));
                FixedGeom FG{nix, niy, d_t, d_t + ntx};
                B2_CUDA(ctx, cudaFuncSetAttribute(fv.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
                const dim3 grid(ntx, nty, n_frames);
                B2_KERNEL(ctx, "resample_fixed_kernel");
                fv.fn<<<grid, fv.nt, fsmem, ctx->stream>>>(in, out, P, FG);
                B2_LAUNCH_CHECK(ctx);
                return B200TIMG_OK;
            }
        }
        // tile shape: largest of a few candidates whose decoded window + intermediate fit in shared memory
        static const int cand[][2] = {{64, 16}, {32, 16}, {32, 8}, {16, 8}, {8, 4}};
        TileGeom G{0, 0, 0, 0};
        size_t smem = 0;
        size_t smem_budget = 74 * 1024;      // 3 CTAs/SM
        if (const char *e = getenv("B200TIMG_TILE_SMEM_KB")) smem_budget = (size_t)atoi(e) * 1024;   // tuning knob
        for (const auto &c : cand) {
            int nix = 1, niy = 1;
            for (int x0 = 0; x0 < ow; x0 += c[0]) {
                int lo = 0x7fffffff, hi = -1;
                for (int x = x0; x < std::min(ow, x0 + c[0]); ++x) { lo = std::min(lo, pl->h.first[x]); hi = std::max(hi, pl->h.first[x] + pl->h.count[x] - 1); }
                nix = std::max(nix, hi - lo + 1);
            }
            for (int y0 = 0; y0 < oh; y0 += c[1]) {
                int lo = 0x7fffffff, hi = -1;
                for (int y = y0; y < std::min(oh, y0 + c[1]); ++y) { lo = std::min(lo, pl->v.first[y]); hi = std::max(hi, pl->v.first[y] + pl->v.count[y] - 1); }
                niy = std::max(niy, hi - lo + 1);
            }
            const size_t need = sizeof(float4) * ((size_t)nix * niy + (pl->vertical_first ? (size_t)c[1] * nix : (size_t)niy * c[0]))
                              + sizeof(int) * (2 * (size_t)c[0] + 2 * (size_t)c[1])
                              + sizeof(float) * ((size_t)c[0] * (pl->h.widest | 1) + (size_t)c[1] * (pl->v.widest | 1));
            if (need <= smem_budget || (&c == &cand[4] && need <= 200 * 1024)) { G = TileGeom{c[0], c[1], nix, niy}; smem = need; break; }
        }
        if (G.tw) {
            TiledFn fn = pl->vertical_first ? pick_h<true>(tap_class(pl->h.widest), tap_class(pl->v.widest))
                                            : pick_h<false>(tap_class(pl->h.widest), tap_class(pl->v.widest));
            B2_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            const dim3 grid((ow + G.tw - 1) / G.tw, (oh + G.th - 1) / G.th, n_frames);
            B2_KERNEL(ctx, "resample_tiled_kernel");
            fn<<<grid, RT, smem, ctx->stream>>>(in, out, P, G);
        } else {
            // extreme ratios whose window does not fit in shared memory: per-pixel kernel
            const dim3 grid((ow + 31) / 32, (oh + 7) / 8, n_frames);
            B2_KERNEL(ctx, "resample_direct_kernel");
            if (pl->vertical_first) resample_direct_kernel<true><<<grid, 256, 0, ctx->stream>>>(in, out, P);
            else resample_direct_kernel<false><<<grid, 256, 0, ctx->stream>>>(in, out, P);
        }
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
