// Shared declarations of the sixel kernels (sixel.cu: palette / LUT / dither, sixel_emit.cu: byte emit).
#pragma once
#include "common.cuh"

namespace b200timg {

constexpr int SIXEL_HDR_CAP = 4672;    // DCS q + raster attributes (<= 29 bytes) + 256 palette definitions of <= 18 bytes

struct SixelFrameHdr {
    uint32_t ncolors, origcolors, diffuse, header_len;
    uint32_t frame_size, pad0, pad1, pad2;
    uint32_t palette[256];             // r | g << 8 | b << 16
};

struct SixelWork {                     // device pointers into ctx->sixel_work
    SixelFrameHdr *hdr;                // [n_frames]
    uint32_t *ent_a, *ent_b;           // [n_frames][ent_cap] median-cut tables (bucket << 16 | count)
    uint8_t *lut;                      // [n_frames][32768]
    uint8_t *index;                    // [n_frames][w*h]
    uint32_t *boundary;                // [n_frames][nb32][w] packed errors of each 32-row band's last row
    uint32_t *band_bytes;              // [n_frames][nbands]  sizes
    uint32_t *band_off;                // [n_frames][nbands]  offset of each band's first byte inside its frame
    char *scratch;                     // [n_frames][nbands][band_cap] band bytes before compaction
    size_t band_cap;
    int ent_cap, nb32, nbands;
    // single-pass emit (sixel_emit.cu)
    char *hdr_bytes;                   // [n_frames][SIXEL_HDR_CAP] header + palette definitions of each frame
    unsigned long long *desc;          // [n_frames * nbands * ntiles] look-back descriptors: flag << 62 | bytes
    uint32_t *ctl;                     // [0] CTA ticket, [1] status (bit 0: output buffer too small)
};

__device__ __forceinline__ uint32_t hash15(uint32_t px) {   // (r>>3)<<10 | (g>>3)<<5 | (b>>3)
    return ((px & 0xf8) << 7) | ((px >> 6) & 0x3e0) | ((px >> 19) & 0x1f);
}
__device__ __forceinline__ uint32_t key5(uint32_t entry, int plane) { return (entry >> (26 - 5 * plane)) & 31; }

// ------------------------------------------------------------------ block helpers (1024 thr)
template <int NT>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *s_w /*[NT/32]*/, uint32_t &total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    uint32_t pre = 0, tot = 0;
    for (int k = 0; k < NT / 32; ++k) { if (k < wid) pre += s_w[k]; tot += s_w[k]; }
    total = tot;
    __syncthreads();
    return pre + inc - v;
}


// ---- number formatting shared by the emitters
__device__ __forceinline__ uint32_t ndig_u(uint32_t v) { uint32_t n = 1; while (v >= 10) { v /= 10; ++n; } return n; }
__device__ __forceinline__ char *put_num_u(char *o, uint32_t v) {
    char tmp[10]; int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *o++ = tmp[--n];
    return o;
}

// The <=6 distinct (colour, bits) pairs of column x of a 6-row band: slot i is valid iff row i is the
// first row showing its colour (fixed slots, so everything stays in registers).  Returns the valid mask.
__device__ __forceinline__ uint32_t column_entries(const uint8_t *__restrict__ idx, int w, int x, uint32_t *c, uint32_t *bits) {
#pragma unroll
    for (int i = 0; i < 6; ++i) c[i] = idx[(long long)i * w + x];
    uint32_t valid = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        bool seen = false;
        uint32_t b = 0;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            if (j < i && c[j] == c[i]) seen = true;
            if (j >= i && c[j] == c[i]) b |= 1u << j;
        }
        bits[i] = b;
        if (!seen) valid |= 1u << i;
    }
    return valid;
}

int launch_sixel_emit(b200timg_ctx *ctx, int w, int h, int n_frames, const SixelWork &W, char *d_out, size_t out_cap,
                      uint64_t *d_offsets);
int launch_sixel_emit3(b200timg_ctx *ctx, int w, int h, int n_frames, const SixelWork &W, char *d_out, size_t out_cap,
                       uint64_t *d_offsets);
size_t sixel_dither_workspace(int w, int h, int n_frames, size_t *o_bnd, size_t *o_prog);
int launch_sixel_dither(b200timg_ctx *ctx, const uint32_t *fb, int w, int h, int n_frames, int n_total, const SixelWork &W, void *d_bnd, void *d_prog);
size_t sixel_emit_workspace(int w, int h, int n_frames, size_t *o_hdr_bytes, size_t *o_desc, size_t *o_ctl);

}  // namespace b200timg
