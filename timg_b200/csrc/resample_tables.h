// Host-side construction of the separable resampling tables the K1 kernel consumes.
//
// The reference's ImageScaler (STB build, src/image-scaler.cc:75-97) delegates to
// third_party/stb/stb_image_resize2.h with default filters and edge clamp.  To be a
// drop-in we must reproduce that library's results to the bit, which are fully
// determined by (a) the per-axis contributor ranges and coefficients, (b) the order in
// which taps are accumulated and (c) which axis is filtered first.  This module restates
// (a) and (c) and exports what the kernel needs for (b).  Line numbers cite
// third_party/stb/stb_image_resize2.h.
#pragma once
#include <cstdint>
#include <vector>

namespace b200timg {

enum class AxisFilter : int { kPoint = 0, kBox = 1, kMitchell = 2 };

struct AxisTable {
    int in_size = 0, out_size = 0;
    float scale = 1.f, inv_scale = 1.f;
    AxisFilter filter = AxisFilter::kPoint;
    int gather_mode = 1;            // 1 enlarging gather, 2 shrinking gather, 0 vertical scatter
    int filter_pixel_width = 1;
    int widest = 1;                 // max taps over all outputs, after edge folding
    std::vector<int32_t> first;     // [out] first contributing input index (>= 0)
    std::vector<int32_t> count;     // [out] number of taps
    std::vector<int32_t> lead;      // [out] leading zero taps the reference's pack step adds (horizontal)
    std::vector<float> coeff;       // [out][widest], zero padded
};

struct ResamplePlan {
    AxisTable h, v;
    bool copy_only = false;         // both axes point-sampled: no arithmetic at all (:6938)
    bool vertical_first = false;    // :6859-6905
    bool h_sequential = false;      // horizontal taps use one accumulator (widest <= 3, :5801-5868)
};

// Returns false if the geometry is degenerate.
bool build_resample_plan(int in_w, int in_h, int out_w, int out_h, ResamplePlan *plan);

}  // namespace b200timg
