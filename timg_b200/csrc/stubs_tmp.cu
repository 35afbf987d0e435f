#include "common.cuh"
namespace b200timg {
int launch_scale(b200timg_ctx *ctx, const uint8_t *, int, int, int, uint8_t *, int, int, int, int) { return ctx->fail(B200TIMG_EINVAL, "scale: not built yet"); }
int launch_sixel(b200timg_ctx *ctx, const uint8_t *, int, int, int, char *, size_t, uint64_t *) { return ctx->fail(B200TIMG_EINVAL, "sixel: not built yet"); }
}
