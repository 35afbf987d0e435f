#include "common.cuh"
namespace b200timg {
int launch_sixel(b200timg_ctx *ctx, const uint8_t *, int, int, int, char *, size_t, uint64_t *) { return ctx->fail(B200TIMG_EINVAL, "sixel: not built yet"); }
}
