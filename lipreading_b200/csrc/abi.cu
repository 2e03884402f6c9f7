// abi.cu — library-level entry points of liblr_b200 (error text, launch accounting).
#include "common.cuh"
#include <string.h>

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void lr_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void lr_count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

extern "C" {
int lr_abi_version(void) { return 1; }
const char* lr_last_error(void) { return g_err; }
uint64_t lr_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
}

// ---- lr_scale_rows: out[b,:] = in[b,:] * scale[b] ---------------------------------------
__global__ void scale_rows_kernel(const float* __restrict__ in, const float* __restrict__ scale,
                                  float* __restrict__ out, int B, int64_t row_elems) {
  int b = blockIdx.y;
  float s = scale[b];
  const float* src = in + (int64_t)b * row_elems;
  float* dst = out + (int64_t)b * row_elems;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < row_elems;
       i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src[i] * s;
}

extern "C" int lr_scale_rows(const float* in, const float* scale, float* out, int B,
                             int64_t row_elems, void* stream) {
  LR_CHECK_ARG(in && scale && out && B > 0 && row_elems > 0, "lr_scale_rows: bad arguments");
  dim3 grid(lr_div_up(row_elems, 256 * 4) > 64 ? 64 : lr_div_up(row_elems, 256 * 4), B);
  scale_rows_kernel<<<grid, 256, 0, lr_stream(stream)>>>(in, scale, out, B, row_elems);
  LR_CHECK_LAUNCH();
  return LR_OK;
}
