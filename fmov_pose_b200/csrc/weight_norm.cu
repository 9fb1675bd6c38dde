// Weight normalisation of every linear layer of both MLPs in ONE launch per direction.
//
// The reference wraps each nn.Linear in nn.utils.weight_norm (dim = 0; models/fields.py:81-82, 160-161): the effective
// weight is W = g * v / ||v|| per output row, recomputed every iteration, and autograd carries dL/dW back to g and v.
// In torch that is one kernel per layer and direction (14 + 14 launches per step plus their temporaries).  Here a warp
// owns one output row of one layer: forward W = v * (g / ||v||), backward (dot = <dW, v>)
//     dg = dot / ||v||,      dv = (g / ||v||) * dW - (g * dot / ||v||^3) * v.
// HBM-bound and tiny (802 K parameters): what matters is the launch count.
#include "fmov_common.cuh"
#include "../../include/fmov_b200.h"

namespace fmov {

constexpr int WN_MAX_LAYERS = 16;
constexpr int WN_WARPS = 8;

struct WnLayer {
  const float* v;       // [rows, cols]
  const float* g;       // [rows] (weight_g is [rows,1])
  float* W;             // fwd: out [rows, cols]
  float* norm;          // fwd: out [rows]; bwd: in
  const float* dW;      // bwd: in [rows, cols]
  float* dv;            // bwd: out [rows, cols]
  float* dg;            // bwd: out [rows]
  int rows, cols;
  int row0;             // first global row index of this layer
};
struct WnArgs {
  int n_layers, total_rows;
  WnLayer layer[WN_MAX_LAYERS];
};

__device__ __forceinline__ float wn_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <bool BWD>
__global__ void __launch_bounds__(WN_WARPS * 32) weight_norm_kernel(const __grid_constant__ WnArgs a) {
  const int lane = threadIdx.x & 31;
  const int grow = blockIdx.x * WN_WARPS + (threadIdx.x >> 5);
  if (grow >= a.total_rows) return;
  int li = 0;
  while (li + 1 < a.n_layers && grow >= a.layer[li + 1].row0) ++li;
  const WnLayer& L = a.layer[li];
  const int r = grow - L.row0;
  const float* __restrict__ v = L.v + (size_t)r * L.cols;
  if (!BWD) {
    float ss = 0.f;
    for (int i = lane; i < L.cols; i += 32) { const float x = v[i]; ss = fmaf(x, x, ss); }
    const float nrm = sqrtf(wn_sum(ss));
    const float sc = L.g[r] / nrm;
    float* __restrict__ w = L.W + (size_t)r * L.cols;
    for (int i = lane; i < L.cols; i += 32) w[i] = v[i] * sc;
    if (lane == 0) L.norm[r] = nrm;
  } else {
    const float* __restrict__ dW = L.dW + (size_t)r * L.cols;
    float dot = 0.f;
    for (int i = lane; i < L.cols; i += 32) dot = fmaf(dW[i], v[i], dot);
    dot = wn_sum(dot);
    const float nrm = L.norm[r], g = L.g[r];
    const float c1 = g / nrm, c2 = g * dot / (nrm * nrm * nrm);
    float* __restrict__ dv = L.dv + (size_t)r * L.cols;
    for (int i = lane; i < L.cols; i += 32) dv[i] = c1 * dW[i] - c2 * v[i];
    if (lane == 0) L.dg[r] = dot / nrm;
  }
}

}  // namespace fmov
using namespace fmov;

static int fill_wn(WnArgs& a, int n_layers, const float* const* v, const float* const* g, const int* rows, const int* cols) {
  FMOV_REQUIRE(n_layers >= 1 && n_layers <= WN_MAX_LAYERS, "weight_norm: 1..%d layers (got %d)", WN_MAX_LAYERS, n_layers);
  FMOV_REQUIRE(v && g && rows && cols, "weight_norm: null argument");
  memset(&a, 0, sizeof(a));
  a.n_layers = n_layers;
  int r0 = 0;
  for (int i = 0; i < n_layers; ++i) {
    FMOV_REQUIRE(v[i] && g[i] && rows[i] >= 1 && cols[i] >= 1, "weight_norm: bad layer %d", i);
    a.layer[i].v = v[i]; a.layer[i].g = g[i]; a.layer[i].rows = rows[i]; a.layer[i].cols = cols[i]; a.layer[i].row0 = r0;
    r0 += rows[i];
  }
  a.total_rows = r0;
  return OK;
}

/* all pointer arguments are HOST arrays of n_layers device pointers; layer i: v [rows[i], cols[i]], g [rows[i]] */
extern "C" int fmov_weight_norm_fwd(int n_layers, const float* const* v, const float* const* g, const int* rows,
                                    const int* cols, float* const* W, float* const* norm, void* stream) {
  WnArgs a;
  int st = fill_wn(a, n_layers, v, g, rows, cols);
  if (st) return st;
  FMOV_REQUIRE(W && norm, "fmov_weight_norm_fwd: null output");
  for (int i = 0; i < n_layers; ++i) {
    FMOV_REQUIRE(W[i] && norm[i], "fmov_weight_norm_fwd: null output for layer %d", i);
    a.layer[i].W = W[i]; a.layer[i].norm = norm[i];
  }
  weight_norm_kernel<false><<<(a.total_rows + WN_WARPS - 1) / WN_WARPS, WN_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
  FMOV_LAUNCH_CHECK("weight_norm_kernel<fwd>");
  return OK;
}

extern "C" int fmov_weight_norm_bwd(int n_layers, const float* const* v, const float* const* g, const int* rows,
                                    const int* cols, const float* const* norm, const float* const* dW, float* const* dv,
                                    float* const* dg, void* stream) {
  WnArgs a;
  int st = fill_wn(a, n_layers, v, g, rows, cols);
  if (st) return st;
  FMOV_REQUIRE(norm && dW && dv && dg, "fmov_weight_norm_bwd: null argument");
  for (int i = 0; i < n_layers; ++i) {
    FMOV_REQUIRE(norm[i] && dW[i] && dv[i] && dg[i], "fmov_weight_norm_bwd: null pointer for layer %d", i);
    a.layer[i].norm = const_cast<float*>(norm[i]); a.layer[i].dW = dW[i]; a.layer[i].dv = dv[i]; a.layer[i].dg = dg[i];
  }
  weight_norm_kernel<true><<<(a.total_rows + WN_WARPS - 1) / WN_WARPS, WN_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
  FMOV_LAUNCH_CHECK("weight_norm_kernel<bwd>");
  return OK;
}
