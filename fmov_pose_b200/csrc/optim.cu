// Optimiser tail of the train iteration (exp_runner.py:258-269, 801-816) in two launches:
//
//   fmov_grad_gather   this step's parameter gradients (separate autograd tensors) -> ONE flat fp32 buffer G laid out like
//                      the optimiser state, followed by one "active" flag per parameter group.  G is what NCCL all-reduces
//                      when the ray batch is sharded (flags included: a group is stepped if ANY rank touched it).
//   fmov_adam_step     torch.optim.Adam (betas, eps, no weight decay / amsgrad) over every parameter tensor of every ACTIVE
//                      group: the reference keeps one Adam for the networks and one per pose MLP and only steps the pose
//                      MLPs of the frames it rendered (exp_runner.py:785-816).  Group learning rates, step counters and the
//                      flags live in device memory, so a CUDA-graph replay follows the LR schedule and the active set.
//
// HBM-bound and small (0.8 M network parameters + 21 K per pose MLP): what matters is the launch count — torch's fused Adam
// with per-parameter step tensors is 6-8 launches plus flatten / unflatten copies around the all-reduce.
#include "fmov_common.cuh"
#include "../../include/fmov_b200.h"

namespace fmov {

constexpr int GATHER_MAX = 160;          // tensors per gather launch (by-value table)
constexpr int ADAM_CHUNK = 2048;         // elements per block

struct GatherArgs {
  int n;
  int n_groups;
  const float* src[GATHER_MAX];
  long long dst_off[GATHER_MAX];
  int numel[GATHER_MAX];
  int blk0[GATHER_MAX + 1];              // first block of tensor i (blocks of ADAM_CHUNK elements)
  unsigned long long active_mask;        // bit g = this rank produced gradients for group g (groups >= 64: see flags_in)
  const float* flags_in;                 // optional device array [n_groups] used instead of the mask
  float* G;
  long long n_total;                     // floats in G before the flags
};

__global__ void __launch_bounds__(256) grad_gather_kernel(const __grid_constant__ GatherArgs a) {
  const int b = blockIdx.x;
  if (b == a.blk0[a.n]) {                // the extra last block writes the group flags
    for (int g = threadIdx.x; g < a.n_groups; g += blockDim.x)
      a.G[a.n_total + g] = a.flags_in ? a.flags_in[g] : (g < 64 && ((a.active_mask >> g) & 1ull) ? 1.f : 0.f);
    return;
  }
  int lo = 0, hi = a.n - 1;              // tensor owning block b
  while (lo < hi) { const int m = (lo + hi + 1) >> 1; if (a.blk0[m] <= b) lo = m; else hi = m - 1; }
  const int t = lo;
  const int e0 = (b - a.blk0[t]) * ADAM_CHUNK;
  const int e1 = min(e0 + ADAM_CHUNK, a.numel[t]);
  const float* __restrict__ s = a.src[t];
  float* __restrict__ d = a.G + a.dst_off[t];
  for (int i = e0 + threadIdx.x; i < e1; i += blockDim.x) d[i] = s[i];
}

struct AdamTable {                        // device-resident, built once per optimiser
  float* const* param;                    // [n_tensors] parameter tensors
  const long long* off;                   // [n_tensors] offset into G / M / V
  const int* numel;                       // [n_tensors]
  const int* group;                       // [n_tensors]
  const int* chunk_tensor;                // [n_chunks]
  const int* chunk_e0;                    // [n_chunks]
};

__global__ void __launch_bounds__(256)
adam_step_kernel(AdamTable tb, int n_chunks, int n_groups, const float* __restrict__ G, long long n_total,
                 float* __restrict__ M, float* __restrict__ V, const float* __restrict__ lr, float* __restrict__ step,
                 float beta1, float beta2, float eps, float grad_scale, unsigned int* done) {
  const int c = blockIdx.x;
  const int t = tb.chunk_tensor[c];
  const int g = tb.group[t];
  const float* __restrict__ flags = G + n_total;
  if (flags[g] > 0.f) {
    // torch.optim.Adam: step += 1; bias corrections 1 - beta^step; p -= lr / bc1 * m / (sqrt(v) / sqrt(bc2) + eps)
    const double st = (double)step[g] + 1.0;
    const float bc1 = (float)(1.0 - pow((double)beta1, st));
    const float bc2s = (float)sqrt(1.0 - pow((double)beta2, st));
    const float step_size = lr[g] / bc1;
    const int e0 = tb.chunk_e0[c];
    const int e1 = min(e0 + ADAM_CHUNK, tb.numel[t]);
    const long long off = tb.off[t];
    float* __restrict__ p = tb.param[t];
    for (int i = e0 + threadIdx.x; i < e1; i += blockDim.x) {
      const float gr = G[off + i] * grad_scale;
      const float m = M[off + i] + (gr - M[off + i]) * (1.f - beta1);          // lerp(m, g, 1 - beta1)
      const float v = beta2 * V[off + i] + (1.f - beta2) * gr * gr;
      M[off + i] = m;
      V[off + i] = v;
      p[i] -= step_size * m / (sqrtf(v) / bc2s + eps);
    }
  }
  // the last block to finish advances the step counters of the active groups (every block has read them by then)
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(done, 1u) == (unsigned)gridDim.x - 1u;
  }
  __syncthreads();
  if (last) {
    for (int k = threadIdx.x; k < n_groups; k += blockDim.x)
      if (flags[k] > 0.f) step[k] += 1.f;
    if (threadIdx.x == 0) *done = 0u;
  }
}

}  // namespace fmov
using namespace fmov;

extern "C" int fmov_adam_chunk(void) { return ADAM_CHUNK; }

/* src / dst_off / numel: HOST arrays of n entries (n <= 160): gradient tensor i -> G[dst_off[i] .. + numel[i]).
 * G[n_total + g] = flag of group g: bit g of active_mask, or flags_in[g] (device) when given.  Ranges of G that no tensor
 * covers are left as they are (zero G first when the buffer is all-reduced). */
extern "C" int fmov_grad_gather(int n, const float* const* src, const long long* dst_off, const int* numel, int n_groups,
                                unsigned long long active_mask, const float* flags_in, float* G, long long n_total,
                                void* stream) {
  FMOV_REQUIRE(n >= 0 && n <= GATHER_MAX && n_groups >= 1 && G && n_total >= 0, "fmov_grad_gather: bad arguments (n=%d, max %d)",
               n, GATHER_MAX);
  FMOV_REQUIRE(n == 0 || (src && dst_off && numel), "fmov_grad_gather: null table");
  FMOV_REQUIRE(flags_in || n_groups <= 64, "fmov_grad_gather: more than 64 groups need a device flag array");
  GatherArgs a;
  memset(&a, 0, sizeof(a));
  a.n = n; a.n_groups = n_groups; a.active_mask = active_mask; a.flags_in = flags_in; a.G = G; a.n_total = n_total;
  int blk = 0;
  for (int i = 0; i < n; ++i) {
    FMOV_REQUIRE(src[i] && numel[i] >= 1 && dst_off[i] >= 0 && dst_off[i] + numel[i] <= n_total,
                 "fmov_grad_gather: bad tensor %d", i);
    a.src[i] = src[i]; a.dst_off[i] = dst_off[i]; a.numel[i] = numel[i]; a.blk0[i] = blk;
    blk += (numel[i] + ADAM_CHUNK - 1) / ADAM_CHUNK;
  }
  a.blk0[n] = blk;
  grad_gather_kernel<<<blk + 1, 256, 0, (cudaStream_t)stream>>>(a);
  FMOV_LAUNCH_CHECK("grad_gather_kernel");
  return OK;
}

/* Device tables (built once by the caller): param [n_tensors] device pointers, off / numel / group [n_tensors],
 * chunk_tensor / chunk_e0 [n_chunks] (chunks of fmov_adam_chunk() elements).  G: gradients + n_groups flags at
 * G[n_total..]; M, V: Adam moments laid out like G; lr, step: [n_groups] device floats; done: zero-initialised device
 * counter owned by the optimiser.  grad_scale multiplies every gradient (1 for SUM-reduced shards of one batch). */
extern "C" int fmov_adam_step(float* const* param, const long long* off, const int* numel, const int* group,
                              const int* chunk_tensor, const int* chunk_e0, int n_chunks, int n_groups, const float* G,
                              long long n_total, float* M, float* V, const float* lr, float* step, float beta1, float beta2,
                              float eps, float grad_scale, unsigned int* done, void* stream) {
  FMOV_REQUIRE(param && off && numel && group && chunk_tensor && chunk_e0 && G && M && V && lr && step && done,
               "fmov_adam_step: null argument");
  FMOV_REQUIRE(n_chunks >= 1 && n_groups >= 1 && n_total >= 1, "fmov_adam_step: bad sizes");
  AdamTable tb{param, off, numel, group, chunk_tensor, chunk_e0};
  adam_step_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(tb, n_chunks, n_groups, G, n_total, M, V, lr, step, beta1,
                                                               beta2, eps, grad_scale, done);
  FMOV_LAUNCH_CHECK("adam_step_kernel");
  return OK;
}
