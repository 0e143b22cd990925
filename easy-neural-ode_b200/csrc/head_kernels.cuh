// The rest of mnist.py's update() around the ODE block, as three small kernels instead of ~40 framework
// launches (they were 8 % of the device time of a train step):
//   PostODE  (mnist.py:225-238): logits = sigmoid(y1) @ W + b
//   _loss_fn (mnist.py:90-94, 449-452): mean softmax cross-entropy, and its gradients wrt y1, W, b
//   optimizers.momentum (mnist.py:530-540): v <- mass * v + g ; x <- x - lr * v
// Deterministic: every sum has a fixed order (no atomics).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "eno_common.cuh"

namespace eno {

constexpr int kHeadMaxClasses = 32;

// one CTA per sample: s = sigmoid(y1), logits, softmax, loss, dlogits, and the cotangent of y1
__global__ void __launch_bounds__(256) k_head_rows(const float* __restrict__ y1, const int64_t* __restrict__ labels,
                                                   const float* __restrict__ W, const float* __restrict__ b, int D, int C,
                                                   float inv_count, float* __restrict__ s_out, float* __restrict__ dl_out,
                                                   float* __restrict__ loss_rows, float* __restrict__ gy) {
  extern __shared__ float sm[];   // [D] sigmoid(y1 row)
  __shared__ float logits[kHeadMaxClasses], dl[kHeadMaxClasses];
  const int r = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float s = sigmoidf_(y1[(size_t)r * D + d]);
    sm[d] = s;
    s_out[(size_t)r * D + d] = s;
  }
  __syncthreads();
  for (int c = warp; c < C; c += nw) {
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) acc = fmaf(sm[d], W[(size_t)d * C + c], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) logits[c] = acc + b[c];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = logits[0];
    for (int c = 1; c < C; ++c) m = fmaxf(m, logits[c]);
    float z = 0.f;
    for (int c = 0; c < C; ++c) z += expf(logits[c] - m);
    const float lz = logf(z);
    const int y = (int)labels[r];
    loss_rows[r] = -(logits[y] - m - lz);
    for (int c = 0; c < C; ++c) {
      const float p = expf(logits[c] - m - lz);
      const float g = (p - (c == y ? 1.f : 0.f)) * inv_count;
      dl[c] = g;
      dl_out[(size_t)r * C + c] = g;
    }
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(dl[c], W[(size_t)d * C + c], acc);
    const float s = sm[d];
    gy[(size_t)r * D + d] = acc * s * (1.f - s);
  }
}

// gW[d][c] = sum_r s[r][d] dl[r][c];  gb[c] = sum_r dl[r][c];  loss = mean_r loss_rows.
// One warp per output: lanes stride the rows (a serial loop over B rows is B dependent L2 latencies), then a
// fixed-shape shuffle tree - deterministic.
__global__ void __launch_bounds__(256) k_head_wgrad(const float* __restrict__ s, const float* __restrict__ dl,
                                                    const float* __restrict__ loss_rows, int B, int D, int C,
                                                    float* __restrict__ gW, float* __restrict__ gb, float* __restrict__ loss_out) {
  const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  float acc = 0.f;
  if (idx < D * C) {
    const int d = idx / C, c = idx - d * C;
    for (int r = lane; r < B; r += 32) acc = fmaf(s[(size_t)r * D + d], dl[(size_t)r * C + c], acc);
  } else if (idx < D * C + C) {
    const int c = idx - D * C;
    for (int r = lane; r < B; r += 32) acc += dl[(size_t)r * C + c];
  } else if (idx == D * C + C) {
    for (int r = lane; r < B; r += 32) acc += loss_rows[r];
  } else {
    return;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane != 0) return;
  if (idx < D * C) gW[idx] = acc;
  else if (idx < D * C + C) gb[idx - D * C] = acc;
  else *loss_out = acc / (float)B;
}

// momentum SGD over up to three parameter blocks in one launch; the update is applied only when both guard
// words (the solves' `done` flags in deferred mode) are non-zero, so garbage gradients never reach the state
struct MomentumBlocks {
  float* p[3];
  float* v[3];
  const float* g[3];
  long long n[3];
};

__global__ void __launch_bounds__(256) k_momentum(MomentumBlocks m, float lr, float mass, float lam_w, const float* lr_dev,
                                                  const int* ok_a, const int* ok_b) {
  const bool ok = (!ok_a || *ok_a != 0) && (!ok_b || *ok_b != 0);
  if (!ok) return;
  if (lr_dev) lr = *lr_dev;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
#pragma unroll
  for (int k = 0; k < 3; ++k)
    for (long long i = tid; i < m.n[k]; i += nth) {
      const float vn = mass * m.v[k][i] + (m.g[k][i] + lam_w * m.p[k][i]);
      m.v[k][i] = vn;
      m.p[k][i] = m.p[k][i] - lr * vn;
    }
}

}  // namespace eno
