// labelanything_b200 — first piece of the training-step row: the reference's default loss,
// focal loss with label-frequency class weighting, forward value and gradient w.r.t. the logits, each in ONE pass
// over [B, C, H, W] (sm_100a, HBM-bound fp32 work).
//
// Reference (label_anything/loss/__init__.py:67-92, loss/focal.py:8-25, loss/utils.py:17-42):
//     wtarget, class_weights = get_weight_matrix_from_labels(target, C)    # w_c = 1 / log(1.1 + n_c / N), ignore -> 0
//     ce = F.cross_entropy(x, target, reduction="none")                    # 0 where target == -100
//     pt = exp(-ce);  focal = (1 - pt) ** gamma * wtarget * ce;  loss = mean(focal)
// = a unique() over the labels, a gathered [B, H, W] weight map, and ~8 elementwise passes over [B, H, W] plus the
// log-softmax over [B, C, H, W]; autograd keeps all of them alive for the backward pass.  Here:
//   la_label_class_weights : label histogram (warp-aggregated atomics) -> w_c                       (reads 8 B / pixel)
//   la_focal_loss          : per pixel max / log-sum-exp over the C planes, ce, focal term, fp32 per-thread sums ->
//                            fp64 per-CTA partials -> the last CTA adds them in a fixed order (deterministic);
//                            with grad_out: d loss / d logits = s/N * w_t [gamma (1-pt)^(gamma-1) pt log pt
//                            - (1-pt)^gamma] (delta_tc - p_c) written in the same pass      (4C + 8 (+4C) B / pixel)
// A target outside [0, C) that is not ignore_index poisons the loss with NaN (torch raises a device assert there).
#include "la_common.cuh"

namespace la {

constexpr int LOSS_THREADS = 256;
constexpr int LOSS_MAX_CTAS = 148 * 8;
constexpr int LOSS_NREG = 8;   // logit planes kept in registers; further planes are re-read (L1 / L2 hits)

struct FocalParams {
  const float* logits;        // [B, C, P]
  const long long* target;    // [B, P]
  const float* class_w;       // [C] or nullptr (all ones)
  const float* grad_scale;    // device scalar (upstream gradient) or nullptr (1)
  float* loss_out;            // [1] or nullptr
  float* grad_out;            // [B, C, P] or nullptr
  float* wtarget_out;         // [B, P] or nullptr
  double* partials;           // workspace: [LOSS_MAX_CTAS] partial sums
  unsigned int* counter;      // workspace: CTAs done (self-resetting)
  long long P;
  int B, C;
  float gamma;
  long long ignore_index;
  int mean;
};

__device__ __forceinline__ float pow_gamma(float base, float gamma) {
  if (gamma == 2.0f) return base * base;   // torch.pow(x, 2) is x * x
  if (gamma == 1.0f) return base;
  if (gamma == 0.0f) return 1.0f;
  return powf(base, gamma);
}

template <int VEC, bool GRAD>
__global__ void __launch_bounds__(LOSS_THREADS) focal_loss_kernel(const FocalParams p) {
  const long long groups_per_item = p.P / VEC;
  const long long total = groups_per_item * p.B;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const float inv_n = p.mean ? 1.0f / (static_cast<float>(p.B) * static_cast<float>(p.P)) : 1.0f;
  float gscale = inv_n;
  if constexpr (GRAD) {
    if (p.grad_scale != nullptr) gscale *= __ldg(p.grad_scale);
  }
  float acc = 0.0f;
  for (long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; q < total; q += stride) {
    const int b = static_cast<int>(q / groups_per_item);
    const long long px = (q - static_cast<long long>(b) * groups_per_item) * VEC;
    const float* src = p.logits + (static_cast<long long>(b) * p.C) * p.P + px;
    float xr[LOSS_NREG][VEC];
#pragma unroll
    for (int c = 0; c < LOSS_NREG; ++c) {
      if (c < p.C && p.logits != nullptr) {
        if constexpr (VEC == 4) {
          const float4 v = *reinterpret_cast<const float4*>(src + static_cast<long long>(c) * p.P);
          xr[c][0] = v.x; xr[c][1] = v.y; xr[c][2] = v.z; xr[c][3] = v.w;
        } else {
          xr[c][0] = src[static_cast<long long>(c) * p.P];
        }
      } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) xr[c][i] = -INFINITY;
      }
    }
    float m[VEC], s[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      m[i] = xr[0][i];
#pragma unroll
      for (int c = 1; c < LOSS_NREG; ++c) m[i] = fmaxf(m[i], xr[c][i]);
    }
    const int c_tail = p.logits != nullptr ? p.C : 0;   // weight-map-only calls carry no logits
    for (int c = LOSS_NREG; c < c_tail; ++c) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) m[i] = fmaxf(m[i], src[static_cast<long long>(c) * p.P + i]);
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      s[i] = 0.0f;
#pragma unroll
      for (int c = 0; c < LOSS_NREG; ++c)
        if (c < p.C) s[i] += __expf(xr[c][i] - m[i]);
    }
    for (int c = LOSS_NREG; c < c_tail; ++c) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) s[i] += __expf(src[static_cast<long long>(c) * p.P + i] - m[i]);
    }
    long long tg[VEC];
    if constexpr (VEC == 4) {
      const longlong2 a = *reinterpret_cast<const longlong2*>(p.target + static_cast<long long>(b) * p.P + px);
      const longlong2 c2 = *(reinterpret_cast<const longlong2*>(p.target + static_cast<long long>(b) * p.P + px) + 1);
      tg[0] = a.x; tg[1] = a.y; tg[2] = c2.x; tg[3] = c2.y;
    } else {
      tg[0] = p.target[static_cast<long long>(b) * p.P + px];
    }
    float coef[VEC], lse[VEC], wt[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      lse[i] = m[i] + logf(s[i]);
      const bool ignored = tg[i] == p.ignore_index;
      const bool in_range = tg[i] >= 0 && tg[i] < p.C;
      float f = 0.0f;
      coef[i] = 0.0f;
      wt[i] = 0.0f;
      if (!ignored && in_range) {
        const int t = static_cast<int>(tg[i]);
        const float xt = p.logits != nullptr ? src[static_cast<long long>(t) * p.P + i] : 0.0f;
        const float w = p.class_w ? __ldg(p.class_w + t) : 1.0f;
        const float ce = lse[i] - xt;
        const float pt = expf(-ce);
        const float om = 1.0f - pt;
        f = pow_gamma(om, p.gamma) * w * ce;
        wt[i] = w;
        if constexpr (GRAD) {
          // d/dx_c [(1-pt)^g w ce] = w [g (1-pt)^(g-1) pt log pt - (1-pt)^g] (delta_tc - p_c),  log pt = -ce
          const float dpow = p.gamma == 0.0f ? 0.0f : p.gamma * pow_gamma(om, p.gamma - 1.0f);
          coef[i] = w * (dpow * pt * (-ce) - pow_gamma(om, p.gamma)) * gscale;
        }
      } else if (!ignored) {
        f = __int_as_float(0x7fc00000);   // label outside [0, C): poison
      }
      acc += f;
    }
    if (p.wtarget_out != nullptr) {
      if constexpr (VEC == 4)
        *reinterpret_cast<float4*>(p.wtarget_out + static_cast<long long>(b) * p.P + px) = make_float4(wt[0], wt[1], wt[2], wt[3]);
      else
        p.wtarget_out[static_cast<long long>(b) * p.P + px] = wt[0];
    }
    if constexpr (GRAD) {
      float* dst = p.grad_out + (static_cast<long long>(b) * p.C) * p.P + px;
      for (int c = 0; c < p.C; ++c) {
        float g[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          float x;
          if (c < LOSS_NREG) {
            x = xr[0][i];
#pragma unroll
            for (int k = 1; k < LOSS_NREG; ++k) x = (c == k) ? xr[k][i] : x;
          } else {
            x = src[static_cast<long long>(c) * p.P + i];
          }
          const float pc = expf(x - lse[i]);
          g[i] = coef[i] * ((tg[i] == c ? 1.0f : 0.0f) - pc);
          if (coef[i] == 0.0f) g[i] = 0.0f;   // ignored pixels: exactly zero even for non-finite logits
        }
        if constexpr (VEC == 4)
          __stcs(reinterpret_cast<float4*>(dst + static_cast<long long>(c) * p.P), make_float4(g[0], g[1], g[2], g[3]));
        else
          dst[static_cast<long long>(c) * p.P] = g[0];
      }
    }
  }
  if (p.loss_out == nullptr) return;
  // deterministic reduction: warp shuffle -> per-CTA fp64 partial -> the last CTA sums the partials in order
  __shared__ double s_part[LOSS_THREADS / 32];
  __shared__ bool s_last;
  double d = static_cast<double>(acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < LOSS_THREADS / 32; ++w) t += s_part[w];
    p.partials[blockIdx.x] = t;
    __threadfence();
    s_last = atomicAdd(p.counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last) {
    // every thread adds a fixed strided subset, then a fixed-order tree: same result on every run
    __shared__ double s_fin[LOSS_THREADS];
    __threadfence();
    double t = 0.0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += LOSS_THREADS) t += *(volatile double*)(p.partials + i);
    s_fin[threadIdx.x] = t;
    __syncthreads();
    for (int o = LOSS_THREADS / 2; o > 0; o >>= 1) {
      if (threadIdx.x < o) s_fin[threadIdx.x] += s_fin[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      *p.loss_out = static_cast<float>(s_fin[0] * static_cast<double>(inv_n));
      *p.counter = 0u;
    }
  }
}

// hist[c] = #labels == c (c < C), hist[C] = #labels == ignore_index, hist[C + 1] = everything else
__global__ void __launch_bounds__(256) label_histogram_kernel(const long long* __restrict__ labels, long long n, int C,
                                                              long long ignore_index, unsigned long long* hist) {
  extern __shared__ unsigned int s_h[];
  for (int i = threadIdx.x; i < C + 2; i += blockDim.x) s_h[i] = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long i0 = warp_global * 32; i0 < n; i0 += n_warps * 32) {
    const long long i = i0 + lane;
    int key = -1;
    if (i < n) {
      const long long v = __ldcs(labels + i);
      key = v == ignore_index ? C : ((v >= 0 && v < C) ? static_cast<int>(v) : C + 1);
    }
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (key >= 0 && lane == __ffs(peers) - 1) atomicAdd(&s_h[key], static_cast<unsigned int>(__popc(peers)));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C + 2; i += blockDim.x)
    if (s_h[i]) atomicAdd(hist + i, static_cast<unsigned long long>(s_h[i]));
}

// loss/utils.py:17-42: w_c = 1 / log(1.1 + n_c / N) for the classes that occur (N = all labels, ignored ones
// included), 1 for the others
__global__ void class_weights_kernel(const unsigned long long* hist, int C, float* class_w) {
  double total = 0.0;
  for (int i = 0; i < C + 2; ++i) total += static_cast<double>(hist[i]);
  const float n_all = static_cast<float>(static_cast<long long>(total));
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const unsigned long long n = hist[c];
    class_w[c] = n ? 1.0f / logf(1.1f + static_cast<float>(static_cast<long long>(n)) / n_all) : 1.0f;
  }
}

}  // namespace la

extern "C" long long la_focal_loss_workspace_bytes(void) {
  return static_cast<long long>(la::LOSS_MAX_CTAS) * sizeof(double) + 16;
}

extern "C" int la_focal_loss(void* stream, const float* logits, const long long* target, const float* class_w,
                             const float* grad_scale, float* loss_out, float* grad_out, float* wtarget_out,
                             void* workspace, int batch, int classes, long long pixels, float gamma,
                             long long ignore_index, int mean) {
  using namespace la;
  LA_CHECK_ARG(target != nullptr, "la_focal_loss: null pointer");
  LA_CHECK_ARG(logits != nullptr || (loss_out == nullptr && grad_out == nullptr),
               "la_focal_loss: the loss and its gradient need the logits");
  LA_CHECK_ARG(batch > 0 && classes > 0 && pixels > 0, "la_focal_loss: empty problem");
  LA_CHECK_ARG(loss_out || grad_out || wtarget_out, "la_focal_loss: nothing to compute");
  LA_CHECK_ARG(loss_out == nullptr || workspace != nullptr, "la_focal_loss: the loss value needs the workspace");
  LA_CHECK_ARG(workspace == nullptr || (reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "la_focal_loss: workspace must be 8-byte aligned");
  LA_CHECK_ARG(gamma >= 0.0f, "la_focal_loss: gamma must be non-negative");
  FocalParams p;
  p.logits = logits;
  p.target = target;
  p.class_w = class_w;
  p.grad_scale = grad_scale;
  p.loss_out = loss_out;
  p.grad_out = grad_out;
  p.wtarget_out = wtarget_out;
  p.partials = static_cast<double*>(workspace);
  p.counter = workspace ? reinterpret_cast<unsigned int*>(static_cast<double*>(workspace) + LOSS_MAX_CTAS) : nullptr;
  p.P = pixels;
  p.B = batch;
  p.C = classes;
  p.gamma = gamma;
  p.ignore_index = ignore_index;
  p.mean = mean;
  auto al = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const bool vec4 = pixels % 4 == 0 && al(logits) && al(target) && al(grad_out) && al(wtarget_out);
  const long long groups = static_cast<long long>(batch) * (vec4 ? pixels / 4 : pixels);
  long long ctas = (groups + LOSS_THREADS - 1) / LOSS_THREADS;
  // one resident wave: 4 CTAs per SM at 64 registers (value), 2 at 118 (gradient)
  long long cap = static_cast<long long>(sm_count()) * (grad_out ? 2 : 4);
  if (cap > LOSS_MAX_CTAS) cap = LOSS_MAX_CTAS;
  if (ctas > cap) ctas = cap;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned g = static_cast<unsigned>(ctas);
  // the "CTAs done" counter starts at zero on every launch: the workspace needs no initialisation by the caller and a
  // launch that was aborted (or raced on the same scratch) cannot leave it in a state where the last-CTA branch never fires
  if (p.counter != nullptr) LA_CHECK_CUDA(cudaMemsetAsync(p.counter, 0, sizeof(unsigned int), st));
  if (vec4) {
    if (grad_out) focal_loss_kernel<4, true><<<g, LOSS_THREADS, 0, st>>>(p);
    else focal_loss_kernel<4, false><<<g, LOSS_THREADS, 0, st>>>(p);
  } else {
    if (grad_out) focal_loss_kernel<1, true><<<g, LOSS_THREADS, 0, st>>>(p);
    else focal_loss_kernel<1, false><<<g, LOSS_THREADS, 0, st>>>(p);
  }
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

extern "C" int la_label_class_weights(void* stream, const long long* labels, long long n, int classes,
                                      long long ignore_index, long long* hist, float* class_w) {
  using namespace la;
  LA_CHECK_ARG(labels && hist && class_w, "la_label_class_weights: null pointer");
  LA_CHECK_ARG(n > 0 && classes > 0 && classes <= 8192, "la_label_class_weights: need n > 0 and 0 < classes <= 8192");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LA_CHECK_CUDA(cudaMemsetAsync(hist, 0, sizeof(long long) * (classes + 2), st));
  long long ctas = (n + 255) / 256;
  if (ctas > static_cast<long long>(sm_count()) * 8) ctas = static_cast<long long>(sm_count()) * 8;
  label_histogram_kernel<<<static_cast<unsigned>(ctas), 256, sizeof(unsigned int) * (classes + 2), st>>>(
      labels, n, classes, ignore_index, reinterpret_cast<unsigned long long*>(hist));
  LA_CHECK_CUDA(cudaGetLastError());
  class_weights_kernel<<<1, 128, 0, st>>>(reinterpret_cast<const unsigned long long*>(hist), classes, class_w);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}
