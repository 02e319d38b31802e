// labelanything_b200 — the step right after the hot path in the reference's train / validation loops
// (SURVEY.md §8 row f4): class prediction, episode-local -> dataset-global label mapping and the confusion-matrix
// update of the mIoU metrics, in ONE pass over the logits (sm_100a, HBM-bound integer work).
//
// Reference sequence (label_anything/experiment/run.py:520-541,696-704):
//     preds = logits.argmax(dim=1)                                          # [B, H, W] int64
//     glob_preds, glob_gt = to_global_multiclass(classes, categories, preds, gt)   # data/utils.py:567-590
//     metric.update(glob_preds, glob_gt)      # torchmetrics MulticlassJaccardIndex: confmat[target, pred] += 1,
//                                             # targets equal to ignore_index (-100) dropped (utils/metrics.py:28-42)
// i.e. one read of the logits, 2 (C-1) read-modify-write passes over two int64 label maps and a bincount.  Here every
// pixel is read once: C fp32 logits + one int64 ground-truth label in, (optionally) two int64 labels out, and the
// confusion matrix is accumulated in a per-CTA shared-memory histogram with warp-aggregated atomics (segmentation
// maps are piecewise constant, so most warps hit one bin), flushed with 64-bit global atomics — integer arithmetic,
// bit-exact and order-independent.
//
// to_global_multiclass applies its `torch.where(t == j + 1, value_j, t)` substitutions SEQUENTIALLY, so a label that
// was already mapped can be mapped again by a later step.  The host composes the steps into one table per episode
// (labelanything_b200/metrics.py::chain_label_map); the kernel applies label_map[b][v] to 0 <= v < map_len and leaves
// every other value (e.g. -100) untouched, which reproduces the reference exactly.
#include "la_common.cuh"

namespace la {

struct LabelParams {
  const float* logits;           // [B, C, P] fp32 or nullptr
  const long long* preds_in;     // [B, P] episode-local labels when logits == nullptr (may be nullptr too)
  const long long* gt;           // [B, P] or nullptr
  const long long* label_map;    // [B, map_len] or nullptr (identity)
  long long* preds_out;          // [B, P] or nullptr
  long long* gt_out;             // [B, P] or nullptr
  unsigned long long* confmat;   // [G, G] (target-major) accumulated, or nullptr
  unsigned long long* invalid;   // [1]: pixels whose (target, pred) fell outside [0, G) (not counted)
  long long P;                   // pixels per item
  int B, C, map_len, G;
  long long ignore_index;
  int hist_in_smem;
};

__device__ __forceinline__ long long map_label(const LabelParams& p, int b, long long v) {
  if (p.label_map != nullptr && v >= 0 && v < p.map_len) return __ldg(p.label_map + static_cast<long long>(b) * p.map_len + v);
  return v;
}

// torch.argmax order: NaN beats everything, the first maximal value wins
__device__ __forceinline__ void argmax_step(float v, int c, float& best, int& arg) {
  if (v > best || (v != v && best == best)) {
    best = v;
    arg = c;
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) label_confusion_kernel(const LabelParams p) {
  extern __shared__ unsigned int s_hist[];
  const int lane = threadIdx.x & 31;
  const int GG = p.G * p.G;
  if (p.confmat != nullptr && p.hist_in_smem) {
    for (int i = threadIdx.x; i < GG; i += blockDim.x) s_hist[i] = 0u;
    __syncthreads();
  }
  const long long groups_per_item = p.P / VEC;            // VEC == 4 requires P % 4 == 0
  const long long total = groups_per_item * p.B;
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  unsigned int bad = 0;

  for (long long q0 = warp_global * 32; q0 < total; q0 += n_warps * 32) {   // warp-uniform trip count
    const long long q = q0 + lane;
    const bool live = q < total;
    const int b = live ? static_cast<int>(q / groups_per_item) : 0;
    const long long px = live ? (q - static_cast<long long>(b) * groups_per_item) * VEC : 0;
    const long long off = static_cast<long long>(b) * p.P + px;

    long long pred[VEC];
    if (p.logits != nullptr) {
      float best[VEC];
      int arg[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) { best[i] = -INFINITY; arg[i] = 0; }
      const float* src = p.logits + (static_cast<long long>(b) * p.C) * p.P + px;
      if (live) {
#pragma unroll 4
        for (int c = 0; c < p.C; ++c) {
          if constexpr (VEC == 4) {
            const float4 v = __ldcs(reinterpret_cast<const float4*>(src + static_cast<long long>(c) * p.P));
            argmax_step(v.x, c, best[0], arg[0]);
            argmax_step(v.y, c, best[1], arg[1]);
            argmax_step(v.z, c, best[2], arg[2]);
            argmax_step(v.w, c, best[3], arg[3]);
          } else {
            argmax_step(__ldcs(src + static_cast<long long>(c) * p.P), c, best[0], arg[0]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < VEC; ++i) pred[i] = arg[i];
    } else if (p.preds_in != nullptr && live) {
      if constexpr (VEC == 4) {
        const longlong2 a = __ldcs(reinterpret_cast<const longlong2*>(p.preds_in + off));
        const longlong2 c = __ldcs(reinterpret_cast<const longlong2*>(p.preds_in + off) + 1);
        pred[0] = a.x; pred[1] = a.y; pred[2] = c.x; pred[3] = c.y;
      } else {
        pred[0] = __ldcs(p.preds_in + off);
      }
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) pred[i] = 0;
    }
    const bool have_pred = p.logits != nullptr || p.preds_in != nullptr;

    long long tgt[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) tgt[i] = p.ignore_index;
    if (p.gt != nullptr && live) {
      if constexpr (VEC == 4) {
        const longlong2 a = __ldcs(reinterpret_cast<const longlong2*>(p.gt + off));
        const longlong2 c = __ldcs(reinterpret_cast<const longlong2*>(p.gt + off) + 1);
        tgt[0] = a.x; tgt[1] = a.y; tgt[2] = c.x; tgt[3] = c.y;
      } else {
        tgt[0] = __ldcs(p.gt + off);
      }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      pred[i] = map_label(p, b, pred[i]);
      tgt[i] = map_label(p, b, tgt[i]);
    }
    if (live) {
      if (p.preds_out != nullptr && have_pred) {
        if constexpr (VEC == 4) {
          __stcs(reinterpret_cast<longlong2*>(p.preds_out + off), make_longlong2(pred[0], pred[1]));
          __stcs(reinterpret_cast<longlong2*>(p.preds_out + off) + 1, make_longlong2(pred[2], pred[3]));
        } else {
          __stcs(p.preds_out + off, pred[0]);
        }
      }
      if (p.gt_out != nullptr && p.gt != nullptr) {
        if constexpr (VEC == 4) {
          __stcs(reinterpret_cast<longlong2*>(p.gt_out + off), make_longlong2(tgt[0], tgt[1]));
          __stcs(reinterpret_cast<longlong2*>(p.gt_out + off) + 1, make_longlong2(tgt[2], tgt[3]));
        } else {
          __stcs(p.gt_out + off, tgt[0]);
        }
      }
    }
    if (p.confmat != nullptr) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        int key = -1;
        if (live && p.gt != nullptr && tgt[i] != p.ignore_index) {
          if (tgt[i] >= 0 && tgt[i] < p.G && pred[i] >= 0 && pred[i] < p.G)
            key = static_cast<int>(tgt[i]) * p.G + static_cast<int>(pred[i]);
          else
            ++bad;
        }
        // one atomic per distinct bin of the warp
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (key >= 0 && lane == __ffs(peers) - 1) {
          if (p.hist_in_smem) atomicAdd(&s_hist[key], static_cast<unsigned int>(__popc(peers)));
          else atomicAdd(p.confmat + key, static_cast<unsigned long long>(__popc(peers)));
        }
      }
    }
  }
  if (p.confmat != nullptr) {
    if (p.hist_in_smem) {
      __syncthreads();
      for (int i = threadIdx.x; i < GG; i += blockDim.x) {
        const unsigned int n = s_hist[i];
        if (n) atomicAdd(p.confmat + i, static_cast<unsigned long long>(n));
      }
    }
    if (bad) atomicAdd(p.invalid, static_cast<unsigned long long>(bad));
  }
}

}  // namespace la

extern "C" int la_label_confusion(void* stream, const float* logits, const long long* preds_in, const long long* gt,
                                  const long long* label_map, long long* preds_out, long long* gt_out,
                                  long long* confmat, long long* invalid, int batch, int classes, long long pixels,
                                  int map_len, int num_classes, long long ignore_index) {
  using namespace la;
  LA_CHECK_ARG(batch > 0 && pixels > 0, "la_label_confusion: empty problem");
  LA_CHECK_ARG(logits == nullptr || preds_in == nullptr, "la_label_confusion: give logits or preds_in, not both");
  LA_CHECK_ARG(logits == nullptr || classes > 0, "la_label_confusion: classes must be positive with logits");
  LA_CHECK_ARG(logits || preds_in || gt, "la_label_confusion: nothing to read");
  LA_CHECK_ARG(label_map == nullptr || map_len > 0, "la_label_confusion: label_map needs map_len > 0");
  LA_CHECK_ARG(confmat == nullptr || (num_classes > 0 && num_classes <= 4096 && invalid != nullptr && gt != nullptr &&
                                      (logits || preds_in)),
               "la_label_confusion: the confusion matrix needs predictions, gt, invalid and 0 < num_classes <= 4096");
  LA_CHECK_ARG(static_cast<long long>(batch) * pixels < (1ll << 40), "la_label_confusion: too many pixels");
  LabelParams p;
  p.logits = logits;
  p.preds_in = preds_in;
  p.gt = gt;
  p.label_map = label_map;
  p.preds_out = preds_out;
  p.gt_out = gt_out;
  p.confmat = reinterpret_cast<unsigned long long*>(confmat);
  p.invalid = reinterpret_cast<unsigned long long*>(invalid);
  p.P = pixels;
  p.B = batch;
  p.C = classes;
  p.map_len = map_len;
  p.G = confmat ? num_classes : 0;
  p.ignore_index = ignore_index;
  const size_t hist_bytes = static_cast<size_t>(p.G) * p.G * sizeof(unsigned int);
  p.hist_in_smem = confmat != nullptr && hist_bytes <= 96 * 1024;
  const size_t smem = p.hist_in_smem ? hist_bytes : 0;
  // 16-byte vectors need every row start aligned: pixels % 4 == 0 and 16/32-byte aligned bases
  auto al = [](const void* q, uintptr_t a) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & (a - 1)) == 0; };
  const bool vec4 = pixels % 4 == 0 && al(logits, 16) && al(preds_in, 16) && al(gt, 16) && al(preds_out, 16) && al(gt_out, 16);
  const long long groups = static_cast<long long>(batch) * (vec4 ? pixels / 4 : pixels);
  long long ctas = (groups + 255) / 256;
  const long long max_ctas = static_cast<long long>(sm_count()) * (smem > 48 * 1024 ? 2 : 4);
  if (ctas > max_ctas) ctas = max_ctas;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (vec4) {
    if (smem > 48 * 1024)
      LA_CHECK_CUDA(cudaFuncSetAttribute(label_confusion_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    label_confusion_kernel<4><<<static_cast<unsigned>(ctas), 256, smem, st>>>(p);
  } else {
    if (smem > 48 * 1024)
      LA_CHECK_CUDA(cudaFuncSetAttribute(label_confusion_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    label_confusion_kernel<1><<<static_cast<unsigned>(ctas), 256, smem, st>>>(p);
  }
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Iterative prompting (second half): `generate_points_from_errors`
// (label_anything/experiment/substitution.py:17-96) -- one corrective point per (episode, class) sampled from the
// pixels where the prediction is wrong: errors = one_hot(gt) - one_hot(argmax(logits)) is +1 at a false negative of the
// class (-> positive point) and -1 at a false positive (-> negative point); `torch.nonzero` lists them in (b, c, h, w)
// order and the reference draws `num_points` indices per (b, c) group with torch.randint.  The reference materialises
// two one-hot [B, C, H, W] int64 tensors, their difference and the full nonzero list (~100 bytes per pixel per class);
// here: one pass counts the errors per (b, c, row), and one block per (b, c) scans the H row counts, finds the row of
// the k-th error and selects the pixel inside it with ballots -- the logits are read twice, nothing is materialised.
// The random draw is an INPUT (rand[b, c, n], any non-negative integers; k = rand mod count), so the caller owns the
// RNG and the result is reproducible; classes without an error get the reference's (0, 0) / label 0 entry; the
// background class keeps its coordinates but gets label 0 (substitution.py:94-95).  Coordinates leave as (x, y) scaled
// per episode like PromptsProcessor.torch_apply_coords (fp32 * fp32, substitution.py:166-171).
// ------------------------------------------------------------------------------------------------------------------
namespace la {

struct ErrParams {
  const float* logits;      // [B, C, H, W]
  const long long* gt;      // [B, H, W]
  int* rowcnt;              // [B, C, H] workspace
  const long long* rnd;     // [B, C, n]
  const float* sx;          // [B] new_w / old_w as fp32
  const float* sy;          // [B]
  float* points;            // [B, C, n, 2] (x, y)
  float* labels;            // [B, C, n]
  int B, C, H, W, n;
  long long ignore_index;
};

__device__ __forceinline__ int pixel_pred(const ErrParams& p, int b, int h, int w) {
  float best = -INFINITY;
  int arg = 0;
  const float* src = p.logits + ((static_cast<long long>(b) * p.C) * p.H + h) * p.W + w;
  for (int c = 0; c < p.C; ++c) argmax_step(__ldg(src + static_cast<long long>(c) * p.H * p.W), c, best, arg);
  return arg;
}

__device__ __forceinline__ int pixel_target(const ErrParams& p, int b, int h, int w) {
  const long long t = __ldg(p.gt + (static_cast<long long>(b) * p.H + h) * p.W + w);
  return t == p.ignore_index ? 0 : static_cast<int>(t);     // substitution.py:33
}

// one warp per image row: rowcnt[b, c, h] = errors of class c in row h (a wrong pixel is an error of BOTH its classes)
__global__ void __launch_bounds__(256) error_row_count_kernel(const ErrParams p) {
  extern __shared__ int s_cnt[];                      // [warps][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* cnt = s_cnt + warp * p.C;
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + warp;
  const bool live = row < static_cast<long long>(p.B) * p.H;
  for (int c = lane; c < p.C; c += 32) cnt[c] = 0;
  __syncwarp();
  if (live) {
    const int b = static_cast<int>(row / p.H), h = static_cast<int>(row % p.H);
    for (int w = lane; w < p.W; w += 32) {
      const int pr = pixel_pred(p, b, h, w), t = pixel_target(p, b, h, w);
      if (pr != t) {
        if (t >= 0 && t < p.C) atomicAdd(cnt + t, 1);
        atomicAdd(cnt + pr, 1);
      }
    }
    __syncwarp();
    for (int c = lane; c < p.C; c += 32) p.rowcnt[(static_cast<long long>(b) * p.C + c) * p.H + h] = cnt[c];
  }
}

// one block per (b, c): scan the row counts, then one warp per requested point
__global__ void __launch_bounds__(128) error_point_select_kernel(const ErrParams p) {
  extern __shared__ int s_pre[];                      // [H + 1] exclusive prefix of the row counts
  const int b = blockIdx.x / p.C, c = blockIdx.x % p.C;
  const int* rc = p.rowcnt + (static_cast<long long>(b) * p.C + c) * p.H;
  if (threadIdx.x == 0) {
    int run = 0;
    for (int h = 0; h < p.H; ++h) {
      s_pre[h] = run;
      run += rc[h];
    }
    s_pre[p.H] = run;
  }
  __syncthreads();
  const int total = s_pre[p.H];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < p.n; i += blockDim.x >> 5) {
    const long long o = (static_cast<long long>(b) * p.C + c) * p.n + i;
    float x = 0.f, y = 0.f, lab = 0.f;
    if (total > 0) {
      const long long r = __ldg(p.rnd + o);
      int k = static_cast<int>((r < 0 ? -r : r) % total);
      int lo = 0, hi = p.H - 1;                       // last row whose prefix <= k
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_pre[mid] <= k) lo = mid; else hi = mid - 1;
      }
      const int h = lo;
      k -= s_pre[h];
      int found_w = -1, found_lab = 0;
      for (int w0 = 0; w0 < p.W && found_w < 0; w0 += 32) {
        const int w = w0 + lane;
        bool err = false;
        int t = 0;
        if (w < p.W) {
          const int pr = pixel_pred(p, b, h, w);
          t = pixel_target(p, b, h, w);
          err = pr != t && (pr == c || t == c);
        }
        const unsigned m = __ballot_sync(0xffffffffu, err);
        const int here = __popc(m);
        if (k < here) {
          // the k-th set bit of m
          unsigned mm = m;
          for (int j = 0; j < k; ++j) mm &= mm - 1;
          const int src_lane = __ffs(mm) - 1;
          found_w = w0 + src_lane;
          found_lab = __shfl_sync(0xffffffffu, t == c ? 1 : -1, src_lane);
        } else {
          k -= here;
        }
      }
      if (found_w >= 0) {
        x = __fmul_rn(static_cast<float>(found_w), __ldg(p.sx + b));
        y = __fmul_rn(static_cast<float>(h), __ldg(p.sy + b));
        lab = c == 0 ? 0.f : static_cast<float>(found_lab);
      }
    }
    if (lane == 0) {
      p.points[2 * o] = x;
      p.points[2 * o + 1] = y;
      p.labels[o] = lab;
    }
  }
}

}  // namespace la

extern "C" long long la_error_points_workspace_bytes(int batch, int classes, int height) {
  return static_cast<long long>(batch) * classes * height * static_cast<long long>(sizeof(int));
}

extern "C" int la_error_points(void* stream, const float* logits, const long long* gt, int batch, int classes, int height,
                               int width, long long ignore_index, const long long* rnd, int n_points, const float* sx,
                               const float* sy, void* workspace, float* points, float* labels) {
  using namespace la;
  LA_CHECK_ARG(logits && gt && rnd && sx && sy && workspace && points && labels, "la_error_points: null pointer");
  LA_CHECK_ARG(batch > 0 && classes > 0 && height > 0 && width > 0 && n_points > 0, "la_error_points: empty problem");
  LA_CHECK_ARG(classes <= 1024 && height <= 8192, "la_error_points: at most 1024 classes and 8192 rows");
  ErrParams p;
  p.logits = logits;
  p.gt = gt;
  p.rowcnt = static_cast<int*>(workspace);
  p.rnd = rnd;
  p.sx = sx;
  p.sy = sy;
  p.points = points;
  p.labels = labels;
  p.B = batch;
  p.C = classes;
  p.H = height;
  p.W = width;
  p.n = n_points;
  p.ignore_index = ignore_index;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long rows = static_cast<long long>(batch) * height;
  error_row_count_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 8 * classes * sizeof(int), st>>>(p);
  LA_CHECK_CUDA(cudaGetLastError());
  error_point_select_kernel<<<batch * classes, 128, (height + 1) * sizeof(int), st>>>(p);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}
