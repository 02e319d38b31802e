// labelanything_b200 — input preprocessing on the GPU: the step right before `Lam.forward`.
//
// Replaces, bit for bit, what the reference's data pipeline does on the host CPU per image:
//   label_anything/data/transforms.py:14-46  CustomResize (= PIL.Image.resize(BILINEAR) through torchvision) ->
//                                            ToTensor -> CustomNormalize (mean / std, zero pad to S x S)
//   label_anything/data/__init__.py:33-61    the non-custom variant Resize((S, S)) -> ToTensor -> Normalize
//   label_anything/data/transforms.py:159-224 PromptsProcessor.apply_masks / apply_coords / apply_boxes
//
// la_preprocess_image_u8: Pillow's 8-bit antialiased triangle-filter resample (src/libImaging/Resample.c) is integer
// arithmetic: per output coordinate a window [xmin, xmin + count) of source pixels and 22-bit fixed-point coefficients,
// a horizontal pass rounded to 8 bits, then a vertical pass.  The coefficient tables are computed on the host exactly
// as Pillow does (double arithmetic; labelanything_b200/transforms.py) and the two passes run here with the same
// 32-bit accumulators, so the result is the same bytes.  The vertical pass is fused with ToTensor + normalisation
// ((v / 255 - mean) / std in IEEE fp32, same operation order as torch) + the zero padding, and writes the CHW fp32
// tensor the image encoder reads.  HBM-bound, integer work: one thread per output pixel, three channels each.
//
// la_rasterize_masks_u8: OR of the instance masks, nearest resize to the preprocess shape, zero pad to S, nearest
// resize to 256 x 256 -- composed into one gather per output pixel with ATen's nearest index arithmetic
// (min(floor(dst * float(in / out)), in - 1)); also raises the "mask present" flag (flag_masks, data/utils.py:218-224).
//
// la_scale_coords_f64: points / box corners scaled in double like numpy, rounded once to fp32 (torch.tensor -> fp32 slot).
#include "la_common.cuh"

namespace la {

constexpr int PRE_PRECISION_BITS = 32 - 8 - 2;   // Resample.c

__device__ __forceinline__ unsigned char clip8(int v) {
  v >>= PRE_PRECISION_BITS;
  return static_cast<unsigned char>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass: src u8 [H, W, 3] -> dst u8 [H, out_w, 3]
__global__ void __launch_bounds__(256)
resample_h_kernel(const unsigned char* __restrict__ src, int H, int W, int out_w, const int* __restrict__ bounds,
                  const int* __restrict__ kk, int ksize, unsigned char* __restrict__ dst) {
  const long long total = static_cast<long long>(H) * out_w;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int xx = static_cast<int>(i % out_w);
    const long long y = i / out_w;
    const int xmin = __ldg(bounds + 2 * xx), cnt = __ldg(bounds + 2 * xx + 1);
    const int* k = kk + static_cast<long long>(xx) * ksize;
    const unsigned char* row = src + (y * W + xmin) * 3;
    int s0 = 1 << (PRE_PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int x = 0; x < cnt; ++x) {
      const int c = __ldg(k + x);
      s0 += static_cast<int>(row[3 * x + 0]) * c;
      s1 += static_cast<int>(row[3 * x + 1]) * c;
      s2 += static_cast<int>(row[3 * x + 2]) * c;
    }
    unsigned char* o = dst + i * 3;
    o[0] = clip8(s0);
    o[1] = clip8(s1);
    o[2] = clip8(s2);
  }
}

// vertical pass + ToTensor + normalise + zero pad: src u8 [in_h, w, 3] -> out fp32 [3, S, S]
__global__ void __launch_bounds__(256)
resample_v_normalize_kernel(const unsigned char* __restrict__ src, int in_h, int w, int new_h,
                            const int* __restrict__ bounds, const int* __restrict__ kk, int ksize, int S, float m0,
                            float m1, float m2, float d0, float d1, float d2, float* __restrict__ out) {
  const long long plane = static_cast<long long>(S) * S;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < plane;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int xx = static_cast<int>(i % S), yy = static_cast<int>(i / S);
    float r0 = 0.f, r1 = 0.f, r2 = 0.f;
    if (yy < new_h && xx < w) {
      unsigned char v0, v1, v2;
      if (bounds == nullptr) {   // height unchanged: Pillow skips the pass
        const unsigned char* p = src + (static_cast<long long>(yy) * w + xx) * 3;
        v0 = p[0];
        v1 = p[1];
        v2 = p[2];
      } else {
        const int ymin = __ldg(bounds + 2 * yy), cnt = __ldg(bounds + 2 * yy + 1);
        const int* k = kk + static_cast<long long>(yy) * ksize;
        int s0 = 1 << (PRE_PRECISION_BITS - 1), s1 = s0, s2 = s0;
        for (int y = 0; y < cnt; ++y) {
          const int c = __ldg(k + y);
          const unsigned char* p = src + (static_cast<long long>(ymin + y) * w + xx) * 3;
          s0 += static_cast<int>(p[0]) * c;
          s1 += static_cast<int>(p[1]) * c;
          s2 += static_cast<int>(p[2]) * c;
        }
        v0 = clip8(s0);
        v1 = clip8(s1);
        v2 = clip8(s2);
      }
      // ToTensor: v / 255 ; Normalize: (x - mean) / std -- IEEE fp32, no contraction
      r0 = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v0), 255.0f), m0), d0);
      r1 = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v1), 255.0f), m1), d1);
      r2 = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v2), 255.0f), m2), d2);
    }
    out[i] = r0;
    out[plane + i] = r1;
    out[2 * plane + i] = r2;
  }
}

__device__ __forceinline__ int nearest_src(int dst, int in_size, int out_size) {
  if (in_size == out_size) return dst;
  const float scale = __fdiv_rn(static_cast<float>(in_size), static_cast<float>(out_size));
  const int s = static_cast<int>(floorf(__fmul_rn(static_cast<float>(dst), scale)));
  return s < in_size - 1 ? s : in_size - 1;
}

__global__ void __launch_bounds__(256)
rasterize_masks_kernel(const unsigned char* __restrict__ masks, int n, int H, int W, int new_h, int new_w,
                       int long_side, int out_side, float* __restrict__ out, unsigned char* __restrict__ flag) {
  const int total = out_side * out_side;
  bool any = false;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ox = i % out_side, oy = i / out_side;
    int sy, sx;
    bool inside = true;
    if (new_h > 0) {   // custom_preprocess: nearest to (new_h, new_w), zero pad to long_side, nearest to out_side
      const int py = nearest_src(oy, long_side, out_side), px = nearest_src(ox, long_side, out_side);
      inside = py < new_h && px < new_w;
      sy = inside ? nearest_src(py, H, new_h) : 0;
      sx = inside ? nearest_src(px, W, new_w) : 0;
    } else {           // straight nearest from (H, W) to out_side
      sy = nearest_src(oy, H, out_side);
      sx = nearest_src(ox, W, out_side);
    }
    unsigned char v = 0;
    if (inside) {
      for (int m = 0; m < n; ++m) v |= masks[(static_cast<long long>(m) * H + sy) * W + sx] != 0;
    }
    out[i] = v ? 1.0f : 0.0f;
    any |= v != 0;
  }
  if (__syncthreads_or(any) && threadIdx.x == 0 && flag != nullptr) *flag = 1;
}

__global__ void scale_coords_kernel(const double* __restrict__ in, long long n, double sx, double sy,
                                    float* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    out[2 * i] = static_cast<float>(__dmul_rn(in[2 * i], sx));
    out[2 * i + 1] = static_cast<float>(__dmul_rn(in[2 * i + 1], sy));
  }
}

static int grid_for(long long total, int block = 256) {
  long long g = (total + block - 1) / block;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace la

extern "C" int la_preprocess_image_u8(void* stream, const void* src, int H, int W, int new_h, int new_w, int S,
                                      const int* bounds_x, const int* kk_x, int ksize_x, const int* bounds_y,
                                      const int* kk_y, int ksize_y, void* tmp, float mean0, float mean1, float mean2,
                                      float std0, float std1, float std2, float* out) {
  using namespace la;
  LA_CHECK_ARG(src && out, "la_preprocess_image_u8: null pointer");
  LA_CHECK_ARG(H > 0 && W > 0 && new_h > 0 && new_w > 0 && S > 0 && new_h <= S && new_w <= S,
               "la_preprocess_image_u8: bad sizes (H %d W %d -> %d x %d in %d)", H, W, new_h, new_w, S);
  LA_CHECK_ARG((new_w == W) == (bounds_x == nullptr) && (new_h == H) == (bounds_y == nullptr),
               "la_preprocess_image_u8: a coefficient table is needed exactly for the axes whose size changes");
  LA_CHECK_ARG(bounds_x == nullptr || (kk_x != nullptr && ksize_x > 0 && tmp != nullptr),
               "la_preprocess_image_u8: horizontal pass needs kk_x and the [H, new_w, 3] scratch");
  LA_CHECK_ARG(bounds_y == nullptr || (kk_y != nullptr && ksize_y > 0), "la_preprocess_image_u8: vertical pass needs kk_y");
  LA_CHECK_ARG(std0 != 0.f && std1 != 0.f && std2 != 0.f, "la_preprocess_image_u8: std must be non-zero");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned char* mid = static_cast<const unsigned char*>(src);
  if (bounds_x != nullptr) {
    resample_h_kernel<<<grid_for(static_cast<long long>(H) * new_w), 256, 0, st>>>(
        static_cast<const unsigned char*>(src), H, W, new_w, bounds_x, kk_x, ksize_x, static_cast<unsigned char*>(tmp));
    LA_CHECK_CUDA(cudaGetLastError());
    mid = static_cast<const unsigned char*>(tmp);
  }
  resample_v_normalize_kernel<<<grid_for(static_cast<long long>(S) * S), 256, 0, st>>>(
      mid, H, new_w, new_h, bounds_y, kk_y, ksize_y, S, mean0, mean1, mean2, std0, std1, std2, out);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

extern "C" int la_rasterize_masks_u8(void* stream, const void* masks, int n, int H, int W, int new_h, int new_w,
                                     int long_side, int out_side, float* out, void* flag) {
  using namespace la;
  LA_CHECK_ARG(out != nullptr && out_side > 0 && long_side > 0, "la_rasterize_masks_u8: bad output");
  LA_CHECK_ARG(n == 0 || (masks != nullptr && H > 0 && W > 0), "la_rasterize_masks_u8: bad masks");
  LA_CHECK_ARG((new_h > 0) == (new_w > 0) && new_h <= long_side && new_w <= long_side,
               "la_rasterize_masks_u8: bad preprocess shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n == 0) {   // no instance of this class in the image: all zeros, flag untouched (transforms.py:198-201)
    LA_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * out_side * out_side, st));
    return LA_OK;
  }
  rasterize_masks_kernel<<<grid_for(static_cast<long long>(out_side) * out_side), 256, 0, st>>>(
      static_cast<const unsigned char*>(masks), n, H, W, new_h, new_w, long_side, out_side, out,
      static_cast<unsigned char*>(flag));
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}

extern "C" int la_scale_coords_f64(void* stream, const void* coords, long long n, double sx, double sy, float* out) {
  using namespace la;
  LA_CHECK_ARG(n >= 0 && (n == 0 || (coords && out)), "la_scale_coords_f64: null pointer");
  if (n == 0) return LA_OK;
  scale_coords_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const double*>(coords), n, sx, sy, out);
  LA_CHECK_CUDA(cudaGetLastError());
  return LA_OK;
}
