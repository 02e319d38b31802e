// labelanything_b200 — multi-head attention for the token <-> image attentions of the two-way transformer and the
// small token-set attentions of the prompt encoder (sm_100a, CUDA cores).
//
// These attentions have one tiny side: n = 1..~60 prompt/class tokens against T = 900..4096 image tokens with
// head_dim 8..64 (hard part H6).  They are HBM/L2-bound streaming reductions, not
// tensor-core shaped contractions, so they run as vectorised CUDA-core kernels:
//
//   la_attention_tokens, mode chosen from the shape:
//     * key-parallel   (few queries, many keys — tokens -> image, transformer.py:311-316,245-250):
//         one warp per (sequence, head, key split); lanes split each key's head slice in 16-byte pieces and walk
//         the keys with an online softmax for up to 4 queries at a time; partial (m, l, acc) of the key groups
//         are merged by shuffles, of the key splits by a second tiny kernel.
//     * query-parallel (many queries or few keys — image -> tokens, transformer.py:322-327; token
//         self-attention, :300-306; AttentionMLPBlock, common.py:151-184): head_dim/8 lanes per (query, head),
//         sequential online softmax over the keys (which stay L1/L2 resident).
//   Projected positional tables can be added on the fly: q' = q + q_add[query index], k' = k + k_add[key index]
//   (the image positional encoding pushed through the projection once per model: Wk·pe + bk), so the image
//   tokens need only ONE projection GEMM with concatenated weights.
//   Masks: the reference's key_mask / query_mask are no-ops (common.py:117-139) -> none here.
#include "la_common.cuh"

namespace la {

struct TokAttParams {
  const __nv_bfloat16* q;
  long long ld_q;
  const __nv_bfloat16* k;
  long long ld_k;
  const __nv_bfloat16* v;
  long long ld_v;
  const float* q_add;  // [nq, ld_qadd] or nullptr (column = head*DH + d)
  long long ld_qadd;
  const float* k_add;  // [nk, ld_kadd] or nullptr
  long long ld_kadd;
  __nv_bfloat16* out;
  long long ld_out;
  float* ws;  // key-parallel partials: [n_seq][splits][heads][nq][DH + 2]
  int n_seq, nq, nk, n_heads, splits;
  float scale_log2;
};

__device__ __forceinline__ float ex2f(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __low2float(h[i]);
    f[2 * i + 1] = __high2float(h[i]);
  }
}
__device__ __forceinline__ void add8(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] += a.x; f[1] += a.y; f[2] += a.z; f[3] += a.w;
  f[4] += b.x; f[5] += b.y; f[6] += b.z; f[7] += b.w;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8], float s) {
  uint4 u;
  u.x = pack_bf16(f[0] * s, f[1] * s);
  u.y = pack_bf16(f[2] * s, f[3] * s);
  u.z = pack_bf16(f[4] * s, f[5] * s);
  u.w = pack_bf16(f[6] * s, f[7] * s);
  *reinterpret_cast<uint4*>(p) = u;
}

// ------------------------------------------------------------------------------------------------------
// query-parallel: G = DH/8 lanes per (sequence, query, head)
// ------------------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(256) attn_query_parallel_kernel(const TokAttParams p) {
  constexpr int G = DH / 8;
  const long long gid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) / G;
  const int sub = threadIdx.x % G;
  const long long total = static_cast<long long>(p.n_seq) * p.nq * p.n_heads;
  const bool live = gid < total;
  const long long g = live ? gid : total - 1;  // dead lanes shadow the last unit so shuffles stay converged
  const int h = static_cast<int>(g % p.n_heads);
  const long long sq = g / p.n_heads;  // seq * nq + query
  const int qi = static_cast<int>(sq % p.nq);
  const long long seq = sq / p.nq;
  const int col = h * DH + sub * 8;

  float q[8];
  load8(p.q + sq * p.ld_q + col, q);
  if (p.q_add) add8(p.q_add + static_cast<long long>(qi) * p.ld_qadd + col, q);
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] *= p.scale_log2;

  float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const __nv_bfloat16* kp = p.k + seq * p.nk * p.ld_k + col;
  const __nv_bfloat16* vp = p.v + seq * p.nk * p.ld_v + col;

  int j = 0;
  for (; j + 2 <= p.nk; j += 2) {
    float k0[8], k1[8], v0[8], v1[8];
    load8(kp + static_cast<long long>(j) * p.ld_k, k0);
    load8(kp + static_cast<long long>(j + 1) * p.ld_k, k1);
    load8(vp + static_cast<long long>(j) * p.ld_v, v0);
    load8(vp + static_cast<long long>(j + 1) * p.ld_v, v1);
    if (p.k_add) {
      add8(p.k_add + static_cast<long long>(j) * p.ld_kadd + col, k0);
      add8(p.k_add + static_cast<long long>(j + 1) * p.ld_kadd + col, k1);
    }
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s0 = fmaf(q[i], k0[i], s0);
      s1 = fmaf(q[i], k1[i], s1);
    }
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    const float mn = fmaxf(m, fmaxf(s0, s1));
    const float c = ex2f(m - mn), p0 = ex2f(s0 - mn), p1 = ex2f(s1 - mn);
    l = fmaf(l, c, p0 + p1);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(acc[i], c, fmaf(p0, v0[i], p1 * v1[i]));
    m = mn;
  }
  if (j < p.nk) {
    float k0[8], v0[8];
    load8(kp + static_cast<long long>(j) * p.ld_k, k0);
    load8(vp + static_cast<long long>(j) * p.ld_v, v0);
    if (p.k_add) add8(p.k_add + static_cast<long long>(j) * p.ld_kadd + col, k0);
    float s0 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s0 = fmaf(q[i], k0[i], s0);
#pragma unroll
    for (int o = 1; o < G; o <<= 1) s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    const float mn = fmaxf(m, s0);
    const float c = ex2f(m - mn), p0 = ex2f(s0 - mn);
    l = fmaf(l, c, p0);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(acc[i], c, p0 * v0[i]);
    m = mn;
  }
  if (live) store8(p.out + sq * p.ld_out + col, acc, 1.0f / l);
}

// ------------------------------------------------------------------------------------------------------
// query-parallel, few keys (<= 16: the image -> token attention of the two-way blocks, 4096 queries against the 1..16
// sparse tokens of their sequence): ONE THREAD per (sequence, query, head), all scores first, then the softmax, then
// the values -- no shuffles, no running rescale; the key / value rows of a sequence are shared by every thread that
// works on it (L1 broadcast).  The G-lanes-per-unit kernel above spent 3.5 ms on 2.5 GB of queries + outputs
// (0.7 TB/s, profiles/r02_ncu_modeb_v1.txt): nine serial online-softmax steps with two shuffles each per lane.
// ------------------------------------------------------------------------------------------------------
constexpr int QR_MAX_KEYS = 16;
constexpr int QR_ITERS = 8;     // query rows per thread: one CTA stages a sequence's keys / values once for 2048 units

// A CTA works on ONE sequence: it stages that sequence's (k + k_add) * scale and v as fp32 in shared memory (rows padded
// by 4 floats per head: the eight heads of a warp's lanes hit eight different bank groups), then every thread walks
// QR_ITERS (query, head) units: the only global traffic of the inner loop is the unit's q row (+ table row) in and its
// output row out.  The first version read k / v through L1 inside the loop and sat at 0.6 TB/s with 7.9 warps stalled
// on loads per issued instruction.
template <int DH>
__global__ void __launch_bounds__(256) attn_query_row_kernel(const TokAttParams p, int chunks) {
  extern __shared__ float s_kv[];                       // [2][nk][heads][DH + 4]
  constexpr int HS = DH + 4;
  const int H = p.n_heads;
  const long long seq = blockIdx.x / chunks;
  const int chunk = blockIdx.x % chunks;
  float* ks = s_kv;
  float* vs = s_kv + p.nk * H * HS;
  for (int i = threadIdx.x; i < p.nk * H * (DH / 8); i += blockDim.x) {
    const int c = i % (DH / 8);
    const int h = (i / (DH / 8)) % H;
    const int j = i / ((DH / 8) * H);
    const int col = h * DH + 8 * c;
    float t[8];
    load8(p.k + (seq * p.nk + j) * p.ld_k + col, t);
    if (p.k_add) add8(p.k_add + static_cast<long long>(j) * p.ld_kadd + col, t);
    float* kd = ks + (j * H + h) * HS + 8 * c;
#pragma unroll
    for (int e = 0; e < 8; ++e) kd[e] = t[e] * p.scale_log2;
    load8(p.v + (seq * p.nk + j) * p.ld_v + col, t);
    float* vd = vs + (j * H + h) * HS + 8 * c;
#pragma unroll
    for (int e = 0; e < 8; ++e) vd[e] = t[e];
  }
  __syncthreads();
  const long long units = static_cast<long long>(p.nq) * H;        // (query, head) units of this sequence
  const long long u0 = static_cast<long long>(chunk) * (256 * QR_ITERS);
  for (int it = 0; it < QR_ITERS; ++it) {
    const long long u = u0 + it * 256 + threadIdx.x;
    if (u >= units) break;
    const int h = static_cast<int>(u % H);
    const int qi = static_cast<int>(u / H);
    const long long sq = seq * p.nq + qi;
    const int col = h * DH;
    float q[DH];
#pragma unroll
    for (int c = 0; c < DH / 8; ++c) {
      float t[8];
      load8(p.q + sq * p.ld_q + col + 8 * c, t);
      if (p.q_add) add8(p.q_add + static_cast<long long>(qi) * p.ld_qadd + col + 8 * c, t);
#pragma unroll
      for (int i = 0; i < 8; ++i) q[8 * c + i] = t[i];
    }
    float s[QR_MAX_KEYS];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < QR_MAX_KEYS; ++j) {
      s[j] = -INFINITY;
      if (j < p.nk) {
        const float4* kr = reinterpret_cast<const float4*>(ks + (j * H + h) * HS);
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < DH / 4; ++c) {
          const float4 t = kr[c];
          acc = fmaf(q[4 * c], t.x, acc);
          acc = fmaf(q[4 * c + 1], t.y, acc);
          acc = fmaf(q[4 * c + 2], t.z, acc);
          acc = fmaf(q[4 * c + 3], t.w, acc);
        }
        s[j] = acc;
        m = fmaxf(m, acc);
      }
    }
    float l = 0.f, o[DH];
#pragma unroll
    for (int i = 0; i < DH; ++i) o[i] = 0.f;
#pragma unroll
    for (int j = 0; j < QR_MAX_KEYS; ++j) {
      if (j < p.nk) {
        const float pj = ex2f(s[j] - m);
        l += pj;
        const float4* vr = reinterpret_cast<const float4*>(vs + (j * H + h) * HS);
#pragma unroll
        for (int c = 0; c < DH / 4; ++c) {
          const float4 t = vr[c];
          o[4 * c] = fmaf(pj, t.x, o[4 * c]);
          o[4 * c + 1] = fmaf(pj, t.y, o[4 * c + 1]);
          o[4 * c + 2] = fmaf(pj, t.z, o[4 * c + 2]);
          o[4 * c + 3] = fmaf(pj, t.w, o[4 * c + 3]);
        }
      }
    }
    const float inv = 1.0f / l;
#pragma unroll
    for (int c = 0; c < DH / 8; ++c) {
      float t[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) t[i] = o[8 * c + i];
      store8(p.out + sq * p.ld_out + col + 8 * c, t, inv);
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// key-parallel: one warp per (sequence, split, head); QB queries per pass
// ------------------------------------------------------------------------------------------------------
// KP_QB: queries handled per pass over the keys (1 for the single-token sequences of mask-only prompts, which keeps
// the register count low enough for KP_UNROLL key groups of loads in flight per warp -- the kernel is a pure HBM
// stream and was latency bound with one group in flight).
// With several tokens per sequence (point / box prompts: 9 tokens; the decoder's class tokens) ALL queries ride in ONE
// pass over the keys when they fit the register budget (KP_QB up to 10 with two key groups in flight): with KP_QB = 4
// nine queries meant three passes, i.e. K and V of all S*T image tokens read three times (7.6 GB instead of 2.5 GB per
// four episodes, profiles/r02_ncu_modeb_v1.txt).
template <int DH, int KP_QB, int KP_UNROLL>
__global__ void __launch_bounds__(256) attn_key_parallel_kernel(const TokAttParams p) {
  constexpr int G = DH / 8;
  constexpr int KG = 32 / G;  // keys per warp iteration
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int sub = lane % G, kg = lane / G;
  const long long seq = blockIdx.x / p.splits;
  const int split = blockIdx.x % p.splits;
  const int per = (p.nk + p.splits - 1) / p.splits;
  const int k_begin = split * per;
  const int k_end = min(p.nk, k_begin + per);

  for (int h = warp; h < p.n_heads; h += nwarps) {
    const int col = h * DH + sub * 8;
    const __nv_bfloat16* kp = p.k + seq * p.nk * p.ld_k + col;
    const __nv_bfloat16* vp = p.v + seq * p.nk * p.ld_v + col;
    for (int q0 = 0; q0 < p.nq; q0 += KP_QB) {
      float q[KP_QB][8], m[KP_QB], l[KP_QB], acc[KP_QB][8];
#pragma unroll
      for (int a = 0; a < KP_QB; ++a) {
        const int qi = min(q0 + a, p.nq - 1);
        load8(p.q + (seq * p.nq + qi) * p.ld_q + col, q[a]);
        if (p.q_add) add8(p.q_add + static_cast<long long>(qi) * p.ld_qadd + col, q[a]);
        m[a] = -INFINITY;
        l[a] = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          q[a][i] *= p.scale_log2;
          acc[a][i] = 0.f;
        }
      }
      for (int j0 = k_begin; j0 < k_end; j0 += KG * KP_UNROLL) {
        // KP_UNROLL key groups: all loads first (raw 16-byte pieces), then the online-softmax updates
        uint4 kraw[KP_UNROLL], vraw[KP_UNROLL];
        float4 ka0[KP_UNROLL], ka1[KP_UNROLL];
        bool on[KP_UNROLL];
#pragma unroll
        for (int u = 0; u < KP_UNROLL; ++u) {
          const int j = j0 + u * KG + kg;
          on[u] = j < k_end;
          const int jj = on[u] ? j : k_end - 1;
          kraw[u] = *reinterpret_cast<const uint4*>(kp + static_cast<long long>(jj) * p.ld_k);
          vraw[u] = *reinterpret_cast<const uint4*>(vp + static_cast<long long>(jj) * p.ld_v);
          if (p.k_add) {
            const float4* ap = reinterpret_cast<const float4*>(p.k_add + static_cast<long long>(jj) * p.ld_kadd + col);
            ka0[u] = __ldg(ap);
            ka1[u] = __ldg(ap + 1);
          }
        }
#pragma unroll
        for (int u = 0; u < KP_UNROLL; ++u) {
          float kk[8], vv[8];
          {
            const __nv_bfloat162* hk = reinterpret_cast<const __nv_bfloat162*>(&kraw[u]);
            const __nv_bfloat162* hv = reinterpret_cast<const __nv_bfloat162*>(&vraw[u]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              kk[2 * i] = __low2float(hk[i]);
              kk[2 * i + 1] = __high2float(hk[i]);
              vv[2 * i] = __low2float(hv[i]);
              vv[2 * i + 1] = __high2float(hv[i]);
            }
          }
          if (p.k_add) {
            kk[0] += ka0[u].x; kk[1] += ka0[u].y; kk[2] += ka0[u].z; kk[3] += ka0[u].w;
            kk[4] += ka1[u].x; kk[5] += ka1[u].y; kk[6] += ka1[u].z; kk[7] += ka1[u].w;
          }
#pragma unroll
          for (int a = 0; a < KP_QB; ++a) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) s = fmaf(q[a][i], kk[i], s);
#pragma unroll
            for (int o = 1; o < G; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (!on[u]) s = -INFINITY;
            const float mn = fmaxf(m[a], s);
            // mn == -inf only while this key group has seen no key at all: keep the state untouched
            const float c = (mn == -INFINITY) ? 1.f : ex2f(m[a] - mn);
            const float pw = (mn == -INFINITY) ? 0.f : ex2f(s - mn);
            l[a] = fmaf(l[a], c, pw);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[a][i] = fmaf(acc[a][i], c, pw * vv[i]);
            m[a] = mn;
          }
        }
      }
      // merge the KG key groups of the warp
#pragma unroll
      for (int a = 0; a < KP_QB; ++a) {
#pragma unroll
        for (int o = G; o < 32; o <<= 1) {
          const float mo = __shfl_xor_sync(0xffffffffu, m[a], o);
          const float lo = __shfl_xor_sync(0xffffffffu, l[a], o);
          const float mn = fmaxf(m[a], mo);
          const float c0 = (mn == -INFINITY) ? 1.f : ex2f(m[a] - mn);
          const float c1 = (mn == -INFINITY) ? 0.f : ex2f(mo - mn);
          l[a] = l[a] * c0 + lo * c1;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float ao = __shfl_xor_sync(0xffffffffu, acc[a][i], o);
            acc[a][i] = acc[a][i] * c0 + ao * c1;
          }
          m[a] = mn;
        }
      }
      if (kg == 0) {
#pragma unroll
        for (int a = 0; a < KP_QB; ++a) {
          const int qi = q0 + a;
          if (qi < p.nq) {
            if (p.splits == 1) {
              store8(p.out + (seq * p.nq + qi) * p.ld_out + col, acc[a], 1.0f / l[a]);
            } else {
              float* w = p.ws + ((((seq * p.splits + split) * p.n_heads + h) * p.nq) + qi) * (DH + 2);
              if (sub == 0) {
                w[0] = m[a];
                w[1] = l[a];
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) w[2 + sub * 8 + i] = acc[a][i];
            }
          }
        }
      }
    }
  }
}

// merge the key splits: one thread per (sequence, head, query, d)
template <int DH>
__global__ void __launch_bounds__(256) attn_merge_splits_kernel(const TokAttParams p) {
  const long long total = static_cast<long long>(p.n_seq) * p.n_heads * p.nq * DH;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(i % DH);
    long long r = i / DH;
    const int qi = static_cast<int>(r % p.nq);
    r /= p.nq;
    const int h = static_cast<int>(r % p.n_heads);
    const long long seq = r / p.n_heads;
    float mx = -INFINITY;
    for (int s = 0; s < p.splits; ++s)
      mx = fmaxf(mx, p.ws[((((seq * p.splits + s) * p.n_heads + h) * p.nq) + qi) * (DH + 2)]);
    float l = 0.f, a = 0.f;
    for (int s = 0; s < p.splits; ++s) {
      const float* w = p.ws + ((((seq * p.splits + s) * p.n_heads + h) * p.nq) + qi) * (DH + 2);
      const float c = (w[0] == -INFINITY) ? 0.f : ex2f(w[0] - mx);
      l = fmaf(w[1], c, l);
      a = fmaf(w[2 + d], c, a);
    }
    p.out[(seq * p.nq + qi) * p.ld_out + h * DH + d] = __float2bfloat16_rn(a / l);
  }
}

template <int DH>
static int launch_tokens(cudaStream_t st, const TokAttParams& p, bool key_parallel) {
  if (key_parallel) {
    const long long grid = static_cast<long long>(p.n_seq) * p.splits;
    const int threads = 32 * (p.n_heads < 8 ? p.n_heads : 8);
    if (p.nq == 1) attn_key_parallel_kernel<DH, 1, 4><<<static_cast<unsigned>(grid), threads, 0, st>>>(p);
    else if (p.nq <= 4) attn_key_parallel_kernel<DH, 4, 4><<<static_cast<unsigned>(grid), threads, 0, st>>>(p);
    else if (p.nq <= 6) attn_key_parallel_kernel<DH, 6, 2><<<static_cast<unsigned>(grid), threads, 0, st>>>(p);
    else attn_key_parallel_kernel<DH, 10, 2><<<static_cast<unsigned>(grid), threads, 0, st>>>(p);
    LA_CHECK_CUDA(cudaGetLastError());
    if (p.splits > 1) {
      const long long total = static_cast<long long>(p.n_seq) * p.n_heads * p.nq * DH;
      long long blocks = (total + 255) / 256;
      const long long cap = static_cast<long long>(sm_count()) * 8;
      if (blocks > cap) blocks = cap;
      attn_merge_splits_kernel<DH><<<static_cast<unsigned>(blocks), 256, 0, st>>>(p);
      LA_CHECK_CUDA(cudaGetLastError());
    }
  } else if (p.nk <= QR_MAX_KEYS && p.nq >= 64 &&
             static_cast<long long>(p.n_seq) * ((static_cast<long long>(p.nq) * p.n_heads + 256 * QR_ITERS - 1) /
                                                (256 * QR_ITERS)) < (1ll << 31)) {
    const int chunks = static_cast<int>((static_cast<long long>(p.nq) * p.n_heads + 256 * QR_ITERS - 1) / (256 * QR_ITERS));
    const size_t smem = sizeof(float) * 2 * p.nk * p.n_heads * (DH + 4);
    auto kern = attn_query_row_kernel<DH>;
    if (smem > 48 * 1024) LA_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<static_cast<unsigned>(p.n_seq * chunks), 256, smem, st>>>(p, chunks);
    LA_CHECK_CUDA(cudaGetLastError());
  } else {
    constexpr int G = DH / 8;
    const long long units = static_cast<long long>(p.n_seq) * p.nq * p.n_heads;
    const long long blocks = (units * G + 255) / 256;
    attn_query_parallel_kernel<DH><<<static_cast<unsigned>(blocks), 256, 0, st>>>(p);
    LA_CHECK_CUDA(cudaGetLastError());
  }
  return LA_OK;
}

}  // namespace la

extern "C" {

int la_attention_tokens_splits(long long n_seq, int nq, int nk) {
  // key-parallel only when the queries are few and the keys many
  if (!(nq <= 64 && nk >= 256 && nk >= 8 * nq)) return 0;
  const long long target = 2ll * la::sm_count();
  long long s = (target + n_seq - 1) / n_seq;
  const long long max_s = nk / 128;
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  return static_cast<int>(s);
}

long long la_attention_tokens_workspace_bytes(long long n_seq, int nq, int nk, int n_heads, int head_dim) {
  const int s = la_attention_tokens_splits(n_seq, nq, nk);
  if (s <= 1) return 0;
  return n_seq * s * n_heads * nq * static_cast<long long>(head_dim + 2) * 4;
}

int la_attention_tokens(void* stream, const void* q, long long ld_q, const void* k, long long ld_k, const void* v,
                        long long ld_v, const float* q_add, long long ld_qadd, const float* k_add,
                        long long ld_kadd, void* out, long long ld_out, long long n_seq, int nq, int nk, int n_heads,
                        int head_dim, float scale, void* workspace) {
  using namespace la;
  LA_CHECK_ARG(q && k && v && out, "la_attention_tokens: null pointer");
  LA_CHECK_ARG(n_seq > 0 && nq > 0 && nk > 0 && n_heads > 0, "la_attention_tokens: empty problem");
  LA_CHECK_ARG(head_dim == 8 || head_dim == 16 || head_dim == 32 || head_dim == 64,
               "la_attention_tokens: head_dim %d unsupported (8, 16, 32, 64)", head_dim);
  LA_CHECK_ARG(ld_q % 8 == 0 && ld_k % 8 == 0 && ld_v % 8 == 0 && ld_out % 8 == 0,
               "la_attention_tokens: row strides must be multiples of 8 elements");
  LA_CHECK_ARG((!q_add || ld_qadd % 4 == 0) && (!k_add || ld_kadd % 4 == 0),
               "la_attention_tokens: table strides must be multiples of 4");
  LA_CHECK_ARG(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                 reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(q_add) |
                 reinterpret_cast<uintptr_t>(k_add)) & 15) == 0,
               "la_attention_tokens: pointers must be 16-byte aligned");
  LA_CHECK_ARG(n_seq * static_cast<long long>(nq) * n_heads * (head_dim / 8) < (1ll << 38),
               "la_attention_tokens: problem too large");
  TokAttParams p;
  p.q = static_cast<const __nv_bfloat16*>(q);
  p.ld_q = ld_q;
  p.k = static_cast<const __nv_bfloat16*>(k);
  p.ld_k = ld_k;
  p.v = static_cast<const __nv_bfloat16*>(v);
  p.ld_v = ld_v;
  p.q_add = q_add;
  p.ld_qadd = ld_qadd;
  p.k_add = k_add;
  p.ld_kadd = ld_kadd;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ld_out = ld_out;
  p.ws = static_cast<float*>(workspace);
  p.n_seq = static_cast<int>(n_seq);
  p.nq = nq;
  p.nk = nk;
  p.n_heads = n_heads;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.splits = la_attention_tokens_splits(n_seq, nq, nk);
  const bool key_parallel = p.splits >= 1;
  LA_CHECK_ARG(n_seq < (1ll << 31) / (p.splits > 0 ? p.splits : 1), "la_attention_tokens: too many sequences");
  LA_CHECK_ARG(p.splits <= 1 || workspace, "la_attention_tokens: workspace required (%d key splits)", p.splits);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (head_dim) {
    case 8: return launch_tokens<8>(st, p, key_parallel);
    case 16: return launch_tokens<16>(st, p, key_parallel);
    case 32: return launch_tokens<32>(st, p, key_parallel);
    default: return launch_tokens<64>(st, p, key_parallel);
  }
}

}  // extern "C"
