// SIMT fp32 engine for the NeRF MLP (SPARF_ENGINE_SIMT_FP32): CUDA-core FFMA GEMMs, layer by layer,
// activations in a caller-provided HBM workspace, processed in row chunks so memory stays bounded.
// It is the bit-level twin of the reference's fp32 path (same per-element arithmetic, fp32 accumulate)
// and serves (a) as the always-available exact engine, (b) as the on-device cross-check for the
// tcgen05 engine (mlp_tc.cu).  Works for any width that is a multiple of 8 and any L_xyz/L_view <= 16.
//
// Reference: NeRF.compute_raw_density / NeRF.forward (source/models/frequency_nerf.py:149-227),
// FrequencyEmbedder (:47-69), positional_encoding (:229-258).
#include <algorithm>

#include "common.cuh"
#include "mlp_simt.cuh"

namespace sparf {

// ------------------------------------------------------------------------------------------------
// encoders
// ------------------------------------------------------------------------------------------------
__global__ void c2f_weights_kernel(C2F c, int L_xyz, int L_view, float* __restrict__ wts /*[32]*/) {
  int j = threadIdx.x;
  if (j < 16) wts[j] = j < L_xyz ? band_weight(c, L_xyz, j) : 0.f;
  else if (j < 32) wts[j] = (j - 16) < L_view ? band_weight(c, L_view, j - 16) : 0.f;
}

// enc[m][0:3] = x = o + t*d ; enc[m][3 + c*2L + {0,L} + j] = w_j * {sin,cos}(x_c * 2^j pi) ; zero pad to E3p
__global__ void encode_xyz_kernel(long long total, int S, int L, int E3p, const float* __restrict__ origins,
                                  const float* __restrict__ dirs, const float* __restrict__ t,
                                  const float* __restrict__ wts, float* __restrict__ enc) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int col = (int)(idx % E3p);
  long long m = idx / E3p;
  long long r = m / S;
  float val = 0.f;
  if (col < 3 + 6 * L) {
    int c = col < 3 ? col : (col - 3) / (2 * L);
    float x = add_rn(origins[r * 3 + c], mul_rn(dirs[r * 3 + c], t[m]));  // camera.py:433-435
    if (col < 3) {
      val = x;
    } else {
      int rem = (col - 3) - c * 2 * L;
      int is_cos = rem >= L;
      int j = rem - is_cos * L;
      float arg = mul_rn(x, band_freq(j));
      val = mul_rn(is_cos ? cosf(arg) : sinf(arg), wts[j]);
    }
  }
  enc[idx] = val;
}

// per-ray view-direction encoding: unit = d / max(|d|, 1e-12) (F.normalize), same layout as above
__global__ void encode_dir_kernel(int total, int L, int Evp, const float* __restrict__ dirs,
                                  const float* __restrict__ wts_view, float* __restrict__ denc) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int col = idx % Evp, r = idx / Evp;
  float val = 0.f;
  if (col < 3 + 6 * L) {
    float dx = dirs[r * 3], dy = dirs[r * 3 + 1], dz = dirs[r * 3 + 2];
    float len = fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-12f);
    int c = col < 3 ? col : (col - 3) / (2 * L);
    float u = __fdiv_rn(c == 0 ? dx : (c == 1 ? dy : dz), len);
    if (col < 3) {
      val = u;
    } else {
      int rem = (col - 3) - c * 2 * L;
      int is_cos = rem >= L;
      int j = rem - is_cos * L;
      float arg = mul_rn(u, band_freq(j));
      val = mul_rn(is_cos ? cosf(arg) : sinf(arg), wts_view[j]);
    }
  }
  denc[idx] = val;
}

// ------------------------------------------------------------------------------------------------
// 128x128x8 SGEMM family.  256 threads, 8x8 outputs per thread laid out as 2x2 blocks of 4x4 so that
// shared-memory reads are conflict-free float4s.
// ------------------------------------------------------------------------------------------------
constexpr int BM = 128, BN = 128, BK = 8;

struct Frag {
  float acc[8][8];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  }
  __device__ __forceinline__ void mma(const float (*As)[BM], const float (*Bs)[BN], int ty, int tx) {
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
};
__device__ __forceinline__ int frag_row(int ty, int i) { return (i < 4 ? 0 : 64) + ty * 4 + (i & 3); }
__device__ __forceinline__ int frag_col(int tx, int j) { return (j < 4 ? 0 : 64) + tx * 4 + (j & 3); }

// ---- NT:  Y[m][n] = act( sum_k X1[m][k] W[n][k] + sum_k X2[m/div2][k] W[n][wcol2+k] + bias[n] )
template <int ACT>
__global__ void __launch_bounds__(256) gemm_nt_kernel(int M, int N, const float* __restrict__ X1, int ld1, int K1,
                                                      int K1v, const float* __restrict__ X2, int ld2, int K2, int K2v,
                                                      int div2, const float* __restrict__ W, int ldw, int wcol2,
                                                      const float* __restrict__ bias, float* __restrict__ Y, int ldy) {
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int lrow = tid >> 1, lk = (tid & 1) * 4;
  Frag f;
  f.zero();
  for (int src = 0; src < 2; ++src) {
    const float* X = src == 0 ? X1 : X2;
    if (!X) continue;
    const int ld = src == 0 ? ld1 : ld2, K = src == 0 ? K1 : K2, Kv = src == 0 ? K1v : K2v;
    const int wc = src == 0 ? 0 : wcol2, dv = src == 0 ? 1 : div2;
    for (int k0 = 0; k0 < K; k0 += BK) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      int m = m0 + lrow;
      if (m < M) a = *reinterpret_cast<const float4*>(X + (size_t)(m / dv) * ld + k0 + lk);
      As[lk + 0][lrow] = a.x; As[lk + 1][lrow] = a.y; As[lk + 2][lrow] = a.z; As[lk + 3][lrow] = a.w;
      int n = n0 + lrow;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int k = k0 + lk + i;
        Bs[lk + i][lrow] = (n < N && k < Kv) ? W[(size_t)n * ldw + wc + k] : 0.f;
      }
      __syncthreads();
      f.mma(As, Bs, ty, tx);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + frag_row(ty, i);
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int n = n0 + frag_col(tx, j);
      if (n >= N) continue;
      float v = f.acc[i][j] + (bias ? bias[n] : 0.f);
      if (ACT == 1) v = fmaxf(v, 0.f);
      Y[(size_t)m * ldy + n] = v;
    }
  }
}

// ---- NN (dgrad):  D[m][k] = mask(m,k) * ( sum_n G[m][n] W[n][wcol+k] + r1_vec[m]*r1_row[k] )  (= or +=)
__global__ void __launch_bounds__(256) gemm_nn_kernel(int M, int N, int Kout, int Kv, const float* __restrict__ G, int ldg,
                                                      const float* __restrict__ W, int ldw, int wcol,
                                                      const float* __restrict__ mask_src, int ldmask,
                                                      const float* __restrict__ r1_vec, const float* __restrict__ r1_row,
                                                      float* __restrict__ D, int ldd, int accumulate) {
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, c0 = blockIdx.x * BN;
  const int lrow = tid >> 1, lk = (tid & 1) * 4;
  const int bn = tid >> 5, bc = (tid & 31) * 4;
  Frag f;
  f.zero();
  for (int nk0 = 0; nk0 < N; nk0 += BK) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    int m = m0 + lrow;
    if (m < M && nk0 + lk < N) a = *reinterpret_cast<const float4*>(G + (size_t)m * ldg + nk0 + lk);  // N % 4 == 0
    As[lk + 0][lrow] = a.x; As[lk + 1][lrow] = a.y; As[lk + 2][lrow] = a.z; As[lk + 3][lrow] = a.w;
    int n = nk0 + bn;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int k = c0 + bc + i;
      Bs[bn][bc + i] = (n < N && k < Kv) ? W[(size_t)n * ldw + wcol + k] : 0.f;
    }
    __syncthreads();
    f.mma(As, Bs, ty, tx);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + frag_row(ty, i);
    if (m >= M) continue;
    float rv = r1_vec ? r1_vec[m] : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int k = c0 + frag_col(tx, j);
      if (k >= Kout) continue;
      float v = f.acc[i][j];
      if (r1_vec && k < Kv) v = fmaf(rv, r1_row[k], v);
      if (mask_src && !(mask_src[(size_t)m * ldmask + k] > 0.f)) v = 0.f;
      float* d = D + (size_t)m * ldd + k;
      *d = accumulate ? (*d + v) : v;
    }
  }
}

// ---- TN (wgrad):  dW[n][wcol+k] += sum_{m in slab} G[m][n] X[m/div][k]      (atomic over slabs)
__global__ void __launch_bounds__(256) gemm_tn_kernel(int M, int N, int K, int Kv, int rows_per_slab,
                                                      const float* __restrict__ G, int ldg, const float* __restrict__ X,
                                                      int ldx, int div, float* __restrict__ dW, int ldw, int wcol) {
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.y * BM, c0 = blockIdx.x * BN;
  const int m_begin = blockIdx.z * rows_per_slab, m_end = min(M, m_begin + rows_per_slab);
  const int lm = tid >> 5, lc = (tid & 31) * 4;
  Frag f;
  f.zero();
  for (int mm0 = m_begin; mm0 < m_end; mm0 += BK) {
    int m = mm0 + lm;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < m_end) {
      if (n0 + lc < N) a = *reinterpret_cast<const float4*>(G + (size_t)m * ldg + n0 + lc);  // N % 4 == 0
      if (c0 + lc < K) b = *reinterpret_cast<const float4*>(X + (size_t)(m / div) * ldx + c0 + lc);  // K % 4 == 0
    }
    *reinterpret_cast<float4*>(&As[lm][lc]) = a;
    *reinterpret_cast<float4*>(&Bs[lm][lc]) = b;
    __syncthreads();
    f.mma(As, Bs, ty, tx);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int n = n0 + frag_row(ty, i);
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int k = c0 + frag_col(tx, j);
      if (k >= Kv) continue;
      atomicAdd(dW + (size_t)n * ldw + wcol + k, f.acc[i][j]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// narrow layers (1 or 3 outputs): one warp per row
// ------------------------------------------------------------------------------------------------
// MODE 0: density row:  raw = X.W[0] + b ; raw_out[m] = raw ; sigma[m] = softplus(raw + noise)
// MODE 1: colour head:  rgb[m][j] = sigmoid(X.W[j] + b[j]), j < 3
template <int MODE>
__global__ void rowdot_kernel(long long M, int K, const float* __restrict__ X, int ldx, const float* __restrict__ W,
                              int ldw, const float* __restrict__ bias, const float* __restrict__ noise,
                              float* __restrict__ raw_out, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  long long m = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  constexpr int NS = MODE == 0 ? 1 : 3;
  float acc[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) acc[j] = 0.f;
  const float* x = X + (size_t)m * ldx;
  for (int k = lane; k < K; k += 32) {
    float xv = x[k];
#pragma unroll
    for (int j = 0; j < NS; ++j) acc[j] = fmaf(xv, W[(size_t)j * ldw + k], acc[j]);
  }
#pragma unroll
  for (int j = 0; j < NS; ++j)
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], s);
  if (lane == 0) {
    if (MODE == 0) {
      float raw = acc[0] + bias[0];
      if (raw_out) raw_out[m] = raw;
      float z = noise ? add_rn(raw, noise[m]) : raw;
      if (out) out[m] = softplus_f(z);
    } else {
#pragma unroll
      for (int j = 0; j < NS; ++j) out[m * 3 + j] = sigmoid_f(acc[j] + bias[j]);
    }
  }
}

// g_pre[m][j] = d_rgb[m][j] * c (1-c) ;  g_raw[m] = d_sigma[m] * softplus'(raw + noise)
__global__ void head_grad_kernel(long long M, const float* __restrict__ d_rgb, const float* __restrict__ rgbv,
                                 const float* __restrict__ d_sigma, const float* __restrict__ raw,
                                 const float* __restrict__ noise, float* __restrict__ g_pre /*[M,4]*/,
                                 float* __restrict__ g_raw) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float c = rgbv[m * 3 + j];
    g_pre[m * 4 + j] = d_rgb[m * 3 + j] * c * (1.f - c);
  }
  g_pre[m * 4 + 3] = 0.f;
  float z = noise ? add_rn(raw[m], noise[m]) : raw[m];
  g_raw[m] = d_sigma[m] * softplus_grad_f(z);
}

// dW[j][k] += sum_m g[m*gs + j] X[m][k] ; db[j] += sum_m g[m*gs+j]   (NS <= 3 narrow outputs)
template <int NS>
__global__ void narrow_wgrad_kernel(long long M, int K, int rows_per_block, const float* __restrict__ g, int gs,
                                    const float* __restrict__ X, int ldx, float* __restrict__ dW, int ldw,
                                    float* __restrict__ db) {
  long long m_begin = (long long)blockIdx.x * rows_per_block;
  long long m_end = m_begin + rows_per_block < M ? m_begin + rows_per_block : M;
  for (int k = threadIdx.x; k < K + 1; k += blockDim.x) {
    float acc[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) acc[j] = 0.f;
    for (long long m = m_begin; m < m_end; ++m) {
      float xv = k < K ? X[(size_t)m * ldx + k] : 1.f;  // column K = the bias
#pragma unroll
      for (int j = 0; j < NS; ++j) acc[j] = fmaf(g[m * gs + j], xv, acc[j]);
    }
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      if (k < K) atomicAdd(dW + (size_t)j * ldw + k, acc[j]);
      else atomicAdd(db + j, acc[j]);
    }
  }
}

// Ghid[m][k] = (hid[m][k] > 0) * sum_j g_pre[m][j] W9[j][k]
__global__ void narrow_dgrad_kernel(long long total, int K, const float* __restrict__ g_pre, const float* __restrict__ W,
                                    int ldw, const float* __restrict__ hid, float* __restrict__ out) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int k = (int)(idx % K);
  long long m = idx / K;
  float v = 0.f;
#pragma unroll
  for (int j = 0; j < 3; ++j) v = fmaf(g_pre[m * 4 + j], W[(size_t)j * ldw + k], v);
  out[idx] = hid[idx] > 0.f ? v : 0.f;
}

// db[n] += sum_m G[m][n]
__global__ void colsum_kernel(long long M, int N, int rows_per_block, const float* __restrict__ G, int ldg,
                              float* __restrict__ db) {
  __shared__ float red[8][33];
  int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  int n = blockIdx.x * 32 + cx;
  long long m_begin = (long long)blockIdx.y * rows_per_block;
  long long m_end = m_begin + rows_per_block < M ? m_begin + rows_per_block : M;
  float acc = 0.f;
  if (n < N)
    for (long long m = m_begin + ry; m < m_end; m += 8) acc += G[(size_t)m * ldg + n];
  red[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && n < N) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v += red[i][cx];
    atomicAdd(db + n, v);
  }
}

// out[r][c] = sum_{k<S} in[(r*S+k)][c]
__global__ void ray_reduce_kernel(int nrays, int S, int C, const float* __restrict__ in, float* __restrict__ out) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nrays * C) return;
  int c = idx % C, r = idx / C;
  float acc = 0.f;
  for (int k = 0; k < S; ++k) acc += in[((size_t)r * S + k) * C + c];
  out[idx] = acc;
}

// positional-encoding backward + reduction over the ray:  d_o += sum_k g_x ; d_d += sum_k t_k g_x
// d/dx [w sin(f x)] = f * (w cos(f x)) = f * enc_cos ; d/dx [w cos(f x)] = -f * enc_sin.  One warp per ray.
__global__ void posenc_bwd_kernel(int nrays, int S, int L, int E3p, const float* __restrict__ enc,
                                  const float* __restrict__ Genc, const float* __restrict__ t,
                                  float* __restrict__ d_o, float* __restrict__ d_d) {
  const int lane = threadIdx.x & 31;
  int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= nrays) return;
  float so[3] = {0.f, 0.f, 0.f}, sd[3] = {0.f, 0.f, 0.f};
  for (int k = lane; k < S; k += 32) {
    size_t m = (size_t)r * S + k;
    const float* e = enc + m * E3p;
    const float* g = Genc + m * E3p;
    float tk = t[m];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float gx = g[c];
      for (int j = 0; j < L; ++j) {
        float f = band_freq(j);
        int is = 3 + c * 2 * L + j, ic = is + L;
        gx += f * (g[is] * e[ic] - g[ic] * e[is]);
      }
      so[c] += gx;
      sd[c] += tk * gx;
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      so[c] += __shfl_xor_sync(0xffffffffu, so[c], s);
      sd[c] += __shfl_xor_sync(0xffffffffu, sd[c], s);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (d_o) d_o[r * 3 + c] += so[c];
      if (d_d) d_d[r * 3 + c] += sd[c];
    }
  }
}

// view-direction encoding backward: g_unit from Gdenc, then through unit = d/|d|
__global__ void direnc_bwd_kernel(int nrays, int L, int Evp, const float* __restrict__ denc,
                                  const float* __restrict__ Gdenc, const float* __restrict__ dirs,
                                  float* __restrict__ d_d) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrays) return;
  const float* e = denc + (size_t)r * Evp;
  const float* g = Gdenc + (size_t)r * Evp;
  float gu[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float gx = g[c];
    for (int j = 0; j < L; ++j) {
      float f = band_freq(j);
      int is = 3 + c * 2 * L + j, ic = is + L;
      gx += f * (g[is] * e[ic] - g[ic] * e[is]);
    }
    gu[c] = gx;
  }
  float dx = dirs[r * 3], dy = dirs[r * 3 + 1], dz = dirs[r * 3 + 2];
  float len = fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-12f);
  float ux = dx / len, uy = dy / len, uz = dz / len;
  float dot = gu[0] * ux + gu[1] * uy + gu[2] * uz;
  d_d[r * 3] += (gu[0] - ux * dot) / len;
  d_d[r * 3 + 1] += (gu[1] - uy * dot) / len;
  d_d[r * 3 + 2] += (gu[2] - uz * dot) / len;
}

// ------------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------------
static inline int pad4(int x) { return (x + 3) / 4 * 4; }
static inline int pad8(int x) { return (x + 7) / 8 * 8; }

SimtDims simt_dims(const SparfMLP* mlp) {
  SimtDims d;
  d.E3 = 3 + 6 * mlp->L_xyz;
  d.Ev = 3 + 6 * mlp->L_view;
  d.E3p = pad8(d.E3);
  d.Evp = pad8(d.Ev);
  d.W = mlp->width;
  d.HW = mlp->head_width;
  d.nt = mlp->n_trunk;
  d.skip = mlp->skip_layer;
  return d;
}

int simt_validate(const SparfMLP* mlp) {
  SPARF_REQUIRE(mlp != nullptr, "mlp is NULL");
  SPARF_REQUIRE(mlp->n_trunk >= 2 && mlp->n_trunk <= SPARF_MAX_TRUNK, "n_trunk=%d unsupported", mlp->n_trunk);
  SPARF_REQUIRE(mlp->width % 8 == 0 && mlp->width >= 8 && mlp->head_width % 8 == 0 && mlp->head_width >= 8,
                "width=%d / head_width=%d must be multiples of 8", mlp->width, mlp->head_width);
  SPARF_REQUIRE(mlp->L_xyz >= 1 && mlp->L_xyz <= SPARF_MAX_L && mlp->L_view >= 1 && mlp->L_view <= SPARF_MAX_L,
                "L_xyz=%d / L_view=%d unsupported", mlp->L_xyz, mlp->L_view);
  SPARF_REQUIRE(mlp->skip_layer < mlp->n_trunk && mlp->skip_layer != 0, "skip_layer=%d unsupported", mlp->skip_layer);
  SPARF_REQUIRE(!mlp->use_c2f || mlp->progress, "use_c2f needs the progress pointer");
  for (int i = 0; i < mlp->n_trunk; ++i)
    SPARF_REQUIRE(mlp->trunk_w[i] && mlp->trunk_b[i], "trunk layer %d has NULL tensors", i);
  SPARF_REQUIRE(mlp->head_w[0] && mlp->head_b[0] && mlp->head_w[1] && mlp->head_b[1], "head has NULL tensors");
  return SPARF_OK;
}

// rows handled per chunk (whole rays)
static int chunk_rays(int S, int backward) {
  int rows = backward ? 32768 : 65536;
  int n = rows / S;
  return n < 1 ? 1 : n;
}

struct Carver {
  char* p;
  size_t used, cap;
  float* take(size_t nfloats) {
    size_t bytes = align_up(nfloats * sizeof(float), 256);
    float* r = reinterpret_cast<float*>(p + used);
    used += bytes;
    return r;
  }
};

size_t simt_workspace_bytes(const SparfMLP* mlp, int R, int S, int backward) {
  SimtDims d = simt_dims(mlp);
  size_t nr = (size_t)std::min(R, chunk_rays(S, backward));
  size_t Mc = nr * S;
  auto a = [](size_t n) { return align_up(n * sizeof(float), 256); };
  size_t total = a(32) + a(Mc * d.E3p) + a(nr * d.Evp) + a(Mc * d.HW) + a(Mc) + a(Mc * 3);
  if (!backward) {
    total += 2 * a(Mc * d.W);
  } else {
    total += (size_t)d.nt * a(Mc * d.W);            // h0..h_{nt-1}
    total += 2 * a(Mc * d.W);                         // G ping-pong
    total += a(Mc * d.E3p) + a(Mc * d.HW) + a(Mc * 4) + a(Mc) + a(Mc * d.Evp) + a(nr * d.Evp);
  }
  return total + 256;
}

static inline int trunk_in_main(const SimtDims& d, int l) { return l == 0 ? d.E3p : d.W; }
static inline int trunk_in_main_valid(const SimtDims& d, int l) { return l == 0 ? d.E3 : d.W; }
static inline int trunk_ldw(const SimtDims& d, int l) {
  return (l == 0 ? d.E3 : d.W) + (l == d.skip ? d.E3 : 0);
}

#define LAUNCH_OK(name) SPARF_CHECK_LAUNCH(name)

// forward through the MLP for one chunk.  H: array of nt activation buffers (may alias in pairs when
// !keep), raw may be NULL.
static int simt_chunk_forward(const SparfMLP* mlp, const SimtDims& d, int nr, int S, const float* origins,
                              const float* dirs, const float* t, const float* noise, float* wts, float* enc,
                              float* denc, float** H, float* raw, float* hid, float* sigma, float* rgb,
                              cudaStream_t st) {
  const long long Mc = (long long)nr * S;
  C2F c2f{mlp->use_c2f, mlp->c2f_start, mlp->c2f_range, mlp->progress};
  c2f_weights_kernel<<<1, 32, 0, st>>>(c2f, mlp->L_xyz, mlp->L_view, wts);
  LAUNCH_OK("c2f_weights_kernel");
  encode_xyz_kernel<<<ceil_div(Mc * d.E3p, 256), 256, 0, st>>>(Mc * d.E3p, S, mlp->L_xyz, d.E3p, origins, dirs, t, wts, enc);
  LAUNCH_OK("encode_xyz_kernel");
  encode_dir_kernel<<<ceil_div((long long)nr * d.Evp, 256), 256, 0, st>>>(nr * d.Evp, mlp->L_view, d.Evp, dirs, wts + 16, denc);
  LAUNCH_OK("encode_dir_kernel");
  const float* in = enc;
  for (int l = 0; l < d.nt; ++l) {
    const bool last = l == d.nt - 1;
    const int ldw = trunk_ldw(d, l);
    const float* Wl = mlp->trunk_w[l] + (last ? ldw : 0);  // last layer: row 0 is the density row
    const float* bl = mlp->trunk_b[l] + (last ? 1 : 0);
    dim3 grid(ceil_div(d.W, BN), ceil_div(Mc, BM));
    const bool sk = l == d.skip;
    gemm_nt_kernel<1><<<grid, 256, 0, st>>>((int)Mc, d.W, in, trunk_in_main(d, l), trunk_in_main(d, l),
                                            trunk_in_main_valid(d, l), sk ? enc : nullptr, d.E3p, d.E3p, d.E3, 1, Wl, ldw,
                                            d.W, bl, H[l], d.W);
    LAUNCH_OK("gemm_nt_kernel");
    if (last) {
      rowdot_kernel<0><<<ceil_div(Mc, 8), 256, 0, st>>>(Mc, d.W, in, d.W, mlp->trunk_w[l], ldw, mlp->trunk_b[l], noise, raw, sigma);
      LAUNCH_OK("rowdot_kernel<0>");
    }
    in = H[l];
  }
  {
    dim3 grid(ceil_div(d.HW, BN), ceil_div(Mc, BM));
    gemm_nt_kernel<1><<<grid, 256, 0, st>>>((int)Mc, d.HW, H[d.nt - 1], d.W, d.W, d.W, denc, d.Evp, d.Evp, d.Ev, S,
                                            mlp->head_w[0], d.W + d.Ev, d.W, mlp->head_b[0], hid, d.HW);
    LAUNCH_OK("gemm_nt_kernel(head)");
    rowdot_kernel<1><<<ceil_div(Mc, 8), 256, 0, st>>>(Mc, d.HW, hid, d.HW, mlp->head_w[1], d.HW, mlp->head_b[1], nullptr, nullptr, rgb);
    LAUNCH_OK("rowdot_kernel<1>");
  }
  return SPARF_OK;
}

int simt_mlp_forward(const SparfMLP* mlp, int R, int S, const float* origins, const float* dirs, const float* t,
                     const float* noise, float* sigma, float* rgb, void* workspace, size_t workspace_bytes,
                     cudaStream_t st) {
  int rc = simt_validate(mlp);
  if (rc) return rc;
  if (workspace_bytes < simt_workspace_bytes(mlp, R, S, 0)) {
    set_error("mlp_forward: workspace %zu < %zu bytes", workspace_bytes, simt_workspace_bytes(mlp, R, S, 0));
    return SPARF_ERR_WORKSPACE;
  }
  SimtDims d = simt_dims(mlp);
  const int nrc = std::min(R, chunk_rays(S, 0));
  const size_t Mc = (size_t)nrc * S;
  Carver cv{reinterpret_cast<char*>(workspace), 0, workspace_bytes};
  float* wts = cv.take(32);
  float* enc = cv.take(Mc * d.E3p);
  float* denc = cv.take((size_t)nrc * d.Evp);
  float* hid = cv.take(Mc * d.HW);
  cv.take(Mc);
  cv.take(Mc * 3);
  float* ha = cv.take(Mc * d.W);
  float* hb = cv.take(Mc * d.W);
  float* H[SPARF_MAX_TRUNK];
  for (int l = 0; l < d.nt; ++l) H[l] = (l & 1) ? hb : ha;
  for (int r0 = 0; r0 < R; r0 += nrc) {
    int nr = std::min(nrc, R - r0);
    size_t m0 = (size_t)r0 * S;
    rc = simt_chunk_forward(mlp, d, nr, S, origins + (size_t)r0 * 3, dirs + (size_t)r0 * 3, t + m0,
                            noise ? noise + m0 : nullptr, wts, enc, denc, H, nullptr, hid, sigma + m0, rgb + m0 * 3, st);
    if (rc) return rc;
  }
  return SPARF_OK;
}

int simt_mlp_backward(const SparfMLP* mlp, int R, int S, const float* origins, const float* dirs, const float* t,
                      const float* noise, const float* d_sigma, const float* d_rgb, const SparfMLPGrad* grad,
                      float* d_origins, float* d_dirs, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  int rc = simt_validate(mlp);
  if (rc) return rc;
  SPARF_REQUIRE(grad != nullptr, "mlp_backward: grad is NULL");
  if (workspace_bytes < simt_workspace_bytes(mlp, R, S, 1)) {
    set_error("mlp_backward: workspace %zu < %zu bytes", workspace_bytes, simt_workspace_bytes(mlp, R, S, 1));
    return SPARF_ERR_WORKSPACE;
  }
  SimtDims d = simt_dims(mlp);
  const bool need_rays = d_origins != nullptr || d_dirs != nullptr;
  const int nrc = std::min(R, chunk_rays(S, 1));
  const size_t Mcap = (size_t)nrc * S;
  Carver cv{reinterpret_cast<char*>(workspace), 0, workspace_bytes};
  float* wts = cv.take(32);
  float* enc = cv.take(Mcap * d.E3p);
  float* denc = cv.take((size_t)nrc * d.Evp);
  float* hid = cv.take(Mcap * d.HW);
  float* raw = cv.take(Mcap);
  float* rgbv = cv.take(Mcap * 3);
  float* H[SPARF_MAX_TRUNK];
  for (int l = 0; l < d.nt; ++l) H[l] = cv.take(Mcap * d.W);
  float* G0 = cv.take(Mcap * d.W);
  float* G1 = cv.take(Mcap * d.W);
  float* Genc = cv.take(Mcap * d.E3p);
  float* Ghid = cv.take(Mcap * d.HW);
  float* gpre = cv.take(Mcap * 4);
  float* graw = cv.take(Mcap);
  float* Gdtmp = cv.take(Mcap * d.Evp);
  float* Gdenc = cv.take((size_t)nrc * d.Evp);

  for (int r0 = 0; r0 < R; r0 += nrc) {
    const int nr = std::min(nrc, R - r0);
    const long long Mc = (long long)nr * S;
    const size_t m0 = (size_t)r0 * S;
    const float* o_c = origins + (size_t)r0 * 3;
    const float* d_c = dirs + (size_t)r0 * 3;
    const float* t_c = t + m0;
    const float* nz = noise ? noise + m0 : nullptr;
    rc = simt_chunk_forward(mlp, d, nr, S, o_c, d_c, t_c, nz, wts, enc, denc, H, raw, hid, nullptr, rgbv, st);
    if (rc) return rc;
    const int slab = 2048;  // rows per wgrad slab
    const int nslab = ceil_div(Mc, slab);

    head_grad_kernel<<<ceil_div(Mc, 256), 256, 0, st>>>(Mc, d_rgb + m0 * 3, rgbv, d_sigma + m0, raw, nz, gpre, graw);
    LAUNCH_OK("head_grad_kernel");
    // colour head, layer 1 (HW -> 3)
    narrow_wgrad_kernel<3><<<ceil_div(Mc, 512), 128, 0, st>>>(Mc, d.HW, 512, gpre, 4, hid, d.HW, grad->head_w[1], d.HW, grad->head_b[1]);
    LAUNCH_OK("narrow_wgrad_kernel<3>");
    narrow_dgrad_kernel<<<ceil_div(Mc * d.HW, 256), 256, 0, st>>>(Mc * d.HW, d.HW, gpre, mlp->head_w[1], d.HW, hid, Ghid);
    LAUNCH_OK("narrow_dgrad_kernel");
    // colour head, layer 0 ([feat | denc] -> HW)
    const int ldw8 = d.W + d.Ev;
    float* feat = H[d.nt - 1];
    gemm_tn_kernel<<<dim3(ceil_div(d.W, BN), ceil_div(d.HW, BM), nslab), 256, 0, st>>>((int)Mc, d.HW, d.W, d.W, slab, Ghid, d.HW, feat, d.W, 1, grad->head_w[0], ldw8, 0);
    LAUNCH_OK("gemm_tn_kernel(head feat)");
    gemm_tn_kernel<<<dim3(ceil_div(d.Evp, BN), ceil_div(d.HW, BM), nslab), 256, 0, st>>>((int)Mc, d.HW, d.Evp, d.Ev, slab, Ghid, d.HW, denc, d.Evp, S, grad->head_w[0], ldw8, d.W);
    LAUNCH_OK("gemm_tn_kernel(head dir)");
    colsum_kernel<<<dim3(ceil_div(d.HW, 32), ceil_div(Mc, 1024)), 256, 0, st>>>(Mc, d.HW, 1024, Ghid, d.HW, grad->head_b[0]);
    LAUNCH_OK("colsum_kernel(head)");
    gemm_nn_kernel<<<dim3(ceil_div(d.W, BN), ceil_div(Mc, BM)), 256, 0, st>>>((int)Mc, d.HW, d.W, d.W, Ghid, d.HW, mlp->head_w[0], ldw8, 0, feat, d.W, nullptr, nullptr, G0, d.W, 0);
    LAUNCH_OK("gemm_nn_kernel(head)");
    if (d_dirs) {
      gemm_nn_kernel<<<dim3(ceil_div(d.Evp, BN), ceil_div(Mc, BM)), 256, 0, st>>>((int)Mc, d.HW, d.Evp, d.Ev, Ghid, d.HW, mlp->head_w[0], ldw8, d.W, nullptr, 0, nullptr, nullptr, Gdtmp, d.Evp, 0);
      LAUNCH_OK("gemm_nn_kernel(head dir)");
      ray_reduce_kernel<<<ceil_div((long long)nr * d.Evp, 256), 256, 0, st>>>(nr, S, d.Evp, Gdtmp, Gdenc);
      LAUNCH_OK("ray_reduce_kernel");
      direnc_bwd_kernel<<<ceil_div(nr, 128), 128, 0, st>>>(nr, mlp->L_view, d.Evp, denc, Gdenc, d_c, d_dirs + (size_t)r0 * 3);
      LAUNCH_OK("direnc_bwd_kernel");
    }
    // trunk, last layer: z = [raw | feat_pre]
    float* G = G0;
    float* Gn = G1;
    bool genc_written = false;
    for (int l = d.nt - 1; l >= 0; --l) {
      const bool last = l == d.nt - 1;
      const int ldw = trunk_ldw(d, l);
      const float* in = l == 0 ? enc : H[l - 1];
      const int Kin = trunk_in_main(d, l), Kinv = trunk_in_main_valid(d, l);
      const int rowoff = last ? 1 : 0;
      float* dWl = grad->trunk_w[l] + (size_t)rowoff * ldw;
      const float* Wl = mlp->trunk_w[l] + (size_t)rowoff * ldw;
      gemm_tn_kernel<<<dim3(ceil_div(Kin, BN), ceil_div(d.W, BM), nslab), 256, 0, st>>>((int)Mc, d.W, Kin, Kinv, slab, G, d.W, in, Kin, 1, dWl, ldw, 0);
      LAUNCH_OK("gemm_tn_kernel(trunk)");
      if (l == d.skip) {
        gemm_tn_kernel<<<dim3(ceil_div(d.E3p, BN), ceil_div(d.W, BM), nslab), 256, 0, st>>>((int)Mc, d.W, d.E3p, d.E3, slab, G, d.W, enc, d.E3p, 1, dWl, ldw, d.W);
        LAUNCH_OK("gemm_tn_kernel(skip)");
      }
      colsum_kernel<<<dim3(ceil_div(d.W, 32), ceil_div(Mc, 1024)), 256, 0, st>>>(Mc, d.W, 1024, G, d.W, grad->trunk_b[l] + rowoff);
      LAUNCH_OK("colsum_kernel(trunk)");
      if (last) {
        narrow_wgrad_kernel<1><<<ceil_div(Mc, 512), 128, 0, st>>>(Mc, d.W, 512, graw, 1, in, d.W, grad->trunk_w[l], ldw, grad->trunk_b[l]);
        LAUNCH_OK("narrow_wgrad_kernel<1>");
      }
      if (l > 0) {
        gemm_nn_kernel<<<dim3(ceil_div(d.W, BN), ceil_div(Mc, BM)), 256, 0, st>>>((int)Mc, d.W, d.W, d.W, G, d.W, Wl, ldw, 0, in, d.W, last ? graw : nullptr, last ? mlp->trunk_w[l] : nullptr, Gn, d.W, 0);
        LAUNCH_OK("gemm_nn_kernel(trunk)");
      }
      if (need_rays && (l == d.skip || l == 0)) {
        gemm_nn_kernel<<<dim3(ceil_div(d.E3p, BN), ceil_div(Mc, BM)), 256, 0, st>>>((int)Mc, d.W, d.E3p, d.E3, G, d.W, Wl, ldw, l == 0 ? 0 : d.W, nullptr, 0, nullptr, nullptr, Genc, d.E3p, genc_written ? 1 : 0);
        LAUNCH_OK("gemm_nn_kernel(enc)");
        genc_written = true;
      }
      float* tmp = G; G = Gn; Gn = tmp;
    }
    if (need_rays) {
      posenc_bwd_kernel<<<ceil_div(nr, 4), 128, 0, st>>>(nr, S, mlp->L_xyz, d.E3p, enc, Genc, t_c,
                                                        d_origins ? d_origins + (size_t)r0 * 3 : nullptr,
                                                        d_dirs ? d_dirs + (size_t)r0 * 3 : nullptr);
      LAUNCH_OK("posenc_bwd_kernel");
    }
  }
  return SPARF_OK;
}

}  // namespace sparf
