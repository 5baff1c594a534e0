// tcgen05 engine for the NeRF MLP (SPARF_ENGINE_TC_3X / TC_1X).
//
// FORWARD: one persistent, warp-specialised kernel.  A CTA owns 128 sample rows at a time (one TMEM lane
// per row) and pushes them through all 9 tensor-core layers without the activations ever leaving the SM:
//
//   warp 0      weight producer : streams pre-packed bf16 (hi | lo) weight blocks, 16 KB each, from L2 into a
//                                 3-deep shared-memory ring with cp.async.bulk + mbarrier complete_tx
//   warp 1      MMA issuer      : one thread issues tcgen05.mma (M=128, N=128, K=16) into one of TWO 256-column
//                                 fp32 TMEM accumulators (ping-pong per layer), commits to mbarriers
//   warps 2-9   epilogue        : positional encoding -> A operand; per layer TMEM -> registers -> bias/ReLU ->
//                                 (hi, lo) bf16 split -> next layer's A operand in shared memory, handed to the
//                                 MMA warp per 64-column K block so layer l+1 starts while layer l drains;
//                                 density row, colour head (128->3) and activations in fp32 on CUDA cores
//
// PRECISION: x*W is evaluated as x_hi*W_hi + x_lo*W_hi + x_hi*W_lo with bf16 operands and fp32 accumulation
// (error-compensated split, ~2^-17 relative per product; SURVEY.md hard part 1).  TC_1X keeps only
// the first term.  The first layer's inputs (x = o + t d, sin/cos of x * 2^j pi) are computed with the
// reference's exact fp32 op sequence before the split.
//
// Reference: NeRF.forward_samples / forward / compute_raw_density (source/models/frequency_nerf.py:149-281).
#include <algorithm>

#include "common.cuh"
#include "mlp_simt.cuh"
#include "mlp_tc.cuh"
#include "tc_common.cuh"

namespace sparf {
using namespace tc;

namespace {

constexpr int kW = 256;          // trunk width
constexpr int kHW = 128;         // head width
constexpr int kL = 10;           // L_xyz
constexpr int kLv = 4;           // L_view
constexpr int kEv = 27;
constexpr int kNumLayers = 9;    // tensor-core layers: trunk 0..7 + head 0
constexpr int kTileM = 128;
constexpr int kStages = 3;
constexpr int kChunkBytes = 16384;  // one [128 x 64] bf16 operand block
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;  // 320
constexpr int kChunksPerTile = 128;
constexpr bool kFwdF16 = true;   // forward operands: fp16 (hi | lo) halves, see tc_common.cuh split2

// K blocks of a layer: enc first (available early), then the 4 activation blocks
__host__ __device__ constexpr int layer_nkb(int l) { return l == 0 ? 1 : (l == 4 ? 5 : 4); }
__host__ __device__ constexpr int layer_nh(int l) { return l == 8 ? 1 : 2; }     // N / 128
__host__ __device__ constexpr bool kb_is_enc(int l, int kbi) { return l == 0 || (l == 4 && kbi == 0); }
__host__ __device__ constexpr int kb_act_index(int l, int kbi) { return l == 4 ? kbi - 1 : kbi; }

// ---- shared memory map (offsets from a 1024-aligned base)
constexpr int kOffAct = 0;                               // 8 blocks: hi kb0..3, lo kb0..3
constexpr int kOffEnc = kOffAct + 8 * kChunkBytes;       // 2 blocks: hi, lo
constexpr int kOffRing = kOffEnc + 2 * kChunkBytes;      // kStages blocks
constexpr int kOffBias = kOffRing + kStages * kChunkBytes;   // 8 x 256 floats (trunk biases; layer 7: rows 1..256)
constexpr int kOffW7r0 = kOffBias + 8 * 256 * 4;         // 256 floats: density row of the last trunk layer
constexpr int kOffW9 = kOffW7r0 + 256 * 4;               // 3 x 128 floats
constexpr int kOffMisc = kOffW9 + 3 * 128 * 4;           // b7[0], b9[0..2], c2f weights [16]
constexpr int kOffPart = kOffMisc + 32 * 4;              // 2 x 128 x 4 floats: cross-warp partial dots
constexpr int kOffBar = kOffPart + 2 * 128 * 4 * 4;      // mbarriers
constexpr int kNumBars = 2 * kStages + 5 + 4;
constexpr int kSmemBytes = kOffBar + kNumBars * 8 + 16;
static_assert(kSmemBytes + 1024 <= 232448, "shared memory budget exceeded");

struct FwdParams {
  const uint8_t* packed;   // kChunksPerTile chunks of 16 KB
  const float* raybias;    // [R,128]: b8 + W8[:,256:283] . dir_enc(ray)
  const float* origins;
  const float* dirs;
  const float* t;
  const float* noise;
  float* sigma;
  float* rgb;
  const float* bias[8];
  const float* w7;         // last trunk layer weight [257,256]
  const float* w9;         // [3,128]
  const float* b9;
  C2F c2f;
  long long M;             // R*S rows
  int S;
  int num_tiles;
  int passes;              // 3 (compensated) or 1
};

// reference column of internal encoder column ic (frequency_nerf.py:65-68 layout), -1 = zero pad
__host__ __device__ inline int enc_ref_col(int ic) {
  if (ic < 3) return ic;
  if (ic == 3) return -1;
  int p = (ic - 4) >> 1, is_cos = (ic - 4) & 1;
  int c = p / kL, j = p % kL;
  return 3 + c * 2 * kL + is_cos * kL + j;
}

// ------------------------------------------------------------------------------------------------
// weight packing: fp32 nn.Linear tensors -> bf16 (hi | lo) SW128 operand blocks in stream order
//   for l: for kb: for nh: for part in {hi, lo}: one 16 KB chunk  [128 (n) x 64 (k)]
// ------------------------------------------------------------------------------------------------
struct PackParams {
  const float* w[9];   // trunk 0..7, head 0
  uint8_t* packed;
};

__global__ void pack_weights_kernel(PackParams pp) {
  // chunk -> (l, kbi, nh, part)
  int chunk = blockIdx.x;
  int l = 0, base = 0;
  for (;; ++l) {
    int n = layer_nkb(l) * layer_nh(l) * 2;
    if (chunk < base + n) break;
    base += n;
  }
  int rel = chunk - base;
  int part = rel & 1, nh = (rel >> 1) % layer_nh(l), kbi = (rel >> 1) / layer_nh(l);
  const bool enc = kb_is_enc(l, kbi);
  const int ldw = l == 0 ? 63 : (l == 4 ? 319 : (l == 8 ? 283 : 256));
  const float* W = pp.w[l];
  uint8_t* dst = pp.packed + (size_t)chunk * kChunkBytes;
  for (int e = threadIdx.x; e < 128 * 64; e += blockDim.x) {
    int n = e >> 6, k = e & 63;
    int row = (l == 7 ? 1 : 0) + nh * 128 + n;
    int col;
    if (enc) {
      int rc = enc_ref_col(k);
      col = rc < 0 ? -1 : (l == 0 ? 0 : kW) + rc;
    } else {
      col = kb_act_index(l, kbi) * 64 + k;
    }
    float v = col < 0 ? 0.f : W[(size_t)row * ldw + col];
    *reinterpret_cast<uint16_t*>(dst + sw128_offset(n, k)) = split1<kFwdF16>(v, part);
  }
}

// per-ray colour-head bias: raybias[r][n] = b8[n] + sum_k W8[n][256+k] * dir_enc(r)[k]   (fp32, exact path)
__global__ void raybias_kernel(int R, const float* __restrict__ dirs, const float* __restrict__ w8,
                               const float* __restrict__ b8, C2F c2f, float* __restrict__ raybias) {
  __shared__ float denc[4][32];
  const int rl = threadIdx.x >> 7, n = threadIdx.x & 127;
  const int r = blockIdx.x * 4 + rl;
  if (n < 32) {
    float val = 0.f;
    if (r < R && n < kEv) {
      float dx = dirs[r * 3], dy = dirs[r * 3 + 1], dz = dirs[r * 3 + 2];
      float len = fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-12f);
      int c = n < 3 ? n : (n - 3) / (2 * kLv);
      float u = __fdiv_rn(c == 0 ? dx : (c == 1 ? dy : dz), len);
      if (n < 3) {
        val = u;
      } else {
        int rem = (n - 3) - c * 2 * kLv;
        int is_cos = rem >= kLv;
        int j = rem - is_cos * kLv;
        float arg = mul_rn(u, band_freq(j));
        val = mul_rn(is_cos ? cosf(arg) : sinf(arg), band_weight(c2f, kLv, j));
      }
    }
    denc[rl][n] = val;
  }
  __syncthreads();
  if (r >= R) return;
  float acc = b8[n];
  const float* wrow = w8 + (size_t)n * (kW + kEv) + kW;
#pragma unroll
  for (int k = 0; k < kEv; ++k) acc = fmaf(wrow[k], denc[rl][k], acc);
  raybias[(size_t)r * kHW + n] = acc;
}

// ------------------------------------------------------------------------------------------------
// the fused forward kernel
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1) tc_mlp_fwd_kernel(const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* s_bias = reinterpret_cast<float*>(smem + kOffBias);
  float* s_w7r0 = reinterpret_cast<float*>(smem + kOffW7r0);
  float* s_w9 = reinterpret_cast<float*>(smem + kOffW9);
  float* s_misc = reinterpret_cast<float*>(smem + kOffMisc);   // [0]=b7[0], [1..3]=b9, [8..23]=c2f weights
  float* s_part = reinterpret_cast<float*>(smem + kOffPart);   // [h][row][4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* w_full = bars;                 // [kStages]
  uint64_t* w_empty = bars + kStages;      // [kStages]
  uint64_t* a_ready = bars + 2 * kStages;  // [5]: act blocks 0..3, enc
  uint64_t* d_full = a_ready + 5;          // [2]
  uint64_t* d_empty = d_full + 2;          // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + kNumBars);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup
  for (int i = tid; i < 8 * 256; i += kThreads) {
    int l = i >> 8, n = i & 255;
    s_bias[i] = p.bias[l][n + (l == 7 ? 1 : 0)];
  }
  for (int i = tid; i < 256; i += kThreads) s_w7r0[i] = p.w7[i];
  for (int i = tid; i < 3 * 128; i += kThreads) s_w9[i] = p.w9[i];
  if (tid == 0) {
    s_misc[0] = p.bias[7][0];
    s_misc[1] = p.b9[0]; s_misc[2] = p.b9[1]; s_misc[3] = p.b9[2];
  }
  if (tid < 16) s_misc[8 + tid] = tid < kL ? band_weight(p.c2f, kL, tid) : 0.f;
  if (tid == 32) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 5; ++i) mbar_init(&a_ready[i], kEpiWarps);
    for (int i = 0; i < 2; ++i) { mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(s_tmem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  const int my_tiles = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    // ============================== weight producer ==============================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        for (int c = 0; c < kChunksPerTile; ++c) {
          // in 1-pass mode the lo chunks (odd) are skipped
          if (p.passes == 1 && (c & 1)) continue;
          mbar_wait(&w_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&w_full[stage], kChunkBytes);
          bulk_g2s(smem + kOffRing + stage * kChunkBytes, p.packed + (size_t)c * kChunkBytes, kChunkBytes, &w_full[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, 128, kFwdF16 ? 0 : 1);
      const uint32_t act_addr = smem_u32(smem + kOffAct), enc_addr = smem_u32(smem + kOffEnc);
      const uint32_t ring_addr = smem_u32(smem + kOffRing);
      uint32_t stage = 0, phase = 0;
      uint32_t a_cnt[5] = {0, 0, 0, 0, 0};
      uint32_t d_cnt[2] = {0, 0};
      for (int it = 0; it < my_tiles; ++it) {
        for (int l = 0; l < kNumLayers; ++l) {
          const int buf = l & 1;
          mbar_wait(&d_empty[buf], (d_cnt[buf] & 1) ^ 1);   // epilogue of the previous user of this accumulator
          ++d_cnt[buf];
          tc_fence_after();
          const int nkb = layer_nkb(l), nh_cnt = layer_nh(l);
          for (int kbi = 0; kbi < nkb; ++kbi) {
            uint32_t a_hi, a_lo;
            if (kb_is_enc(l, kbi)) {
              if (l == 0) { mbar_wait(&a_ready[4], a_cnt[4] & 1); ++a_cnt[4]; }
              a_hi = enc_addr; a_lo = enc_addr + kChunkBytes;
            } else {
              int a = kb_act_index(l, kbi);
              mbar_wait(&a_ready[a], a_cnt[a] & 1);
              ++a_cnt[a];
              a_hi = act_addr + a * kChunkBytes; a_lo = act_addr + (4 + a) * kChunkBytes;
            }
            tc_fence_after();
            for (int nh = 0; nh < nh_cnt; ++nh) {
              const uint32_t d_addr = tmem_base + (uint32_t)(buf * 256 + nh * 128);
              for (int part = 0; part < (p.passes == 1 ? 1 : 2); ++part) {
                mbar_wait(&w_full[stage], phase);
                tc_fence_after();
                const uint32_t b_addr = ring_addr + stage * kChunkBytes;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                  const uint64_t db = make_smem_desc(b_addr + ks * 32);
                  const uint32_t first = (kbi == 0 && part == 0 && ks == 0) ? 0u : 1u;
                  umma_ss(d_addr, make_smem_desc(a_hi + ks * 32), db, idesc, first);
                  if (part == 0 && p.passes != 1) umma_ss(d_addr, make_smem_desc(a_lo + ks * 32), db, idesc, 1u);
                }
                umma_commit(&w_empty[stage]);   // frees the ring slot when these MMAs have read it
                if (++stage == kStages) { stage = 0; phase ^= 1; }
              }
            }
          }
          umma_commit(&d_full[buf]);            // accumulator of layer l complete
        }
      }
    }
  } else {
    // ============================== epilogue warps ==============================
    const int e = warp - 2;
    const int q = warp & 3;           // TMEM lane quadrant this warp may access
    const int h = e >> 2;             // which 32-column half of every 64-column block
    const int row = q * 32 + lane;
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    uint32_t d_cnt[2] = {0, 0};
    uint8_t* act_hi = smem + kOffAct;
    uint8_t* act_lo = smem + kOffAct + 4 * kChunkBytes;
    const float* wts = s_misc + 8;

    for (int it = 0; it < my_tiles; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const long long m = (long long)tile * kTileM + row;
      const bool valid = m < p.M;
      const long long ray = valid ? m / p.S : 0;

      // ---------------- positional encoding -> A_enc (internal column order: x y z 0 | (sin,cos) pairs)
      {
        float x[3] = {0.f, 0.f, 0.f};
        if (valid) {
          float tv = p.t[m];
#pragma unroll
          for (int c = 0; c < 3; ++c) x[c] = add_rn(p.origins[ray * 3 + c], mul_rn(p.dirs[ray * 3 + c], tv));
        }
        float vals[32];
        if (h == 0) { vals[0] = x[0]; vals[1] = x[1]; vals[2] = x[2]; vals[3] = 0.f; }
        const int p0 = h == 0 ? 0 : 14, np = h == 0 ? 14 : 16, v0 = h == 0 ? 4 : 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (i < np) {
            int pr = p0 + i;
            int c = pr / kL, j = pr - c * kL;
            float arg = mul_rn(c == 0 ? x[0] : (c == 1 ? x[1] : x[2]), band_freq(j));
            float sn, cs;
            sincosf(arg, &sn, &cs);
            float w = wts[j];
            vals[v0 + 2 * i] = mul_rn(sn, w);
            vals[v0 + 2 * i + 1] = mul_rn(cs, w);
          }
        }
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) split2<kFwdF16>(vals[2 * i], vals[2 * i + 1], hi[i], lo[i]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t off = sw128_offset(row, h * 32 + c * 8);
          *reinterpret_cast<uint4*>(smem + kOffEnc + off) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
          *reinterpret_cast<uint4*>(smem + kOffEnc + kChunkBytes + off) = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[4]);
      }

      // ---------------- layers
      for (int l = 0; l < kNumLayers; ++l) {
        const int buf = l & 1;
        mbar_wait(&d_full[buf], d_cnt[buf] & 1);
        ++d_cnt[buf];
        tc_fence_after();
        const int nchunk = l == 8 ? 2 : 4;
        float dot0 = 0.f, dot1 = 0.f, dot2 = 0.f;   // density row (l == 6) or rgb rows (l == 8)
        for (int j = 0; j < nchunk; ++j) {
          uint32_t v[32];
          const int col0 = j * 64 + h * 32;
          tmem_ld32(tmem_base + t_lane + (uint32_t)(buf * 256 + col0), v);
          tmem_ld_wait();
          float f[32];
          if (l < 8) {
            const float* b = s_bias + l * 256 + col0;
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = fmaxf(__uint_as_float(v[i]) + b[i], 0.f);
          } else {
            const float* b = p.raybias + (size_t)ray * kHW + col0;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              float4 bb = *reinterpret_cast<const float4*>(b + i);
              f[i] = fmaxf(__uint_as_float(v[i]) + bb.x, 0.f);
              f[i + 1] = fmaxf(__uint_as_float(v[i + 1]) + bb.y, 0.f);
              f[i + 2] = fmaxf(__uint_as_float(v[i + 2]) + bb.z, 0.f);
              f[i + 3] = fmaxf(__uint_as_float(v[i + 3]) + bb.w, 0.f);
            }
          }
          if (l == 6) {
#pragma unroll
            for (int i = 0; i < 32; ++i) dot0 = fmaf(f[i], s_w7r0[col0 + i], dot0);
          }
          if (l == 8) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              dot0 = fmaf(f[i], s_w9[col0 + i], dot0);
              dot1 = fmaf(f[i], s_w9[128 + col0 + i], dot1);
              dot2 = fmaf(f[i], s_w9[256 + col0 + i], dot2);
            }
          } else {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) split2<kFwdF16>(f[2 * i], f[2 * i + 1], hi[i], lo[i]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t off = (uint32_t)j * kChunkBytes + sw128_offset(row, h * 32 + c * 8);
              *reinterpret_cast<uint4*>(act_hi + off) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
              *reinterpret_cast<uint4*>(act_lo + off) = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_ready[j]);
          }
        }
        // accumulator drained: hand it back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&d_empty[buf]);

        if (l == 6 || l == 8) {
          // combine the two column halves of each row (warps q and q+4) through shared memory
          float* pr = s_part + ((size_t)h * 128 + row) * 4;
          pr[0] = dot0; pr[1] = dot1; pr[2] = dot2;
          named_bar_sync(1, kEpiWarps * 32);
          if (h == 0 && valid) {
            const float* o = s_part + ((size_t)128 + row) * 4;
            if (l == 6) {
              float raw = dot0 + o[0] + s_misc[0];
              float z = p.noise ? add_rn(raw, p.noise[m]) : raw;
              p.sigma[m] = softplus_f(z);
            } else {
              p.rgb[m * 3 + 0] = sigmoid_f(dot0 + o[0] + s_misc[1]);
              p.rgb[m * 3 + 1] = sigmoid_f(dot1 + o[1] + s_misc[2]);
              p.rgb[m * 3 + 2] = sigmoid_f(dot2 + o[2] + s_misc[3]);
            }
          }
          named_bar_sync(1, kEpiWarps * 32);
        }
      }
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
bool tc_supports(const SparfMLP* mlp) {
  return mlp && mlp->n_trunk == 8 && mlp->width == kW && mlp->head_width == kHW && mlp->skip_layer == 4 &&
         mlp->L_xyz == kL && mlp->L_view == kLv;
}

bool tc_backward_available() { return false; }

size_t tc_workspace_bytes(const SparfMLP* mlp, int R, int S, int backward, int engine) {
  if (backward) return simt_workspace_bytes(mlp, R, S, 1);
  return align_up((size_t)kChunksPerTile * kChunkBytes, 256) + align_up((size_t)R * kHW * sizeof(float), 256) + 256;
}

int tc_mlp_forward(const SparfMLP* mlp, int engine, int R, int S, const float* origins, const float* dirs,
                   const float* t, const float* noise, float* sigma, float* rgb, void* workspace,
                   size_t workspace_bytes, cudaStream_t st) {
  int rc = simt_validate(mlp);
  if (rc) return rc;
  if (!tc_supports(mlp)) {
    set_error("tcgen05 engine: unsupported MLP shape (needs 8x256 trunk, skip 4, L_xyz=10, L_view=4, head 128)");
    return SPARF_ERR_UNSUPPORTED;
  }
  if (workspace_bytes < tc_workspace_bytes(mlp, R, S, 0, engine)) {
    set_error("tc_mlp_forward: workspace %zu < %zu bytes", workspace_bytes, tc_workspace_bytes(mlp, R, S, 0, engine));
    return SPARF_ERR_WORKSPACE;
  }
  uint8_t* packed = reinterpret_cast<uint8_t*>(workspace);
  float* raybias = reinterpret_cast<float*>(packed + align_up((size_t)kChunksPerTile * kChunkBytes, 256));

  PackParams pp;
  for (int l = 0; l < 8; ++l) pp.w[l] = mlp->trunk_w[l];
  pp.w[8] = mlp->head_w[0];
  pp.packed = packed;
  pack_weights_kernel<<<kChunksPerTile, 256, 0, st>>>(pp);
  SPARF_CHECK_LAUNCH("pack_weights_kernel");

  C2F c2f{mlp->use_c2f, mlp->c2f_start, mlp->c2f_range, mlp->progress};
  raybias_kernel<<<ceil_div(R, 4), 512, 0, st>>>(R, dirs, mlp->head_w[0], mlp->head_b[0], c2f, raybias);
  SPARF_CHECK_LAUNCH("raybias_kernel");

  FwdParams p;
  p.packed = packed;
  p.raybias = raybias;
  p.origins = origins; p.dirs = dirs; p.t = t; p.noise = noise;
  p.sigma = sigma; p.rgb = rgb;
  for (int l = 0; l < 8; ++l) p.bias[l] = mlp->trunk_b[l];
  p.w7 = mlp->trunk_w[7];
  p.w9 = mlp->head_w[1];
  p.b9 = mlp->head_b[1];
  p.c2f = c2f;
  p.M = (long long)R * S;
  p.S = S;
  p.num_tiles = (int)((p.M + kTileM - 1) / kTileM);
  p.passes = engine == SPARF_ENGINE_TC_1X ? 1 : 3;
  static bool attr_set = false;
  if (!attr_set) {
    SPARF_CHECK_CUDA(cudaFuncSetAttribute(tc_mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes + 1024));
    attr_set = true;
  }
  int grid = std::min(p.num_tiles, num_sms());
  tc_mlp_fwd_kernel<<<grid, kThreads, kSmemBytes + 1024, st>>>(p);
  SPARF_CHECK_LAUNCH("tc_mlp_fwd_kernel");
  return SPARF_OK;
}

int tc_mlp_backward(const SparfMLP* mlp, int engine, int R, int S, const float* origins, const float* dirs,
                    const float* t, const float* noise, const float* d_sigma, const float* d_rgb,
                    const SparfMLPGrad* grad, float* d_origins, float* d_dirs, void* workspace,
                    size_t workspace_bytes, cudaStream_t st) {
  // until the tcgen05 dgrad / wgrad kernels land, gradients come from the fp32 SIMT engine
  return simt_mlp_backward(mlp, R, S, origins, dirs, t, noise, d_sigma, d_rgb, grad, d_origins, d_dirs, workspace,
                           workspace_bytes, st);
}

}  // namespace sparf
