// Minimal tcgen05 GEMM used to validate, on hardware, every primitive the tensor-core engine relies on:
// SWIZZLE_128B operand images (written by generic stores AND pre-packed + bulk-copied), shared-memory /
// instruction descriptors, K stepping inside the swizzle atom, commit -> mbarrier, TMEM loads.
//   D[128, 128] = A[128, K] * B[128, K]^T,  K = 128, operands rounded to bf16, fp32 accumulate.
#include "common.cuh"
#include "tc_common.cuh"

namespace sparf {
using namespace tc;

// B [128, K] fp32 row-major -> packed bf16 SW128 blocks, block kb = columns [64 kb, 64 kb + 64)
__global__ void selftest_pack_kernel(const float* __restrict__ B, int K, uint8_t* __restrict__ packed) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 128 * K) return;
  int n = idx / K, k = idx % K;
  int kb = k / kBlockK, kl = k % kBlockK;
  __nv_bfloat16 v = __float2bfloat16_rn(B[idx]);
  *reinterpret_cast<__nv_bfloat16*>(packed + (size_t)kb * 16384 + sw128_offset(n, kl)) = v;
}

// a_in_tmem: the A operand is written to tensor memory with tcgen05.st (columns 128..) and consumed from there
// (the operand path planned for the chain kernels: no shared-memory A blocks, no proxy fence in the epilogue)
__global__ void __launch_bounds__(128) selftest_gemm_kernel(const float* __restrict__ A, const uint8_t* __restrict__ Bpacked,
                                                             int K, float* __restrict__ D, int a_in_tmem) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int nkb = K / kBlockK;
  uint8_t* sA = smem;                       // nkb blocks of 16 KB
  uint8_t* sB = smem + (size_t)nkb * 16384;  // nkb blocks of 16 KB
  __shared__ uint64_t bar_b, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    mbar_init(&bar_b, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 256);
    tmem_relinquish();
  }
  if (a_in_tmem) {
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t a_base = tmem_base_s + 128u + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < K / 2; c0 += 16) {       // 16 columns = 32 consecutive K elements of this thread's row
      uint32_t w[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        float a0 = A[(size_t)tid * K + 2 * (c0 + e)], a1 = A[(size_t)tid * K + 2 * (c0 + e) + 1];
        __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);      // low half = even K element
        w[e] = *reinterpret_cast<uint32_t*>(&h);
      }
      tmem_st16(a_base + (uint32_t)c0, w);
    }
    tmem_st_wait();
  }
  // A: thread = row, generic 16-byte stores of 8 bf16 at the swizzled position
  for (int kb = 0; kb < nkb; ++kb) {
    for (int c = 0; c < 8; ++c) {
      uint32_t w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float a0 = A[(size_t)tid * K + kb * 64 + c * 8 + 2 * e], a1 = A[(size_t)tid * K + kb * 64 + c * 8 + 2 * e + 1];
        __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);
        w[e] = *reinterpret_cast<uint32_t*>(&h);
      }
      *reinterpret_cast<uint4*>(sA + (size_t)kb * 16384 + sw128_offset(tid, c * 8)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (tid == 0) {
    mbar_arrive_expect_tx(&bar_b, (uint32_t)nkb * 16384u);
    for (int kb = 0; kb < nkb; ++kb) bulk_g2s(sB + (size_t)kb * 16384, Bpacked + (size_t)kb * 16384, 16384u, &bar_b);
    mbar_wait(&bar_b, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc(128, 128, 1);
    for (int kb = 0; kb < nkb; ++kb) {
      for (int ks = 0; ks < 4; ++ks) {
        uint64_t da = make_smem_desc(smem_u32(sA + (size_t)kb * 16384) + ks * 32);
        uint64_t db = make_smem_desc(smem_u32(sB + (size_t)kb * 16384) + ks * 32);
        if (a_in_tmem) umma_ts(tmem_base, tmem_base + 128u + (uint32_t)(kb * 32 + ks * 8), db, idesc, (kb | ks) != 0);
        else umma_ss(tmem_base, da, db, idesc, (kb | ks) != 0);
      }
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) D[(size_t)tid * 128 + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// "TN" GEMM with MN-major operands (the weight-gradient shape):
//   D[128 (m), 128 (n)] = sum_{r < rows} G[r][m] * X[r][n],   G, X: [rows, 128] fp32 row-major, rows in {64, 128}
// Both operands are written as the forward A-operand image ([rows x 64] blocks, sw128_offset(row, col)) and
// consumed through MN-major descriptors.
// x_fp16 != 0: X is stored as fp16 while G stays bf16 (mixed operand formats in one kind::f16 MMA)
__global__ void __launch_bounds__(128) selftest_tn_kernel(const float* __restrict__ G, const float* __restrict__ X, int rows,
                                                           float* __restrict__ D, int x_fp16) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sG = smem;              // 2 feature blocks x 16 KB
  uint8_t* sX = smem + 2 * 16384;  // 2 feature blocks x 16 KB
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 128);
    tmem_relinquish();
  }
  // thread = sample row: 16-byte stores of 8 consecutive features
  for (int src = 0; src < 2; ++src) {
    const float* S = src == 0 ? G : X;
    uint8_t* dst = src == 0 ? sG : sX;
    for (int fb = 0; fb < 2; ++fb) {
      for (int c = 0; c < 8; ++c) {
        uint32_t w[4] = {0, 0, 0, 0};
        if (tid < rows) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float v0 = S[(size_t)tid * 128 + fb * 64 + c * 8 + 2 * e], v1 = S[(size_t)tid * 128 + fb * 64 + c * 8 + 2 * e + 1];
            if (src == 1 && x_fp16) {
              __half2 h = __floats2half2_rn(v0, v1);
              w[e] = *reinterpret_cast<uint32_t*>(&h);
            } else {
              __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
              w[e] = *reinterpret_cast<uint32_t*>(&h);
            }
          }
        }
        *reinterpret_cast<uint4*>(dst + (size_t)fb * 16384 + sw128_offset(tid, c * 8)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_ab(128, 128, 1, x_fp16 ? 0 : 1, 1, 1);
    for (int ks = 0; ks < rows / 16; ++ks) {   // 16 sample rows = two 8-row atoms = 2048 bytes
      uint64_t da = make_smem_desc_mn(smem_u32(sG) + ks * 2048, 16384);
      uint64_t db = make_smem_desc_mn(smem_u32(sX) + ks * 2048, 16384);
      umma_ss(tmem_base, da, db, idesc, ks != 0);
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) D[(size_t)tid * 128 + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}

}  // namespace sparf

using namespace sparf;

extern "C" int sparf_tc_selftest_tn(const float* G, const float* X, int32_t rows, float* D, sparf_stream_t stream) {
  SPARF_REQUIRE(rows == 64 || rows == 128, "tc_selftest_tn: rows=%d", rows);
  size_t smem = (size_t)4 * 16384 + 1024;
  SPARF_CHECK_CUDA(cudaFuncSetAttribute(selftest_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  selftest_tn_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(G, X, rows, D, 0);
  SPARF_CHECK_LAUNCH("selftest_tn_kernel");
  return SPARF_OK;
}

// Probe: the same GEMM with G in bf16 and X in fp16 (a_format != b_format).  Debug entry (tools/probe_mixed_formats.py).
extern "C" int sparf_tc_selftest_tn_mixed(const float* G, const float* X, int32_t rows, float* D, sparf_stream_t stream) {
  SPARF_REQUIRE(rows == 64 || rows == 128, "tc_selftest_tn_mixed: rows=%d", rows);
  size_t smem = (size_t)4 * 16384 + 1024;
  SPARF_CHECK_CUDA(cudaFuncSetAttribute(selftest_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  selftest_tn_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(G, X, rows, D, 1);
  SPARF_CHECK_LAUNCH("selftest_tn_kernel");
  return SPARF_OK;
}

// A [128,K], B [128,K] fp32 device, K in {64,128,192,256}; packed: >= 128*K*2 bytes scratch; D [128,128] out.
extern "C" int sparf_tc_selftest(const float* A, const float* B, int32_t K, void* packed, float* D, sparf_stream_t stream) {
  SPARF_REQUIRE(K % 64 == 0 && K >= 64 && K <= 256, "tc_selftest: K=%d", K);
  cudaStream_t st = (cudaStream_t)stream;
  selftest_pack_kernel<<<ceil_div(128 * K, 256), 256, 0, st>>>(B, K, (uint8_t*)packed);
  SPARF_CHECK_LAUNCH("selftest_pack_kernel");
  size_t smem = (size_t)2 * (K / 64) * 16384 + 1024;
  SPARF_CHECK_CUDA(cudaFuncSetAttribute(selftest_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  selftest_gemm_kernel<<<1, 128, smem, st>>>(A, (const uint8_t*)packed, K, D, 0);
  SPARF_CHECK_LAUNCH("selftest_gemm_kernel");
  return SPARF_OK;
}

// Same GEMM with the A operand in tensor memory (tcgen05.st + tcgen05.mma [d], [a], b-desc).
extern "C" int sparf_tc_selftest_ts(const float* A, const float* B, int32_t K, void* packed, float* D, sparf_stream_t stream) {
  SPARF_REQUIRE(K % 64 == 0 && K >= 64 && K <= 256, "tc_selftest_ts: K=%d", K);
  cudaStream_t st = (cudaStream_t)stream;
  selftest_pack_kernel<<<ceil_div(128 * K, 256), 256, 0, st>>>(B, K, (uint8_t*)packed);
  SPARF_CHECK_LAUNCH("selftest_pack_kernel");
  size_t smem = (size_t)2 * (K / 64) * 16384 + 1024;
  SPARF_CHECK_CUDA(cudaFuncSetAttribute(selftest_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  selftest_gemm_kernel<<<1, 128, smem, st>>>(A, (const uint8_t*)packed, K, D, 1);
  SPARF_CHECK_LAUNCH("selftest_gemm_kernel");
  return SPARF_OK;
}

// ------------------------------------------------------------------------------------------------
// Micro-benchmark: per-SM throughput of cp.async.bulk global(L2) -> shared as a function of the number of
// copies in flight and the copy size (sizes the weight ring of the fused kernels).  One CTA per SM streams
// `iters` chunks round-robin from a `src_bytes` buffer; a chunk slot is re-armed as soon as its copy landed.
namespace sparf {
using namespace tc;
__global__ void __launch_bounds__(128, 1) bulkcopy_probe_kernel(const uint8_t* __restrict__ src, uint32_t src_bytes, int stages,
                                                               uint32_t chunk, int iters, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  // issuers: lane 0 of each of the first `nissue` warps (encoded in the high bits of `iters`); issuer w owns the
  // slots s with s % nissue == w
  const int nissue = max(1, (iters >> 24) & 15);
  const int lanes_code = (iters >> 28) & 15;
  const int nlanes = lanes_code == 0 ? 1 : lanes_code * 4;     // lanes per issuer warp, each copies chunk / nlanes bytes
  iters &= 0xFFFFFF;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (w < nissue) {
    const uint32_t nchunks = src_bytes / chunk;
    const uint32_t piece = chunk / nlanes;
    const int active = chunk / piece;
    long long t0 = clock64();
    for (int i = w; i < iters + stages; i += nissue) {
      const int s = i % stages;
      if (i >= stages) mbar_wait(&bars[s], ((i / stages) - 1) & 1);   // previous copy into this slot landed
      if (i < iters) {
        if (lane == 0) mbar_arrive_expect_tx(&bars[s], chunk);
        __syncwarp();
        if (lane < active)
          bulk_g2s(smem + (size_t)s * chunk + lane * piece, src + (size_t)((i + blockIdx.x * 7) % nchunks) * chunk + lane * piece,
                   piece, &bars[s]);
      }
    }
    if (w == 0 && lane == 0) cycles[blockIdx.x] = clock64() - t0;
  }
}
}  // namespace sparf

extern "C" int sparf_tc_bulkcopy_probe(const void* src, uint32_t src_bytes, int32_t stages, uint32_t chunk, int32_t iters,
                                       int32_t grid, long long* cycles, sparf_stream_t stream) {
  SPARF_REQUIRE(stages >= 1 && stages <= 16 && chunk % 16 == 0 && (size_t)stages * chunk <= 200 * 1024, "bulkcopy_probe: bad shape");
  size_t smem = (size_t)stages * chunk + 1024;
  SPARF_CHECK_CUDA(cudaFuncSetAttribute(sparf::bulkcopy_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sparf::bulkcopy_probe_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>((const uint8_t*)src, src_bytes, stages, chunk, iters, cycles);
  SPARF_CHECK_LAUNCH("bulkcopy_probe_kernel");
  return SPARF_OK;
}
