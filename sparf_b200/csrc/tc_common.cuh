// Blackwell (sm_100a) primitives used by the tensor-core MLP engine: mbarrier, bulk async copy,
// tcgen05 (TMEM allocation, UMMA descriptors, MMA issue, commit, TMEM loads) and the shared-memory
// operand layout.  Inline PTX only; no CUTLASS dependency.
//
// OPERAND LAYOUT.  Every MMA operand block, in shared memory AND (pre-packed) in global memory, is the
// canonical K-major SWIZZLE_128B image of a [rows x 64] bf16 tile:
//     byte(r, k) = (r / 8) * 1024 + (r % 8) * 128 + (((k / 8) ^ (r % 8)) * 16) + (k % 8) * 2
// i.e. 8-row x 128-byte atoms stacked along M/N with stride 1024 B (SBO), 16-byte chunks XOR-swizzled
// by the row index.  Blocks are 1024-byte aligned.  One tcgen05.mma consumes K = 16 elements = 32 bytes
// of each row; stepping K inside the 128-byte row = adding 32 bytes to the descriptor start address.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace sparf {
namespace tc {

constexpr int kBlockK = 64;                 // bf16 elements per operand row (128 bytes)
constexpr int kAtomBytes = 1024;            // 8 rows x 128 B
constexpr int kUmmaK = 16;                  // K of one tcgen05.mma (kind::f16)

__host__ __device__ __forceinline__ constexpr uint32_t sw128_offset(uint32_t r, uint32_t k) {
  return (r >> 3) * 1024u + (r & 7u) * 128u + ((((k >> 3) ^ (r & 7u))) << 4) + (k & 7u) * 2u;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU box.  try_wait sleeps in hardware between polls,
// so the bound (~2^26 polls, seconds) is never approached by a correct run.  On timeout: flag + trap.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("sparf tc: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x,
             smem_u32(bar), parity);
      __trap();
    }
  }
}

// both of two local barriers (each with its own parity); the two polls are in flight together
__device__ __forceinline__ void mbar_wait_two(uint64_t* bar_a, uint32_t parity_a, uint64_t* bar_b, uint32_t parity_b) {
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q, [%3], %4;\n\t"
        "and.pred p, p, q;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar_a)), "r"(parity_a), "r"(smem_u32(bar_b)), "r"(parity_b)
        : "memory");
    if (ok) return;
    if (++spins > (1u << 26)) {
      printf("sparf tc: double mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

// generic-proxy writes (st.shared) -> visible to the async proxy (tensor core / bulk copy)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------- bulk copy
// 1-D bulk async copy global -> shared, completion signalled on an mbarrier (complete_tx::bytes).
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// 1-D bulk async copy shared -> global (bulk-group completion).  The shared-memory source must have been released to
// the async proxy (fence.proxy.async by its writers) before the issuing thread learns that it is complete.
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed groups have finished READING shared memory (the source may be overwritten)
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// true in exactly one lane of a converged warp.  The MMA / bulk-copy issue loops run warp-uniform and predicate only
// the issuing instruction on this, so their operands (descriptors, addresses) stay in uniform registers; under a
// plain `if (lane == 0)` the compiler emits an ELECT + 5x R2UR "waterfall" loop around every tcgen05.mma.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 64-bit shared-memory matrix descriptor: K-major, SWIZZLE_128B, dense 8-row atoms (SBO = 1024 B).
// (cute::UMMA::SmemDescriptor: start[0,14) | LBO[16,30) | SBO[32,46) | version=1 [46,48) | layout[61,64)=2)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major), canonical 1
  d |= (uint64_t)(kAtomBytes >> 4) << 32;    // stride byte offset between 8-row atoms
  d |= (uint64_t)1 << 46;                    // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}

// MN-major SWIZZLE_128B descriptor: the SAME byte image as above ([rows x 64] with 128-byte rows), but read
// with the 64 contiguous elements of a row as the M/N dimension and the rows as K (used by the weight-
// gradient GEMM dW = G^T X, whose reduction runs over sample rows).  In CUTLASS terms
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units: LBO = byte distance between consecutive 64-element
// groups along M/N (here: between [rows x 64] blocks), SBO = distance between 8-row groups along K = 1024.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(kAtomBytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// 32-bit instruction descriptor, kind::f16, fp32 accumulate.  fmt: 0 = f16, 1 = bf16; major: 0 = K, 1 = MN.
// (cute::UMMA::InstrDescriptor: c_format[4,6)=1 | a_format[7,10) | b_format[10,13) | a_major 15 | b_major 16 |
//  n>>3 [17,23) | m>>4 [24,29))
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int fmt, int a_mn = 0, int b_mn = 0) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// same with different 16-bit formats for A and B (probe: does the hardware accept f16 x bf16 in one kind::f16 MMA?)
__host__ __device__ constexpr uint32_t make_idesc_ab(int M, int N, int a_fmt, int b_fmt, int a_mn = 0, int b_mn = 0) {
  return (1u << 4) | ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// TMEM -> registers: 32 lanes (this warp's quadrant) x 32 consecutive fp32 columns; thread i gets lane i.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// same for 16 consecutive columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// registers -> TMEM: thread i writes 16 consecutive 32-bit columns of lane i (its warp's quadrant)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]^T: the A operand read from tensor memory (lane = row, two 16-bit K elements per column)
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------- splitting
// Error-compensated operand representation v = hi + lo.
//   kF16 (forward): fp16 halves, 11 + 11 significant bits -> |v - hi - lo| <~ 2^-22 |v| (fp32-like); the
//                   conversion saturates at +-65504 (NeRF activations / weights are O(1..1e2)).
//   bf16 (backward): bf16 halves, 8 + 8 bits, 2^-17 relative, full fp32 exponent range for tiny gradients.
template <bool kF16>
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  if (kF16) {
    uint32_t h;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(b), "f"(a));   // low half = a
    hi = h;
    float2 hf = __half22float2(*reinterpret_cast<__half2*>(&h));
    __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    lo = *reinterpret_cast<uint32_t*>(&l);
  } else {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);            // .x = a (low 16 bits), .y = b
    hi = *reinterpret_cast<uint32_t*>(&h);
    float ra = a - __uint_as_float(hi << 16);
    float rb = b - __uint_as_float(hi & 0xFFFF0000u);
    __nv_bfloat162 l = __floats2bfloat162_rn(ra, rb);
    lo = *reinterpret_cast<uint32_t*>(&l);
  }
}
// scalar version for the weight packers: part 0 = hi, 1 = lo; returns the 16-bit pattern
template <bool kF16>
__device__ __forceinline__ uint16_t split1(float v, int part) {
  if (kF16) {
    __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    __half o = part == 0 ? h : __float2half_rn(v - __half2float(h));
    return *reinterpret_cast<uint16_t*>(&o);
  } else {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    __nv_bfloat16 o = part == 0 ? h : __float2bfloat16_rn(v - __bfloat162float(h));
    return *reinterpret_cast<uint16_t*>(&o);
  }
}

}  // namespace tc
}  // namespace sparf
