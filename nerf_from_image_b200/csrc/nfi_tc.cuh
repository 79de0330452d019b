// tcgen05 / TMEM / mbarrier / TMA-bulk primitives used by the tensor-core
// variant of the decoder MLP (sm_100a only; hand-written PTX, no CUTLASS).
//
// Layer 1 of the decoder (models/generator.py:294-299: Linear(32->64)) runs as
// D[128 points x 64] = A[128 x 32] * W1^T on the 5th-generation tensor cores
// with fp32 accuracy recovered by the hi/lo split ("3xTF32"):
//     a = a_hi + a_lo,  a_hi = a with the low 13 mantissa bits cleared (exactly
//     representable in TF32), a_lo = a - a_hi (exact in fp32);
//     A*W ~= A_lo*W_hi + A_hi*W_lo + A_hi*W_hi       (error ~2^-21 relative)
// Operands live in shared memory in the canonical K-major SWIZZLE_128B layout
// (a row of 32 fp32 = 128 B = one swizzle row; 8 rows = one 1024-B atom); the
// accumulator lives in TMEM (lane = point, column = hidden unit) and is read
// back with tcgen05.ld for the softplus / layer-2 epilogue.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nfi {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x20000u)  // suspend-time hint (ns): sleep, do not poll
      : "memory");
  return ok != 0;
}
// Waits for the phase with the given parity.  try_wait suspends the thread in
// hardware for a bounded time; a wait that has not succeeded after ~2 s of SM
// clock is a pipeline bug and traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (((++spins) & 63u) == 0 && clock64() - t0 > 4000000000LL) __trap();
  }
}
// Same, with a fixed back-off between polls (timing experiments: NFI_WAIT_NS).
template <int NS>
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  if (NS <= 0) return mbar_wait(bar, parity);
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (true) {
    __nanosleep(NS);
    if (mbar_try_wait(bar, parity)) return;
    if (((++spins) & 63u) == 0 && clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
// TMA bulk copy global -> shared (SASS: UBLKCP), completion on an mbarrier.
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ------------------------------------------------------------- fences
// generic-proxy shared-memory writes -> visible to the async proxy (UMMA reads)
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// named barrier over `count` threads (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void bar_sync(uint32_t id, uint32_t count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// ------------------------------------------------------------- TMEM
// One warp allocates `ncols` (power of two >= 32) and publishes the base
// address through shared memory.
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}

// 16 consecutive fp32 columns of this thread's TMEM lane (lane = 32*(warp%4)+laneid)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// same without the wait: issue several, then tmem_wait_ld() once
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
      "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
      "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
      "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
      "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------- UMMA
// Shared-memory matrix descriptor, K-major, SWIZZLE_128B, fp32 rows of 128 B:
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4 (=1)
//   bits [32,46) stride byte offset >> 4 (1024 B between 8-row groups)
//   bits [46,48) version = 1 (sm_100)      bits [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor, kind::tf32, fp32 accumulate, A and B K-major:
//   [4,6) c_format=1 (F32)  [7,10) a_format=2 (TF32)  [10,13) b_format=2
//   [17,23) N>>3            [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void umma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T : A read from TMEM (lane = row, one 32-bit
// column per K element), no output lane disabled.
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// All MMAs issued so far by this thread arrive on `bar` when they complete.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::
                   "r"(smem_u32(bar))
               : "memory");
}

// byte offset of (row, 16-byte chunk) inside a [rows x 128 B] SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}

// byte offset of (row, 16-byte chunk) inside a [rows x 64 B] SWIZZLE_64B tile (chunk 0..3) and a
// [rows x 32 B] SWIZZLE_32B tile (chunk 0..1): address bits [4,6) / [4,5) XOR bits [7,9) / [7,8)
__device__ __forceinline__ uint32_t sw64_offset(int row, int chunk) {
  return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}
__device__ __forceinline__ uint32_t sw32_offset(int row, int chunk) {
  return (uint32_t)(row * 32 + ((chunk ^ ((row >> 2) & 1)) << 4));
}

__device__ __forceinline__ float tf32_hi(float x) {
  return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
}

// accumulate / overwrite variants without the setp (the flag is known at compile time)
template <bool ACC>
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                        uint32_t idesc) {
  if (ACC)
    asm volatile(
        "{\n.reg .pred p;\nsetp.eq.u32 p, 0, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc)
        : "memory");
  else
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.u32 p, 0, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc)
        : "memory");
}
template <bool ACC>
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                        uint32_t idesc) {
  if (ACC)
    asm volatile(
        "{\n.reg .pred p;\nsetp.eq.u32 p, 0, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%4, %4, %4, %4}, p;\n}\n" ::"r"(
            d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(0u)
        : "memory");
  else
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.u32 p, 0, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%4, %4, %4, %4}, p;\n}\n" ::"r"(
            d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(0u)
        : "memory");
}

// Issues the 12 MMAs of one 128x64x32 3xTF32 product (small terms first).
// Arguments are descriptor BASES (umma_desc_sw128 of the tile start); a K-step
// of 8 tf32 = 32 bytes adds 2 to the start-address field.
__device__ __forceinline__ void issue_layer1_d(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo,
                                               uint64_t w_hi, uint64_t w_lo) {
  constexpr uint32_t idesc = umma_idesc_tf32(128, 64);
  umma_ss<false>(d_tmem, a_lo, w_hi, idesc);
#pragma unroll
  for (int ks = 1; ks < 4; ++ks) umma_ss<true>(d_tmem, a_lo + 2 * ks, w_hi + 2 * ks, idesc);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) umma_ss<true>(d_tmem, a_hi + 2 * ks, w_lo + 2 * ks, idesc);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) umma_ss<true>(d_tmem, a_hi + 2 * ks, w_hi + 2 * ks, idesc);
}
// D2 = H_lo*W2_hi (TS) + H_hi*W2_lo (SS) + H_hi*W2_hi (SS); H_hi k-block 1 is
// 16 KB (1024 descriptor units) after k-block 0, W2 k-block 1 is 2 KB (128) after.
__device__ __forceinline__ void issue_layer2_d(uint32_t d2_tmem, uint32_t hlo_tmem, uint64_t hhi,
                                               uint64_t w2_hi, uint64_t w2_lo) {
  constexpr uint32_t idesc = umma_idesc_tf32(128, 16);
  umma_ts<false>(d2_tmem, hlo_tmem, w2_hi, idesc);
#pragma unroll
  for (int ks = 1; ks < 8; ++ks)
    umma_ts<true>(d2_tmem, hlo_tmem + 8 * ks, w2_hi + (ks >> 2) * 128 + (ks & 3) * 2, idesc);
#pragma unroll
  for (int ks = 0; ks < 8; ++ks)
    umma_ss<true>(d2_tmem, hhi + (ks >> 2) * 1024 + (ks & 3) * 2,
                  w2_lo + (ks >> 2) * 128 + (ks & 3) * 2, idesc);
#pragma unroll
  for (int ks = 0; ks < 8; ++ks)
    umma_ss<true>(d2_tmem, hhi + (ks >> 2) * 1024 + (ks & 3) * 2,
                  w2_hi + (ks >> 2) * 128 + (ks & 3) * 2, idesc);
}

// Issues the 12 MMAs of one 128x64x32 3xTF32 product (small terms first).
// a_hi/a_lo: [128 x 32] tiles, w_hi/w_lo: [64 x 32] tiles (shared addresses).
__device__ __forceinline__ void issue_layer1(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo,
                                             uint32_t w_hi, uint32_t w_lo) {
  constexpr uint32_t idesc = umma_idesc_tf32(128, 64);
  uint32_t acc = 0;
#pragma unroll
  for (int term = 0; term < 3; ++term) {
    const uint32_t a = (term == 0) ? a_lo : a_hi;
    const uint32_t w = (term == 1) ? w_lo : w_hi;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {  // K = 32 = 4 x (8 tf32 = 32 bytes)
      umma_tf32_ss(d_tmem, umma_desc_sw128(a + 32 * ks), umma_desc_sw128(w + 32 * ks), idesc,
                   acc);
      acc = 1;
    }
  }
}

// Layer 2 (64 -> 16 padded outputs), 3xTF32 with the hidden activations split
// as H = H_hi + H_lo: H_hi lives in shared memory (two [128 x 32] K-blocks in
// the space of the layer-1 A tiles), H_lo lives in TMEM where D1 was.
//   D2 = H_lo*W2_hi (TS) + H_hi*W2_lo (SS) + H_hi*W2_hi (SS)
// w2_hi / w2_lo: [16 x 64] fp32 as two [16 x 32] K-blocks of 2048 bytes.
__device__ __forceinline__ void issue_layer2(uint32_t d2_tmem, uint32_t hlo_tmem, uint32_t hhi_k0,
                                             uint32_t hhi_k1, uint32_t w2_hi, uint32_t w2_lo) {
  constexpr uint32_t idesc = umma_idesc_tf32(128, 16);
  uint32_t acc = 0;
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {  // K = 64 = 8 x 8
    const uint32_t wb = w2_hi + (ks >> 2) * 2048 + (ks & 3) * 32;
    umma_tf32_ts(d2_tmem, hlo_tmem + 8 * ks, umma_desc_sw128(wb), idesc, acc);
    acc = 1;
  }
#pragma unroll
  for (int term = 0; term < 2; ++term) {
    const uint32_t w = (term == 0) ? w2_lo : w2_hi;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const uint32_t a = ((ks >> 2) ? hhi_k1 : hhi_k0) + (ks & 3) * 32;
      const uint32_t wb = w + (ks >> 2) * 2048 + (ks & 3) * 32;
      umma_tf32_ss(d2_tmem, umma_desc_sw128(a), umma_desc_sw128(wb), idesc, 1);
    }
  }
}

__device__ __forceinline__ void prefetch_l1(const void* p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ------------------------------------------------------------- 16-bit operands (kind::f16)
// two floats -> packed bf16 pair (round-to-nearest-even), `lo` in the low half (lower address)
__device__ __forceinline__ uint32_t bf16x2_rn(float lo, float hi) {
  uint32_t y;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo));
  return y;
}
// Shared-memory matrix descriptors for 16-bit operands.  `layout`: 2 = SWIZZLE_128B (128-byte
// rows), 4 = SWIZZLE_64B (64-byte rows), 6 = SWIZZLE_32B (32-byte rows).  K-major: rows = M / N,
// `sbo` = bytes between 8-row groups.  MN-major: rows = K, a row holds 64 / 32 / 16 elements of
// M / N; `lbo` = bytes between such atoms along M / N, `sbo` = bytes between 8-row K groups (a
// K16 instruction reads two of them).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t layout, uint32_t lbo,
                                              uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
// Instruction descriptor, kind::f16 with bf16 operands, fp32 accumulate:
//   [4,6) c_format=1 (F32)  [7,10) a_format=1 (BF16)  [10,13) b_format=1
//   [15] A MN-major  [16] B MN-major  [17,23) N>>3  [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <bool ACC>
__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc) {
  if (ACC)
    asm volatile(
        "{\n.reg .pred p;\nsetp.eq.u32 p, 0, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc)
        : "memory");
  else
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.u32 p, 0, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc)
        : "memory");
}

}  // namespace tc
}  // namespace nfi
