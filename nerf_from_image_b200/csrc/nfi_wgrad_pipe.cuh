// Decoder-weight gradients on tcgen05 (the GAN generator step,
// run.py:1044: loss.backward() with the decoder trainable).
//
//   dW1[j][c] = sum_points dpre[p][j] F[p][c]      db1[j] = sum_points dpre[p][j]
//   dW2[o][j] = sum_points dOut[p][o] H[p][j]      db2[o] = sum_points dOut[p][o]
//
// are contractions over the POINTS, i.e. over what the other GEMMs of the path keep as the M
// dimension (TMEM lane = point).  tcgen05 reads MN-major ("transposed") operands from shared
// memory -- for 16-bit types with every swizzle; TF32 would need the 32-byte-atomicity layout,
// which a K-major feature tile cannot double as -- so every operand of a step is laid down
// point-major as a bf16 hi/lo PAIR (2^-17: these sums are dominated by a few very large terms,
// single bf16 operands left 1e-3 at every size; with pairs 3e-5 against the fp32 kernel,
// profiles/r2_wgrad_tc_accuracy.txt) and five accumulating kind::f16 MMAs per 16 points form
//
//   XA = [ dpre_hi (64) | dpre_lo (64) ]   XB = [ H_hi (64) | H_lo (64) ]      (A operands, M = 128)
//   [448,480) += XA^T F_hi,  += XA^T F_lo      rows j and 64 + j together: dW1[j][.]
//   [480,496) += XB^T dOut_hi, += XB^T dOut_lo   rows j and 64 + j together: dW2[.][j]
//   [496,504) += XA^T 1                          db1[j]
//
// in 56 TMEM columns, cut into chains of kWgFlush steps that are banked in fp32 (below) and added
// to the global gradients once per CTA.  db2 is summed in fp32 registers by the shading threads.
//
// To make room for the pair tiles, layer 1 of the recompute chain runs on bf16 pairs too (as the
// synthesis convolutions do): the gather writes the features as two [128][32 bf16] SWIZZLE_64B
// tiles (16 KB per stage instead of 32) that are BOTH the K-major A operand of MMA1 and, rows = K,
// the MN-major B operand of the dW1 products.  The rest of the chain is render_backward_pipe's
// (MMA2 / MMA3 in 3xTF32, softplus and its reverse, reverse compositing).  Two forms:
//   PLANES = true  (the generator step: decoder AND planes / palette / beta / alpha gradients,
//                   cameras are data): MMA4 and the scatter of render_backward_pipe are folded
//                   in -- ONE sweep, 26.0 ms per 32 images at config 2 against 21.7 + 16.7 ms of
//                   the two kernels (profiles/r2_time_wgrad.txt);
//   PLANES = false (decoder gradients only, or a pose gradient is wanted too): this kernel beside
//                   render_backward_pipe, which then sees the decoder as a constant.
// TMEM: two slots of 224 columns
//   [0,64) D1 -> H_lo (-> D4)   [64,128) H_hi (-> dpre_hi)   [128,144) D2 -> dOut_hi
//   [144,160) dOut_lo   [160,224) D3 (-> dpre_lo)
// and the accumulators at [448,504).
#pragma once
#include <cuda_bf16.h>

#include "nfi_backward_pipe.cuh"

namespace nfi {

constexpr int kWgSlots = 2;
constexpr int kWgSlotCols = 224;
constexpr int kWgAccCol = 448;
constexpr int kWgStages = 3;
// The tensor core adds into a TMEM accumulator with truncation, a bias that grows with the length
// of the chain (finding 3: 1.35e-4 after 4,608 terms in the synthesis
// convolutions); a CTA of config 2 would chain 28,000 accumulating MMAs = 450,000 terms.  The
// chain is therefore cut every kWgFlush steps (2,048 terms): the accumulator is added, in fp32
// round-to-nearest, to a per-CTA row buffer in global memory (L2-resident, single writer per
// element) and the next chain starts from zero.
constexpr int kWgFlush = 16;
constexpr size_t kWgAccBytesPerCta = 128 * 64 * sizeof(float);

// PLANES: the kernel also forms dL/d(texel features) (MMA4) and scatters it into the plane
// gradient, and accumulates the palette / beta / alpha gradients -- i.e. it is the WHOLE backward
// of the GAN generator step in one sweep (no pose gradient: cameras are data there).
template <bool PLANES>
struct WgCfgT {
  static constexpr int P = 2;
  static constexpr int kThreadsTotal = 384 + 128 * P;
  static constexpr int kStageBytes = 16384;                           // F_hi | F_lo, 8 KB each
  static constexpr int kWbBytes = PLANES ? 32768 : 16384;             // W2^T hi/lo (+ W1^T/3 hi/lo)
  static constexpr int kSmWb = 25600;
  static constexpr int kSmW1 = kSmWb + kWbBytes;                      // W1 bf16 hi | lo (8 KB)
  static constexpr int kSmA = kSmW1 + 8192;
  static constexpr int kSmX = kSmA + kWgStages * kStageBytes;         // 4 x 16 KB
  static constexpr int kSmE = kSmX + 65536;                           // dOut hi | lo
  static constexpr int kSmOnes = kSmE + 8192;                         // 1 KB of 1.0
  static constexpr int kStageWarp = 32 * 36 * 4;                      // per-warp scatter staging
  static constexpr int kSmStage = kSmOnes + 1024;
  static constexpr int kSmPal = kSmStage + (PLANES ? P * 4 * kStageWarp : 0);
  static constexpr int kSmFrac = kSmPal + 48 * 4;
  static constexpr int kSmBars = kSmFrac + 128 * 4;
  // full[3], a_free[3], per slot: d1_full, h_ready, d2_full, dout_ready, d3_full, slot_free,
  // dpre_ready, d4_full; x_ready, x_free; weights x2
  static constexpr int kNumBars = 2 * kWgStages + 8 * kWgSlots + 2 + 2;
  static constexpr int kSmTmemPtr = kSmBars + kNumBars * 8;
  static constexpr int kSmBytes = kSmTmemPtr + 16;
  // 640 x 96 = 61440 = 256 x 96 + 128 x (128 + 136 + 24) = 256 x 112 + 128 x (96 + 136 + 24)
  static constexpr int kProdRegs = PLANES ? 112 : 96;
  static constexpr int kActRegs = PLANES ? 96 : 128;
  static constexpr int kShadeRegs = 136;
  static constexpr int kAuxRegs = 24;
};
using WgCfg = WgCfgT<false>;
static_assert(WgCfgT<true>::kSmBytes <= 227 * 1024, "shared memory budget");
static_assert((WgCfgT<false>::kSmA & 1023) == 0 && (WgCfgT<false>::kSmX & 1023) == 0 &&
                  (WgCfgT<false>::kSmE & 1023) == 0 && (WgCfgT<false>::kSmW1 & 1023) == 0 &&
                  (WgCfgT<true>::kSmA & 1023) == 0 && (WgCfgT<true>::kSmX & 1023) == 0 &&
                  (WgCfgT<true>::kSmE & 1023) == 0 && (WgCfgT<true>::kSmW1 & 1023) == 0,
              "swizzled tiles are 1024-byte aligned");


template <int NOUT_PAD, bool PLANES = false>
__global__ void __launch_bounds__(WgCfgT<PLANES>::kThreadsTotal, 1)
render_wgrad_pipe(const nfi_render_params p, const nfi_render_grads g,
                  const unsigned char* __restrict__ wimg, float* __restrict__ acc_ws) {
  using Cfg = WgCfgT<PLANES>;
  constexpr int P = Cfg::P;
  constexpr int NA = NOUT_PAD - 1;
  constexpr int NS = kWgStages;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = smem_raw;
  const int tid = threadIdx.x, lane = tid & 31;
  const int hw_wg = __shfl_sync(kFull, tid >> 7, 0);
  // logical role: 0 activation, 1 shading, 2 MMA issuers, 3.. producer sets
  const int wg = (hw_wg < P) ? hw_wg + 3 : (P + 2 - hw_wg);
  const int gt = tid & 127;
  const int wig = __shfl_sync(kFull, gt >> 5, 0);
  const int S = p.num_samples;
  const int n_total = (p.fine_sampling ? 2 : 1) * S;  // steps per tile

  uint64_t* bars = reinterpret_cast<uint64_t*>(base + Cfg::kSmBars);
  uint64_t* full = bars;                      // [3] stage gathered (4 warps)
  uint64_t* a_free = full + NS;               // [3] stage read by the dW MMAs (commit)
  uint64_t* d1_full = a_free + NS;            // [2] commit
  uint64_t* h_ready = d1_full + kWgSlots;     // [2] 4 warps
  uint64_t* d2_full = h_ready + kWgSlots;     // [2] commit
  uint64_t* dout_ready = d2_full + kWgSlots;  // [2] 4 warps (dOut in TMEM and in the E tiles)
  uint64_t* d3_full = dout_ready + kWgSlots;  // [2] commit
  uint64_t* slot_free = d3_full + kWgSlots;   // [2] 4 warps (D3 / H read)
  uint64_t* dpre_ready = slot_free + kWgSlots;  // [2] 4 warps (PLANES: dpre hi/lo in TMEM)
  uint64_t* d4_full = dpre_ready + kWgSlots;    // [2] commit (PLANES)
  uint64_t* x_ready = d4_full + kWgSlots;     // X tile written (4 warps)
  uint64_t* x_free = x_ready + 1;             // X and E tiles read by the dW MMAs (commit)
  uint64_t* wbar = x_free + 1;                // [2] weight images landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(base + Cfg::kSmTmemPtr);
  const float* b1s = reinterpret_cast<const float*>(base + kWiB1);
  const float* b2s = reinterpret_cast<const float*>(base + kWiB2);
  float* pal = reinterpret_cast<float*>(base + Cfg::kSmPal);
  float* frac = reinterpret_cast<float*>(base + Cfg::kSmFrac);
  if (tid < 128) frac[tid] = (float)tid / (float)S;
  // constants in shared memory: 1 KB of bf16 ones (the B operand of db1) and layer 1's weights
  // (x log2 e, as the forward weight image has them) as a bf16 hi / lo pair, [64 rows = hidden
  // unit][32 k] K-major SWIZZLE_64B
  for (int i = tid; i < 256; i += Cfg::kThreadsTotal)
    reinterpret_cast<uint32_t*>(base + Cfg::kSmOnes)[i] = 0x3F803F80u;
  for (int i = tid; i < kHid * kC; i += Cfg::kThreadsTotal) {
    const int j = i / kC, c = i % kC;
    const float w = p.w1[i] * kLog2e;
    const __nv_bfloat16 hi = __float2bfloat16_rn(w);
    const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
    const uint32_t off = tc::sw64_offset(j, c >> 3) + (c & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(base + Cfg::kSmW1 + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(base + Cfg::kSmW1 + 4096 + off) = lo;
  }
  tc::fence_async_smem();
  if (tid == 0) {
    if (tc::smem_u32(base) & 1023u) __trap();
    for (int i = 0; i < NS; ++i) {
      tc::mbar_init(&full[i], 4);
      tc::mbar_init(&a_free[i], 1);
    }
    for (int i = 0; i < kWgSlots; ++i) {
      tc::mbar_init(&d1_full[i], 1);
      tc::mbar_init(&h_ready[i], kWarps);
      tc::mbar_init(&d2_full[i], 1);
      tc::mbar_init(&dout_ready[i], kWarps);
      tc::mbar_init(&d3_full[i], 1);
      tc::mbar_init(&slot_free[i], kWarps);
      tc::mbar_init(&dpre_ready[i], kWarps);
      tc::mbar_init(&d4_full[i], 1);
    }
    tc::mbar_init(x_ready, kWarps);
    tc::mbar_init(x_free, 1);
    tc::mbar_init(&wbar[0], 1);
    tc::mbar_init(&wbar[1], 1);
    tc::fence_mbar_init();
  }
  if (tid < 32) tc::tmem_alloc(tmem_ptr, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(kFull, *tmem_ptr, 0);
  if (tid == 0) {
    tc::mbar_expect_tx(&wbar[0], kWiBytes);
    tc::tma_bulk_g2s(base, wimg, kWiBytes, &wbar[0]);
    tc::mbar_expect_tx(&wbar[1], Cfg::kWbBytes);
    tc::tma_bulk_g2s(base + Cfg::kSmWb, wimg + 32768, Cfg::kWbBytes, &wbar[1]);  // W2^T (| W1^T / 3)
  }
  tc::mbar_wait(&wbar[0], 0);
  tc::mbar_wait(&wbar[1], 0);

  const uint32_t base_s = tc::smem_u32(base);
  const int tiles_x = (p.width + kTileW - 1) / kTileW;
  const int tiles_y = (p.height + kTileH - 1) / kTileH;
  const int n_tiles = tiles_x * tiles_y * p.batch;
  const int my_tiles =
      ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const uint32_t total_steps = (uint32_t)my_tiles * (uint32_t)n_total;
  const int R = p.plane_res;
  const float inv_range = 1.f / p.scene_range;
  const uint32_t plane_bytes = (uint32_t)R * (uint32_t)R * 128u;
  const uint32_t lane_addr = (uint32_t)(32 * wig) << 16;

  if (wg >= 3) {
    // ================================ PRODUCERS (gather; PLANES: also the scatter) ================
    if (PLANES) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::kProdRegs));
    const int set = wg - 3;
    float* Dw = reinterpret_cast<float*>(base + Cfg::kSmStage + (set * 4 + wig) * Cfg::kStageWarp);
    const int q = lane >> 3, kq = lane & 7;
    const uint32_t row_units = (uint32_t)R * 8u;
    uint32_t n0 = 0;  // ring position of the tile's first step
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, n0 += (uint32_t)n_total) {
      const TileCoord tcd = tile_coord(tile, tiles_x, tiles_y);
      const int b = tcd.b;
      int px, py;
      tile_pixel(tcd.tile_x, tcd.tile_y, wig, lane, px, py);
      px = min(px, p.width - 1);
      py = min(py, p.height - 1);
      const size_t ray = ((size_t)b * p.height + py) * p.width + px;
      Ray r;
      setup_ray(p, b, py, px, r);
      const unsigned char* planes_b =
          reinterpret_cast<const unsigned char*>(p.planes) + (size_t)b * 3 * plane_bytes;
      float* gplanes_b = (PLANES && g.grad_planes)
                             ? g.grad_planes + (size_t)b * 3 * (plane_bytes >> 2) : nullptr;
      MergeWalk mw;
      mw.init(p, r, ray, frac);
      int walked = 0;
      ByteTaps cur, nxt;  // taps of the (up to) two steps this set has between "gathered" and "scattered"
      auto gather_step = [&](int i, ByteTaps& tp) {
        float z = 0.f;
        while (walked <= i) {
          z = mw.pop();
          ++walked;
        }
        const float x0 = (r.ox + r.dx * z) * inv_range, x1 = (r.oy + r.dy * z) * inv_range,
                    x2 = (r.oz + r.dz * z) * inv_range;
        byte_taps(x0, x1, R, 0u, tp.o[0], tp.fx[0], tp.fy[0]);
        byte_taps(x0, x2, R, plane_bytes >> 4, tp.o[1], tp.fx[1], tp.fy[1]);
        byte_taps(x1, x2, R, plane_bytes >> 3, tp.o[2], tp.fx[2], tp.fy[2]);
        const uint32_t m = n0 + (uint32_t)i;
        const uint32_t st = m % NS, u = m / NS;
        unsigned char* const stage = base + Cfg::kSmA + st * Cfg::kStageBytes;
        NFI_STEP_WAIT(&a_free[st], (u & 1) ^ 1);
        gather_to_tiles_lean<true>(planes_b, R, tp, stage, stage + 8192, 32 * wig, lane);
        tc::fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[st]);
      };
      if constexpr (!PLANES) {
        for (int i = set; i < n_total; i += P) gather_step(i, cur);
      } else {
        if (set < n_total) gather_step(set, cur);
        for (int i = set; i < n_total; i += P) {
          if (i + P < n_total) gather_step(i + P, nxt);
          // ---- scatter step i: D4 -> plane gradient (as render_backward_pipe, no pose gradient)
          const uint32_t m = n0 + (uint32_t)i;
          const uint32_t sl = m % kWgSlots, v = m / kWgSlots;
          const uint32_t d4 = tmem_base + sl * kWgSlotCols + lane_addr;
          NFI_STEP_WAIT(&d4_full[sl], v & 1);
          tc::tc_fence_after();
          {
            uint32_t ra[16], rb[16];
            tc::tmem_ld16_nowait(d4, ra);
            tc::tmem_ld16_nowait(d4 + 16, rb);
            tc::tmem_wait_ld();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&slot_free[sl]);
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              *reinterpret_cast<float4*>(Dw + lane * 36 + 4 * c4) =
                  make_float4(__uint_as_float(ra[4 * c4]), __uint_as_float(ra[4 * c4 + 1]),
                              __uint_as_float(ra[4 * c4 + 2]), __uint_as_float(ra[4 * c4 + 3]));
              *reinterpret_cast<float4*>(Dw + lane * 36 + 16 + 4 * c4) =
                  make_float4(__uint_as_float(rb[4 * c4]), __uint_as_float(rb[4 * c4 + 1]),
                              __uint_as_float(rb[4 * c4 + 2]), __uint_as_float(rb[4 * c4 + 3]));
            }
          }
          __syncwarp();
          if (gplanes_b != nullptr) {
#pragma unroll 1
            for (int gq = 0; gq < 8; ++gq) {
              const int src = 4 * gq + q;
              const float4 d4v = *reinterpret_cast<const float4*>(Dw + src * 36 + 4 * kq);
#pragma unroll
              for (int pl = 0; pl < 3; ++pl) {
                const uint32_t a00 = __shfl_sync(kFull, cur.o[pl], src) | (uint32_t)kq;
                const float fx = __shfl_sync(kFull, cur.fx[pl], src);
                const float fy = __shfl_sync(kFull, cur.fy[pl], src);
                const float gx0 = 1.f - fx, gy0 = 1.f - fy;
                const float w00 = gx0 * gy0, w01 = fx * gy0, w10 = gx0 * fy, w11 = fx * fy;
                float* gp = gplanes_b;
                red_add_v4(gp + (size_t)a00 * 4, d4v.x * w00, d4v.y * w00, d4v.z * w00, d4v.w * w00);
                red_add_v4(gp + (size_t)(a00 + 8u) * 4, d4v.x * w01, d4v.y * w01, d4v.z * w01,
                           d4v.w * w01);
                red_add_v4(gp + (size_t)(a00 + row_units) * 4, d4v.x * w10, d4v.y * w10,
                           d4v.z * w10, d4v.w * w10);
                red_add_v4(gp + (size_t)(a00 + row_units + 8u) * 4, d4v.x * w11, d4v.y * w11,
                           d4v.z * w11, d4v.w * w11);
              }
            }
          }
          __syncwarp();
          cur = nxt;
        }
      }
    }
  } else if (wg == 2) {
    // ================================ MMA ISSUERS ================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::kAuxRegs));
    if (wig == 0) {
      // layer 1 on bf16 pairs: D1 = F_lo W1_hi + F_hi W1_lo + F_hi W1_hi, K = 32 = two K16 steps
      constexpr uint32_t idesc1 = tc::umma_idesc_bf16(128, 64, false, false);
      const uint64_t dsc_w1_hi = tc::umma_desc(base_s + Cfg::kSmW1, 4, 16, 512);
      const uint64_t dsc_w1_lo = tc::umma_desc(base_s + Cfg::kSmW1 + 4096, 4, 16, 512);
      uint32_t st = 0, u = 0, sl = 0, v = 0;
      for (uint32_t m = 0; m < total_steps; ++m) {
        NFI_STEP_WAIT(&full[st], u & 1);
        NFI_STEP_WAIT(&slot_free[sl], (v & 1) ^ 1);
        if (elect_one()) {
          tc::tc_fence_after();
          const uint32_t stage_s = base_s + Cfg::kSmA + st * Cfg::kStageBytes;
          const uint64_t a_hi = tc::umma_desc(stage_s, 4, 16, 512);
          const uint64_t a_lo = tc::umma_desc(stage_s + 8192, 4, 16, 512);
          const uint32_t d1 = tmem_base + sl * kWgSlotCols;
          tc::umma_f16_ss<false>(d1, a_lo, dsc_w1_hi, idesc1);
          tc::umma_f16_ss<true>(d1, a_lo + 2, dsc_w1_hi + 2, idesc1);
          tc::umma_f16_ss<true>(d1, a_hi, dsc_w1_lo, idesc1);
          tc::umma_f16_ss<true>(d1, a_hi + 2, dsc_w1_lo + 2, idesc1);
          tc::umma_f16_ss<true>(d1, a_hi, dsc_w1_hi, idesc1);
          tc::umma_f16_ss<true>(d1, a_hi + 2, dsc_w1_hi + 2, idesc1);
          tc::umma_commit(&d1_full[sl]);
        }
        __syncwarp();
        if (++st == NS) { st = 0; ++u; }
        if (++sl == kWgSlots) { sl = 0; ++v; }
      }
    } else if (wig == 1) {
      const uint64_t dsc_w2_hi = tc::umma_desc_sw128(base_s + kWiW2Hi);
      const uint64_t dsc_w2_lo = tc::umma_desc_sw128(base_s + kWiW2Lo);
      uint32_t sl = 0, v = 0;
      for (uint32_t m = 0; m < total_steps; ++m) {
        NFI_STEP_WAIT(&h_ready[sl], v & 1);
        if (elect_one()) {
          tc::tc_fence_after();
          const uint32_t d = tmem_base + sl * kWgSlotCols;
          issue_layer2_tt(d + 128, d, d + 64, dsc_w2_hi, dsc_w2_lo);
          tc::umma_commit(&d2_full[sl]);
        }
        __syncwarp();
        if (++sl == kWgSlots) { sl = 0; ++v; }
      }
    } else if (wig == 2) {
      const uint64_t b_hi = tc::umma_desc_sw128(base_s + Cfg::kSmWb + kWbW2tHi);
      const uint64_t b_lo = tc::umma_desc_sw128(base_s + Cfg::kSmWb + kWbW2tLo);
      uint32_t sl = 0, v = 0;
      for (uint32_t m = 0; m < total_steps; ++m) {
        NFI_STEP_WAIT(&dout_ready[sl], v & 1);
        if (elect_one()) {
          tc::tc_fence_after();
          const uint32_t d = tmem_base + sl * kWgSlotCols;
          issue_mma3(d + 160, d + 128, d + 144, b_hi, b_lo);
          tc::umma_commit(&d3_full[sl]);
        }
        __syncwarp();
        if (++sl == kWgSlots) { sl = 0; ++v; }
      }
    } else {
      // the weight-gradient chain: 8 k-steps of 16 points, five MMAs each
      constexpr uint32_t id_w1 = tc::umma_idesc_bf16(128, 32, true, true);
      constexpr uint32_t id_w2 = tc::umma_idesc_bf16(128, 16, true, true);
      constexpr uint32_t id_b1 = tc::umma_idesc_bf16(128, 8, true, false);
      const uint64_t dsc_xa = tc::umma_desc(base_s + Cfg::kSmX, 2, 16384, 1024);
      const uint64_t dsc_xb = tc::umma_desc(base_s + Cfg::kSmX + 32768, 2, 16384, 1024);
      const uint64_t dsc_ehi = tc::umma_desc(base_s + Cfg::kSmE, 6, 16, 256);
      const uint64_t dsc_elo = tc::umma_desc(base_s + Cfg::kSmE + 4096, 6, 16, 256);
      const uint64_t dsc_one = tc::umma_desc(base_s + Cfg::kSmOnes, 6, 16, 256);  // K-major [8][16]
      const uint32_t d_w1 = tmem_base + kWgAccCol, d_w2 = d_w1 + 32, d_b1 = d_w1 + 48;
      const uint64_t w1t_hi = tc::umma_desc_sw128(base_s + Cfg::kSmWb + kWbW1tHi);
      const uint64_t w1t_lo = tc::umma_desc_sw128(base_s + Cfg::kSmWb + kWbW1tLo);
      uint32_t st = 0, u = 0, sl = 0, v = 0;
      for (uint32_t m = 0; m < total_steps; ++m) {
        if constexpr (PLANES) {  // MMA4: D4 = dpre (W1 / 3) over D1 / H_lo, 3xTF32 as render_backward_pipe
          NFI_STEP_WAIT(&dpre_ready[sl], v & 1);
          if (elect_one()) {
            tc::tc_fence_after();
            const uint32_t d = tmem_base + sl * kWgSlotCols;
            issue_mma4(d, d + 64, d + 160, w1t_hi, w1t_lo);
            tc::umma_commit(&d4_full[sl]);
          }
          __syncwarp();
        }
        NFI_STEP_WAIT(&full[st], u & 1);
        NFI_STEP_WAIT(&dout_ready[sl], v & 1);
        NFI_STEP_WAIT(x_ready, m & 1);
        if (elect_one()) {
          tc::tc_fence_after();
          const uint32_t stage_s = base_s + Cfg::kSmA + st * Cfg::kStageBytes;
          const uint64_t dsc_fhi = tc::umma_desc(stage_s, 4, 16, 512);        // MN-major: rows = K
          const uint64_t dsc_flo = tc::umma_desc(stage_s + 8192, 4, 16, 512);
          const bool fresh = (m % kWgFlush == 0);  // new chain: the first MMA into each accumulator overwrites
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {  // 16 points: 2048 B of X, 1024 B of F, 512 B of dOut
            const uint64_t xa = dsc_xa + 128 * ks, xb = dsc_xb + 128 * ks;
            if (ks == 0 && fresh) {
              tc::umma_f16_ss<false>(d_w1, xa, dsc_fhi, id_w1);
              tc::umma_f16_ss<false>(d_w2, xb, dsc_ehi, id_w2);
              tc::umma_f16_ss<false>(d_b1, xa, dsc_one, id_b1);
            } else {
              tc::umma_f16_ss<true>(d_w1, xa, dsc_fhi + 64 * ks, id_w1);
              tc::umma_f16_ss<true>(d_w2, xb, dsc_ehi + 32 * ks, id_w2);
              tc::umma_f16_ss<true>(d_b1, xa, dsc_one, id_b1);
            }
            tc::umma_f16_ss<true>(d_w1, xa, dsc_flo + 64 * ks, id_w1);
            tc::umma_f16_ss<true>(d_w2, xb, dsc_elo + 32 * ks, id_w2);
          }
          tc::umma_commit(&a_free[st]);
          tc::umma_commit(x_free);
        }
        __syncwarp();
        if (++st == NS) { st = 0; ++u; }
        if (++sl == kWgSlots) { sl = 0; ++v; }
      }
    }
  } else if (wg == 0) {
    // ================================ ACTIVATION (forward and reverse) ================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::kActRegs));
    unsigned char* const xrow = base + Cfg::kSmX;
    // Accumulator row of this thread (TMEM lane gt), hidden unit j = gt % 64 (rows 64.. are the
    // products of the lo halves): columns 0..31 dW1[j][.], 32..47 dW2[.][j], 48 db1[j].
    // `first`: the row buffer is written, not added to; `final`: the total goes to the global
    // gradients instead.
    float* const myrow = acc_ws + ((size_t)blockIdx.x * 128 + gt) * 64;
    const int nout_w = 1 + (p.n_attention > 0 ? p.n_attention : 3);
    auto flush = [&](bool first, bool final) {
      const uint32_t dacc = tmem_base + kWgAccCol + lane_addr;
      const int j = gt & (kHid - 1);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float v16[16];
        tc::tmem_ld16(dacc + 16 * c, v16);  // (c == 3: only column 48 is an accumulator)
        float4* row4 = reinterpret_cast<float4*>(myrow + 16 * c);
        if (!first) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (c == 3 && q > 0) break;
            const float4 o = row4[q];
            v16[4 * q] += o.x;
            v16[4 * q + 1] += o.y;
            v16[4 * q + 2] += o.z;
            v16[4 * q + 3] += o.w;
          }
        }
        if (!final) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (c == 3 && q > 0) break;
            row4[q] = make_float4(v16[4 * q], v16[4 * q + 1], v16[4 * q + 2], v16[4 * q + 3]);
          }
        } else if (c < 2) {
          if (g.grad_w1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) atomicAdd(g.grad_w1 + j * kC + 16 * c + i, v16[i]);
          }
        } else if (c == 2) {
          if (g.grad_w2) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i < nout_w) atomicAdd(g.grad_w2 + i * kHid + j, v16[i]);
          }
        } else if (g.grad_b1) {
          atomicAdd(g.grad_b1 + j, v16[0]);
        }
      }
    };
    auto act_fwd = [&](uint32_t m) {
      const uint32_t sl = m % kWgSlots, v = m / kWgSlots;
      const uint32_t d1 = tmem_base + sl * kWgSlotCols + lane_addr;
      NFI_STEP_WAIT(&d1_full[sl], v & 1);
      tc::tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float lo[16], hi[16];
        tc::tmem_ld16(d1 + 16 * c, lo);
        softplus_split16(lo, hi, b1s + 16 * c);
        tc::tmem_st16(d1 + 16 * c, lo);
        tc::tmem_st16(d1 + 64 + 16 * c, hi);
      }
      tc::tmem_wait_st();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&h_ready[sl]);
    };
    // dpre = dH * sigmoid(pre),  sigmoid(pre) = 1 - exp(-softplus(pre)); dpre and H go to the
    // X tiles as bf16 hi / lo pairs (row = point = this thread, MN-major atoms of 64 columns)
    auto act_bwd = [&](uint32_t m) {
      const uint32_t sl = m % kWgSlots, v = m / kWgSlots;
      const uint32_t d = tmem_base + sl * kWgSlotCols + lane_addr;
      NFI_STEP_WAIT(&d3_full[sl], v & 1);
      NFI_STEP_WAIT(x_free, (m & 1) ^ 1);  // the previous step's dW MMAs have read the X tile
      tc::tc_fence_after();
      // the chain of steps [m - kWgFlush, m) is complete and the next one cannot start before
      // this step's x_ready: the accumulator is quiescent -> bank it
      if (m > 0 && m % kWgFlush == 0) flush(m == kWgFlush, false);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r3[16], rh[16], rl[16];
        tc::tmem_ld16_nowait(d + 160 + 16 * c, r3);
        tc::tmem_ld16_nowait(d + 64 + 16 * c, rh);
        tc::tmem_ld16_nowait(d + 16 * c, rl);
        tc::tmem_wait_ld();
        float dp[16], hh[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float h = __uint_as_float(rh[i]) + __uint_as_float(rl[i]);
          const float sg = 1.f - tc::ex2_approx(-h * kLog2e);
          dp[i] = __uint_as_float(r3[i]) * sg;
          hh[i] = h;
        }
        if constexpr (PLANES) {  // dpre as a TF32 pair for MMA4: hi over H_hi, lo over D3
          float thi[16], tlo[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            thi[i] = tc::tf32_hi(dp[i]);
            tlo[i] = dp[i] - thi[i];
          }
          tc::tmem_st16(d + 160 + 16 * c, tlo);
          tc::tmem_st16(d + 64 + 16 * c, thi);
        }
        // columns 16c .. 16c+15 of dpre_hi / dpre_lo (tiles 0, 1) and H_hi / H_lo (tiles 2, 3):
        // two 16-byte chunks of this point's 128-byte row in each
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const uint32_t off = tc::sw128_offset(gt, 2 * c + q);
          uint32_t dh[4], dl[4], hh2[4], hl[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float a0 = dp[8 * q + 2 * i], a1 = dp[8 * q + 2 * i + 1];
            dh[i] = tc::bf16x2_rn(a0, a1);
            dl[i] = tc::bf16x2_rn(a0 - __uint_as_float(dh[i] << 16),
                                  a1 - __uint_as_float(dh[i] & 0xFFFF0000u));
            const float b0 = hh[8 * q + 2 * i], b1v = hh[8 * q + 2 * i + 1];
            hh2[i] = tc::bf16x2_rn(b0, b1v);
            hl[i] = tc::bf16x2_rn(b0 - __uint_as_float(hh2[i] << 16),
                                  b1v - __uint_as_float(hh2[i] & 0xFFFF0000u));
          }
          *reinterpret_cast<uint4*>(xrow + off) = make_uint4(dh[0], dh[1], dh[2], dh[3]);
          *reinterpret_cast<uint4*>(xrow + 16384 + off) = make_uint4(dl[0], dl[1], dl[2], dl[3]);
          *reinterpret_cast<uint4*>(xrow + 32768 + off) = make_uint4(hh2[0], hh2[1], hh2[2], hh2[3]);
          *reinterpret_cast<uint4*>(xrow + 49152 + off) = make_uint4(hl[0], hl[1], hl[2], hl[3]);
        }
      }
      if constexpr (PLANES) tc::tmem_wait_st();
      tc::tc_fence_before();
      tc::fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(x_ready);
        if constexpr (PLANES) mbar_arrive(&dpre_ready[sl]);  // (the slot is freed by the scatter)
        else mbar_arrive(&slot_free[sl]);
      }
    };
    if (total_steps > 0) act_fwd(0);
    for (uint32_t m = 0; m < total_steps; ++m) {
      if (m + 1 < total_steps) act_fwd(m + 1);
      act_bwd(m);
    }
    // ---- the last chain + the banked ones -> global gradients (one atomic per entry per CTA)
    if (total_steps > 0) {
      NFI_STEP_WAIT(x_free, (total_steps - 1) & 1);
      tc::tc_fence_after();
      flush(total_steps <= (uint32_t)kWgFlush, true);
      tc::tc_fence_before();
    }
  } else {
    // ================================ SHADING (forward and reverse) ================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::kShadeRegs));
    FieldConst fc;
    fc.A = p.n_attention;
    fc.use_sdf = p.use_sdf;
    const float beta = p.use_sdf ? p.beta[0] : 1.f;
    fc.inv_beta = p.use_sdf ? 1.f / beta : 0.f;
    fc.inv_alpha = p.use_sdf ? 1.f / p.alpha[0] : 0.f;
    float acc_b2[NOUT_PAD];
#pragma unroll
    for (int o = 0; o < NOUT_PAD; ++o) acc_b2[o] = 0.f;
    uint32_t m = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const TileCoord tcd = tile_coord(tile, tiles_x, tiles_y);
      const int b = tcd.b;
      int px, py;
      tile_pixel(tcd.tile_x, tcd.tile_y, wig, lane, px, py);
      const bool valid = (px < p.width) && (py < p.height);
      px = min(px, p.width - 1);
      py = min(py, p.height - 1);
      const size_t ray = ((size_t)b * p.height + py) * p.width + px;
      Ray r;
      setup_ray(p, b, py, px, r);
      tc::bar_sync(1, kThreads);
      if (gt < 48)
        pal[gt] = (p.n_attention > 0 && gt < p.n_attention * 3)
                      ? p.palette[(size_t)b * p.n_attention * 3 + gt]
                      : 0.f;
      tc::bar_sync(1, kThreads);

      // upstream gradients of this ray (zero for padding lanes)
      const float vz = valid ? 1.f : 0.f;
      const float g_r = vz * g.g_rgb[ray * 3 + 0], g_g = vz * g.g_rgb[ray * 3 + 1],
                  g_b = vz * g.g_rgb[ray * 3 + 2];
      float g_m = (g.g_mask ? vz * g.g_mask[ray] : 0.f);
      const float out_m = g.out_mask[ray];
      float o_r = g.out_rgb[ray * 3 + 0], o_g = g.out_rgb[ray * 3 + 1],
            o_b = g.out_rgb[ray * 3 + 2];
      if (p.white_background) {
        g_m -= (g_r + g_g + g_b);
        const float bg = 1.f - out_m;
        o_r -= bg;
        o_g -= bg;
        o_b -= bg;
      }
      const float total = (g_r * o_r + g_g * o_g + g_b * o_b) + g_m * out_m;
      float accP[PLANES ? NA : 1];  // PLANES: palette / beta / alpha gradients of this ray
#pragma unroll
      for (int a = 0; a < (PLANES ? NA : 1); ++a) accP[a] = 0.f;
      float acc_beta = 0.f, acc_alpha = 0.f;

      MergeWalk mw;
      mw.init(p, r, ray, frac);
      float z = mw.pop();
      float T = 1.f, prefix = 0.f;
      for (int i = 0; i < n_total; ++i, ++m) {
        const bool has_next = (i + 1 < n_total);
        const float zn = has_next ? mw.pop() : z;
        const float delta = has_next ? (zn - z) * r.dn : 0.f;
        const uint32_t sl = m % kWgSlots, v = m / kWgSlots;
        const uint32_t d = tmem_base + sl * kWgSlotCols + lane_addr;
        const float wx = r.ox + r.dx * z, wy = r.oy + r.dy * z, wz = r.oz + r.dz * z;
        const float x0 = wx * inv_range, x1 = wy * inv_range, x2 = wz * inv_range;
        const float keep = (fabsf(x0) > 1.f || fabsf(x1) > 1.f || fabsf(x2) > 1.f) ? 0.f : 1.f;
        NFI_STEP_WAIT(&d2_full[sl], v & 1);
        tc::tc_fence_after();
        float o16[16];
        tc::tmem_ld16(d + 128, o16);
        float out[NOUT_PAD];
#pragma unroll
        for (int o = 0; o < NOUT_PAD; ++o) out[o] = o16[o] + b2s[o];
        // density (models/generator.py:629-636) and its derivative wrt out[0]
        float sigma, dsig_dout0, nd = 0.f, e_sdf = 0.f, sg = 0.f;
        if (fc.use_sdf) {
          nd = -out[0];
          e_sdf = tc::ex2_approx(-fabsf(nd) * (fc.inv_beta * kLog2e));
          sg = (nd > 0.f) ? 1.f : ((nd < 0.f) ? -1.f : 0.f);
          sigma = fc.inv_alpha * ((0.5f + 0.5f * sg * (1.f - e_sdf)) * keep);
          // analytic derivative also AT the zero crossing (see nfi_backward_pipe.cuh)
          dsig_dout0 = -(fc.inv_alpha * keep) * 0.5f * e_sdf * fc.inv_beta;
        } else {
          const float x = out[0] - 1.f;
          sigma = (x > 20.f ? x : log1pf(expf(x))) * keep;
          dsig_dout0 = keep * sigmoid_fast(x);
        }
        float probs[NA];
        float cr, cg, cb;
        if (fc.A > 0) {
          float mx = out[1];
#pragma unroll
          for (int a = 1; a < NA; ++a) mx = fmaxf(mx, out[1 + a]);
          float s = 0.f;
#pragma unroll
          for (int a = 0; a < NA; ++a) {
            probs[a] = tc::ex2_approx(out[1 + a] - mx);
            s += probs[a];
          }
          const float inv = __fdividef(1.f, s);
          cr = cg = cb = 0.f;
#pragma unroll
          for (int a = 0; a < NA; ++a) {
            probs[a] *= inv;
            cr = fmaf(probs[a], pal[3 * a + 0], cr);
            cg = fmaf(probs[a], pal[3 * a + 1], cg);
            cb = fmaf(probs[a], pal[3 * a + 2], cb);
          }
        } else {
          cr = sigmoid_fast(out[1]) * 2.004f - 1.002f;
          cg = sigmoid_fast(out[2]) * 2.004f - 1.002f;
          cb = sigmoid_fast(out[3]) * 2.004f - 1.002f;
#pragma unroll
          for (int a = 0; a < NA; ++a) probs[a] = 0.f;
        }
        // ---- compositing, forward and reverse (nfi_backward.cuh header)
        const float e_sd = __expf(-sigma * delta);
        const float a = 1.f - e_sd;
        const float w = a * T;
        const float s_i = (g_r * cr + g_g * cg + g_b * cb) + g_m;
        prefix = fmaf(w, s_i, prefix);
        const float one_m_a = 1.f - a;
        const float dsig = delta * one_m_a * (T * s_i - (total - prefix) / (one_m_a + 1e-10f));
        T = T * (one_m_a + 1e-10f);
        // ---- field head, reverse
        float dOut[16];
#pragma unroll
        for (int o = 0; o < 16; ++o) dOut[o] = 0.f;
        dOut[0] = dsig * dsig_dout0;
        if (PLANES && fc.use_sdf) {
          acc_beta = fmaf(dsig, fc.inv_alpha * keep * (-0.5f * sg * e_sdf * fabsf(nd) * fc.inv_beta *
                                                       fc.inv_beta), acc_beta);
          acc_alpha = fmaf(dsig, -sigma * fc.inv_alpha, acc_alpha);
        }
        const float wr = w * g_r, wg2 = w * g_g, wb = w * g_b;
        if (fc.A > 0) {
          float dp[NA];
          float dot = 0.f;
#pragma unroll
          for (int q = 0; q < NA; ++q) {
            const float vv = wr * pal[3 * q + 0] + wg2 * pal[3 * q + 1] + wb * pal[3 * q + 2];
            dp[q] = vv;
            dot = fmaf(probs[q], vv, dot);
            if (PLANES) accP[q] = fmaf(w, probs[q], accP[q]);
          }
#pragma unroll
          for (int q = 0; q < NA; ++q) dOut[1 + q] = probs[q] * (dp[q] - dot);
        } else {
          const float sr = (cr + 1.002f) / 2.004f, sg2 = (cg + 1.002f) / 2.004f,
                      sb = (cb + 1.002f) / 2.004f;
          dOut[1] = wr * 2.004f * sr * (1.f - sr);
          dOut[2] = wg2 * 2.004f * sg2 * (1.f - sg2);
          dOut[3] = wb * 2.004f * sb * (1.f - sb);
        }
#pragma unroll
        for (int o = 0; o < NOUT_PAD; ++o) acc_b2[o] += dOut[o];
        // ---- hand dOut to the tensor core: TF32 hi/lo into TMEM (MMA3), a bf16 hi/lo pair into
        //      the two [128][16] SWIZZLE_32B tiles the dW2 products read (once the previous
        //      step's chain has read them)
        {
          float hi[16], lo[16];
#pragma unroll
          for (int o = 0; o < 16; ++o) {
            hi[o] = tc::tf32_hi(dOut[o]);
            lo[o] = dOut[o] - hi[o];
          }
          tc::tmem_st16(d + 128, hi);
          tc::tmem_st16(d + 144, lo);
          NFI_STEP_WAIT(x_free, (m & 1) ^ 1);
          unsigned char* const et = base + Cfg::kSmE;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            uint32_t eh[4], el[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float a0 = dOut[8 * q + 2 * i], a1 = dOut[8 * q + 2 * i + 1];
              eh[i] = tc::bf16x2_rn(a0, a1);
              el[i] = tc::bf16x2_rn(a0 - __uint_as_float(eh[i] << 16),
                                    a1 - __uint_as_float(eh[i] & 0xFFFF0000u));
            }
            const uint32_t off = tc::sw32_offset(gt, q);
            *reinterpret_cast<uint4*>(et + off) = make_uint4(eh[0], eh[1], eh[2], eh[3]);
            *reinterpret_cast<uint4*>(et + 4096 + off) = make_uint4(el[0], el[1], el[2], el[3]);
          }
          tc::tmem_wait_st();
          tc::tc_fence_before();
          tc::fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&dout_ready[sl]);
        }
        z = zn;
      }
      if constexpr (PLANES) {  // per-tile write-out, as render_backward_pipe
        if (g.grad_palette != nullptr && p.n_attention > 0) {
#pragma unroll
          for (int a = 0; a < NA; ++a) {
            const float pr = warp_sum(accP[a] * g_r), pg = warp_sum(accP[a] * g_g),
                        pb = warp_sum(accP[a] * g_b);
            if (lane == 0 && a < p.n_attention) {
              float* gp = g.grad_palette + ((size_t)b * p.n_attention + a) * 3;
              atomicAdd(gp + 0, pr);
              atomicAdd(gp + 1, pg);
              atomicAdd(gp + 2, pb);
            }
          }
        }
        if (p.use_sdf) {
          const float sb = warp_sum(acc_beta), sa = warp_sum(acc_alpha);
          if (lane == 0) {
            if (g.grad_beta) atomicAdd(g.grad_beta, sb);
            if (g.grad_alpha) atomicAdd(g.grad_alpha, sa);
          }
        }
      }
    }
    if (g.grad_b2 != nullptr) {
      const int nout = 1 + (p.n_attention > 0 ? p.n_attention : 3);
#pragma unroll
      for (int o = 0; o < NOUT_PAD; ++o) {
        const float sb = warp_sum(acc_b2[o]);
        if (lane == 0 && o < nout) atomicAdd(g.grad_b2 + o, sb);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (tid < 32) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace nfi
