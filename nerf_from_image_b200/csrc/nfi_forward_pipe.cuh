// Forward render kernel, pipelined tensor-core variant (NFI_MLP_TC_PIPE, the default).
//
// Same arithmetic as nfi_forward_tc.cuh (3xTF32 decoder on tcgen05, thread = ray
// for everything per-ray), but the per-step chain
//     gather -> MMA1 -> softplus/split -> MMA2 -> density/colour/composite
// is cut so that NO resource is held across more than one link of it, and every
// role only ever does its own kind of work:
//
//   persistent CTA (one per SM), one 16x8-pixel tile (128 rays) in flight:
//     warpgroup 0   ACTIVATION  thread = TMEM lane: D1 -> softplus -> H_hi / H_lo (stateless,
//                   throughput-bound: FMA pipe + one MUFU per hidden unit)
//     warpgroup 1   SHADING     thread = ray = TMEM lane: D2 -> density / colour -> coarse
//                   weights or sorted merge + compositing (latency-bound chains);
//                   the two groups put two independent instruction streams on
//                   every SM sub-partition
//     warpgroup 2   warp 8 issues every layer-1 tcgen05.mma, warp 9 every layer-2
//                   tcgen05.mma (the ~25-50 cycles an MMA takes to issue are
//                   nobody else's problem); warps 10-11 idle
//     warpgroups 3+ P PRODUCER sets of 4 warps; set q gathers the steps n = q (mod P)
//                   into A stage n mod (P+1) (A_hi/A_lo, 32 KB, SWIZZLE_128B)
//
//   stage   : producers --full[q]--> MMA1 --a_free[q] (tcgen05.commit)--> producers
//             (a stage is released as soon as the tensor core has READ it; the
//             hidden activations never come back to shared memory)
//   TMEM    : three slots of 160 columns: [0,64) D1 then H_lo, [64,128) H_hi,
//             [128,144) D2;  MMA1 --d1_full--> activation --h_ready--> MMA2
//             --d2_full--> shading --slot_free--> MMA1
//   softplus: ln2 * (max(x', 0) + lg2(1 + 2^-|x'|)) with x' = x log2 e: two MUFU ops (ex2, lg2)
//             per hidden unit, 8 unit pairs side by side; log2(e) is folded into W1/b1 by
//             prep_weight_image (a degree-8 FMA-pipe log1p instead of lg2 was measured slower:
//             DESIGN.md section 5);
//   resample: the S uniforms of a ray are sorted by a bitonic network across the
//             warp (2 per lane) and pushed through the inverse CDF with shuffle
//             binary searches: one warp-pass per ray instead of a serial per-thread
//             walk (run.py:259-281, lib/nerf_utils.py:183-222).
// What bounds the kernel after this: the L1 data pipe (12 texel lines of 128 bytes per point plus
// the A-tile stores) at 65 %, issue slots at 55 %, and a consumer chain that alone needs two
// thirds of a step.
#pragma once
#include <type_traits>

#include "nfi_forward_tc.cuh"

namespace nfi {

// back-off between mbarrier polls of the per-step hand-offs (0 = hardware-suspended try_wait only)
#ifndef NFI_WAIT_NS
#define NFI_WAIT_NS 0
#endif
#define NFI_STEP_WAIT(bar, par) tc::mbar_wait_backoff<NFI_WAIT_NS>(bar, par)

// ------------------------------------------------------------------ small helpers
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

__device__ __forceinline__ float ld_relaxed(const float* p) {
  float v;
  asm volatile("ld.relaxed.cta.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

// Shell sort (gaps 23, 10, 4, 1) of a thread-private shared-memory column.
__device__ __forceinline__ void column_sort(float* col, int n, int stride) {
  const int gaps[4] = {23, 10, 4, 1};
#pragma unroll 1
  for (int gi = 0; gi < 4; ++gi) {
    const int gap = gaps[gi];
    for (int i = gap; i < n; ++i) {
      const float v = col[i * stride];
      int j = i - gap;
      while (j >= 0 && col[j * stride] > v) {
        col[(j + gap) * stride] = col[j * stride];
        j -= gap;
      }
      col[(j + gap) * stride] = v;
    }
  }
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(pred));
  return pred != 0;
}

struct TileCoord {
  int b, tile_x, tile_y;
};

__device__ __forceinline__ TileCoord tile_coord(int tile, int tiles_x, int tiles_y) {
  TileCoord c;
  const int per_img = tiles_x * tiles_y;
  c.b = tile / per_img;
  const int r = tile % per_img;
  // 2x2 blocks of tiles are consecutive: the two groups of a CTA (and the next
  // CTA) work on neighbouring tiles of the same image -> shared texels in L1/L2
  const int bx = (tiles_x + 1) / 2;
  const int blk = r / 4, in = r % 4;
  int tx = 2 * (blk % bx) + (in & 1), ty = 2 * (blk / bx) + (in >> 1);
  if ((tiles_x & 1) || (tiles_y & 1)) {  // odd tile grids: plain row-major order
    tx = r % tiles_x;
    ty = r / tiles_x;
  }
  c.tile_x = tx;
  c.tile_y = ty;
  return c;
}

constexpr int kPipeSlots = 3;
constexpr int kPipeSlotCols = 160;  // [0,64) D1 then H_lo, [64,128) H_hi, [128,144) D2
constexpr int kPipeStageBytes = 32768;
// NFI_L1_BF16: layer 1 of the forward kernel on bf16 hi/lo pairs (operands exact to 2^-17
// instead of 3xTF32's 2^-22, as nfi_wgrad_pipe.cuh and the synthesis convolutions): the A stage
// shrinks from 32 KB to 16 KB (two [128][32 bf16] SWIZZLE_64B tiles: half the stores of the
// gather, 64 KB of shared memory back to the L1), six K16 MMAs instead of twelve K8.
// Measured (profiles/r2_ab_forward_bf16_l1.txt): 8.52 vs 8.69 ms, 1.9 % -- and the importance-
// resampled depths move from <= 2e-5 to > 2e-5 of the oracle's (the coarse densities feed an
// inverse CDF), which costs the gradient tests their margin.  Not worth it: off by default; the
// weight-gradient kernel, which resamples nothing, uses the bf16 form (nfi_wgrad_pipe.cuh).
#ifndef NFI_L1_BF16
#define NFI_L1_BF16 0
#endif
constexpr int kFwdStageBytes = NFI_L1_BF16 ? 16384 : 32768;

// scratch per CTA: the tc_scratch layout plus NES parked extras per coarse sample
__host__ __device__ inline size_t pipe_scratch_floats(int S, int nes) {
  return tc_scratch_floats_per_group(S) + (size_t)S * kThreads * nes;
}

template <int P>
struct PipeCfg {
  static constexpr int kThreadsTotal = 384 + 128 * P;
  static constexpr int kStages = P + 1;  // one spare: a set never waits for its own MMA
  static constexpr int kSmA = 25600;
  static constexpr int kSmPal = kSmA + kStages * kFwdStageBytes;
  static constexpr int kSmFrac = kSmPal + 48 * 4;  // s / S for s < 128 (one IEEE division each)
  static constexpr int kSmBars = kSmFrac + 128 * 4;
  // full[P], a_free[P], d1_full[3], h_ready[3], d2_full[3], slot_free[3], cw_ready, zf_ready,
  // weights
  static constexpr int kNumBars = 2 * kStages + 4 * kPipeSlots + 3;
  static constexpr int kSmTmemPtr = kSmBars + kNumBars * 8;
  static constexpr int kSmBytes = kSmTmemPtr + 16;
  // setmaxnreg moves registers inside the CTA's launch allocation (threads x launch regs):
  //   P = 3: 768 x 80 = 61440 = 128 x (88 + 80 + 24) + 384 x 96 (12 texel loads in flight per
  //   producer warp: 48 data registers + 64-bit addresses + taps).  A 256-bit variant
  //   (ld.global.nc.v8.f32, 4 lanes per texel, 8 points per iteration: 23 instead of 37
  //   instructions per point) assembles for sm_100a but traps as an illegal instruction
  //   on the B200, so 128-bit loads it is.
#ifndef NFI_ACT_REGS
#define NFI_ACT_REGS 88
#define NFI_SHADE_REGS 80
#define NFI_PROD_REGS 96
#endif
  static constexpr int kActRegs = NFI_ACT_REGS;
  static constexpr int kShadeRegs = NFI_SHADE_REGS;
  static constexpr int kAuxRegs = 24;
  static constexpr int kProducerRegs = NFI_PROD_REGS;
};

// this kernel's weight image has log2(e) folded into layer 1 (the softplus works
// on x' = x log2 e) and padded colour logits pushed to -1e30 (softmax needs no mask)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kPadLogit = -1e30f;

// D2 = H_lo*W2_hi + H_hi*W2_lo + H_hi*W2_hi, both halves of H read from TMEM.
__device__ __forceinline__ void issue_layer2_tt(uint32_t d2_tmem, uint32_t hlo_tmem,
                                                uint32_t hhi_tmem, uint64_t w2_hi,
                                                uint64_t w2_lo) {
  constexpr uint32_t idesc = tc::umma_idesc_tf32(128, 16);
  tc::umma_ts<false>(d2_tmem, hlo_tmem, w2_hi, idesc);
#pragma unroll
  for (int ks = 1; ks < 8; ++ks)
    tc::umma_ts<true>(d2_tmem, hlo_tmem + 8 * ks, w2_hi + (ks >> 2) * 128 + (ks & 3) * 2, idesc);
#pragma unroll
  for (int ks = 0; ks < 8; ++ks)
    tc::umma_ts<true>(d2_tmem, hhi_tmem + 8 * ks, w2_lo + (ks >> 2) * 128 + (ks & 3) * 2, idesc);
#pragma unroll
  for (int ks = 0; ks < 8; ++ks)
    tc::umma_ts<true>(d2_tmem, hhi_tmem + 8 * ks, w2_hi + (ks >> 2) * 128 + (ks & 3) * 2, idesc);
}

__device__ __forceinline__ float ldcg(const float* p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// Hidden activations of 16 TMEM columns: v = log2(e) * (W1 f) (the accumulator),
// bias16 pre-scaled by log2(e).  Returns h split into TF32 hi (hi[]) and the
// remainder (in place, v[]).
//   softplus(x) = ln2 * (max(x', 0) + lg2(1 + 2^-|x'|)),  x' = x log2 e
// 13 instructions per unit pair (2 ex2 + 2 lg2 on the MUFU pipe, which the
// gather does not use); the 8 pairs advance side by side through every stage.
__device__ __forceinline__ void softplus_split16(float (&v)[16], float (&hi)[16],
                                                 const float* __restrict__ bias16) {
  float2 x[8], l[8];
#pragma unroll
  for (int i4 = 0; i4 < 4; ++i4) {
    const float4 bb = *reinterpret_cast<const float4*>(bias16 + 4 * i4);
    x[2 * i4] = __fadd2_rn(make_float2(v[4 * i4], v[4 * i4 + 1]), make_float2(bb.x, bb.y));
    x[2 * i4 + 1] = __fadd2_rn(make_float2(v[4 * i4 + 2], v[4 * i4 + 3]), make_float2(bb.z, bb.w));
  }
#pragma unroll
  for (int k = 0; k < 8; ++k)
    l[k] = make_float2(tc::ex2_approx(-fabsf(x[k].x)), tc::ex2_approx(-fabsf(x[k].y)));
#pragma unroll
  for (int k = 0; k < 8; ++k) l[k] = __fadd2_rn(l[k], make_float2(1.f, 1.f));
#pragma unroll
  for (int k = 0; k < 8; ++k) l[k] = make_float2(tc::lg2_approx(l[k].x), tc::lg2_approx(l[k].y));
#pragma unroll
  for (int k = 0; k < 8; ++k)
    l[k] = __fadd2_rn(l[k], make_float2(fmaxf(x[k].x, 0.f), fmaxf(x[k].y, 0.f)));
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float2 h = __fmul2_rn(l[k], make_float2(kLn2, kLn2));
    hi[2 * k] = tc::tf32_hi(h.x);
    hi[2 * k + 1] = tc::tf32_hi(h.y);
    const float2 lo2 = __fadd2_rn(h, make_float2(-hi[2 * k], -hi[2 * k + 1]));
    v[2 * k] = lo2.x;
    v[2 * k + 1] = lo2.y;
  }
}

// Density and colour from the decoder outputs, branch-free over the palette
// (padded logits arrive as -1e30, padded palette rows are zero).
// models/generator.py:625-679.
template <int NOUT_PAD>
__device__ __forceinline__ void field_head_fast(const float (&out)[NOUT_PAD], const FieldConst& fc,
                                                const float* __restrict__ pal, float keep,
                                                float& sigma, float& cr, float& cg, float& cb,
                                                float* probs = nullptr) {
  constexpr int NA = NOUT_PAD - 1;
  const float d = out[0];
  if (fc.use_sdf) {
    const float nd = -d;
    const float e = tc::ex2_approx(-fabsf(nd) * (fc.inv_beta * kLog2e));
    const float sg = (nd > 0.f) ? 1.f : ((nd < 0.f) ? -1.f : 0.f);
    const float cdf = 0.5f + 0.5f * sg * (1.f - e);
    sigma = fc.inv_alpha * (cdf * keep);
  } else {
    const float x = d - 1.f;
    sigma = (x > 20.f ? x : log1pf(expf(x))) * keep;
  }
  if (fc.A > 0) {
    // colour logits are in log2 units (prep_weight_image scales rows >= 1 of W2/b2)
    float m = out[1];
#pragma unroll
    for (int a = 1; a < NA; ++a) m = fmaxf(m, out[1 + a]);
    float pv[((3 * NA + 3) / 4) * 4];
#pragma unroll
    for (int i = 0; i < (3 * NA + 3) / 4; ++i) {
      const float4 t = *reinterpret_cast<const float4*>(pal + 4 * i);
      pv[4 * i] = t.x;
      pv[4 * i + 1] = t.y;
      pv[4 * i + 2] = t.z;
      pv[4 * i + 3] = t.w;
    }
    float s = 0.f, r = 0.f, g = 0.f, b = 0.f;
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      const float e = tc::ex2_approx(out[1 + a] - m);
      if (probs != nullptr) probs[a] = e;
      s += e;
      r = fmaf(e, pv[3 * a + 0], r);
      g = fmaf(e, pv[3 * a + 1], g);
      b = fmaf(e, pv[3 * a + 2], b);
    }
    const float inv = __fdividef(1.f, s);
    cr = r * inv;
    cg = g * inv;
    cb = b * inv;
    if (probs != nullptr) {
#pragma unroll
      for (int a = 0; a < NA; ++a) probs[a] *= inv;
    }
  } else {
    cr = sigmoid_fast(out[1]) * 2.004f - 1.002f;
    cg = sigmoid_fast(out[2]) * 2.004f - 1.002f;
    cb = sigmoid_fast(out[3]) * 2.004f - 1.002f;
  }
}

// ------------------------------------------------------------------ resampling
// Warp-per-ray importance resampling (run.py:266-281, lib/nerf_utils.py:183-222).
// Per-ray arrays of up to 32 N entries live N per lane: element e = 32 i + lane in v[i]
// (N = 2 for S <= 64, N = 4 for S <= 128).
template <int N>
struct LaneVec {
  float v[N];
};
// y_e = x_{e+1}; the element past the end is `pad`
template <int N>
__device__ __forceinline__ LaneVec<N> shift_down1(const LaneVec<N>& x, float pad, int lane) {
  LaneVec<N> y;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float a = __shfl_down_sync(kFull, x.v[i], 1);
    const float nxt0 = (i + 1 < N) ? __shfl_sync(kFull, x.v[(i + 1 < N) ? i + 1 : i], 0) : pad;
    y.v[i] = (lane == 31) ? nxt0 : a;
  }
  return y;
}
template <int N>
__device__ __forceinline__ float pick(const LaneVec<N>& x, int idx) {
  float r = __shfl_sync(kFull, x.v[0], idx & 31);
#pragma unroll
  for (int i = 1; i < N; ++i) {
    const float t = __shfl_sync(kFull, x.v[i], idx & 31);
    r = ((idx >> 5) == i) ? t : r;
  }
  return r;
}
// ascending bitonic sort of 32 N values
template <int N>
__device__ __forceinline__ void bitonic_sort(LaneVec<N>& x, int lane) {
#pragma unroll
  for (int k = 2; k <= 32 * N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {  // partner is another register of the same lane
        const int dj = j >> 5;
#pragma unroll
        for (int i = 0; i < N; ++i) {
          if ((i & dj) == 0) {
            const int ip = i | dj;
            const bool asc = (k >= 32 * N) ? true : (((32 * i) & k) == 0);
            const float lo = fminf(x.v[i], x.v[ip]), hi = fmaxf(x.v[i], x.v[ip]);
            x.v[i] = asc ? lo : hi;
            x.v[ip] = asc ? hi : lo;
          }
        }
      } else {
        const bool lower = (lane & j) == 0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
          const float o = __shfl_xor_sync(kFull, x.v[i], j);
          // element index 32 i + lane: bit of k taken from the lane (k < 32) or from i
          const bool asc = (k >= 32 * N) ? true : (k < 32 ? ((lane & k) == 0) : (((32 * i) & k) == 0));
          x.v[i] = (lower == asc) ? fminf(x.v[i], o) : fmaxf(x.v[i], o);
        }
      }
    }
  }
}

// One ray.  w: coarse weights (elements >= S are don't-care), t: coarse depths,
// u: uniforms (padded with 2.0 beyond S; sorted inside unless `sorted`).  Returns
// the S fine depths in ascending order.
template <int N>
__device__ __forceinline__ LaneVec<N> resample_ray(const LaneVec<N>& w, const LaneVec<N>& t,
                                                   LaneVec<N> u, bool sorted, int S, int lane) {
  const float inf = __int_as_float(0x7f800000);
  // smoothed pdf p_m, m = 0 .. S-3  (run.py:266-272, + 1e-5 of sample_pdf)
  const LaneVec<N> w1 = shift_down1<N>(w, 0.f, lane);
  const LaneVec<N> w2 = shift_down1<N>(w1, 0.f, lane);
  LaneVec<N> q;
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float pm = ((fmaxf(w.v[i], w1.v[i]) + fmaxf(w1.v[i], w2.v[i])) * 0.5f + 0.01f) + 1e-5f;
    if (32 * i + lane >= S - 2) pm = 0.f;
    q.v[i] = pm;
    sum += pm;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(kFull, sum, o);
  // inclusive scan of q_m = p_m / sum over the 32 N slots
#pragma unroll
  for (int i = 0; i < N; ++i) q.v[i] = q.v[i] / sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const float y = __shfl_up_sync(kFull, q.v[i], o);
      if (lane >= o) q.v[i] += y;
    }
  }
#pragma unroll
  for (int i = 1; i < N; ++i) q.v[i] += __shfl_sync(kFull, q.v[i - 1], 31);
  // cdf c_j, j = 0 .. S-2: c_0 = 0, c_j = scan_{j-1}; +inf beyond
  LaneVec<N> c;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float up = __shfl_up_sync(kFull, q.v[i], 1);
    const float prev31 = (i > 0) ? __shfl_sync(kFull, q.v[(i > 0) ? i - 1 : 0], 31) : 0.f;
    c.v[i] = (lane == 0) ? prev31 : up;
    if (32 * i + lane > S - 2) c.v[i] = inf;
  }
  // bins b_j = (t_j + t_{j+1}) / 2, j = 0 .. S-2
  const LaneVec<N> t1 = shift_down1<N>(t, 0.f, lane);
  LaneVec<N> bn;
#pragma unroll
  for (int i = 0; i < N; ++i) bn.v[i] = 0.5f * (t1.v[i] + t.v[i]);
  if (!sorted) bitonic_sort<N>(u, lane);
  LaneVec<N> z;
#pragma unroll
  for (int h = 0; h < N; ++h) {
    const float uu = u.v[h];
    int pos = 0;  // number of cdf entries <= u (searchsorted right=True)
#pragma unroll
    for (int s = 16 * N; s > 0; s >>= 1) {
      const float val = pick<N>(c, pos + s - 1);
      if (val <= uu) pos += s;
    }
    const int below = max(pos - 1, 0), above = min(pos, S - 2);
    const float c0 = pick<N>(c, below), c1 = pick<N>(c, above);
    const float b0 = pick<N>(bn, below), b1 = pick<N>(bn, above);
    float den = c1 - c0;
    if (den < 1e-5f) den = 1.f;
    z.v[h] = b0 + (uu - c0) / den * (b1 - b0);
  }
  return z;
}

// Importance resampling of the rays [first, first + count) of a warp's 32 rows, one
// warp-pass per ray.  (tnear, tfar, ray, valid) are this lane's own row's.  N = 2 slots per
// lane serve S <= 64, N = 4 (a separate kernel instantiation, so that its register needs
// do not touch the allocation of the common case's per-step loops) S <= 128.
struct ResampleArgs {
  const float* sc_w;
  float* sc_zf;
  float* z_fine;
  const float* noise_t;
  const float* noise_u;
  const float* frac;
  int S, wig;
  bool explicit_noise;
};
template <int N>
__device__ __forceinline__ void resample_rows_impl(const ResampleArgs& a, int first, int count,
                                                   float tnear, float tfar, size_t ray, bool valid,
                                                   int lane) {
  const int S = a.S;
  if (count <= 0) return;
  // inputs of ray j+1 are loaded while ray j is resampled
  struct In {
    LaneVec<N> w, n, u;
    size_t rayj;
  };
  auto load = [&](int j, In& in) {
    in.rayj = ((size_t)__shfl_sync(kFull, (unsigned)(ray >> 32), j) << 32) |
              (size_t)__shfl_sync(kFull, (unsigned)ray, j);
    const int col = 32 * a.wig + j;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const int e = 32 * i + lane;
      in.w.v[i] = (e < S) ? ldcg(a.sc_w + e * kThreads + col) : 0.f;
      in.n.v[i] = 0.f;
      if (a.explicit_noise) {
        in.n.v[i] = (e < S) ? a.noise_t[in.rayj * S + e] : 0.f;
        in.u.v[i] = (e < S) ? a.noise_u[in.rayj * S + e] : 2.f;
      } else {
        in.u.v[i] = (e < S) ? linspace01(e, S) : 2.f;
      }
    }
  };
  In cur, nxt;
  load(first, cur);
#pragma unroll 1
  for (int j = first; j < first + count; ++j) {
    load(j + 1 < first + count ? j + 1 : j, nxt);
    const float nearj = __shfl_sync(kFull, tnear, j), farj = __shfl_sync(kFull, tfar, j);
    const bool validj = __shfl_sync(kFull, (int)valid, j) != 0;
    const int col = 32 * a.wig + j;
    const float spanj = farj - nearj;
    LaneVec<N> t;
#pragma unroll
    for (int i = 0; i < N; ++i)
      t.v[i] = lerp_torch(nearj, farj, a.frac[32 * i + lane]) + cur.n.v[i] * (spanj / (float)S);
    const LaneVec<N> z = resample_ray<N>(cur.w, t, cur.u, !a.explicit_noise, S, lane);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const int e = 32 * i + lane;
      if (e < S) a.sc_zf[e * kThreads + col] = z.v[i];
      if (a.z_fine != nullptr && validj && e < S) a.z_fine[cur.rayj * S + e] = z.v[i];
    }
    cur = nxt;
  }
}
template <int NOUT_PAD, int EXTRA, bool FINE, int P, bool DBG, int NSLOT = 2>
__global__ void __launch_bounds__(PipeCfg<P>::kThreadsTotal, 1)
render_forward_pipe(const nfi_render_params p, const unsigned char* __restrict__ wimg,
                    float* __restrict__ scratch) {
  using Cfg = PipeCfg<P>;
  // extras composited with the weights: 3 world coordinates (EXTRA 1, recomputed from the
  // depth) or the NA attention probabilities (EXTRA 2, parked in scratch for the coarse samples)
  constexpr int NA_ = NOUT_PAD - 1;
  constexpr int NE = (EXTRA == 1) ? 3 : (EXTRA == 2 ? NA_ : 0);
  constexpr int NES = (EXTRA == 2) ? NA_ : 0;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = smem_raw;
  const int tid = threadIdx.x, lane = tid & 31;
  // Role by warpgroup.  The arbiter favours high warp ids, so the two consumer roles --
  // one warp each per sub-partition, nothing to hide their latency behind -- sit above the
  // producers: hardware wg 0..P-1 producer sets, P MMA issuers, P+1 shading, P+2 activation.
  // `wg` below is the logical role: 0 activation, 1 shading, 2 MMA issuers, 3.. producer sets.
  const int hw_wg = __shfl_sync(kFull, tid >> 7, 0);
  const int wg = (hw_wg < P) ? hw_wg + 3 : (P + 2 - hw_wg);
  const int gt = tid & 127;                        // row of the tile = ray = TMEM lane
  const int wig = __shfl_sync(kFull, gt >> 5, 0);  // warp in warpgroup = TMEM lane quadrant
  const int S = p.num_samples;

  uint64_t* bars = reinterpret_cast<uint64_t*>(base + Cfg::kSmBars);
  constexpr int NS = Cfg::kStages;
  uint64_t* full = bars;                         // [NS] stage gathered            (4 warps)
  uint64_t* a_free = full + NS;                  // [NS] stage read by the tensor core (commit)
  uint64_t* d1_full = a_free + NS;               // [3]  layer-1 accumulator ready (commit)
  uint64_t* h_ready = d1_full + kPipeSlots;      // [3]  H_hi/H_lo in TMEM         (4 warps)
  uint64_t* d2_full = h_ready + kPipeSlots;      // [3]  layer-2 accumulator ready (commit)
  uint64_t* slot_free = d2_full + kPipeSlots;    // [3]  D2 read                   (4 warps)
  uint64_t* cw_ready = slot_free + kPipeSlots;   //      coarse weights of the tile written (4 warps)
  uint64_t* zf_ready = cw_ready + 1;             //      fine depths of the tile written (every resampling warp)
  uint64_t* wbar = zf_ready + 1;                 //      weight image landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(base + Cfg::kSmTmemPtr);
  const float* b1s = reinterpret_cast<const float*>(base + kWiB1);
  const float* b2s = reinterpret_cast<const float*>(base + kWiB2);
  float* pal = reinterpret_cast<float*>(base + Cfg::kSmPal);
  float* frac = reinterpret_cast<float*>(base + Cfg::kSmFrac);
  if (tid < 128) frac[tid] = (float)tid / (float)S;

  if (tid == 0) {
    if (tc::smem_u32(base) & 1023u) __trap();
    for (int i = 0; i < NS; ++i) {
      tc::mbar_init(&full[i], 4);
      tc::mbar_init(&a_free[i], 1);
    }
    for (int i = 0; i < kPipeSlots; ++i) {
      tc::mbar_init(&d1_full[i], 1);
      // consumer barriers count WARPS: a warp-wide mbarrier.arrive is 32 serialised
      // shared-memory atomics on one word (L1 data-pipe wavefronts the gather needs)
      tc::mbar_init(&h_ready[i], kWarps);
      tc::mbar_init(&d2_full[i], 1);
      tc::mbar_init(&slot_free[i], kWarps);
    }
    tc::mbar_init(cw_ready, kWarps);
    tc::mbar_init(zf_ready, (2 + P) * kWarps);
    tc::mbar_init(wbar, 1);
    tc::fence_mbar_init();
  }
  if (tid < 32) tc::tmem_alloc(tmem_ptr, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(kFull, *tmem_ptr, 0);
  if (tid == 0) {
    tc::mbar_expect_tx(wbar, kWiBytes);
    tc::tma_bulk_g2s(base, wimg, kWiBytes, wbar);
  }
  tc::mbar_wait(wbar, 0);
#if NFI_L1_BF16
  // layer 1's weights (x log2 e, as the image has them in TF32) as a bf16 hi / lo pair over the
  // image's W1 region: [64 rows = hidden unit][32 k] K-major SWIZZLE_64B, 4 KB each
  __syncthreads();
  for (int i = tid; i < kHid * kC; i += Cfg::kThreadsTotal) {
    const int j = i / kC, c = i % kC;
    const float w = p.w1[i] * kLog2e;
    const uint32_t hi = tc::bf16x2_rn(w, 0.f) & 0xFFFFu;
    const uint32_t lo = tc::bf16x2_rn(w - __uint_as_float(hi << 16), 0.f) & 0xFFFFu;
    const uint32_t off = tc::sw64_offset(j, c >> 3) + (c & 7) * 2;
    *reinterpret_cast<unsigned short*>(base + kWiW1Hi + off) = (unsigned short)hi;
    *reinterpret_cast<unsigned short*>(base + kWiW1Hi + 4096 + off) = (unsigned short)lo;
  }
  tc::fence_async_smem();
  __syncthreads();
#endif

  const uint32_t base_s = tc::smem_u32(base);
  const int tiles_x = (p.width + kTileW - 1) / kTileW;
  const int tiles_y = (p.height + kTileH - 1) / kTileH;
  const int n_tiles = tiles_x * tiles_y * p.batch;
  const int my_tiles =
      ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const uint32_t total_steps = (uint32_t)my_tiles * (uint32_t)S * (FINE ? 2u : 1u);
  const int R = p.plane_res;
  const float inv_range = 1.f / p.scene_range;
  const bool explicit_noise = (p.noise_mode == NFI_NOISE_EXPLICIT);
  // timing experiments only (bench.py --mlp-mode 0x104 / 0x204 / 0x804): results are garbage
  const bool dbg_skip_gather = (p.mlp_mode & 0x100) != 0;
  const bool dbg_skip_consumer = (p.mlp_mode & 0x200) != 0;
  const bool dbg_window = (p.mlp_mode & 0x800) != 0;
  // phase timers (DBG instantiation, mlp_mode & 0x1000, buffer in p.normals): lane 0 of
  // producer set 0 warp 0 -> [0..], activation warp 0 -> [16..], shading warp 0 -> [32..],
  // issuer warps -> [48..], [64..]
  const bool dbg_time = DBG && (p.mlp_mode & 0x1000) && p.normals != nullptr && blockIdx.x == 0 &&
                        lane == 0 && (wg <= 3) && (wg == 2 ? wig < 2 : wig == 0);
  long long tacc[DBG ? 10 : 1];
  long long tprev = 0;
  if (DBG)
    for (int i = 0; i < 10; ++i) tacc[i] = 0;
#define NFI_T(i)                         \
  if (DBG && dbg_time) {                 \
    const long long now__ = clock64();   \
    tacc[i] += now__ - tprev;            \
    tprev = now__;                       \
  }
  float* slab = scratch + (size_t)blockIdx.x * pipe_scratch_floats(S, NES);
  float4* sc_srgb = reinterpret_cast<float4*>(slab);   // [S][128] coarse (sigma, r, g, b)
  float* sc_t = slab + (size_t)4 * S * kThreads;       // [S][128] coarse depths
  float* sc_w = sc_t + (size_t)S * kThreads;           // [S][128] coarse weights
  float* sc_zf = sc_w + (size_t)S * kThreads;          // [S][128] fine depths, ascending
  float* sc_e = sc_zf + (size_t)S * kThreads;          // [S][NES][128] coarse attention probabilities

  // Importance resampling of the rays [first, first + count) of this warp's 32 rows:
  // one warp-pass per ray.  (tnear, tfar, ray index, valid) are this lane's own row's.
  // The 32 rows of a quadrant are split over the 2 + P warps that serve it (shading,
  // activation and one warp of every producer set -- the producers would otherwise idle
  // between the coarse and the fine pass): a warp-pass is a chain of dependent shuffles at
  // ~0.1 IPC, so five warps per sub-partition get through the tile's 128 rays sooner than
  // two (measured per tile: 78 k -> 46 k cycles until the fine depths are complete; kernel
  // 8.78 -> 8.65 ms at config 2; profiles/r1_phase_times_v8_pipe.txt).
  constexpr int kRsParts = 2 + P;
  auto rs_first = [](int part) {  // part: 0 shading, 1 activation, 2.. producer sets
    constexpr int base_rows = 32 / kRsParts, rem = 32 % kRsParts;
    return part * base_rows + min(max(part - 2, 0), rem);
  };
  auto resample_rows = [&](int part, float tnear, float tfar, size_t ray, bool valid) {
    const int first = rs_first(part), count = rs_first(part + 1) - first;
    ResampleArgs ra;
    ra.sc_w = sc_w;
    ra.sc_zf = sc_zf;
    ra.z_fine = p.z_fine;
    ra.noise_t = p.noise_t;
    ra.noise_u = p.noise_u;
    ra.frac = frac;
    ra.S = S;
    ra.wig = wig;
    ra.explicit_noise = explicit_noise;
    resample_rows_impl<NSLOT>(ra, first, count, tnear, tfar, ray, valid, lane);
  };

  // The roles never share code after setmaxnreg: ptxas budgets registers per
  // region, and a block reachable from two branches gets the smaller count.
  if (wg >= 3) {
    // ================================ PRODUCER ================================
    if (Cfg::kProducerRegs * Cfg::kThreadsTotal > 65536)  // above the launch allocation per thread
      asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::kProducerRegs));
    else
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::kProducerRegs));
    const int set = wg - 3;
    uint32_t n = 0;        // ring position at the start of the pass, identical in every role
    uint32_t tile_it = 0;  // tiles done by this CTA (parity of zf_ready)
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
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
      const float span = r.tfar - r.tnear;
      const uint32_t plane_bytes = (uint32_t)R * (uint32_t)R * 128u;
      const unsigned char* planes_b =
          reinterpret_cast<const unsigned char*>(p.planes) + (size_t)b * 3 * plane_bytes;
      for (int pass = 0; pass < (FINE ? 2 : 1); ++pass) {
        if (pass == 1) {
          // this set's share of the quadrant's rows
          tc::mbar_wait(cw_ready, tile_it & 1);
          resample_rows(2 + set, r.tnear, r.tfar, ray, valid);
          __threadfence_block();
          __syncwarp();
          if (lane == 0) mbar_arrive(zf_ready);
          tc::mbar_wait(zf_ready, tile_it & 1);
        }
        // first step of this pass that belongs to this set; its sample value is
        // fetched one step ahead (the load is never waited on)
        int s = (int)(((uint32_t)set + (uint32_t)P - (n % P)) % P);
        auto fetch = [&](int ss) -> float {
          if (ss >= S) return 0.f;
          if (pass == 0) return explicit_noise ? p.noise_t[ray * S + ss] : 0.f;
          return ld_relaxed(sc_zf + ss * kThreads + gt);
        };
        float nxt = fetch(s);
        for (; s < S; s += P) {
          const uint32_t st = (n + (uint32_t)s) % NS, u = (n + (uint32_t)s) / NS;
          unsigned char* const stage = base + Cfg::kSmA + st * kFwdStageBytes;
          const float cur = nxt;
          nxt = fetch(s + P);
          if (DBG && dbg_time) tprev = clock64();
          NFI_STEP_WAIT(&a_free[st], (u & 1) ^ 1);  // tensor core has read the previous fill
          NFI_T(0)
          float t;
          if (pass == 0)
            t = lerp_torch(r.tnear, r.tfar, frac[s]) + cur * (span / (float)S);
          else
            t = cur;
          const float x0 = (r.ox + r.dx * t) * inv_range, x1 = (r.oy + r.dy * t) * inv_range,
                      x2 = (r.oz + r.dz * t) * inv_range;
          ByteTaps tp;
          byte_taps(x0, x1, R, 0u, tp.o[0], tp.fx[0], tp.fy[0]);
          byte_taps(x0, x2, R, plane_bytes >> 4, tp.o[1], tp.fx[1], tp.fy[1]);
          byte_taps(x1, x2, R, plane_bytes >> 3, tp.o[2], tp.fx[2], tp.fy[2]);
          NFI_T(1)
          if (dbg_window) {  // every tap inside a 32 KB window: L1 hits only
            tp.o[0] &= 0x7FFu;
            tp.o[1] &= 0x7FFu;
            tp.o[2] &= 0x7FFu;
          }
          if (!dbg_skip_gather) {
#if NFI_L1_BF16
            gather_to_tiles_lean<true>(planes_b, R, tp, stage, stage + 8192, 32 * wig, lane);
#else
            gather_to_tiles_lean(planes_b, R, tp, stage, stage + 16384, 32 * wig, lane);
#endif
          }
          NFI_T(2)
          tc::fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&full[st]);
          NFI_T(3)
          if (DBG && dbg_time) tacc[9] += 1;
        }
        n += (uint32_t)S;
      }
    }
    if (DBG && dbg_time)
      for (int i = 0; i < 10; ++i) p.normals[i] = (float)tacc[i];
  } else if (wg == 2) {
    // ================================ MMA ISSUERS ================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::kAuxRegs));
    if (wig == 0) {
      // layer 1: full[set] + slot_free[slot] -> 12 MMAs -> d1_full[slot], a_free[set]
#if NFI_L1_BF16
      constexpr uint32_t idesc1 = tc::umma_idesc_bf16(128, 64, false, false);
      const uint64_t dsc_w1_hi = tc::umma_desc(base_s + kWiW1Hi, 4, 16, 512);
      const uint64_t dsc_w1_lo = tc::umma_desc(base_s + kWiW1Hi + 4096, 4, 16, 512);
      const uint64_t dsc_a0 = tc::umma_desc(base_s + Cfg::kSmA, 4, 16, 512);
#else
      const uint64_t dsc_w1_hi = tc::umma_desc_sw128(base_s + kWiW1Hi);
      const uint64_t dsc_w1_lo = tc::umma_desc_sw128(base_s + kWiW1Lo);
      const uint64_t dsc_a0 = tc::umma_desc_sw128(base_s + Cfg::kSmA);
#endif
      uint32_t st = 0, u = 0, sl = 0, v = 0;
      for (uint32_t m = 0; m < total_steps; ++m) {
        if (DBG && dbg_time && tprev == 0) tprev = clock64();
        NFI_STEP_WAIT(&full[st], u & 1);
        NFI_T(0)
        NFI_STEP_WAIT(&slot_free[sl], (v & 1) ^ 1);
        NFI_T(1)
        if (elect_one()) {
          tc::tc_fence_after();
          const uint64_t dsc_a = dsc_a0 + (uint64_t)st * (kFwdStageBytes >> 4);
#if NFI_L1_BF16
          // D1 = F_lo W1_hi + F_hi W1_lo + F_hi W1_hi, K = 32 = two K16 steps of 32 bytes
          const uint32_t d1 = tmem_base + sl * kPipeSlotCols;
          const uint64_t a_lo = dsc_a + (8192 >> 4);
          tc::umma_f16_ss<false>(d1, a_lo, dsc_w1_hi, idesc1);
          tc::umma_f16_ss<true>(d1, a_lo + 2, dsc_w1_hi + 2, idesc1);
          tc::umma_f16_ss<true>(d1, dsc_a, dsc_w1_lo, idesc1);
          tc::umma_f16_ss<true>(d1, dsc_a + 2, dsc_w1_lo + 2, idesc1);
          tc::umma_f16_ss<true>(d1, dsc_a, dsc_w1_hi, idesc1);
          tc::umma_f16_ss<true>(d1, dsc_a + 2, dsc_w1_hi + 2, idesc1);
#else
          tc::issue_layer1_d(tmem_base + sl * kPipeSlotCols, dsc_a, dsc_a + (16384 >> 4),
                             dsc_w1_hi, dsc_w1_lo);
#endif
          tc::umma_commit(&d1_full[sl]);
          tc::umma_commit(&a_free[st]);
        }
        __syncwarp();
        NFI_T(2)
        if (++st == NS) { st = 0; ++u; }
        if (++sl == kPipeSlots) { sl = 0; ++v; }
      }
      if (DBG && dbg_time)
        for (int i = 0; i < 10; ++i) p.normals[48 + i] = (float)tacc[i];
    } else if (wig == 1) {
      // layer 2: h_ready[slot] -> 24 MMAs (A from TMEM) -> d2_full[slot]
      const uint64_t dsc_w2_hi = tc::umma_desc_sw128(base_s + kWiW2Hi);
      const uint64_t dsc_w2_lo = tc::umma_desc_sw128(base_s + kWiW2Lo);
      uint32_t sl = 0, v = 0;
      for (uint32_t m = 0; m < total_steps; ++m) {
        if (DBG && dbg_time && tprev == 0) tprev = clock64();
        NFI_STEP_WAIT(&h_ready[sl], v & 1);
        NFI_T(0)
        if (elect_one()) {
          tc::tc_fence_after();
          const uint32_t d_col = tmem_base + sl * kPipeSlotCols;
          issue_layer2_tt(d_col + 128, d_col, d_col + 64, dsc_w2_hi, dsc_w2_lo);
          tc::umma_commit(&d2_full[sl]);
        }
        __syncwarp();
        NFI_T(1)
        if (++sl == kPipeSlots) { sl = 0; ++v; }
      }
      if (DBG && dbg_time)
        for (int i = 0; i < 10; ++i) p.normals[64 + i] = (float)tacc[i];
    }
  } else if (wg == 0) {
    // ================================ ACTIVATION ================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::kActRegs));
    uint32_t sl = 0, v = 0;
    const uint32_t lane_addr = (uint32_t)(32 * wig) << 16;
    // D1 of the slot -> bias + softplus -> H_lo over D1, H_hi next to it -> h_ready
    auto activate = [&]() {
      const uint32_t d1 = tmem_base + sl * kPipeSlotCols + lane_addr;
      if (DBG && dbg_time && tprev == 0) tprev = clock64();
      NFI_T(2)
      NFI_STEP_WAIT(&d1_full[sl], v & 1);
      tc::tc_fence_after();
      NFI_T(0)
      if (!dbg_skip_consumer) {
        uint32_t ra[16], rb[16];
        float lo[16], hi[16];
        tc::tmem_ld16_nowait(d1, ra);
        tc::tmem_ld16_nowait(d1 + 16, rb);
        tc::tmem_wait_ld();
#pragma unroll 1
        for (int c2 = 0; c2 < 2; ++c2) {
#pragma unroll
          for (int i = 0; i < 16; ++i) lo[i] = __uint_as_float(ra[i]);
          if (c2 == 0) tc::tmem_ld16_nowait(d1 + 32, ra);  // next 32 columns while these compute
          softplus_split16(lo, hi, b1s + 32 * c2);
          tc::tmem_st16(d1 + 32 * c2, lo);
          tc::tmem_st16(d1 + 64 + 32 * c2, hi);
#pragma unroll
          for (int i = 0; i < 16; ++i) lo[i] = __uint_as_float(rb[i]);
          if (c2 == 0) tc::tmem_ld16_nowait(d1 + 48, rb);
          softplus_split16(lo, hi, b1s + 32 * c2 + 16);
          tc::tmem_st16(d1 + 32 * c2 + 16, lo);
          tc::tmem_st16(d1 + 64 + 32 * c2 + 16, hi);
          if (c2 == 0) tc::tmem_wait_ld();
        }
        tc::tmem_wait_st();
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&h_ready[sl]);
      if (++sl == kPipeSlots) { sl = 0; ++v; }
      NFI_T(1)
      if (DBG && dbg_time) tacc[9] += 1;
    };
    uint32_t tile_it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
      for (int s = 0; s < S; ++s) activate();
      if (FINE) {
        // second half of this warp's rows is resampled here, the first half by the
        // shading warp that owns them
        const TileCoord tcd = tile_coord(tile, tiles_x, tiles_y);
        int px, py;
        tile_pixel(tcd.tile_x, tcd.tile_y, wig, lane, px, py);
        const bool valid = (px < p.width) && (py < p.height);
        px = min(px, p.width - 1);
        py = min(py, p.height - 1);
        const size_t ray = ((size_t)tcd.b * p.height + py) * p.width + px;
        Ray r;
        setup_ray(p, tcd.b, py, px, r);
        tc::mbar_wait(cw_ready, tile_it & 1);
        NFI_T(3)
        resample_rows(1, r.tnear, r.tfar, ray, valid);
        __threadfence_block();
        __syncwarp();
        if (lane == 0) mbar_arrive(zf_ready);
        NFI_T(4)
        for (int s = 0; s < S; ++s) activate();
      }
    }
    if (DBG && dbg_time)
      for (int i = 0; i < 10; ++i) p.normals[16 + i] = (float)tacc[i];
  } else {
    // ================================ SHADING ================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::kShadeRegs));
    uint32_t sl = 0, v = 0;
    const uint32_t lane_addr = (uint32_t)(32 * wig) << 16;
    FieldConst fc;
    fc.A = p.n_attention;
    fc.use_sdf = p.use_sdf;
    fc.inv_beta = p.use_sdf ? 1.f / p.beta[0] : 0.f;
    fc.inv_alpha = p.use_sdf ? 1.f / p.alpha[0] : 0.f;

    uint32_t tile_it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
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
      const float span = r.tfar - r.tnear;

      tc::bar_sync(1, kThreads);  // previous tile's palette no longer in use
      if (gt < 48)
        pal[gt] = (p.n_attention > 0 && gt < p.n_attention * 3)
                      ? p.palette[(size_t)b * p.n_attention * 3 + gt]
                      : 0.f;
      tc::bar_sync(1, kThreads);
      Compositor<NE, true> comp;
      comp.init();

      // decoder outputs of the next ring position -> density and colour at depth t
      auto shade = [&](float t, float& sigma, float& cr, float& cg, float& cb, float* ex) {
        const uint32_t d2 = tmem_base + sl * kPipeSlotCols + 128 + lane_addr;
        if (DBG && dbg_time && tprev == 0) tprev = clock64();
        NFI_T(2)
        NFI_STEP_WAIT(&d2_full[sl], v & 1);
        tc::tc_fence_after();
        NFI_T(0)
        float o16[16];
        tc::tmem_ld16(d2, o16);
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&slot_free[sl]);
        if (++sl == kPipeSlots) { sl = 0; ++v; }
        const float wx = r.ox + r.dx * t, wy = r.oy + r.dy * t, wz = r.oz + r.dz * t;
        const float x0 = wx * inv_range, x1 = wy * inv_range, x2 = wz * inv_range;
        const float keep =
            (fabsf(x0) > 1.f || fabsf(x1) > 1.f || fabsf(x2) > 1.f) ? 0.f : 1.f;
        float out[NOUT_PAD];
#pragma unroll
        for (int o4 = 0; o4 < NOUT_PAD / 4; ++o4) {
          const float4 bb = *reinterpret_cast<const float4*>(b2s + 4 * o4);
          out[4 * o4] = o16[4 * o4] + bb.x;
          out[4 * o4 + 1] = o16[4 * o4 + 1] + bb.y;
          out[4 * o4 + 2] = o16[4 * o4 + 2] + bb.z;
          out[4 * o4 + 3] = o16[4 * o4 + 3] + bb.w;
        }
        if (dbg_skip_consumer) {
          sigma = out[0];
          cr = cg = cb = out[1];
        } else {
          field_head_fast<NOUT_PAD>(out, fc, pal, keep, sigma, cr, cg, cb,
                                    EXTRA == 2 ? ex : nullptr);
        }
        if (EXTRA == 1) {
          ex[0] = wx;
          ex[1] = wy;
          ex[2] = wz;
        }
        NFI_T(1)
        if (DBG && dbg_time) tacc[9] += 1;
      };

      // ---------------- coarse pass ----------------
      {
        float wT = 1.f, prev_t = 0.f, prev_s = 0.f;
        // jitter: four steps per load, the next four in flight (S % 4 == 0)
        const float4* nz4 = reinterpret_cast<const float4*>(p.noise_t + ray * S);
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 nzc = explicit_noise ? nz4[0] : zero4;
        float4 nzn = (explicit_noise && S > 4) ? nz4[1] : zero4;
        const float jit = span / (float)S;
        for (int s = 0; s < S; ++s) {
          const int sq = s & 3;
          const float nz = sq == 0 ? nzc.x : (sq == 1 ? nzc.y : (sq == 2 ? nzc.z : nzc.w));
          const float t = lerp_torch(r.tnear, r.tfar, frac[s]) + nz * jit;
          if (sq == 3) {
            nzc = nzn;
            nzn = (explicit_noise && s + 5 < S) ? nz4[(s + 5) >> 2] : zero4;
          }
          float sigma, cr, cg, cb;
          float ex[NE > 0 ? NE : 1];
          shade(t, sigma, cr, cg, cb, ex);
          if (FINE) {
            sc_srgb[s * kThreads + gt] = make_float4(sigma, cr, cg, cb);
            sc_t[s * kThreads + gt] = t;
#pragma unroll
            for (int a = 0; a < NES; ++a) sc_e[((size_t)s * NES + a) * kThreads + gt] = ex[a];
            if (s > 0) {
              const float delta = (t - prev_t) * r.dn;
              const float a = 1.f - __expf(-prev_s * delta);
              sc_w[(s - 1) * kThreads + gt] = a * wT;
              wT = wT * ((1.f - a) + 1e-10f);
            }
            prev_t = t;
            prev_s = sigma;
          } else {
            comp.push(t, sigma, cr, cg, cb, ex, r.dn);
          }
          NFI_T(3)
        }
      }

      if (FINE) {
        sc_w[(S - 1) * kThreads + gt] = 0.f;
        __threadfence_block();
        __syncwarp();
        if (lane == 0) mbar_arrive(cw_ready);  // the activation group may resample its half
        tc::mbar_wait(cw_ready, tile_it & 1);
        NFI_T(4)
        resample_rows(0, r.tnear, r.tfar, ray, valid);
        __threadfence_block();
        __syncwarp();
        if (lane == 0) mbar_arrive(zf_ready);  // (with the activation group's 4) producers may start
        tc::mbar_wait(zf_ready, tile_it & 1);
        NFI_T(5)

        // ------- fine pass + sorted merge + compositing -------
        // The next TWO coarse samples wait in registers, so taking one never
        // stalls on the (L2) load of its successor.
        int c = 0;
        float ct0 = ldcg(sc_t + gt), ct1 = ldcg(sc_t + kThreads + gt);
        float4 cq0 = __ldcg(sc_srgb + gt), cq1 = __ldcg(sc_srgb + kThreads + gt);
        auto take_coarse = [&]() {
          float ce[NE > 0 ? NE : 1];
          if (EXTRA == 1) {
            ce[0] = r.ox + r.dx * ct0;
            ce[1] = r.oy + r.dy * ct0;
            ce[2] = r.oz + r.dz * ct0;
          }
#pragma unroll
          for (int a = 0; a < NES; ++a) ce[a] = ldcg(sc_e + ((size_t)c * NES + a) * kThreads + gt);
          comp.push(ct0, cq0.x, cq0.y, cq0.z, cq0.w, ce, r.dn);
          ++c;
          ct0 = ct1;
          cq0 = cq1;
          if (c + 1 < S) {
            ct1 = ldcg(sc_t + (c + 1) * kThreads + gt);
            cq1 = __ldcg(sc_srgb + (c + 1) * kThreads + gt);
          }
        };
        float z0 = ldcg(sc_zf + gt);
        float z1 = (S > 1) ? ldcg(sc_zf + kThreads + gt) : 0.f;
        for (int k = 0; k < S; ++k) {
          const float z = z0;
          z0 = z1;
          z1 = (k + 2 < S) ? ldcg(sc_zf + (k + 2) * kThreads + gt) : 0.f;
          float sigma, cr, cg, cb;
          float ex[NE > 0 ? NE : 1];
          shade(z, sigma, cr, cg, cb, ex);
          while (c < S && ct0 <= z) take_coarse();
          comp.push(z, sigma, cr, cg, cb, ex, r.dn);
          NFI_T(3)
        }
        while (c < S) take_coarse();
        NFI_T(6)
      }

      if (valid) {
        float bg = 0.f;
        if (p.white_background) bg = 1.f - comp.am;
        p.rgb[ray * 3 + 0] = comp.ar + bg;
        p.rgb[ray * 3 + 1] = comp.ag + bg;
        p.rgb[ray * 3 + 2] = comp.ab + bg;
        p.depth[ray] = comp.ad;
        p.mask[ray] = comp.am;
        // the exchange step of the multi-GPU form, fused: the same five floats go to this ray's
        // place in every peer's full-batch buffers over NVLink (posted stores; the ranks meet at
        // a barrier after the kernel, parallel.PeerExchange)
        for (int q = 0; q < p.n_peers; ++q) {
          float* pr = p.peer_rgb[q] + ray * 3;
          pr[0] = comp.ar + bg;
          pr[1] = comp.ag + bg;
          pr[2] = comp.ab + bg;
          p.peer_depth[q][ray] = comp.ad;
          p.peer_mask[q][ray] = comp.am;
        }
        if (EXTRA == 1 && p.extra != nullptr)
          for (int a = 0; a < 3; ++a) p.extra[ray * 3 + a] = comp.ae[a];
        if (EXTRA == 2 && p.extra != nullptr)
          for (int a = 0; a < NE; ++a)
            if (a < p.n_attention) p.extra[ray * p.n_attention + a] = comp.ae[a];
      }
    }
    if (DBG && dbg_time)
      for (int i = 0; i < 10; ++i) p.normals[32 + i] = (float)tacc[i];
  }
#undef NFI_T
  tc::tc_fence_before();
  __syncthreads();
  if (p.n_peers > 0 && p.peer_done != nullptr && tid == 0) {
    // Completion handshake of the fused exchange.  Every CTA: its threads' peer stores are ordered
    // before the bar.sync above, the system-scope fence publishes them, then it counts itself.
    // The last CTA of this rank signals every peer and waits for their signals: a completed
    // kernel means "all ranks' tiles have landed here".
    __threadfence_system();
    const unsigned done = atomicAdd(p.peer_done, 1u);
    if (done == gridDim.x - 1) {
      __threadfence();
      *p.peer_done = 0u;
      for (int q = 0; q < p.n_peers; ++q)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.peer_signal[q]), "r"(p.peer_epoch)
                     : "memory");
      const long long t0 = clock64();
      for (int q = 0; q < p.n_peers; ++q) {
        const uint32_t* flag = p.peer_signal_self + p.peer_rank[q];
        uint32_t v;
        do {
          asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
          if ((int)(v - p.peer_epoch) >= 0) break;
          __nanosleep(200);
          if (clock64() - t0 > 60000000000LL) __trap();  // (~30 s) a peer never arrived: fail, do not hang
        } while (true);
      }
    }
  }
  if (tid < 32) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace nfi
