// Forward render kernel, warp-specialised tensor-core variant (the default).
//
// Same arithmetic and tiles as nfi_forward_tc.cuh, but the per-step chain
//   gather -> MMA1 -> softplus/split -> MMA2 -> density/colour/composite
// is cut into two roles that run concurrently and are coupled only by
// mbarriers over a 2-stage ring (shared-memory A/H tiles + TMEM columns):
//
//   CTA = 768 threads = 2 groups; a group = one 16x8-pixel tile (128 rays) served
//   by 4 consumer warps and TWO sets of 4 producer warps (set p gathers the
//   steps whose ring stage is p, so two steps of a tile are gathered at once).
//   Register budget by setmaxnreg: producers 72, consumers 112 per thread.
//     PRODUCERS.  Lane l of producer warp w owns
//        ray 32w+l only to place its sample and compute the bilinear taps;
//        the warp gathers cooperatively (8 lanes per 128-byte texel) into the
//        stage's A_hi/A_lo tiles; producer thread 0 issues the 12 layer-1
//        tcgen05.mma and commits to d1_full[stage].  Sample positions do not
//        depend on the network's output inside a pass, so producers run up
//        to one full step ahead of the consumers and never wait on an MMA.
//     CONSUMERS (thread = ray).  Wait d1_full,
//        read their TMEM lane, bias + softplus, hand H_hi (shared memory, in
//        the stage's A tiles) and H_lo (TMEM, over D1) back; consumer thread 0
//        issues the 24 layer-2 tcgen05.mma -> d2_full[stage]; read D2,
//        release the stage (stage_free), then density / colour / compositing
//        and all per-ray state (coarse weights, resampling, sorted merge).
//
// The kernel is persistent: grid = #SMs, every group walks a strided list of
// tiles, so the coarse-sample scratch slabs are per (SM, group) and stay in L2
// (148 x 2 x 229 KB = 68 MB at S = 64).
#pragma once
#include "nfi_forward_tc.cuh"

namespace nfi {

constexpr int kWsGroups = 2;
constexpr int kWsThreads = 768;   // 2 consumer + 4 producer warpgroups
constexpr int kWsStages = 2;
// setmaxnreg moves registers only inside the CTA's own launch allocation
// (768 threads x 80): what the 512 producer threads give up, 512 x (80 - 72) = 4,096,
// is exactly what the 256 consumer threads take, 256 x (96 - 80).
constexpr int kWsProducerRegs = 72;
constexpr int kWsConsumerRegs = 96;
constexpr int kWsStageBytes = 32768;  // A_hi + A_lo (later H_hi k-blocks 0/1)
constexpr int kWsSmA = 25600;
constexpr int kWsSmPal = kWsSmA + kWsGroups * kWsStages * kWsStageBytes;  // 156672
constexpr int kWsSmBars = kWsSmPal + kWsGroups * 48 * 4;
// per group: d1_full[2], d2_full[2], stage_free[2], zf_ready  (7) ; + weights barrier
constexpr int kWsBarsPerGroup = 8;
constexpr int kWsSmTmemPtr = kWsSmBars + (kWsGroups * kWsBarsPerGroup + 1) * 8;
constexpr int kWsSmBytes = kWsSmTmemPtr + 16;
// TMEM columns: group g, stage s at 256 g + 128 s: [0,64) D1 / H_lo, [64,80) D2

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

template <int NOUT_PAD, int EXTRA, bool FINE>
__global__ void __launch_bounds__(kWsThreads, 1)
render_forward_ws(const nfi_render_params p, const unsigned char* __restrict__ wimg,
                  float* __restrict__ scratch) {
  constexpr int NE = (EXTRA == 1) ? 3 : 0;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = smem_raw;
  const int tid = threadIdx.x, lane = tid & 31;
  // warpgroups: 0,1 = consumers of group 0,1 ; 2,3 = producer sets 0,1 of group 0 ; 4,5 = of group 1
  const int wg = __shfl_sync(kFull, tid >> 7, 0);  // warp-uniform, and the compiler knows it
  const int role = (wg < 2) ? 1 : 0;            // 0 producer, 1 consumer
  const int g = (wg < 2) ? wg : ((wg - 2) >> 1);  // group
  const int pset = (wg - 2) & 1;                // producer set = ring stage it fills
  const int gt = tid & 127;          // thread within its warpgroup
  const int wig = __shfl_sync(kFull, gt >> 5, 0);  // warp within warpgroup (= TMEM lane quadrant)
  const int S = p.num_samples;

  uint64_t* bars = reinterpret_cast<uint64_t*>(base + kWsSmBars);
  uint64_t* gb = bars + g * kWsBarsPerGroup;
  uint64_t* d1_full = gb;        // [2]
  uint64_t* d2_full = gb + 2;    // [2]
  uint64_t* stage_free = gb + 4; // [2]
  uint64_t* zf_ready = gb + 6;
  uint64_t* wbar = bars + kWsGroups * kWsBarsPerGroup;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(base + kWsSmTmemPtr);
  const float* b1s = reinterpret_cast<const float*>(base + kWiB1);
  const float* b2s = reinterpret_cast<const float*>(base + kWiB2);
  float* pal = reinterpret_cast<float*>(base + kWsSmPal) + g * 48;

  if (tid == 0) {
    if (tc::smem_u32(base) & 1023u) __trap();
    for (int i = 0; i < kWsGroups; ++i) {
      uint64_t* q = bars + i * kWsBarsPerGroup;
      tc::mbar_init(q + 0, 1);
      tc::mbar_init(q + 1, 1);
      tc::mbar_init(q + 2, 1);
      tc::mbar_init(q + 3, 1);
      tc::mbar_init(q + 4, kThreads);
      tc::mbar_init(q + 5, kThreads);
      tc::mbar_init(q + 6, kThreads);
    }
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

  if (role == 0)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kWsProducerRegs));
  else
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kWsConsumerRegs));

  const uint32_t base_s = tc::smem_u32(base);
  const uint32_t w1_hi_s = base_s + kWiW1Hi, w1_lo_s = base_s + kWiW1Lo;
  const uint32_t w2_hi_s = base_s + kWiW2Hi, w2_lo_s = base_s + kWiW2Lo;
  unsigned char* const stage0 = base + kWsSmA + g * kWsStages * kWsStageBytes;
  const uint32_t stage0_s = tc::smem_u32(stage0);
  const uint32_t d_tmem0 = tmem_base + g * 256;
  auto stage_ptr = [&](int st) { return stage0 + st * kWsStageBytes; };
  auto stage_s = [&](int st) { return stage0_s + (uint32_t)st * kWsStageBytes; };
  auto d_tmem = [&](int st) { return d_tmem0 + (uint32_t)st * 128; };
  // UMMA descriptor bases, built once (a stage is 32 KB = 2048 descriptor units)
  const uint64_t dsc_stage0 = tc::umma_desc_sw128(stage0_s);
  const uint64_t dsc_w1_hi = tc::umma_desc_sw128(w1_hi_s), dsc_w1_lo = tc::umma_desc_sw128(w1_lo_s);
  const uint64_t dsc_w2_hi = tc::umma_desc_sw128(w2_hi_s), dsc_w2_lo = tc::umma_desc_sw128(w2_lo_s);

  const int tiles_x = (p.width + kTileW - 1) / kTileW;
  const int tiles_y = (p.height + kTileH - 1) / kTileH;
  const int n_tiles = tiles_x * tiles_y * p.batch;
  const int R = p.plane_res;
  const float inv_range = 1.f / p.scene_range;
  const bool explicit_noise = (p.noise_mode == NFI_NOISE_EXPLICIT);
  // timing experiments only (bench.py --mlp-mode 0x102 / 0x202): results are garbage
  const bool dbg_skip_gather = (p.mlp_mode & 0x100) != 0;
  const bool dbg_skip_consumer = (p.mlp_mode & 0x200) != 0;
  float* slab = scratch + ((size_t)blockIdx.x * kWsGroups + g) * tc_scratch_floats_per_group(S);
  float4* sc_srgb = reinterpret_cast<float4*>(slab);
  float* sc_t = slab + (size_t)4 * S * kThreads;
  float* sc_w = sc_t + (size_t)S * kThreads;
  float* sc_zf = sc_w + (size_t)S * kThreads;

  // phase timers (timing experiments: mlp_mode & 0x1000, buffer passed in p.normals)
  const bool dbg_time = (p.mlp_mode & 0x1000) && p.normals != nullptr && blockIdx.x == 0 &&
                        g == 0 && gt == 0 && (role == 1 || pset == 0);
  long long tacc[6] = {0, 0, 0, 0, 0, 0};
  long long tprev = 0;
#define NFI_T(i)                              \
  if (dbg_time) {                             \
    const long long now__ = clock64();        \
    tacc[i] += now__ - tprev;                 \
    tprev = now__;                            \
  }
  uint32_t n = 0;       // steps issued so far by this thread's role (ring position)
  uint32_t tile_it = 0; // tiles done (parity of zf_ready)

  for (int tile = blockIdx.x * kWsGroups + g; tile < n_tiles;
       tile += gridDim.x * kWsGroups, ++tile_it) {
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

    if (role == 0) {
      // ============================ PRODUCER ============================
      const float* planes_b = p.planes + (size_t)b * 3 * R * R * kC;
      for (int pass = 0; pass < (FINE ? 2 : 1); ++pass) {
        if (pass == 1) tc::mbar_wait(zf_ready, tile_it & 1);
        for (int s = 0; s < S; ++s, ++n) {
          const int st = n & 1;
          if (st != pset) continue;  // the other producer set fills that stage
          if (dbg_time) tprev = clock64();
          tc::mbar_wait(&stage_free[st], ((n >> 1) & 1) ^ 1);
          NFI_T(0)
          float t;
          if (pass == 0) {
            t = lerp_torch(r.tnear, r.tfar, (float)s / (float)S);
            if (explicit_noise) t = t + p.noise_t[ray * S + s] * (span / (float)S);
          } else {
            t = ld_relaxed(sc_zf + s * kThreads + gt);
          }
          const float x0 = (r.ox + r.dx * t) * inv_range, x1 = (r.oy + r.dy * t) * inv_range,
                      x2 = (r.oz + r.dz * t) * inv_range;
          PackedTaps tp;
          pack_taps(x0, x1, R, tp.o[0], tp.fx[0], tp.fy[0]);
          pack_taps(x0, x2, R, tp.o[1], tp.fx[1], tp.fy[1]);
          pack_taps(x1, x2, R, tp.o[2], tp.fx[2], tp.fy[2]);
          if (p.mlp_mode & 0x800) {  // timing experiment: every tap inside a 24 KB window
            tp.o[0] &= 0x3Fu;
            tp.o[1] &= 0x3Fu;
            tp.o[2] &= 0x3Fu;
          }
          NFI_T(1)
          if (!dbg_skip_gather)
            gather_to_tiles_deep(planes_b, R, tp, stage_ptr(st), stage_ptr(st) + 16384, 32 * wig,
                                 lane);
          NFI_T(2)
          tc::fence_async_smem();
          tc::bar_sync(3 + 2 * g + pset, kThreads);
          NFI_T(3)
          if (wig == 0) {
            if (elect_one()) {
              tc::tc_fence_after();
              const uint64_t a_hi_d = dsc_stage0 + (uint64_t)st * (kWsStageBytes >> 4);
              tc::issue_layer1_d(d_tmem(st), a_hi_d, a_hi_d + (16384 >> 4), dsc_w1_hi, dsc_w1_lo);
              tc::umma_commit(&d1_full[st]);
            }
            __syncwarp();
          }
          NFI_T(4)
        }
      }
    } else {
      // ============================ CONSUMER ============================
      tc::bar_sync(1 + g, kThreads);  // previous tile's palette no longer in use
      if (gt < 48)
        pal[gt] = (p.n_attention > 0 && gt < p.n_attention * 3)
                      ? p.palette[(size_t)b * p.n_attention * 3 + gt]
                      : 0.f;
      tc::bar_sync(1 + g, kThreads);
      FieldConst fc;
      fc.A = p.n_attention;
      fc.use_sdf = p.use_sdf;
      fc.inv_beta = p.use_sdf ? 1.f / p.beta[0] : 0.f;
      fc.inv_alpha = p.use_sdf ? 1.f / p.alpha[0] : 0.f;
      Compositor<NE, true> comp;
      comp.init();

      // the decoder outputs of this thread's point for ring position n
      auto consume = [&](float (&out)[NOUT_PAD]) {
        const int st = n & 1;
        const uint32_t par = (n >> 1) & 1;
        const uint32_t d_lane = d_tmem(st) + ((uint32_t)(32 * wig) << 16);
        if (dbg_time && tprev == 0) tprev = clock64();
        NFI_T(5)
        tc::mbar_wait(&d1_full[st], par);
        tc::tc_fence_after();
        NFI_T(0)
        if (dbg_skip_consumer) {
          tc::tc_fence_before();
          mbar_arrive(&stage_free[st]);
#pragma unroll
          for (int o = 0; o < NOUT_PAD; ++o) out[o] = 0.f;
          ++n;
          return;
        }
#pragma unroll 1
        for (int c2 = 0; c2 < 2; ++c2) {  // two 16-column TMEM loads in flight
          uint32_t ra[16], rb[16];
          tc::tmem_ld16_nowait(d_lane + 32 * c2, ra);
          tc::tmem_ld16_nowait(d_lane + 32 * c2 + 16, rb);
          tc::tmem_wait_ld();
          unsigned char* hrow = stage_ptr(st) + ((c2 == 0) ? 0 : 16384);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(half ? rb[i] : ra[i]);
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const float4 bb =
                  *reinterpret_cast<const float4*>(b1s + 32 * c2 + 16 * half + 4 * i4);
              const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
              float hi[4];
#pragma unroll
              for (int i = 0; i < 4; i += 2) {
                // softplus(x) = max(x,0) + ln2 * lg2(1 + 2^(-|x| log2 e)), two lanes packed
                const float2 x = __fadd2_rn(make_float2(v[4 * i4 + i], v[4 * i4 + i + 1]),
                                            make_float2(bv[i], bv[i + 1]));
                const float2 ax = __fmul2_rn(make_float2(fabsf(x.x), fabsf(x.y)),
                                             make_float2(-1.4426950408889634f, -1.4426950408889634f));
                const float2 l = make_float2(tc::lg2_approx(1.f + tc::ex2_approx(ax.x)),
                                             tc::lg2_approx(1.f + tc::ex2_approx(ax.y)));
                const float2 h = __ffma2_rn(l, make_float2(0.6931471805599453f, 0.6931471805599453f),
                                            make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)));
                hi[i] = tc::tf32_hi(h.x);
                hi[i + 1] = tc::tf32_hi(h.y);
                const float2 lo2 = __fadd2_rn(h, make_float2(-hi[i], -hi[i + 1]));
                v[4 * i4 + i] = lo2.x;
                v[4 * i4 + i + 1] = lo2.y;
              }
              const uint32_t off = tc::sw128_offset(gt, half * 4 + i4);
              *reinterpret_cast<float4*>(hrow + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            }
            tc::tmem_st16(d_lane + 32 * c2 + 16 * half, v);
          }
        }
        NFI_T(1)
        tc::tmem_wait_st();
        tc::fence_async_smem();
        tc::tc_fence_before();
        tc::bar_sync(1 + g, kThreads);
        NFI_T(2)
        if (wig == 0) {
          if (elect_one()) {
            tc::tc_fence_after();
            tc::issue_layer2_d(d_tmem(st) + 64, d_tmem(st),
                               dsc_stage0 + (uint64_t)st * (kWsStageBytes >> 4), dsc_w2_hi,
                               dsc_w2_lo);
            tc::umma_commit(&d2_full[st]);
          }
          __syncwarp();
        }
        NFI_T(3)
        tc::mbar_wait(&d2_full[st], par);
        tc::tc_fence_after();
        NFI_T(4)
        float v[16];
        tc::tmem_ld16(d_lane + 64, v);
        tc::tc_fence_before();
        mbar_arrive(&stage_free[st]);  // A/H tiles and TMEM columns of the stage are free
#pragma unroll
        for (int o = 0; o < NOUT_PAD; ++o) out[o] = v[o] + b2s[o];
        ++n;
      };
      auto shade = [&](float t, float& sigma, float& cr, float& cg, float& cb, float* ex) {
        const float wx = r.ox + r.dx * t, wy = r.oy + r.dy * t, wz = r.oz + r.dz * t;
        const float x0 = wx * inv_range, x1 = wy * inv_range, x2 = wz * inv_range;
        const float keep =
            (fabsf(x0) > 1.f || fabsf(x1) > 1.f || fabsf(x2) > 1.f) ? 0.f : 1.f;
        float out[NOUT_PAD];
        consume(out);
        float probs[NOUT_PAD];
        field_head<NOUT_PAD, true>(out, fc, pal, keep, sigma, cr, cg, cb, probs);
        if (EXTRA == 1) {
          ex[0] = wx;
          ex[1] = wy;
          ex[2] = wz;
        }
      };

      // ---------------- coarse pass ----------------
      float wT = 1.f, prev_t = 0.f, prev_s = 0.f;
      for (int s = 0; s < S; ++s) {
        float t = lerp_torch(r.tnear, r.tfar, (float)s / (float)S);
        if (explicit_noise) t = t + p.noise_t[ray * S + s] * (span / (float)S);
        float sigma, cr, cg, cb;
        float ex[NE > 0 ? NE : 1];
        shade(t, sigma, cr, cg, cb, ex);
        if (FINE) {
          sc_srgb[s * kThreads + gt] = make_float4(sigma, cr, cg, cb);
          sc_t[s * kThreads + gt] = t;
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
      }

      if (FINE) {
        sc_w[(S - 1) * kThreads + gt] = 0.f;
        // producers are parked on zf_ready and both stages are drained: the
        // group's 64 KB of tile memory serve as S x 128 float columns
        float* col = reinterpret_cast<float*>(stage_ptr(0)) + gt;
        float sum = 0.f;
        {
          float wa = sc_w[gt], wb = sc_w[kThreads + gt], wc;
          for (int m = 0; m + 2 < S; ++m) {
            wc = sc_w[(m + 2) * kThreads + gt];
            sum += ((fmaxf(wa, wb) + fmaxf(wb, wc)) * 0.5f + 0.01f) + 1e-5f;
            wa = wb;
            wb = wc;
          }
        }
        if (explicit_noise) {
          for (int k = 0; k < S; ++k) col[k * kThreads] = p.noise_u[ray * S + k];
          column_sort(col, S, kThreads);
        } else {
          for (int k = 0; k < S; ++k) col[k * kThreads] = linspace01(k, S);
        }
        {
          int k = 0;
          float c_prev = 0.f;
          float wa = sc_w[gt], wb = sc_w[kThreads + gt], wc;
          float t_lo = sc_t[gt], t_mid = sc_t[kThreads + gt];
          float z0 = 0.5f * (t_mid + t_lo);
          for (int i = 1; i + 1 < S; ++i) {
            wc = sc_w[(i + 1) * kThreads + gt];
            const float pw = ((fmaxf(wa, wb) + fmaxf(wb, wc)) * 0.5f + 0.01f) + 1e-5f;
            wa = wb;
            wb = wc;
            const float c_i = c_prev + pw / sum;
            const float t_hi = sc_t[(i + 1) * kThreads + gt];
            const float z1 = 0.5f * (t_hi + t_mid);
            float den = c_i - c_prev;
            if (den < 1e-5f) den = 1.f;
            while (k < S) {
              const float u = col[k * kThreads];
              if (!(u < c_i)) break;
              col[k * kThreads] = z0 + (u - c_prev) / den * (z1 - z0);
              ++k;
            }
            c_prev = c_i;
            t_mid = t_hi;
            z0 = z1;
          }
          while (k < S) {
            col[k * kThreads] = z0;
            ++k;
          }
        }
        if (p.z_fine != nullptr && valid)
          for (int k = 0; k < S; ++k) p.z_fine[ray * S + k] = col[k * kThreads];
        for (int k = 0; k < S; ++k) sc_zf[k * kThreads + gt] = col[k * kThreads];
        __threadfence_block();
        mbar_arrive(zf_ready);  // producers may start the fine pass

        // ------- fine pass + sorted merge + compositing -------
        // The next TWO coarse samples wait in registers, so taking one never
        // stalls on the (L2) load of its successor.
        int c = 0;
        float ct0 = sc_t[gt], ct1 = sc_t[kThreads + gt];
        float4 cq0 = sc_srgb[gt], cq1 = sc_srgb[kThreads + gt];
        auto take_coarse = [&]() {
          float ce[NE > 0 ? NE : 1];
          if (EXTRA == 1) {
            ce[0] = r.ox + r.dx * ct0;
            ce[1] = r.oy + r.dy * ct0;
            ce[2] = r.oz + r.dz * ct0;
          }
          comp.push(ct0, cq0.x, cq0.y, cq0.z, cq0.w, ce, r.dn);
          ++c;
          ct0 = ct1;
          cq0 = cq1;
          if (c + 1 < S) {
            ct1 = sc_t[(c + 1) * kThreads + gt];
            cq1 = sc_srgb[(c + 1) * kThreads + gt];
          }
        };
        for (int k = 0; k < S; ++k) {
          const float z = sc_zf[k * kThreads + gt];
          float sigma, cr, cg, cb;
          float ex[NE > 0 ? NE : 1];
          shade(z, sigma, cr, cg, cb, ex);
          while (c < S && ct0 <= z) take_coarse();
          comp.push(z, sigma, cr, cg, cb, ex, r.dn);
        }
        while (c < S) take_coarse();
      }

      if (valid) {
        float bg = 0.f;
        if (p.white_background) bg = 1.f - comp.am;
        p.rgb[ray * 3 + 0] = comp.ar + bg;
        p.rgb[ray * 3 + 1] = comp.ag + bg;
        p.rgb[ray * 3 + 2] = comp.ab + bg;
        p.depth[ray] = comp.ad;
        p.mask[ray] = comp.am;
        if (EXTRA == 1 && p.extra != nullptr)
          for (int a = 0; a < 3; ++a) p.extra[ray * 3 + a] = comp.ae[a];
      }
    }
  }
  if (dbg_time) {
    float* dbg = p.normals + (role == 1 ? 8 : 0);
    for (int i = 0; i < 6; ++i) dbg[i] = (float)tacc[i];
    dbg[6] = (float)n;
  }
#undef NFI_T
  tc::tc_fence_before();
  __syncthreads();
  if (tid < 32) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace nfi
