// Persistent CTA-pair GEMM / implicit-GEMM 3x3 convolution with a TMA-store epilogue (sm_100a).
//
//   C[M,N] = epilogue( A[M,K] * B[N,K]^T )     bf16 operands, fp32 accumulation in TMEM.
//
// Why this kernel exists (profiles/r02_gemm_epilogue.txt): the earlier persistent kernels drained TMEM with
// thread = row, every thread storing 16 bytes at a row stride -- 32 separate 128-byte lines per store instruction.  ncu
// showed the epilogue warps parked on lg_throttle and the tensor pipe waiting on them (tile time 9.5 us with a bf16
// output, 15.8 us with an fp32 output, same main loop).  Here the epilogue never touches global memory with LD/ST:
//   TMEM --tcgen05.ld--> registers --(column ops)--> 128B-swizzled smem chunk --TMA store / TMA reduce-add--> HBM
// and a bf16 residual operand arrives the same way (TMA load into swizzled smem, prefetched two chunks ahead).
//
// One cluster of two CTAs (one SM pair) loops over 256 x BN output tiles (BN = 64/128/192/256, or 384 with a single
// accumulator for the N = 768 projections, whose 58-60 tiles then fill one round of SM pairs instead of two half-empty ones):
//   warp 0 / 2  : TMA producers, one lane each: warp 0 loads this CTA's 128 rows of A, warp 2 this CTA's HALF of the BN
//                 weight rows of B -> kStages ring (a single issuing thread paced the main loop, see below)
//   warp 1      : TMEM allocator; in the leader CTA the single tcgen05.mma.cta_group::2 issuer (M = 256)
//   warps 3..10 : two epilogue warpgroups; chunk g (64 columns of a tile) belongs to warpgroup g & 1.  Each warpgroup owns
//                 two 16 KB staging boxes and issues its own TMA stores: elected thread waits for the bulk group that last
//                 read the box, 128-thread named barrier, everybody writes, proxy fence, barrier, elected thread stores.
//                 (r02 first version: a separate store-DMA thread behind two mbarriers per chunk -- the timeline showed
//                 300-450 ns per hand-off and 1.6 us per fp32 chunk, 6.5 us of un-overlapped tail on a one-tile launch.)
// TMEM holds two accumulators (2 x BN columns): the epilogue of tile i overlaps the main loop of tile i+1.
//   full[s]/empty[s]      smem ring across tiles (full lives in the leader: both CTAs' bytes are credited there)
//   tmem_full[2]          multicast tcgen05.commit per finished accumulator
//   tmem_empty[2] (leader) one arrival per epilogue warp of both CTAs once its last tcgen05.ld of the tile retired
//   res_full[8][2]        per epilogue warp: its 32-row slice of a bf16 residual chunk has landed
// fp32 residual == C (the transformer's  x += fc2(...)): no residual read at all, the store is a TMA reduce-add
// (cp.reduce.async.bulk .add.f32), which rounds exactly like acc + x.
// Conv mode: the A tile of filter tap (ky,kx) is a 4-D TMA box of the NHWC input shifted by the tap offset (TMA zero fill
// = padding); the output chunk goes back through a 4-D box [64 ch, bw, bh, 1], clipped at the image border by TMA.
#include <stdlib.h>
#include <string.h>

#include "gemm_common.cuh"

namespace sgf {

static constexpr int kTsEpiWarps = 8;
static constexpr int kTsThreads = 96 + 32 * kTsEpiWarps;
static constexpr int kTsMaxSmem = 232448;  // 227 KB

template <int BN, int kEpi>
struct TsCfg {
  static constexpr bool kOutF32 = (kEpi & kEpiOutF32) != 0;
  static constexpr bool kResB16 = (kEpi & kEpiResBf16) != 0;
  static constexpr int kChunks = BN / 64;
  static constexpr int kNSplit = BN > 256 ? 2 : 1;           // tcgen05.mma N <= 256: a 384-wide tile is two N = 192 MMAs
  static constexpr int kUmmaN = BN / kNSplit;
  static constexpr int kAccBufs = 2 * BN <= 512 ? 2 : 1;     // 512 TMEM columns: BN = 384 has ONE accumulator
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = (BN / 2) * BK * 2;
  static constexpr int kBSplitBytes = (kUmmaN / 2) * BK * 2;  // this CTA's half of one N-split of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBox = 16384;  // one TMA-store box: 128 rows x 128 B (64 bf16 / 32 fp32 columns)
  // boxes per warpgroup: fp32 two (the store of one 32-column half runs under the arithmetic of the other), bf16 one (a
  // 64-column chunk is one box; its store has long been read when the next chunk's arithmetic is done) -- every 32 KB not
  // spent here is one more pipeline stage, and the main loop is bound by bytes in flight (r02_gemm_ts_k_sweep.txt)
  static constexpr int kWgBoxes = kOutF32 ? 2 : 1;
  static constexpr int kResBytes = kResB16 ? 4 * 16384 : 0;  // two 128 x 64 bf16 chunks per warpgroup
  static constexpr int kEpiBytes = 2 * kWgBoxes * kBox + kResBytes;
  static constexpr int kBarBytes = 512;
  static constexpr int kStagesFit = (kTsMaxSmem - 1024 - kEpiBytes - kBarBytes) / kStageBytes;
  static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
  static constexpr int kSmem = 1024 + kStages * kStageBytes + kEpiBytes + kBarBytes;
  static constexpr int kTmemCols = kAccBufs * BN <= 128 ? 128 : (kAccBufs * BN <= 256 ? 256 : 512);
  static_assert(kStages >= 3, "pipeline too shallow");
};

struct TsBars {
  uint64_t full[8], empty[8], tmem_full[2], tmem_empty[2], res_full[kTsEpiWarps][2];
  uint32_t tmem_slot;
};
static_assert(sizeof(TsBars) <= 512, "barrier block");

template <int BN, bool kConv, int kEpi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTsThreads, 1)
    gemm_ts_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const GemmShape shp,
                   const GemmEpilogue ep, const int m_tiles, const int num_m_pairs, const int num_tiles,
                   unsigned long long* trace, const int dual_producers) {
  using Cfg = TsCfg<BN, kEpi>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kChunks = Cfg::kChunks;
  constexpr bool kOutF32 = Cfg::kOutF32;
  constexpr bool kResB16 = Cfg::kResB16;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_out = smem + kStages * Cfg::kStageBytes;                // [2 warpgroups][kWgBoxes][16 KB box]
  uint8_t* smem_res = smem_out + 2 * Cfg::kWgBoxes * Cfg::kBox;         // [2 warpgroups][2][16 KB]
  TsBars* bars = reinterpret_cast<TsBars*>(smem_out + Cfg::kEpiBytes);

  pdl_trigger();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  // debug timeline (sgf_debug_set_gemm_trace): CTA 0 and the last CTA record globaltimer (ns) at 64 event slots per role
  const int tslot = !trace ? -1 : (blockIdx.x == 0 ? 0 : (blockIdx.x == gridDim.x - 2 ? 1 : -1));
  auto tr = [&](int role, int ev) {
    if (tslot >= 0 && ev < 64) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      trace[(tslot * 8 + role) * 64 + ev] = t;
    }
  };
  if (threadIdx.x == 0) tr(0, 0);
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int num_kb = (shp.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    if constexpr (kResB16) tma_prefetch_desc(&tmR);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bars->full[s], 2);  // the A producer and the B producer of the leader each arrive with their bytes
      mbar_init(&bars->empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->tmem_full[i], 1);
      // BN = 64: one chunk per tile, tile i is drained by warpgroup i & 1 alone (4 warps in each CTA)
      mbar_init(&bars->tmem_empty[i], kChunks == 1 ? kTsEpiWarps : 2 * kTsEpiWarps);
    }
    for (int w = 0; w < kTsEpiWarps; ++w) {
      mbar_init(&bars->res_full[w][0], 1);
      mbar_init(&bars->res_full[w][1], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2cta<Cfg::kTmemCols>(&bars->tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;
  if (threadIdx.x == 0) tr(0, 1);
  pdl_wait();
  if (threadIdx.x == 0) tr(0, 2);

  // tile -> this CTA's 128-row m-tile (may be one past the end when the number of m-tiles is odd: loads are then fully
  // out of bounds = zero fill, stores are skipped) and the first column
  auto decode = [&](int t, int& mt, int& n0, int& m0, int& img, int& h0, int& w0) {
    const int nt = t / num_m_pairs;
    mt = (t - nt * num_m_pairs) * 2 + static_cast<int>(rank);
    n0 = nt * BN;
    m0 = mt * BM;
    img = 0; h0 = 0; w0 = 0;
    if constexpr (kConv) {
      const int per_img = shp.tiles_w * shp.tiles_h;
      img = mt / per_img;
      const int r = mt - img * per_img;
      h0 = (r / shp.tiles_w) * shp.bh;
      w0 = (r % shp.tiles_w) * shp.bw;
    }
  };

  if (warp == 0 || warp == 2) {
    // ================================ TMA producers: warp 0 loads A, warp 2 loads B ================================
    // (one thread issuing every load paced the main loop: a cp.async.bulk.tensor costs the issuing thread ~80 ns, three per
    //  k-block is 500 clocks against 2 x BN = 256..512 tensor-core clocks -- profiles/r02_gemm_ts_timeline.txt)
    if (lane == 0 && (dual_producers || warp == 0)) {
      const bool load_a = warp == 0;
      const bool load_b = warp == 2 || !dual_producers;
      uint32_t g = 0;  // running k-block counter across tiles
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        int mt, n0, m0, img, h0, w0;
        decode(t, mt, n0, m0, img, h0, w0);
#pragma unroll 1
        for (int kb = 0; kb < num_kb; ++kb, ++g) {
          const int s = g % kStages;
          mbar_wait(&bars->empty[s], ((g / kStages) & 1) ^ 1);
          uint8_t* sa = smem + s * Cfg::kStageBytes;
          if (load_a) {
            tr(1, static_cast<int>(g));
            if (rank == 0) mbar_expect_tx(&bars->full[s], 2 * Cfg::kABytes);
            if constexpr (kConv) {
              const int tap = kb / shp.cin_blocks;
              const int kc = kb - tap * shp.cin_blocks;
              const int ky = tap / shp.kw, kx = tap - ky * shp.kw;
              tma_load_4d_2cta(sa, &tmA, &bars->full[s], kc * BK, w0 * shp.stride_w + kx - shp.pad_w,
                               h0 * shp.stride_h + ky - shp.pad_h, img);
            } else {
              tma_load_2d_2cta(sa, &tmA, &bars->full[s], kb * BK, m0);
            }
          }
          if (load_b) {
            if (rank == 0) mbar_expect_tx(&bars->full[s], 2 * Cfg::kBBytes);
#pragma unroll
            for (int hs = 0; hs < Cfg::kNSplit; ++hs)
              tma_load_2d_2cta(sa + Cfg::kABytes + hs * Cfg::kBSplitBytes, &tmB, &bars->full[s], kb * BK,
                               n0 + hs * Cfg::kUmmaN + static_cast<int>(rank) * (Cfg::kUmmaN / 2));
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA) ================================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, Cfg::kUmmaN, 0, 0);
      uint32_t g = 0;
      int it = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
        const int ab = it % Cfg::kAccBufs;
        mbar_wait(&bars->tmem_empty[ab], ((it / Cfg::kAccBufs) & 1) ^ 1);  // both CTAs drained this accumulator
        tc_fence_after();
#pragma unroll 1
        for (int kb = 0; kb < num_kb; ++kb, ++g) {
          const int s = g % kStages;
          mbar_wait(&bars->full[s], (g / kStages) & 1);
          tc_fence_after();
          tr(2, static_cast<int>(g));
          const uint32_t sa = smem_u32(smem + s * Cfg::kStageBytes);
          const uint64_t da = make_smem_desc_sw128(sa);
          const uint64_t db = make_smem_desc_sw128(sa + Cfg::kABytes);
#pragma unroll
          for (int hs = 0; hs < Cfg::kNSplit; ++hs)
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_f16_2cta(tmem_base + ab * BN + hs * Cfg::kUmmaN, da + 2 * k, db + hs * (Cfg::kBSplitBytes >> 4) + 2 * k, idesc,
                            (kb | k) != 0 ? 1u : 0u);
          umma_commit_2cta(&bars->empty[s], 0b11);
        }
        umma_commit_2cta(&bars->tmem_full[ab], 0b11);
        tr(3, it);
      }
    }
  } else {
    // ================================ epilogue warpgroups ================================
    const int ew = warp - 3;
    const int wg = ew >> 2;
    const int quarter = warp & 3;  // the TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;
    uint8_t* stage0 = smem_out + wg * Cfg::kWgBoxes * Cfg::kBox;  // this warpgroup's boxes
    const bool elected = (ew & 3) == 0 && lane == 0;  // issues this warpgroup's TMA stores (bulk groups are per thread)
    uint32_t bx = 0;                                  // running box counter of this warpgroup
    int st_w0 = 0, st_h0 = 0, st_img = 0;
    bool tile_valid = false;
    uint8_t* res_slice = smem_res + wg * 2 * 16384 + quarter * 4096;  // + (k & 1) * 16384
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    constexpr int csz = kOutF32 ? 4 : 2;
    (void)csz;

    // chunk sequence of this warpgroup: g = wg, wg + 2, ...  ->  (tile iteration, chunk in tile)
    auto chunk_at = [&](uint32_t g, int& it, int& c, int& t) {
      it = static_cast<int>(g / kChunks);
      c = static_cast<int>(g - static_cast<uint32_t>(it) * kChunks);
      t = cluster_id + it * num_clusters;
    };
    auto issue_residual = [&](uint32_t g, int rb) {  // lane 0: this warp's 32 rows of the residual chunk g
      int it, c, t;
      chunk_at(g, it, c, t);
      if (t >= num_tiles) return;
      int mt, n0, m0, img, h0, w0;
      decode(t, mt, n0, m0, img, h0, w0);
      mbar_expect_tx(&bars->res_full[ew][rb], 4096);
      tma_load_2d(res_slice + rb * 16384, &tmR, &bars->res_full[ew][rb], n0 + c * 64, m0 + quarter * 32);
    };
    if constexpr (kResB16) {
      if (lane == 0) {
        issue_residual(wg, 0);
        issue_residual(wg + 2, 1);
      }
    }

    int cur_it = -1;
    int n0 = 0, m0 = 0;
    int64_t row = 0;
    bool row_ok = false;
    float rn_mean = 0.f, rn_rstd = 1.f;
#pragma unroll 1
    for (uint32_t g = wg, k = 0;; g += 2, ++k) {
      int it, c, t;
      chunk_at(g, it, c, t);
      if (t >= num_tiles) break;
      const int ab = it % Cfg::kAccBufs;
      if (it != cur_it) {
        cur_it = it;
        int mt, img, h0, w0;
        decode(t, mt, n0, m0, img, h0, w0);
        st_w0 = w0; st_h0 = h0; st_img = img;
        tile_valid = mt < m_tiles;
        if constexpr (kConv) {
          const int hh = h0 + r / shp.bw, ww = w0 + r % shp.bw;
          row_ok = mt < m_tiles && hh < shp.H && ww < shp.W;
          row = (static_cast<int64_t>(img) * shp.H + hh) * shp.W + ww;
        } else {
          row = m0 + r;
          row_ok = row < shp.M;
        }
        if constexpr ((kEpi & kEpiRowNorm) != 0) {
          const int64_t rrow = row_ok ? row : 0;
          const float4* sp = reinterpret_cast<const float4*>(ep.rownorm_stats) + rrow * (ep.rownorm_parts / 2);
          float s0 = 0.f, s1 = 0.f;
#pragma unroll 4
          for (int q = 0; q < ep.rownorm_parts / 2; ++q) {
            const float4 u = __ldg(sp + q);
            s0 += u.x; s1 += u.y; s0 += u.z; s1 += u.w;
          }
          rn_mean = s0 * ep.rownorm_inv_dim;
          rn_rstd = rsqrtf(fmaxf(s1 * ep.rownorm_inv_dim - rn_mean * rn_mean, 0.f) + 1e-5f);
        }
        mbar_wait(&bars->tmem_full[ab], (it / Cfg::kAccBufs) & 1);
        tc_fence_after();
        if (lane == 0 && quarter == 3) tr(6, it * 2 + wg);
      }
      const bool last_in_tile = c + 2 >= kChunks;  // this warpgroup's last chunk of the tile
      const int col0 = n0 + c * 64;

      uint4 rq[8];
      if constexpr (kResB16) {
        const int rb = k & 1;
        mbar_wait(&bars->res_full[ew][rb], (k >> 1) & 1);
        const uint8_t* rrow_s = res_slice + rb * 16384 + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) rq[j] = lds128(rrow_s + ((static_cast<uint32_t>(j) ^ sw) << 4));
      }

      float st_sum = 0.f, st_sq = 0.f;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int c0 = col0 + hf * 32;
        uint32_t acc[32];
        tmem_ld_32x32(tmem_base + ab * BN + (static_cast<uint32_t>(quarter * 32) << 16) + c * 64 + hf * 32, acc);
        tmem_ld_wait();
        if (hf == 1 && last_in_tile) {  // accumulator drained by this warp: hand it back to the MMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&bars->tmem_empty[ab]);
        }
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
        if constexpr ((kEpi & kEpiRowNorm) != 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 u4 = __ldg(reinterpret_cast<const float4*>(ep.rownorm_u + c0 + j));
            v[j] = rn_rstd * (v[j] - rn_mean * u4.x); v[j + 1] = rn_rstd * (v[j + 1] - rn_mean * u4.y);
            v[j + 2] = rn_rstd * (v[j + 2] - rn_mean * u4.z); v[j + 3] = rn_rstd * (v[j + 3] - rn_mean * u4.w);
          }
        }
        if constexpr ((kEpi & kEpiScale) != 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 s4 = __ldg(reinterpret_cast<const float4*>(ep.col_scale + c0 + j));
            v[j] *= s4.x; v[j + 1] *= s4.y; v[j + 2] *= s4.z; v[j + 3] *= s4.w;
          }
        }
        if constexpr ((kEpi & kEpiBias) != 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.col_bias + c0 + j));
            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
          }
        }
        if constexpr ((kEpi & kEpiAlpha) != 0) {
          if (ep.alpha_cols >= c0 + 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= ep.alpha;
          } else if (ep.alpha_cols > c0) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < ep.alpha_cols) v[j] *= ep.alpha;
          }
        }
        if constexpr ((kEpi & kEpiGelu) != 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float2 y = gelu_erf2(make_float2(v[j], v[j + 1]));
            v[j] = y.x;
            v[j + 1] = y.y;
          }
        }
        if constexpr (kResB16) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 r4 = rq[4 * hf + j];
            const float2 a = unpack_bf16x2(r4.x), b2 = unpack_bf16x2(r4.y), c2 = unpack_bf16x2(r4.z), d = unpack_bf16x2(r4.w);
            v[8 * j] += a.x; v[8 * j + 1] += a.y; v[8 * j + 2] += b2.x; v[8 * j + 3] += b2.y;
            v[8 * j + 4] += c2.x; v[8 * j + 5] += c2.y; v[8 * j + 6] += d.x; v[8 * j + 7] += d.y;
          }
        }
        if constexpr ((kEpi & kEpiRelu) != 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        // ---- box hand-over inside the warpgroup (no DMA thread, no mbarrier round trip): the elected thread makes sure the
        // store that last read this box has finished (all but the latest bulk group), a 128-thread named barrier publishes
        // that, everybody writes, a second barrier after the proxy fence, the elected thread issues the TMA store.
        // bf16: one box per 64-column chunk (both halves); fp32: one box per 32-column half.
        uint8_t* box = stage0 + (bx % Cfg::kWgBoxes) * Cfg::kBox;
        if (kOutF32 || hf == 0) {
          if (elected) bulk_wait_group_read<Cfg::kWgBoxes - 1>();
          asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
        }
        if constexpr (kOutF32) {
          uint8_t* dst = box + r * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(dst + ((static_cast<uint32_t>(j) ^ sw) << 4)) =
                make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
          uint8_t* dst = box + r * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_bf16x2(v[8 * j], v[8 * j + 1]); o.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
            o.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]); o.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
            *reinterpret_cast<uint4*>(dst + ((static_cast<uint32_t>(4 * hf + j) ^ sw) << 4)) = o;
            if constexpr ((kEpi & kEpiRowStats) != 0) {  // statistics of exactly what is stored
              const float2 a = unpack_bf16x2(o.x), b2 = unpack_bf16x2(o.y), c2 = unpack_bf16x2(o.z), d = unpack_bf16x2(o.w);
              st_sum += ((a.x + a.y) + (b2.x + b2.y)) + ((c2.x + c2.y) + (d.x + d.y));
              st_sq += ((a.x * a.x + a.y * a.y) + (b2.x * b2.x + b2.y * b2.y)) +
                       ((c2.x * c2.x + c2.y * c2.y) + (d.x * d.x + d.y * d.y));
            }
          }
        }
        if (kOutF32 || hf == 1) {
          fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA store (async proxy)
          asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
          if (elected) {
            if (tile_valid) {
              const int col = kOutF32 ? c0 : col0;
              if constexpr (kConv) {
                tma_store_4d(&tmC, box, col, st_w0, st_h0, st_img);
              } else if constexpr ((kEpi & kEpiResF32) != 0) {  // x += ... : reduce-add at L2, no residual read
                tma_reduce_add_2d(&tmC, box, col, m0);
              } else {
                tma_store_2d(&tmC, box, col, m0);
              }
            }
            bulk_commit_group();  // (also for a skipped phantom tile: the group count paces the box reuse)
            if (wg == 0) tr(4, static_cast<int>(bx));
          }
          ++bx;
        }
      }
      if constexpr (kResB16) {
        // The slice is refilled only now, after every value read from it has been CONSUMED (the adds above feed the staged
        // stores).  Issuing the refill right after the loads is a write-after-read race between the TMA (async proxy) and
        // loads that are still in flight: it showed up as stale residuals in the last 16-32 columns of a chunk at BN = 192
        // (tests/test_gemm_ts_gpu.py::test_ts_epilogues[33000-192]).  One warpgroup chunk (~0.6 us) remains to land it.
        __syncwarp();
        if (lane == 0) issue_residual(g + 4, k & 1);  // this warpgroup's chunk after next
      }
      if constexpr ((kEpi & kEpiRowStats) != 0) {
        if (row_ok) {  // one deterministic (sum, sumsq) slot per 64-column block of the row
          const int nparts = shp.N / 64;
          reinterpret_cast<float2*>(ep.rowstats_out)[row * nparts + col0 / 64] = make_float2(st_sum, st_sq);
        }
      }
    }
    if (elected) {
      bulk_wait_group_read<0>();  // smem may be released; the writes themselves complete with the grid
      if (wg == 0) tr(0, 3);
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (threadIdx.x == 0) tr(0, 4);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta<Cfg::kTmemCols>(tmem_base);
  }
}

// ----------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------
static unsigned long long* g_gemm_trace = nullptr;

struct TsMaps {
  CUtensorMap a, b, c, r;
};

template <int BN, bool kConv, int kEpi>
static int launch_ts(const TsMaps& tm, const GemmShape& shp, const GemmEpilogue& ep, int m_tiles, int max_clusters,
                     cudaStream_t st) {
  using Cfg = TsCfg<BN, kEpi>;
  auto kern = gemm_ts_kernel<BN, kConv, kEpi>;
  static bool configured = false;
  if (!configured) {
    SGF_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    configured = true;
  }
  const int num_m_pairs = (m_tiles + 1) / 2;
  const int num_tiles = num_m_pairs * (shp.N / BN);
  // equal share per cluster: 87 tiles on 74 SM pairs take two rounds anyway -- run them on 44 clusters and leave the
  // L2 bandwidth of the idle pairs to the busy ones
  const int rounds = (num_tiles + max_clusters - 1) / max_clusters;
  const int clusters = (num_tiles + rounds - 1) / rounds;
  SGF_CHECK_CUDA(launch_pdl(kern, dim3(2 * clusters), dim3(kTsThreads), Cfg::kSmem, st, tm.a, tm.b, tm.c, tm.r, shp, ep,
                            m_tiles, num_m_pairs, num_tiles, g_gemm_trace,
                            getenv("SGF_GEMM_TS_PRODUCERS") ? atoi(getenv("SGF_GEMM_TS_PRODUCERS")) == 2 : 1));
  count_launch();
  return SGF_OK;
}

#define SGF_TS_CASE(MASK) \
  case (MASK):            \
    return launch_ts<BN, kConv, (MASK)>(tm, shp, ep, m_tiles, max_clusters, st);

template <int BN, bool kConv>
static int dispatch_ts_epi(int mask, const TsMaps& tm, const GemmShape& shp, const GemmEpilogue& ep, int m_tiles,
                           int max_clusters, cudaStream_t st) {
  switch (mask) {
    SGF_TS_CASE(kEpiScale | kEpiBias | kEpiRelu)  // stem conv + BN + ReLU
    SGF_TS_CASE(kEpiScale | kEpiBias)             // downsample conv + BN
    default: break;
  }
  if constexpr (!kConv && BN <= 256) {  // (four residual boxes next to 40 KB stages would leave BN = 384 two stages)
    switch (mask) {
      SGF_TS_CASE(kEpiScale | kEpiBias | kEpiRelu | kEpiResBf16)          // bottleneck conv3 + BN + residual + ReLU
      default: break;
    }
  }
  if constexpr (!kConv) {
    switch (mask) {
      SGF_TS_CASE(kEpiBias | kEpiAlpha)                                   // fused QKV / cross q
      SGF_TS_CASE(kEpiBias | kEpiOutF32)                                  // out_proj, image_proj
      SGF_TS_CASE(kEpiBias)                                               // cross k/v, position projections
      SGF_TS_CASE(kEpiBias | kEpiGelu | kEpiRowStats)                     // fc1 (+ ffn_layernorm statistics)
      SGF_TS_CASE(kEpiBias | kEpiGelu)                                    // fc1
      SGF_TS_CASE(kEpiBias | kEpiRowNorm | kEpiResF32 | kEpiOutF32)       // fc2 with folded ffn_layernorm, x += ...
      SGF_TS_CASE(kEpiBias | kEpiResF32 | kEpiOutF32)                     // fc2, x += ...
      SGF_TS_CASE(0)                                                      // dX = dY W (bf16)
      SGF_TS_CASE(kEpiOutF32)                                             // d(encoder_out)
      default: break;
    }
  }
  return -1;
}

// Tile width.  Measured (profiles/r02_gemm_ts_k_sweep.txt): a k-block costs ~580 clocks at every BN <= 256 (operand delivery,
// not the 2*BN tensor-core clocks, paces the main loop) and 2*BN beyond; a 128 x 64 output chunk costs an epilogue
// warpgroup ~1000 (bf16) / ~1300 (fp32) clocks, two warpgroups work in parallel, and only BN <= 256 (two accumulators) hides
// the epilogue of tile i under the main loop of tile i+1.
static int pick_bn_ts(int m_tiles, int N, int num_kb, int max_clusters, bool out_f32) {
  const char* e = getenv("SGF_GEMM_TS_BN");
  const int forced = e ? atoi(e) : 0;
  int best = 0;
  double best_cost = 1e30;
  const int pairs = (m_tiles + 1) / 2;
  for (int bn : {384, 256, 192, 128, 64}) {
    if (N % bn != 0) continue;
    if (forced == bn) return bn;
    const int tiles = pairs * (N / bn);
    const int rounds = (tiles + max_clusters - 1) / max_clusters;
    const double kblock = 2.0 * bn > 580.0 ? 2.0 * bn : 580.0;
    const double drain = ((bn / 64 + 1) / 2) * (out_f32 ? 1300.0 : 1000.0);
    const bool overlapped = 2 * bn <= 512;
    const double cost = rounds * (num_kb * kblock + (overlapped ? 0.0 : drain)) + (overlapped ? drain : 0.0);
    if (cost < best_cost) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

int gemm_ts_dispatch(const GemmShape& shp, const GemmEpilogue& ep, const void* a, int64_t lda, const void* b, int64_t ldb,
                     const ConvInput* cv, cudaStream_t st) {
  const bool conv = cv != nullptr;
  if (shp.N % 64 != 0 || shp.K < 1) return -1;
  const int mask = epilogue_mask(ep);
  if ((mask & kEpiResF32) && !(ep.residual == ep.c && ep.ldr == ep.ldc && (mask & kEpiOutF32))) return -1;  // in place only
  if ((mask & kEpiRowStats) && ep.c_dtype != SGF_BF16) return -1;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    SGF_CHECK_CUDA(cudaGetDevice(&dev));
    SGF_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int max_clusters = num_sms / 2;
  const int m_tiles = conv ? cv->n * shp.tiles_w * shp.tiles_h : (shp.M + BM - 1) / BM;
  const int num_kb = (shp.K + BK - 1) / BK;
  const bool out_f32 = ep.c_dtype == SGF_F32;
  const int bn = pick_bn_ts(m_tiles, shp.N, num_kb, max_clusters, out_f32);
  if (!bn) return -1;

  TsMaps tm;
  memset(&tm, 0, sizeof(tm));
  if (conv) {
    // input: (c, w, h, n) view with the caller's strides; a strided convolution samples every stride-th pixel through the
    // TMA element strides (box extent = samples x stride), out-of-bounds coordinates are zero fill = the padding
    uint64_t dims[4] = {static_cast<uint64_t>(cv->cin), static_cast<uint64_t>(cv->w), static_cast<uint64_t>(cv->h),
                        static_cast<uint64_t>(cv->n)};
    uint64_t strides[3] = {static_cast<uint64_t>(cv->w_stride) * 2, static_cast<uint64_t>(cv->h_stride) * 2,
                           static_cast<uint64_t>(cv->n_stride) * 2};
    uint32_t box[4] = {BK, static_cast<uint32_t>(shp.bw * shp.stride_w), static_cast<uint32_t>(shp.bh * shp.stride_h), 1};
    uint32_t est[4] = {1, static_cast<uint32_t>(shp.stride_w), static_cast<uint32_t>(shp.stride_h), 1};
    if (box[1] > 256 || box[2] > 256) return -1;
    if (int rc = encode_tmap(&tm.a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, cv->x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, est))
      return rc;
    uint64_t cdims[4] = {static_cast<uint64_t>(shp.N), static_cast<uint64_t>(shp.W), static_cast<uint64_t>(shp.H),
                         static_cast<uint64_t>(cv->n)};
    uint64_t cstr[3] = {static_cast<uint64_t>(shp.N) * 2, static_cast<uint64_t>(shp.N) * shp.W * 2,
                        static_cast<uint64_t>(shp.N) * shp.W * shp.H * 2};
    uint32_t cbox[4] = {64, static_cast<uint32_t>(shp.bw), static_cast<uint32_t>(shp.bh), 1};
    if (int rc = encode_tmap(&tm.c, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, ep.c, cdims, cstr, cbox, CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  } else {
    uint64_t dims[2] = {static_cast<uint64_t>(shp.K), static_cast<uint64_t>(shp.M)};
    uint64_t strides[1] = {static_cast<uint64_t>(lda) * 2};
    uint32_t box[2] = {BK, BM};
    if (int rc = encode_tmap(&tm.a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
    uint64_t cdims[2] = {static_cast<uint64_t>(shp.N), static_cast<uint64_t>(shp.M)};
    uint64_t cstr[1] = {static_cast<uint64_t>(ep.ldc) * (out_f32 ? 4 : 2)};
    uint32_t cbox[2] = {out_f32 ? 32u : 64u, BM};
    if (int rc = encode_tmap(&tm.c, out_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ep.c,
                             cdims, cstr, cbox, CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
    if (mask & kEpiResBf16) {
      uint64_t rstr[1] = {static_cast<uint64_t>(ep.ldr) * 2};
      uint32_t rbox[2] = {64, 32};
      if (int rc = encode_tmap(&tm.r, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ep.residual, cdims, rstr, rbox,
                               CU_TENSOR_MAP_SWIZZLE_128B))
        return rc;
    }
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(shp.K), static_cast<uint64_t>(shp.N)};
    uint64_t strides[1] = {static_cast<uint64_t>(ldb) * 2};
    uint32_t box[2] = {BK, static_cast<uint32_t>(bn > 256 ? bn / 4 : bn / 2)};  // this CTA's half of one N-split
    if (int rc = encode_tmap(&tm.b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, b, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
#define SGF_TS_BN(BNV)                                                                                         \
  case BNV:                                                                                                    \
    return conv ? dispatch_ts_epi<BNV, true>(mask, tm, shp, ep, m_tiles, max_clusters, st)                     \
                : dispatch_ts_epi<BNV, false>(mask, tm, shp, ep, m_tiles, max_clusters, st);
  switch (bn) {
    SGF_TS_BN(64)
    SGF_TS_BN(128)
    SGF_TS_BN(192)
    SGF_TS_BN(256)
    SGF_TS_BN(384)
    default: return -1;
  }
#undef SGF_TS_BN
}

}  // namespace sgf

extern "C" void sgf_debug_set_gemm_trace(void* buf) { sgf::g_gemm_trace = reinterpret_cast<unsigned long long*>(buf); }
