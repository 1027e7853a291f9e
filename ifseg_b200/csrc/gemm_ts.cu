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
// One cluster of two CTAs (one SM pair) loops over 256 x BN output tiles (BN = 64/128/192/256):
//   warp 0      : TMA producer (A: 128 rows of this CTA; B: this CTA's HALF of the BN weight rows) -> kStages ring
//   warp 1      : TMEM allocator; in the leader CTA the single tcgen05.mma.cta_group::2 issuer (M = 256)
//   warp 2      : store DMA: waits for a finished 128 x 64 output chunk in smem, issues the TMA store, recycles buffers
//   warps 3..10 : two epilogue warpgroups; chunk g (64 columns of a tile) belongs to warpgroup g & 1 and its own
//                 staging buffer, so one warpgroup computes while the other one's chunk is being stored
// TMEM holds two accumulators (2 x BN columns): the epilogue of tile i overlaps the main loop of tile i+1.
//   full[s]/empty[s]      smem ring across tiles (full lives in the leader: both CTAs' bytes are credited there)
//   tmem_full[2]          multicast tcgen05.commit per finished accumulator
//   tmem_empty[2] (leader) one arrival per epilogue warp of both CTAs once its last tcgen05.ld of the tile retired
//   stage_ready[2]        4 warp arrivals: chunk written + fence.proxy.async  -> store DMA
//   stage_free[2]         store DMA: the TMA store has finished reading the buffer (cp.async.bulk.wait_group.read)
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
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = (BN / 2) * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kOutBuf = kOutF32 ? 32768 : 16384;  // one 128 x 64 chunk
  static constexpr int kResBytes = kResB16 ? 4 * 16384 : 0;  // two 128 x 64 bf16 chunks per warpgroup
  static constexpr int kEpiBytes = 2 * kOutBuf + kResBytes;
  static constexpr int kBarBytes = 512;
  static constexpr int kStagesFit = (kTsMaxSmem - 1024 - kEpiBytes - kBarBytes) / kStageBytes;
  static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
  static constexpr int kSmem = 1024 + kStages * kStageBytes + kEpiBytes + kBarBytes;
  static constexpr int kTmemCols = 2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512);
  static_assert(kStages >= 3, "pipeline too shallow");
};

struct TsBars {
  uint64_t full[8], empty[8], tmem_full[2], tmem_empty[2], stage_ready[2], stage_free[2], res_full[kTsEpiWarps][2];
  uint32_t tmem_slot;
};
static_assert(sizeof(TsBars) <= 512, "barrier block");

template <int BN, bool kConv, int kEpi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTsThreads, 1)
    gemm_ts_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const GemmShape shp,
                   const GemmEpilogue ep, const int m_tiles, const int num_m_pairs, const int num_tiles) {
  using Cfg = TsCfg<BN, kEpi>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kChunks = Cfg::kChunks;
  constexpr bool kOutF32 = Cfg::kOutF32;
  constexpr bool kResB16 = Cfg::kResB16;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_out = smem + kStages * Cfg::kStageBytes;      // [2][kOutBuf]
  uint8_t* smem_res = smem_out + 2 * Cfg::kOutBuf;            // [2 warpgroups][2][16 KB]
  TsBars* bars = reinterpret_cast<TsBars*>(smem_out + Cfg::kEpiBytes);

  pdl_trigger();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int num_kb = (shp.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    if constexpr (kResB16) tma_prefetch_desc(&tmR);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->tmem_full[i], 1);
      // BN = 64: one chunk per tile, tile i is drained by warpgroup i & 1 alone (4 warps in each CTA)
      mbar_init(&bars->tmem_empty[i], kChunks == 1 ? kTsEpiWarps : 2 * kTsEpiWarps);
      mbar_init(&bars->stage_ready[i], 4);
      mbar_init(&bars->stage_free[i], 1);
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
  pdl_wait();

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

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      uint32_t g = 0;  // running k-block counter across tiles
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        int mt, n0, m0, img, h0, w0;
        decode(t, mt, n0, m0, img, h0, w0);
#pragma unroll 1
        for (int kb = 0; kb < num_kb; ++kb, ++g) {
          const int s = g % kStages;
          mbar_wait(&bars->empty[s], ((g / kStages) & 1) ^ 1);
          uint8_t* sa = smem + s * Cfg::kStageBytes;
          if (rank == 0) mbar_expect_tx(&bars->full[s], 2 * Cfg::kStageBytes);
          if constexpr (kConv) {
            const int tap = kb / shp.cin_blocks;
            const int kc = kb - tap * shp.cin_blocks;
            tma_load_4d_2cta(sa, &tmA, &bars->full[s], kc * BK, w0 + tap % 3 - 1, h0 + tap / 3 - 1, img);
          } else {
            tma_load_2d_2cta(sa, &tmA, &bars->full[s], kb * BK, m0);
          }
          tma_load_2d_2cta(sa + Cfg::kABytes, &tmB, &bars->full[s], kb * BK, n0 + static_cast<int>(rank) * (BN / 2));
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA) ================================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN, 0, 0);
      uint32_t g = 0;
      int it = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
        const int ab = it & 1;
        mbar_wait(&bars->tmem_empty[ab], ((it >> 1) & 1) ^ 1);  // both CTAs drained this accumulator
        tc_fence_after();
#pragma unroll 1
        for (int kb = 0; kb < num_kb; ++kb, ++g) {
          const int s = g % kStages;
          mbar_wait(&bars->full[s], (g / kStages) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * Cfg::kStageBytes);
          const uint64_t da = make_smem_desc_sw128(sa);
          const uint64_t db = make_smem_desc_sw128(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_f16_2cta(tmem_base + ab * BN, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2cta(&bars->empty[s], 0b11);
        }
        umma_commit_2cta(&bars->tmem_full[ab], 0b11);
      }
    }
  } else if (warp == 2) {
    // ================================ store DMA ================================
    if (lane == 0) {
      uint32_t g = 0;  // running chunk counter across tiles; chunk g lives in staging buffer g & 1
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        int mt, n0, m0, img, h0, w0;
        decode(t, mt, n0, m0, img, h0, w0);
        const bool valid = mt < m_tiles;
#pragma unroll 1
        for (int c = 0; c < kChunks; ++c, ++g) {
          const int buf = g & 1;
          mbar_wait(&bars->stage_ready[buf], (g >> 1) & 1);
          const uint8_t* src = smem_out + buf * Cfg::kOutBuf;
          const int col = n0 + c * 64;
          if (valid) {
            if constexpr (kConv) {
              tma_store_4d(&tmC, src, col, w0, h0, img);
            } else if constexpr ((kEpi & kEpiResF32) != 0) {  // x += ... : reduce-add at L2, no residual read
              tma_reduce_add_2d(&tmC, src, col, m0);
              tma_reduce_add_2d(&tmC, src + 16384, col + 32, m0);
            } else if constexpr (kOutF32) {
              tma_store_2d(&tmC, src, col, m0);
              tma_store_2d(&tmC, src + 16384, col + 32, m0);
            } else {
              tma_store_2d(&tmC, src, col, m0);
            }
          }
          bulk_commit_group();
          bulk_wait_group_read<1>();  // the store of chunk g-1 has finished reading its buffer
          if (g >= 1) mbar_arrive(&bars->stage_free[buf ^ 1]);
        }
      }
      bulk_wait_group_read<0>();  // smem may be released; the writes themselves complete with the grid
    }
  } else {
    // ================================ epilogue warpgroups ================================
    const int ew = warp - 3;
    const int wg = ew >> 2;
    const int quarter = warp & 3;  // the TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;
    uint8_t* stage = smem_out + wg * Cfg::kOutBuf;
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
      const int ab = it & 1;
      if (it != cur_it) {
        cur_it = it;
        int mt, img, h0, w0;
        decode(t, mt, n0, m0, img, h0, w0);
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
        mbar_wait(&bars->tmem_full[ab], (it >> 1) & 1);
        tc_fence_after();
      }
      const bool last_in_tile = c + 2 >= kChunks;  // this warpgroup's last chunk of the tile
      const int col0 = n0 + c * 64;

      uint4 rq[8];
      if constexpr (kResB16) {
        const int rb = k & 1;
        mbar_wait(&bars->res_full[ew][rb], (k >> 1) & 1);
        const uint8_t* rrow_s = res_slice + rb * 16384 + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) rq[j] = *reinterpret_cast<const uint4*>(rrow_s + ((static_cast<uint32_t>(j) ^ sw) << 4));
        __syncwarp();
        if (lane == 0) issue_residual(g + 4, rb);  // this warpgroup's chunk after next, into the slice just read
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
        // the staging buffer of this warpgroup is free once the store of its previous chunk has read it
        if (hf == 0) mbar_wait(&bars->stage_free[wg], (k & 1) ^ 1);
        if constexpr (kOutF32) {
          uint8_t* dst = stage + hf * 16384 + r * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(dst + ((static_cast<uint32_t>(j) ^ sw) << 4)) =
                make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
          uint8_t* dst = stage + r * 128;
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
      }
      fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA store (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->stage_ready[wg]);
      if constexpr ((kEpi & kEpiRowStats) != 0) {
        if (row_ok) {  // one deterministic (sum, sumsq) slot per 64-column block of the row
          const int nparts = shp.N / 64;
          reinterpret_cast<float2*>(ep.rowstats_out)[row * nparts + col0 / 64] = make_float2(st_sum, st_sq);
        }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta<Cfg::kTmemCols>(tmem_base);
  }
}

// ----------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------
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
                            m_tiles, num_m_pairs, num_tiles));
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
  if constexpr (!kConv) {
    switch (mask) {
      SGF_TS_CASE(kEpiScale | kEpiBias | kEpiRelu | kEpiResBf16)          // bottleneck conv3 + BN + residual + ReLU
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

// Tile width: per k-block a CTA pulls 16 KB of A and 64 * BN bytes of B through an L2 that delivers ~43 B/clk/SM when
// every SM pulls (B300_MICROARCH: LTS cap ~6300 B/clk), against 2 * BN tensor-core clocks -- the main loop is L2-bound
// at every BN, less so the wider the tile.  cost = rounds * (k-blocks * (385 + 1.5 BN) + drain of the last tile).
static int pick_bn_ts(int m_tiles, int N, int num_kb, int max_clusters, bool out_f32) {
  const char* e = getenv("SGF_GEMM_TS_BN");
  const int forced = e ? atoi(e) : 0;
  int best = 0;
  double best_cost = 1e30;
  const int pairs = (m_tiles + 1) / 2;
  for (int bn : {256, 192, 128, 64}) {
    if (N % bn != 0) continue;
    if (forced == bn) return bn;
    const int tiles = pairs * (N / bn);
    const int rounds = (tiles + max_clusters - 1) / max_clusters;
    const double kblock = 385.0 + 1.5 * bn;
    const double drain = (bn / 64) * (out_f32 ? 900.0 : 600.0);
    const double cost = rounds * (num_kb * kblock + 0.25 * drain) + drain + 1500.0;
    if (cost < best_cost) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

int gemm_ts_dispatch(const GemmShape& shp, const GemmEpilogue& ep, const void* a, int64_t lda, const void* b, int64_t ldb,
                     bool conv, int conv_n, int conv_cin, cudaStream_t st) {
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
  const int m_tiles = conv ? conv_n * shp.tiles_w * shp.tiles_h : (shp.M + BM - 1) / BM;
  const int num_kb = (shp.K + BK - 1) / BK;
  const bool out_f32 = ep.c_dtype == SGF_F32;
  const int bn = pick_bn_ts(m_tiles, shp.N, num_kb, max_clusters, out_f32);
  if (!bn) return -1;

  TsMaps tm;
  memset(&tm, 0, sizeof(tm));
  if (conv) {
    uint64_t dims[4] = {static_cast<uint64_t>(conv_cin), static_cast<uint64_t>(shp.W), static_cast<uint64_t>(shp.H),
                        static_cast<uint64_t>(conv_n)};
    uint64_t strides[3] = {static_cast<uint64_t>(conv_cin) * 2, static_cast<uint64_t>(conv_cin) * shp.W * 2,
                           static_cast<uint64_t>(conv_cin) * shp.W * shp.H * 2};
    uint32_t box[4] = {BK, static_cast<uint32_t>(shp.bw), static_cast<uint32_t>(shp.bh), 1};
    if (int rc = encode_tmap(&tm.a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
    uint64_t cdims[4] = {static_cast<uint64_t>(shp.N), static_cast<uint64_t>(shp.W), static_cast<uint64_t>(shp.H),
                         static_cast<uint64_t>(conv_n)};
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
    uint32_t box[2] = {BK, static_cast<uint32_t>(bn / 2)};
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
    default: return -1;
  }
#undef SGF_TS_BN
}

}  // namespace sgf
