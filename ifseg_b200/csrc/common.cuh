// Shared device-side PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld / fences).  Hand-written inline PTX; no CUTLASS.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/segofa_b200.h"

#define SGF_DEVICE __device__ __forceinline__

namespace sgf {

// ----------------------------------------------------------------------------------------
// host-side error plumbing: every extern "C" entry returns an int status (0 == ok)
// ----------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
void count_launch();
#define SGF_CHECK_CUDA(expr)                                                                  \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      sgf::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return _e == cudaErrorMemoryAllocation ? SGF_ERR_OOM : SGF_ERR_CUDA;                   \
    }                                                                                         \
  } while (0)
#define SGF_REQUIRE(cond, ...)               \
  do {                                       \
    if (!(cond)) {                           \
      sgf::set_last_error(__VA_ARGS__);      \
      return SGF_ERR_INVALID;                \
    }                                        \
  } while (0)

// tensor-map encode through the driver entry point (no link-time libcuda dependency)
int encode_tmap(CUtensorMap* out, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                const uint64_t* strides_bytes /*rank-1*/, const uint32_t* box, CUtensorMapSwizzle swz,
                const uint32_t* elem_strides = nullptr);

// ----------------------------------------------------------------------------------------
// Programmatic dependent launch: every kernel of the library is launched with the
// programmatic-stream-serialization attribute, signals `launch_dependents` as its first
// instruction and executes `griddepcontrol.wait` before its first global-memory access.  The
// next kernel's launch latency and prologue (barrier init, TMEM allocation, descriptor
// prefetch) then overlap the tail of the previous kernel; after the wait the previous grid has
// completed and its writes are visible, so data hazards are exactly those of a plain stream.
// ----------------------------------------------------------------------------------------
SGF_DEVICE void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
SGF_DEVICE void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  static const bool no_pdl = getenv("SGF_NO_PDL") != nullptr;  // A/B switch: plain stream order (every kernel drains first)
  cfg.numAttrs = no_pdl ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ----------------------------------------------------------------------------------------
// small helpers
// ----------------------------------------------------------------------------------------
SGF_DEVICE uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// 16-byte shared-memory load in the shared state space (a plain C++ load through a pointer derived from the generic
// dynamic-smem base compiles to a generic LD.E)
SGF_DEVICE uint4 lds128(const void* p) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)) : "memory");
  return v;
}

SGF_DEVICE uint32_t elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\t"
      "elect.sync %%rx|%%px, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, %%px;\n\t}\n"
      : "=r"(pred));
  return pred;
}

// ----------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------
SGF_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
SGF_DEVICE void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
SGF_DEVICE void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

SGF_DEVICE void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
SGF_DEVICE void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
SGF_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug turns into a trap (launch failure reported to the host) instead
// of a hung GPU.  try_wait itself suspends for a HW-defined interval, so the bound is seconds.
#ifndef SGF_MBAR_SPIN_LIMIT
#define SGF_MBAR_SPIN_LIMIT (1u << 22)
#endif
SGF_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > SGF_MBAR_SPIN_LIMIT) __trap();
  }
}

// ----------------------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------------------
SGF_DEVICE void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 1-D bulk copy global -> shared (no tensor map): size and both addresses multiples of 16 bytes
SGF_DEVICE void bulk_load_1d(void* smem, const void* gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem)),
               "l"(reinterpret_cast<uint64_t>(gmem)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
SGF_DEVICE void tma_load_2d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
SGF_DEVICE void tma_load_3d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
SGF_DEVICE void tma_load_4d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ----------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------------
template <uint32_t kCols>
SGF_DEVICE void tmem_alloc(uint32_t* dst_smem) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
SGF_DEVICE void tmem_dealloc(uint32_t taddr) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
SGF_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
SGF_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]   (kind::f16 : bf16/fp16 inputs, fp32 accumulate)
SGF_DEVICE void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (M = 128 rows on the 128 lanes, two 16-bit K elements per 32-bit
// column, 8 columns per K = 16 step) is read from tensor memory
SGF_DEVICE void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
SGF_DEVICE void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- CTA-pair (cta_group::2) variants: two CTAs of a cluster form one 256-row MMA ----
SGF_DEVICE uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
SGF_DEVICE void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <uint32_t kCols>
SGF_DEVICE void tmem_alloc_2cta(uint32_t* dst_smem) {  // one full warp in EACH CTA of the pair, same smem offset
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
SGF_DEVICE void tmem_dealloc_2cta(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// issued by ONE thread of the leader CTA (rank 0); A/B descriptors name the leader's smem, the peer CTA
// supplies its half from the same offsets; each CTA's TMEM receives its 128 rows of D
SGF_DEVICE void umma_f16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this smem offset in every CTA of `cta_mask` once the MMAs issued so far retire
SGF_DEVICE void umma_commit_2cta(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
// arrive on the LEADER CTA's copy of a barrier (same smem offset) from either CTA of the pair
SGF_DEVICE void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & 0xFEFFFFFFu) : "memory");
}
// TMA load executed by either CTA of the pair; the transaction bytes are credited to the LEADER's barrier
// (peer bit 24 of the shared::cluster address cleared)
SGF_DEVICE void tma_load_3d_2cta(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  const uint32_t bar_leader = smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5}], [%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_leader), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

SGF_DEVICE void tma_load_2d_2cta(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  const uint32_t bar_leader = smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_leader), "r"(c0), "r"(c1)
      : "memory");
}
SGF_DEVICE void tma_load_4d_2cta(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  const uint32_t bar_leader = smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5, %6}], [%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_leader), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ---- TMA stores (shared -> global through a tensor map; out-of-bounds parts of the box are dropped) ----
// Bulk async-groups are per THREAD: the thread that issues the stores commits and waits on them.
SGF_DEVICE void tma_store_2d(const CUtensorMap* m, const void* smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1)
               : "memory");
}
SGF_DEVICE void tma_store_4d(const CUtensorMap* m, const void* smem, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// global[box] += shared[box] (element type of the tensor map; fp32 here), performed at L2
SGF_DEVICE void tma_reduce_add_2d(const CUtensorMap* m, const void* smem, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1)
               : "memory");
}
SGF_DEVICE void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int kPending>
SGF_DEVICE void bulk_wait_group_read() {  // all but the kPending most recent groups have finished READING shared memory
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
}
template <int kPending>
SGF_DEVICE void bulk_wait_group() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(kPending) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets row (lane base + t)
SGF_DEVICE void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
SGF_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
SGF_DEVICE void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns
SGF_DEVICE void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
SGF_DEVICE void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
SGF_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------------------
// UMMA descriptors (bit layouts: PTX ISA "tcgen05 matrix/instruction descriptor")
// ----------------------------------------------------------------------------------------
// Shared-memory operand, 128-byte swizzle, rows of 128 B (64 bf16) packed densely; groups of
// 8 rows are 1024 B apart (SBO).  Valid for K-major (row = M/N index, 128 B = 64 K elements)
// and for MN-major (row = K index, 128 B = 64 M/N elements) tiles; LBO is unused when the
// swizzled extent covers the tile (set to 1 as CUTLASS does).
SGF_DEVICE uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);  // start address  [0,14)
  d |= static_cast<uint64_t>(1) << 16;                      // LBO            [16,30)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;              // SBO            [32,46)
  d |= static_cast<uint64_t>(1) << 46;                      // version = 1 (sm_100)
  d |= static_cast<uint64_t>(2) << 61;                      // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32.  a_mn / b_mn: 1 = MN-major operand.
// fp16 x fp16 -> fp32 (format fields 0)
__host__ __device__ constexpr uint32_t make_idesc_f16(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
  return (1u << 4) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ----------------------------------------------------------------------------------------
// numerics helpers
// ----------------------------------------------------------------------------------------
SGF_DEVICE float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
SGF_DEVICE float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Packed fp32x2 arithmetic (sm_100: FFMA2/FMUL2/FADD2 — two lanes per FMA-pipe issue slot).  The row kernels are
// issue-bound on their elementwise chains, so everything that is not a MUFU or a sign trick goes through these.
SGF_DEVICE float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(reinterpret_cast<uint64_t&>(d))
      : "l"(reinterpret_cast<const uint64_t&>(a)), "l"(reinterpret_cast<const uint64_t&>(b)),
        "l"(reinterpret_cast<const uint64_t&>(c)));
  return d;
}
SGF_DEVICE float2 mul2(float2 a, float2 b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<uint64_t&>(d))
      : "l"(reinterpret_cast<const uint64_t&>(a)), "l"(reinterpret_cast<const uint64_t&>(b)));
  return d;
}
SGF_DEVICE float2 add2(float2 a, float2 b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<uint64_t&>(d))
      : "l"(reinterpret_cast<const uint64_t&>(a)), "l"(reinterpret_cast<const uint64_t&>(b)));
  return d;
}
SGF_DEVICE float2 splat2(float a) { return make_float2(a, a); }

// erf-GELU with the Abramowitz-Stegun 7.1.26 rational approximation of erf (|error| <= 1.5e-7,
// far below the bf16 rounding of the result): 2 MUFU (rcp, ex2) + ~12 FMA-pipe ops, no branches.
SGF_DEVICE float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = fast_rcp(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = fast_exp2(-1.4426950408889634f * z * z);
  const float erf_abs = fmaf(-p * t, e, 1.0f);  // erf(|x|/sqrt2)
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}
// Two elements at a time.  Returns Phi(x) (the normal CDF: gelu(x) = x * Phi(x)) and e = exp(-x^2/2)
// (gelu'(x) = Phi(x) + x * e / sqrt(2 pi)).  Per pair: 10 packed FMA-pipe ops, 4 MUFU, 4 sign/abs ALU ops.
SGF_DEVICE float2 gelu_cdf2(float2 x, float2& e) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 den = fma2(splat2(0.3275911f * 0.70710678118654752440f), ax, splat2(1.0f));
  const float2 t = make_float2(fast_rcp(den.x), fast_rcp(den.y));
  float2 p = fma2(splat2(-1.061405429f), t, splat2(1.453152027f));  // -(A&S polynomial): erf = 1 + (p t) e
  p = fma2(p, t, splat2(-1.421413741f));
  p = fma2(p, t, splat2(0.284496736f));
  p = fma2(p, t, splat2(-0.254829592f));
  const float2 arg = mul2(mul2(x, x), splat2(-0.72134752044448170368f));  // -x^2/2 * log2(e)
  e = make_float2(fast_exp2(arg.x), fast_exp2(arg.y));
  const float2 erf_abs = fma2(mul2(p, t), e, splat2(1.0f));
  const float2 s = make_float2(copysignf(erf_abs.x, x.x), copysignf(erf_abs.y, x.y));
  return fma2(splat2(0.5f), s, splat2(0.5f));
}
SGF_DEVICE float2 gelu_erf2(float2 x) {
  float2 e;
  return mul2(x, gelu_cdf2(x, e));
}
// ----------------------------------------------------------------------------------------
// counter-based dropout / DropPath masks: a pure function of (seed, step, site, row, column), so the forward and
// the adjoint kernels regenerate identical masks and a CUDA-graph replay draws new ones (step lives on the device)
// ----------------------------------------------------------------------------------------
SGF_DEVICE uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
struct DropCtx {
  uint64_t key;
  uint32_t thresh16;   // element dropped when its 16 random bits < thresh16
  uint32_t thresh24;   // sample's path dropped when its 24 random bits < thresh24
  float inv_keep, inv_keep_path;
  int rows_per_sample;
  bool on;
};
SGF_DEVICE DropCtx make_drop_ctx(float drop_p, float droppath_p, uint32_t seed, uint32_t site, const int32_t* step_dev,
                                 int rows_per_sample) {
  DropCtx c;
  c.on = drop_p > 0.f || droppath_p > 0.f;
  const uint64_t step = step_dev ? static_cast<uint64_t>(step_dev[0]) : 0ull;
  c.key = mix64((static_cast<uint64_t>(seed) << 32) ^ (step * 0x632BE59BD9B4E019ull) ^ (static_cast<uint64_t>(site) << 8));
  c.thresh16 = static_cast<uint32_t>(drop_p * 65536.0f + 0.5f);
  c.thresh24 = static_cast<uint32_t>(droppath_p * 16777216.0f + 0.5f);
  c.inv_keep = drop_p > 0.f ? 1.0f / (1.0f - static_cast<float>(c.thresh16) * (1.0f / 65536.0f)) : 1.0f;
  c.inv_keep_path = droppath_p > 0.f ? 1.0f / (1.0f - static_cast<float>(c.thresh24) * (1.0f / 16777216.0f)) : 1.0f;
  c.rows_per_sample = rows_per_sample > 0 ? rows_per_sample : 1;
  return c;
}
// multipliers of the 8 elements of chunk `chunk` of (output) row `row`: element keep / (1-p) times path keep / (1-dp)
SGF_DEVICE void drop_mult8(const DropCtx& c, int64_t row, int chunk, float (&m)[8]) {
  float path = 1.0f;
  if (c.thresh24) {
    const uint64_t b = static_cast<uint64_t>(row / c.rows_per_sample);
    path = (static_cast<uint32_t>(mix64(c.key ^ 0xD1B54A32D192ED03ull ^ (b << 1)) >> 40) >= c.thresh24) ? c.inv_keep_path : 0.f;
  }
  if (c.thresh16) {
    const uint64_t base = c.key ^ (static_cast<uint64_t>(row) << 22) ^ (static_cast<uint64_t>(chunk) << 1);
    // two independent 64-bit streams per chunk: the second one is keyed by a distinct odd constant (base | 1 would
    // coincide with base whenever the key's bit 0 is set, tying elements j and j+4)
    const uint64_t a = mix64(base), b2 = mix64(base ^ 0xA0761D6478BD642Full);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      m[j] = (static_cast<uint32_t>(a >> (16 * j)) & 0xFFFFu) >= c.thresh16 ? c.inv_keep * path : 0.f;
      m[4 + j] = (static_cast<uint32_t>(b2 >> (16 * j)) & 0xFFFFu) >= c.thresh16 ? c.inv_keep * path : 0.f;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = path;
  }
}

// vector fp32 reduction into global memory (no return value): one 16-byte L2 atomic
SGF_DEVICE void red_add_f32x4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
SGF_DEVICE void red_add_f32(float* addr, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}
SGF_DEVICE uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
SGF_DEVICE float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
SGF_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
SGF_DEVICE float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace sgf
