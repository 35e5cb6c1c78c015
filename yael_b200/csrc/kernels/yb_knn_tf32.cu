// yb_knn_tf32.cu -- the distance GEMM of knn_full / nn_single_full (yael/nn.c:54-67,92-129,
// 383-525) as a tcgen05 tensor-core kernel with the per-query top-k fused into its epilogue.  The
// nq x nb distance matrix never reaches HBM.  (The file keeps its first name: the kernel started
// with TF32 operands only.)
//
// What it computes: for every query q and database row b the SCORE s = |b|^2 - 2 <q,b> with
// reduced-precision operands and FP32 accumulation (|q|^2 is constant per query and irrelevant
// for the ranking), and per (query, database range) the k' smallest scores with their row ids.
// The caller (yb_knn.cu) re-ranks those in exact FP32 and certifies the result.  Operand kinds
// (template parameter KIND, one instantiation per kind and epilogue mode):
//   OP_TF32  FP32 rows read as TF32 (kind::tf32, K = 8 per MMA): fallback, streamed host path
//   OP_F16   FP16 rows (kind::f16, K = 16): the k = 1 margin mode of k-means
//   OP_F16N  FP16 rows that carry |b|^2 in 16 extra K elements (folded norms): the epilogue is a
//            MAX tree over raw accumulators -- top-k', sampling and dump modes of k-NN
//   OP_F8    E4M3 +-1 rows (kind::f8f6f4, K = 32): Hamming distances, exact (yb_hamming_tc.cu)
//
// Shape of the kernel (one persistent CTA per SM, 320 threads, warp-specialised):
//   warp 8   TMA producer: query tile A (128 rows x d, resident in smem for a whole work item),
//            database chunks B (256 rows x 128 bytes, 4-stage ring), |b|^2 tiles (1-D bulk copy)
//   warp 9   MMA issuer: tcgen05.mma.cta_group::1, M=128 (queries) x N=256 (rows) x 32 bytes of K
//            per instruction, SWIZZLE_128B K-major operands straight from the TMA layout,
//            accumulators in TMEM, double buffered (2 x 256 columns) so the epilogue of tile t
//            overlaps the MMAs of tile t+1
//   warps 0-7 epilogue: two warps per scheduler.  Thread (w, lane) owns query 32*(w%4)+lane of
//            the tile (its TMEM lane) for the column half w/4 of every tile.  tcgen05.ld brings
//            16 columns at a time; a group whose best score beats thr (the query's admission
//            threshold, a register) is scanned and its candidates appended to the query's list in
//            L2-resident scratch.  When a list fills up the warp compacts it cooperatively (keys
//            in registers, 4-pass radix select) and tightens thr.  This replaces the reference's
//            binheap (yael/binheap.c:139-156) with the same strict '<' admission rule.
// Work items are (query tile, database range) pairs walked range-major so that the CTAs that
// run concurrently stream the same database range and share it through L2.
#include <cuda.h>
#include <stdlib.h>

#include <cuda_fp16.h>

#include "yb_common.cuh"
#include "yb_internal.cuh"

namespace yb {

// ------------------------------------------------------------------ constants
constexpr int TM = 128;          // queries per tile (MMA M)
constexpr int TN = 256;          // database rows per tile (MMA N)
constexpr int KC = 32;           // floats per K chunk = one 128-byte swizzle span
constexpr int MAX_NKC = 4;       // d <= 128
constexpr int STAGES = 4;        // B ring depth
constexpr int NBN = 4;           // |b|^2 ring depth
constexpr int A_CHUNK_BYTES = TM * KC * 4;  // 16 KB
constexpr int B_CHUNK_BYTES = TN * KC * 4;  // 32 KB
// Epilogue warps come in TEAMS of 8 (4 TMEM lane quarters x 2 column halves).  With two teams,
// team b drains accumulator buffer b (every other tile): 4 instead of 2 epilogue warps per
// scheduler to hide the dependent-issue latency of the score / min chains.
#ifndef YB_EPI_TEAMS
#define YB_EPI_TEAMS 1
#endif
constexpr int EPI_TEAMS = YB_EPI_TEAMS;
constexpr int TEAM_WARPS = 8;
constexpr int EPI_WARPS = TEAM_WARPS * EPI_TEAMS;
constexpr int EPI_THREADS = EPI_WARPS * 32;
// 8 epilogue warps (two warpgroups) + one warpgroup holding the TMA producer, the MMA issuer and two
// idle warps: whole warpgroups, so that setmaxnreg can move registers from the third group (56 per
// thread are plenty for the two single-lane loops) to the epilogue warps (224 per thread instead of
// the 168 a 384-thread block gets statically: room for wider TMEM loads without spills)
constexpr int TF32_THREADS = EPI_THREADS + 128;
constexpr int REGS_EPI = 224, REGS_AUX = 56;
__device__ __forceinline__ void regs_epilogue() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_EPI));
}
__device__ __forceinline__ void regs_aux() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_AUX));
}
constexpr int HALF_N = TN / 2;   // columns per epilogue warp and tile

struct Smem {  // offsets inside the 1024-byte aligned dynamic shared memory block
  static constexpr int a_off = 0;
  static constexpr int b_off = MAX_NKC * A_CHUNK_BYTES;
  static constexpr int bn_off = b_off + STAGES * B_CHUNK_BYTES;
  static constexpr int hist_off = bn_off + NBN * TN * 4;
  static constexpr int bar_off = hist_off + EPI_WARPS * 256 * 4;
  // barriers (8 bytes each)
  static constexpr int a_full = 0, a_empty = 1, b_full = 2, b_empty = b_full + STAGES,
                       n_full = b_empty + STAGES, n_empty = n_full + NBN,
                       t_full = n_empty + NBN, t_empty = t_full + 2, x_full = t_empty + 2,
                       x_empty = x_full + 2, nbar = x_empty + 2;
  // folded-norm FP16 with at most two data chunks: the database-side extras (256 rows x 32 bytes)
  // have their own 2-slot ring in the unused fourth query-chunk slot, so that a tile takes two
  // stages of the B ring, not three
  static constexpr int xb_off = a_off + 3 * A_CHUNK_BYTES;
  static constexpr int XB_BYTES = TN * 32;
  static constexpr int tmem_ptr_off = bar_off + nbar * 8;
  static constexpr int total = tmem_ptr_off + 16;
};
constexpr int TF32_SMEM_BYTES = Smem::total;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
// the same with mbarrier.test_wait (pure spinning, no hardware suspend): bring-up comparison
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// the same box lands at the same shared offset in every CTA of `mask`, and completes the same
// mbarrier offset in each of them
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap *map, uint32_t bar,
                                               int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
          "r"(bar),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void *src, uint32_t bytes,
                                             uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {  // arrives on bar when prior MMAs retire
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, TF32 inputs, FP32 accumulate
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The same two, for a warp that walks the issue loop with ALL lanes (warp-uniform control flow,
// operands identical in every lane) and lets elect.sync pick the issuing lane inside the
// instruction's own predicate.  Under `if (lane == 0)` ptxas cannot tell that one lane is
// active and wraps every tcgen05 instruction in a vote/broadcast loop (ELECT, 5 x
// R2UR.BROADCAST, BRA.U.ANY): 170 instructions per 4 MMAs, which made the ISSUE warp the
// kernel's critical path whenever the epilogue warps of its scheduler were busy.  elect.sync
// returns the same leader for the same mask every time, so one thread owns all MMAs and commits.
__device__ __forceinline__ void tc_mma_tf32_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                  uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f8f6f4 with E4M3 operands (one byte per element, K = 32 per instruction, FP32
// accumulate): the Hamming path feeds it +-1.0 (0x38 / 0xB8), for which every product and every
// partial sum is an exactly representable integer.
__device__ __forceinline__ void tc_mma_f8_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f16 with FP16 operands (K = 16 per instruction, FP32 accumulate): the same 10 explicit
// mantissa bits as TF32 at half the operand bytes -- half as many MMAs, shared-memory operand
// reads and accumulator passes per tile.
__device__ __forceinline__ void tc_mma_f16_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                 uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}" ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void tc_commit_mc_elect(uint32_t bar, uint16_t mask) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t"
      "}" ::"r"(bar),
      "h"(mask)
      : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
        "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns -> 16 registers per thread
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 64 consecutive 32-bit columns -> 64 registers per thread
__device__ __forceinline__ void tc_ld64(uint32_t taddr, uint32_t (&v)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 128 consecutive columns holding 16-bit accumulators (D = F16), packed two per register:
// register r = columns (2 r, 2 r + 1) in its (low, high) half -> 64 registers per thread
__device__ __forceinline__ void tc_ld128p(uint32_t taddr, uint32_t (&v)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.pack::16b.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(taddr)
      : "memory");
}
template <int W>
__device__ __forceinline__ void tc_ldw(uint32_t taddr, uint32_t (&v)[W]) {
  if constexpr (W == 16) tc_ld16(taddr, v);
  else if constexpr (W == 32) tc_ld32(taddr, v);
  else tc_ld64(taddr, v);
}
__device__ __forceinline__ void tc_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor: rows of 128 bytes, 8-row groups 1024
// bytes apart (SBO), descriptor version 1 (sm_100), layout type 2.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  uint64_t desc = 0;
  desc |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address, 16-byte units
  desc |= (uint64_t)1 << 16;                         // leading byte offset (unused with swizzle)
  desc |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset
  desc |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
  desc |= (uint64_t)2 << 61;                         // SWIZZLE_128B
  return desc;
}
// K-major, SWIZZLE_32B descriptor: rows of 32 bytes (one MMA K step), 8-row groups 256 bytes
// apart -- the layout a TMA box of 16 halfs x rows with CU_TENSOR_MAP_SWIZZLE_32B produces.  Used
// for the 16 extra K elements of the folded-norm FP16 operands, so that they cost 32 bytes per
// row of shared-memory traffic instead of a whole (zero-filled) 128-byte chunk.
__device__ __forceinline__ uint64_t smem_desc_sw32(uint32_t saddr) {
  uint64_t desc = 0;
  desc |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  desc |= (uint64_t)1 << 16;
  desc |= (uint64_t)(256 >> 4) << 32;
  desc |= (uint64_t)1 << 46;
  desc |= (uint64_t)6 << 61;  // SWIZZLE_32B
  return desc;
}
// instruction descriptor, kind::tf32: D=F32, A=B=TF32, both K-major, N=256, M=128
constexpr uint32_t IDESC_TF32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) |
                                ((uint32_t)(TM >> 4) << 24);

// instruction descriptor, kind::f8f6f4 with E4M3 operands and kind::f16 with FP16 operands: D=F32,
// A=B=format 0 (E4M3 resp. F16), both K-major, N=256, M=128
constexpr uint32_t IDESC_F8 = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
constexpr uint32_t IDESC_F16 = IDESC_F8;

// operand kinds (Tf32Plan::kind).  The kind is a template parameter of the kernel: the issue
// loop of each instantiation carries no kind branches.
enum : int { OP_TF32 = 0, OP_F8 = 1, OP_F16 = 2, OP_F16N = 3 };  // F16N: FP16 with |b|^2 folded into K
constexpr int OP_F8P = 4;  // planning only (tf32_plan_tiles): E4M3 in the packed Hamming modes -- may pair up
// E4M3 rows whose |b|^2 is the SAME constant for every row (+-1 Hamming operands: 2 * bits): no |b|^2
// tiles, score = asc * acc + c0, the MAX-tree epilogue of the folded-norm kind; padding rows are E4M3
// NaN (a NaN accumulator is never admitted)
constexpr int OP_F8C = 5;
// OP_F8H: the OP_F8C operands with FP16 ACCUMULATORS (D = F16).  The Hamming contraction of +-1
// operands yields integers of magnitude <= 512, exact in FP16 at every step of the accumulation; the
// epilogue then drains half the registers per tile (tcgen05.ld ... .pack::16b) and runs its max tree
// on half2 pairs.  Kernel-side only: plans and operand layouts are those of OP_F8C.
constexpr int OP_F8H = 6;
template <int KIND>
__device__ __forceinline__ void tc_mma_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                             uint32_t accumulate) {
  if (KIND == OP_F8 || KIND == OP_F8C || KIND == OP_F8H)
    tc_mma_f8_elect(d_tmem, a_desc, b_desc, IDESC_F8, accumulate);
  else if (KIND == OP_F16 || KIND == OP_F16N)
    tc_mma_f16_elect(d_tmem, a_desc, b_desc, IDESC_F16, accumulate);
  else
    tc_mma_tf32_elect(d_tmem, a_desc, b_desc, IDESC_TF32, accumulate);
}

// bring-up instrumentation (YAEL_B200_TF32_DEBUG bit 512): clock64() deltas of one epilogue warp and
// of the MMA issuer per CTA: [0] issuer waits for a free accumulator, [1] for operand stages, [2]
// for the extras, [3] issuer total, [4] epilogue warp 0 waits for an accumulator, [5] drains it,
// [6] hands it back, [7] epilogue total, [8] tiles
__device__ long long g_tf32_clk[160][16];

// ------------------------------------------------------------------ parameters
struct Tf32Params {
  int nq, nb, d;
  int nkc;            // K chunks of 32 floats
  int last_k8;        // K=8 steps in the last chunk
  int tiles_q;        // query tiles
  int nbt;            // database tiles (256 rows)
  int range_tiles;    // database tiles per range
  int splits;         // ranges
  int lists;          // shortlists per query = 2 * splits (one per range and column half)
  int items;          // tiles_q * splits
  int kprime, cap;
  const float *bnorm;  // [nbt*256], padded with +inf
  float2 *scratch;     // [gridDim.x][256][cap] append lists (score, id bits)
  float *out_score;    // [nq][lists][kprime]
  int *out_id;         // [nq][lists][kprime]
  float *dump;         // [nq][dump_ld] raw scores by LOGICAL column (lists are not produced)
  long dump_ld;
  const float *thr_init;  // [nq] initial admission threshold per query (NULL = +inf): rows whose
                          // score is >= thr_init[q] are known not to matter
  float *out_thr;      // [nq][lists] final admission threshold of every list: each row of the
                       // list's range that is NOT in the list has a score >= this value
  int tile_stride;     // logical tile j covers database tile j * tile_stride (sampling pass)
  float *gmin;         // [nq][gmin_ld] group-minimum mode (lists are not produced): the smallest
                       // score of every run of gsize (16, 32, 64 or 128) LOGICAL columns
  long gmin_ld;
  int gsize;
  int *out_cnt;        // [nq][lists_ld] entries published per list (NULL: not wanted)
  int lists_ld;        // lists per query in the OUTPUT arrays (several passes may share them)
  int list0;           // first output list of this pass
  int id0;             // added to every published row id (pass over a row range of the database)
  const float *k1_margin;  // [nq] k = 1 mode (NULL = top-k' mode): admit s < best_so_far + margin
  int pair;            // 1: CTAs run as clusters of 2 that share every database chunk: each CTA
                       // fetches half of it and multicasts it to both (L2->SM traffic halves)
  int tiles_q2;        // query-tile pairs when pair != 0
  // packed Hamming modes (EPI_HAMP / EPI_HAMG): every accumulator carries ham_slots distances of
  // ham_slots consecutive database rows, acc = dot_0 + 2^8 dot_1 + 2^16 dot_2 (yb_hamming_tc.cu)
  int xk;              // folded-norm FP16: index of the extras chunk (= number of data chunks), its
  int xcol;            // first column (halfs); -1 / 0 otherwise.  nkc counts the extras chunk too.
  int xring;           // 1: the database-side extras travel through their own ring (Smem::xb_off)
  int order;           // work items: 0 = range-major (CTAs of a wave stream the SAME database range),
                       // 1 = query-tile-major (a wave covers every range: ~148/splits CTAs per range)
  int ham_slots;       // 2 or 3
  int ham_nb;          // real database rows (ids >= ham_nb are padding)
  float ham_magic;     // 2^23 + (bits/2)(1 + 2^8 [+ 2^16]): fma(acc, -0.5, magic) holds ham_i in byte i
  float c0;            // OP_F8C: the constant |b|^2 of every row (0 for the folded-norm FP16 kind)
  const float *acc_scale;  // device scalar: score = acc * (*acc_scale) + |b|^2 (NULL = -2: operands
                           // unscaled); FP16 operands are scaled by 2^sigma, *acc_scale = -2^(1-2 sigma)
  int debug;           // bring-up switches (YAEL_B200_TF32_DEBUG): 1 skip epilogue math, 2 skip MMAs
  int nka;             // 2-SM kernel, resident query tile: chunk slots it occupies (MAX_NKC .. MAX_NKC_WIDE)
  int stages2;         // ... and the database ring stages that fit behind it (STAGES2 .. 5)
  int dephase;         // experiment: cycles the column-half-1 epilogue warps wait once per work item
  int cross_tma;       // EPI_CROSS: tiles leave through shared memory and TMA stores ...
  uint32_t cross_stage0, cross_stage1;  // ... staged in these rings (byte offsets, one per column half)
};

// ------------------------------------------------------------------ warp-cooperative compaction
// Reduce one query's list (n <= MAXL entries, global memory, SoA) to its kp smallest scores in
// place.  Returns the new admission threshold (the kp-th smallest score): every dropped entry
// and every later candidate with score >= thr is not among the kp smallest.
//
// The keys are pulled into registers once (MAXL/32 per lane, coalesced, all loads in flight
// together), a 4 x 8-bit radix select runs on the registers with a per-warp shared-memory
// histogram, and one more sweep moves the survivors' ids.
constexpr int MAXL = 1024;
constexpr int KPL = MAXL / 32;  // keys per lane

__device__ __forceinline__ float warp_select_compact(float2 *list, int n, int kp, int *hist) {
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  uint32_t keys[KPL];
#pragma unroll
  for (int j = 0; j < KPL; j++) {
    const int i = j * 32 + lane;
    keys[j] = i < n ? float_key(list[i].x) : 0xffffffffu;
  }
  uint32_t prefix = 0, mask = 0;
  int rem = kp;
#pragma unroll 1
  for (int shift = 24; shift >= 0; shift -= 8) {
#pragma unroll
    for (int u = 0; u < 8; u++) hist[lane * 8 + u] = 0;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < KPL; j++)
      if ((keys[j] & mask) == prefix) atomicAdd(&hist[(keys[j] >> shift) & 255], 1);
    __syncwarp();
    int loc[8], sum = 0;
#pragma unroll
    for (int u = 0; u < 8; u++) {
      loc[u] = hist[lane * 8 + u];
      sum += loc[u];
    }
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    int off = inc - sum;
    int found = -1, newrem = 0;
#pragma unroll
    for (int u = 0; u < 8; u++) {
      if (found < 0 && off < rem && rem <= off + loc[u]) {
        found = lane * 8 + u;
        newrem = rem - off;
      }
      off += loc[u];
    }
    const unsigned who = __ballot_sync(0xffffffffu, found >= 0);
    const int src = __ffs(who) - 1;
    found = __shfl_sync(0xffffffffu, found, src);
    rem = __shfl_sync(0xffffffffu, newrem, src);
    prefix |= (uint32_t)found << shift;
    mask |= 0xffu << shift;
    __syncwarp();
  }
  const uint32_t pivot = prefix;
  int out = 0, eq_taken = 0;
#pragma unroll
  for (int j = 0; j < KPL; j++) {
    const uint32_t key = keys[j];
    const bool eq = key == pivot;
    const unsigned eqb = __ballot_sync(0xffffffffu, eq);
    const bool keep = key < pivot || (eq && eq_taken + __popc(eqb & lt_mask) < rem);
    const unsigned kb = __ballot_sync(0xffffffffu, keep);
    eq_taken += __popc(eqb);
    if (kb == 0) continue;
    float2 e = make_float2(0.f, 0.f);
    if (keep) e = list[j * 32 + lane];
    __syncwarp();
    if (keep) list[out + __popc(kb & lt_mask)] = e;
    out += __popc(kb);
  }
  __syncwarp();
  const uint32_t bits = (pivot & 0x80000000u) ? (pivot & 0x7fffffffu) : ~pivot;
  return __uint_as_float(bits);
}

// (d0, d1) = (a0, a1) * (sc, sc) + (c0, c1), one packed instruction (sc = -2 for unscaled operands)
__device__ __forceinline__ void ffma2_m2(float &d0, float &d1, uint32_t a0, uint32_t a1, float c0,
                                         float c1, float sc) {
  asm("{\n\t"
      ".reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%6, %6};\n\t"
      "mov.b64 rc, {%4, %5};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t"
      "}"
      : "=f"(d0), "=f"(d1)
      : "r"(a0), "r"(a1), "f"(c0), "f"(c1), "f"(sc));
}

#define YB_SC16_PARAMS float s0, float s1, float s2, float s3, float s4, float s5, float s6, float s7, \
                       float s8, float s9, float s10, float s11, float s12, float s13, float s14, float s15
#define YB_SC16_ARGS(a) a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], \
                        a[12], a[13], a[14], a[15]

__device__ __noinline__ int slow_append(YB_SC16_PARAMS, float thr, float2 *mylist, int cnt, int id0) {
  const float sc[16] = {s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15};
#pragma unroll
  for (int c = 0; c < 16; c++) {
    if (sc[c] < thr) {
      mylist[cnt] = make_float2(sc[c], __int_as_float(id0 + c));
      cnt++;
    }
  }
  return cnt;
}

// k = 1: everything within `margin` of the best score seen so far stays a candidate.
// margin >= 2 * (TF32 error bound), so the exact nearest row can never be refused.
// state[0] = thr, state[1] = best (in/out through a small per-thread struct in registers)
struct K1State {
  float thr, best;
  int cnt;
};
__device__ __noinline__ K1State slow_append_k1(YB_SC16_PARAMS, float m, K1State st, float margin,
                                               float2 *mylist, int cap, int id0) {
  const float sc[16] = {s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15};
#pragma unroll
  for (int c = 0; c < 16; c++) {
    if (sc[c] < st.thr) {
      mylist[st.cnt] = make_float2(sc[c], __int_as_float(id0 + c));
      st.cnt++;
    }
  }
  if (m < st.best) {
    st.best = m;
    st.thr = m + margin;
  }
  if (st.cnt > cap - 16) {  // drop what the tighter threshold no longer admits
    int j = 0;
    for (int e = 0; e < st.cnt; e++) {
      const float2 x = mylist[e];
      if (x.x < st.thr) mylist[j++] = x;
    }
    st.cnt = j;
    if (st.cnt > cap - 16) {  // a dense cluster of near-ties: give up on this list
      st.cnt = 0;
      st.best = __uint_as_float(0x7fc00000u);  // NaN marks the overflow
      st.thr = __uint_as_float(0xff800000u);
    }
  }
  return st;
}

// 16 accumulator columns of one query.  Fast path: scores and their minimum (one FFMA and one
// FMNMX3 lane per candidate); only when the minimum beats the query's admission threshold --
// rare once the threshold is tight -- are the 16 candidates tested one by one.
template <bool K1>
__device__ __forceinline__ void process_group(const uint32_t (&v)[16], const float *bn, float &thr,
                                              float &best, float margin, float2 *mylist, int &cnt,
                                              int cap, int id0, float asc) {
  float sc[16];
#pragma unroll
  for (int c4 = 0; c4 < 4; c4++) {
    const float4 b4 = *reinterpret_cast<const float4 *>(bn + c4 * 4);
    // two packed FMAs (fma.rn.f32x2: same rounding as the scalar fmaf, half the issue slots)
    ffma2_m2(sc[c4 * 4 + 0], sc[c4 * 4 + 1], v[c4 * 4 + 0], v[c4 * 4 + 1], b4.x, b4.y, asc);
    ffma2_m2(sc[c4 * 4 + 2], sc[c4 * 4 + 3], v[c4 * 4 + 2], v[c4 * 4 + 3], b4.z, b4.w, asc);
  }
  // fminf ignores NaN operands, which is what we want: a NaN score is never admitted
  float m01 = fminf(fminf(sc[0], sc[1]), fminf(sc[2], sc[3]));
  float m23 = fminf(fminf(sc[4], sc[5]), fminf(sc[6], sc[7]));
  float m45 = fminf(fminf(sc[8], sc[9]), fminf(sc[10], sc[11]));
  float m67 = fminf(fminf(sc[12], sc[13]), fminf(sc[14], sc[15]));
  const float m = fminf(fminf(m01, m23), fminf(m45, m67));
  if (m < thr) {
    if (K1) {
      K1State st = {thr, best, cnt};
      st = slow_append_k1(YB_SC16_ARGS(sc), m, st, margin, mylist, cap, id0);
      thr = st.thr;
      best = st.best;
      cnt = st.cnt;
    } else {
      cnt = slow_append(YB_SC16_ARGS(sc), thr, mylist, cnt, id0);
    }
  }
}

// FP16 operands carry |b|^2 INSIDE the contraction (yb_knn.cu, center_operands_h: three extra K
// elements per row hold -2^(2 sigma - 1) |b|^2 split into FP16 pieces, the query side 2^15), so the
// accumulator is acc' = 2^(2 sigma) (<q,b> - |b|^2 / 2) and the score is asc * acc' with
// asc = -2^(1 - 2 sigma) < 0: no |b|^2 staging, no FMA -- the group test is a MAX tree over the
// raw accumulators against thr / asc (exact: asc is a power of two).  NaN accumulators (padding
// rows, NaN data) never win a max and never pass a comparison.
__device__ __forceinline__ float max3f(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }  // one FMNMX3
__device__ __forceinline__ float group_max16(const uint32_t (&v)[16]) {
  // eight three-input maxima (FMNMX3 retires two values per instruction; the pairwise tree needed ten)
  const float a = max3f(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]));
  const float b = max3f(__uint_as_float(v[3]), __uint_as_float(v[4]), __uint_as_float(v[5]));
  const float c = max3f(__uint_as_float(v[6]), __uint_as_float(v[7]), __uint_as_float(v[8]));
  const float d = max3f(__uint_as_float(v[9]), __uint_as_float(v[10]), __uint_as_float(v[11]));
  const float e = max3f(__uint_as_float(v[12]), __uint_as_float(v[13]), __uint_as_float(v[14]));
  return fmaxf(max3f(a, b, c), max3f(d, e, __uint_as_float(v[15])));
}

template <bool K1>
__device__ __forceinline__ void process_group_nf(const uint32_t (&v)[16], float &thr, float &thrp,
                                                 float &best, float margin, float2 *mylist, int &cnt,
                                                 int cap, int id0, float asc, float inv_asc, float c0) {
  const float M = group_max16(v);
  if (M > thrp) {  // <=> asc * M + c0 < thr: some score of the group beats the admission threshold
    float sc[16];
#pragma unroll
    for (int c = 0; c < 16; c++) sc[c] = fmaf(asc, __uint_as_float(v[c]), c0);  // (c0 = 0: == asc * acc)
    if (K1) {
      K1State st = {thr, best, cnt};
      st = slow_append_k1(YB_SC16_ARGS(sc), fmaf(asc, M, c0), st, margin, mylist, cap, id0);
      thr = st.thr;
      best = st.best;
      cnt = st.cnt;
      thrp = __fmul_rn(thr - c0, inv_asc);
    } else {
      cnt = slow_append(YB_SC16_ARGS(sc), thr, mylist, cnt, id0);
    }
  }
}

// W = 32 or 64 accumulator columns of one query (folded norms): ONE threshold test for the whole
// load, the 16-column groups are only looked at when it fires.  Every tcgen05.ld + wait::ld round
// trip costs a warp ~230 cycles however many columns it brings (8 epilogue warps contend for the
// TMEM read port), so a 128-column half tile is drained in 4 (W = 32) or 2 (W = 64) round trips
// instead of 8.
template <bool K1, int W>
__device__ __forceinline__ void process_wide_nf(const uint32_t (&v)[W], float &thr, float &thrp,
                                                float &best, float margin, float2 *mylist, int &cnt,
                                                int cap, int id0, float asc, float inv_asc, float c0) {
  float gm[W / 16];
#pragma unroll
  for (int s = 0; s < W / 16; s++) gm[s] = group_max16(reinterpret_cast<const uint32_t(&)[16]>(v[16 * s]));
  float M = gm[0];
#pragma unroll
  for (int s = 1; s < W / 16; s++) M = fmaxf(M, gm[s]);
  if (M > thrp) {
#pragma unroll
    for (int s = 0; s < W / 16; s++) {
      if (gm[s] > thrp) {
        float sc[16];
#pragma unroll
        for (int c = 0; c < 16; c++) sc[c] = fmaf(asc, __uint_as_float(v[16 * s + c]), c0);
        if (K1) {
          K1State st = {thr, best, cnt};
          st = slow_append_k1(YB_SC16_ARGS(sc), fmaf(asc, gm[s], c0), st, margin, mylist, cap, id0 + 16 * s);
          thr = st.thr;
          best = st.best;
          cnt = st.cnt;
          thrp = __fmul_rn(thr - c0, inv_asc);
        } else {
          cnt = slow_append(YB_SC16_ARGS(sc), thr, mylist, cnt, id0 + 16 * s);
        }
      }
    }
  }
}

// The same for the top-k' mode with the COMMON admission handled inline: a fired group almost always
// holds exactly one value beyond the threshold -- its maximum, which the tree already produced -- so
// a 16-bit mask of `v > thrp` (independent compares, no chain through the list counter) finds its
// column and the entry is written without the out-of-line call (whose 16 dependent test / store /
// count steps cost a warp ~230 clk, about 1.6 times per tile at 400 admissions per query).  Two or
// more admissions in one group of 16 take the old path.
template <int W>
__device__ __forceinline__ void process_wide_lists(const uint32_t (&v)[W], float thr, float thrp,
                                                   float2 *mylist, int &cnt, int id0, float asc, float c0) {
  float gm[W / 16];
#pragma unroll
  for (int s = 0; s < W / 16; s++) gm[s] = group_max16(reinterpret_cast<const uint32_t(&)[16]>(v[16 * s]));
  float M = gm[0];
#pragma unroll
  for (int s = 1; s < W / 16; s++) M = fmaxf(M, gm[s]);
  if (M > thrp) {
#pragma unroll
    for (int s = 0; s < W / 16; s++) {
      if (gm[s] > thrp) {
        unsigned m = 0u;
#pragma unroll
        for (int c = 0; c < 16; c++) m |= (__uint_as_float(v[16 * s + c]) > thrp) ? (1u << c) : 0u;
        if (__popc(m) == 1) {
          mylist[cnt] = make_float2(fmaf(asc, gm[s], c0), __int_as_float(id0 + 16 * s + __ffs(m) - 1));
          cnt++;
        } else {
          float sc[16];
#pragma unroll
          for (int c = 0; c < 16; c++) sc[c] = fmaf(asc, __uint_as_float(v[16 * s + c]), c0);
          cnt = slow_append(YB_SC16_ARGS(sc), thr, mylist, cnt, id0 + 16 * s);
        }
      }
    }
  }
}

// The top-k' test on 128 columns of FP16 accumulators, two per register (OP_F8H): the max tree runs
// on half2 pairs (one HMNMX2 per two columns), a fired group of 16 columns (8 registers) is unpacked
// and handled like process_wide_lists does.  NaN accumulators (padding rows) never win a maximum.
__device__ __forceinline__ float h2max(uint32_t x) {
  const __half2 h = *reinterpret_cast<const __half2 *>(&x);
  return fmaxf(__low2float(h), __high2float(h));
}
__device__ __forceinline__ uint32_t h2m(uint32_t a, uint32_t b) {
  const __half2 r = __hmax2(*reinterpret_cast<const __half2 *>(&a), *reinterpret_cast<const __half2 *>(&b));
  return *reinterpret_cast<const uint32_t *>(&r);
}
__device__ __forceinline__ void process_wide_lists_h(const uint32_t (&v)[64], float thr, float thrp,
                                                     float2 *mylist, int &cnt, int id0, float asc, float c0) {
  uint32_t gm[8];
#pragma unroll
  for (int s = 0; s < 8; s++)
    gm[s] = h2m(h2m(h2m(v[8 * s], v[8 * s + 1]), h2m(v[8 * s + 2], v[8 * s + 3])),
                h2m(h2m(v[8 * s + 4], v[8 * s + 5]), h2m(v[8 * s + 6], v[8 * s + 7])));
  const float M = h2max(h2m(h2m(h2m(gm[0], gm[1]), h2m(gm[2], gm[3])), h2m(h2m(gm[4], gm[5]), h2m(gm[6], gm[7]))));
  if (M > thrp) {
#pragma unroll
    for (int s = 0; s < 8; s++) {
      const float g = h2max(gm[s]);
      if (g > thrp) {
        float x[16];
#pragma unroll
        for (int r = 0; r < 8; r++) {
          const __half2 h = *reinterpret_cast<const __half2 *>(&v[8 * s + r]);
          x[2 * r] = __low2float(h);
          x[2 * r + 1] = __high2float(h);
        }
        unsigned m = 0u;
#pragma unroll
        for (int c = 0; c < 16; c++) m |= (x[c] > thrp) ? (1u << c) : 0u;
        if (__popc(m) == 1) {
          mylist[cnt] = make_float2(fmaf(asc, g, c0), __int_as_float(id0 + 16 * s + __ffs(m) - 1));
          cnt++;
        } else {
          float sc[16];
#pragma unroll
          for (int c = 0; c < 16; c++) sc[c] = fmaf(asc, x[c], c0);
          cnt = slow_append(YB_SC16_ARGS(sc), thr, mylist, cnt, id0 + 16 * s);
        }
      }
    }
  }
}

// smallest score of 16 accumulator columns (sampling pass)
__device__ __forceinline__ float group_min16(const uint32_t (&v)[16], const float *bn, float asc) {
  float sc[16];
#pragma unroll
  for (int c4 = 0; c4 < 4; c4++) {
    const float4 b4 = *reinterpret_cast<const float4 *>(bn + c4 * 4);
    sc[c4 * 4 + 0] = fmaf(__uint_as_float(v[c4 * 4 + 0]), asc, b4.x);
    sc[c4 * 4 + 1] = fmaf(__uint_as_float(v[c4 * 4 + 1]), asc, b4.y);
    sc[c4 * 4 + 2] = fmaf(__uint_as_float(v[c4 * 4 + 2]), asc, b4.z);
    sc[c4 * 4 + 3] = fmaf(__uint_as_float(v[c4 * 4 + 3]), asc, b4.w);
  }
  float m01 = fminf(fminf(sc[0], sc[1]), fminf(sc[2], sc[3]));
  float m23 = fminf(fminf(sc[4], sc[5]), fminf(sc[6], sc[7]));
  float m45 = fminf(fminf(sc[8], sc[9]), fminf(sc[10], sc[11]));
  float m67 = fminf(fminf(sc[12], sc[13]), fminf(sc[14], sc[15]));
  return fminf(fminf(m01, m23), fminf(m45, m67));
}

// k = 1 margin mode, the common admission inline (as process_wide_lists): a fired group nearly always
// holds ONE value within the margin -- its maximum.  It is appended, and when it beats the best
// score so far the threshold follows (slow_append_k1's rule: append against the OLD threshold, then
// best = min, thr = best + margin).  Several admissions in one group, or a list close to its
// capacity (pruning / overflow handling), take the out-of-line path.
template <int W>
__device__ __forceinline__ void process_wide_k1(const uint32_t (&v)[W], float &thr, float &thrp, float &best,
                                                float margin, float2 *mylist, int &cnt, int cap, int id0,
                                                float asc, float inv_asc, float c0) {
  float gm[W / 16];
#pragma unroll
  for (int s = 0; s < W / 16; s++) gm[s] = group_max16(reinterpret_cast<const uint32_t(&)[16]>(v[16 * s]));
  float M = gm[0];
#pragma unroll
  for (int s = 1; s < W / 16; s++) M = fmaxf(M, gm[s]);
  if (M > thrp) {
#pragma unroll
    for (int s = 0; s < W / 16; s++) {
      if (gm[s] > thrp) {
        unsigned m = 0u;
#pragma unroll
        for (int c = 0; c < 16; c++) m |= (__uint_as_float(v[16 * s + c]) > thrp) ? (1u << c) : 0u;
        if (__popc(m) == 1 && cnt < cap - 16) {
          const float sc = fmaf(asc, gm[s], c0);
          mylist[cnt] = make_float2(sc, __int_as_float(id0 + 16 * s + __ffs(m) - 1));
          cnt++;
          if (sc < best) {
            best = sc;
            thr = sc + margin;
            thrp = __fmul_rn(thr - c0, inv_asc);
          }
        } else {
          float sc[16];
#pragma unroll
          for (int c = 0; c < 16; c++) sc[c] = fmaf(asc, __uint_as_float(v[16 * s + c]), c0);
          K1State st = {thr, best, cnt};
          st = slow_append_k1(YB_SC16_ARGS(sc), fmaf(asc, gm[s], c0), st, margin, mylist, cap, id0 + 16 * s);
          thr = st.thr;
          best = st.best;
          cnt = st.cnt;
          thrp = __fmul_rn(thr - c0, inv_asc);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ packed Hamming helpers
#define YB_U16_PARAMS uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4, uint32_t w5, \
                      uint32_t w6, uint32_t w7, uint32_t w8, uint32_t w9, uint32_t w10, uint32_t w11, \
                      uint32_t w12, uint32_t w13, uint32_t w14, uint32_t w15

// admission threshold of the packed modes: thr is in score units (4 * ham); a row is admitted iff
// ham < tau = ceil(thr / 4).  The byte test below needs tau <= 128, so thr is capped at 512 (rows
// at distance >= 128 are then "refused at 512", which the list threshold reports faithfully).
__device__ __forceinline__ uint32_t ham_tau3(float &thr) {
  thr = fminf(thr, 512.0f);
  const int tau = thr > 0.f ? (int)ceilf(thr * 0.25f) : 0;
  return (uint32_t)tau * 0x010101u;
}

// w = bits of fma(acc, -0.5, magic): byte i = ham_i.  Scan the 16 words of a group; row id of
// (word c, slot i) = row0 + c * slots + i.
__device__ __noinline__ int slow_append_ham(YB_U16_PARAMS, uint32_t tau3, uint32_t mask, int slots,
                                            float2 *mylist, int cnt, int row0, int nb_real) {
  const uint32_t w[16] = {w0, w1, w2, w3, w4, w5, w6, w7, w8, w9, w10, w11, w12, w13, w14, w15};
  const int tau = (int)(tau3 & 0xffu);
#pragma unroll
  for (int c = 0; c < 16; c++) {
    const uint32_t x = w[c];
    if ((((x - tau3) & ~x) & mask) == 0u) continue;
    for (int i = 0; i < slots; i++) {
      const int h = (int)((x >> (8 * i)) & 0xffu);
      const int row = row0 + c * slots + i;
      if (h < tau && row < nb_real) {
        mylist[cnt] = make_float2((float)(4 * h), __int_as_float(row));
        cnt++;
      }
    }
  }
  return cnt;
}

// 16 packed accumulators of one query: per accumulator half an FFMA2, one IADD and one LOP3 decide
// whether ANY of its (up to three) distances is below the threshold ("has a byte less than tau":
// ((x - tau * 0x010101) & ~x & 0x808080) != 0, exact for the existence test, tau <= 128)
__device__ __forceinline__ void process_group_ham(const uint32_t (&v)[16], float magic, uint32_t tau3,
                                                  uint32_t mask, int slots, float2 *mylist, int &cnt,
                                                  int row0, int nb_real) {
  uint32_t w[16];
  uint32_t orr = 0u;
#pragma unroll
  for (int c = 0; c < 16; c += 2) {
    float y0, y1;
    ffma2_m2(y0, y1, v[c], v[c + 1], magic, magic, -0.5f);
    w[c] = __float_as_uint(y0);
    w[c + 1] = __float_as_uint(y1);
    orr |= (w[c] - tau3) & ~w[c];
    orr |= (w[c + 1] - tau3) & ~w[c + 1];
  }
  if (orr & mask)
    cnt = slow_append_ham(w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], w[8], w[9], w[10], w[11],
                          w[12], w[13], w[14], w[15], tau3, mask, slots, mylist, cnt, row0, nb_real);
}

// W = 64 packed accumulators of one query, converted IN PLACE (v[c] becomes the word whose byte i
// is ham_i): per accumulator half an FFMA2, one IADD and one LOP3 (acc | ((x - tau3) & ~x)); the
// 16-word groups are only decoded when their test fires.
template <int W>
__device__ __forceinline__ void process_wide_ham(uint32_t (&v)[W], float magic, uint32_t tau3,
                                                 uint32_t mask, int slots, float2 *mylist, int &cnt,
                                                 int row0, int nb_real) {
  uint32_t gm[W / 16];
#pragma unroll
  for (int s = 0; s < W / 16; s++) gm[s] = 0u;
#pragma unroll
  for (int c = 0; c < W; c += 2) {
    float y0, y1;
    ffma2_m2(y0, y1, v[c], v[c + 1], magic, magic, -0.5f);
    v[c] = __float_as_uint(y0);
    v[c + 1] = __float_as_uint(y1);
    gm[c / 16] |= (v[c] - tau3) & ~v[c];
    gm[c / 16] |= (v[c + 1] - tau3) & ~v[c + 1];
  }
  uint32_t any = gm[0];
#pragma unroll
  for (int s = 1; s < W / 16; s++) any |= gm[s];
  if (any & mask) {
#pragma unroll
    for (int s = 0; s < W / 16; s++) {
      if (gm[s] & mask)
        cnt = slow_append_ham(v[16 * s + 0], v[16 * s + 1], v[16 * s + 2], v[16 * s + 3], v[16 * s + 4],
                              v[16 * s + 5], v[16 * s + 6], v[16 * s + 7], v[16 * s + 8], v[16 * s + 9],
                              v[16 * s + 10], v[16 * s + 11], v[16 * s + 12], v[16 * s + 13],
                              v[16 * s + 14], v[16 * s + 15], tau3, mask, slots, mylist, cnt,
                              row0 + 16 * s * slots, nb_real);
    }
  }
}

// smallest distance among the slots of 16 packed accumulators, as a score (sampling pass)
__device__ __forceinline__ float ham_group_min16(const uint32_t (&v)[16], float magic, int slots) {
  uint32_t m = 255u;
#pragma unroll
  for (int c = 0; c < 16; c++) {
    const uint32_t x = __float_as_uint(fmaf(__uint_as_float(v[c]), -0.5f, magic));
    m = min(m, x & 0xffu);
    m = min(m, (x >> 8) & 0xffu);
    if (slots == 3) m = min(m, (x >> 16) & 0x7fu);
  }
  return (float)(4u * m);
}

// ------------------------------------------------------------------ epilogue role
struct EpiCtx {
  unsigned char *smem;
  uint32_t sbase, tmem_base;
  int warp, lane;
  int first_item, item_step, tq_div;
  int pair;            // 0: independent CTAs; 1: multicast pairs
  uint32_t crank;
  uint32_t t_empty_addr0, t_empty_addr1;  // where to signal "accumulator buffer drained" (local or leader CTA)
  int t_empty_remote;        // the address is a shared::cluster address of the peer CTA
  int n_full0, n_empty0, t_full0;  // barrier indices (the two kernels have different ring depths)
  const CUtensorMap *map_out;      // EPI_CROSS with staged TMA stores: the [nb][ld] output matrix
};

// arrive on a barrier of ANOTHER CTA of the cluster (shared::cluster address from mapa).  Default
// semantics (release at CTA scope), as CUTLASS' ClusterBarrier::arrive(cta_id) does: what has to be
// ordered here are tcgen05 operations, which the tcgen05.fence pair around the barrier orders; an
// explicit .release.cluster made every hand-back cost ~1200 cycles (measured with the clock
// attribution of YAEL_B200_TF32_DEBUG bit 512).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// TMA store of a staged tile (shared -> global through the output's tensor map: rows and queries
// beyond the matrix are clipped by the hardware), bulk-group bookkeeping, the generic -> async proxy
// fence that makes the threads' shared-memory writes visible to it, and a named barrier for the
// four warps that share a staging buffer.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// EPI_CROSS staging: per column half a ring of three [8 rows][128 queries] float buffers.  Half 0
// uses the |b|^2 ring + histogram area (12 KB, idle in this mode), half 1 sits behind the barrier
// block (CROSS_STAGE1_OFF, defined with Smem2): the operand area keeps its full ring.
constexpr int CROSS_STAGE_BYTES = 8 * TM * 4;

// MODE selects what the epilogue does with a tile (one kernel instantiation per mode: the hot
// loop of each stays compact and contiguous in the instruction cache, and carries no per-tile
// mode branches)
enum : int { EPI_LISTS = 0, EPI_NEAREST = 1, EPI_DUMP = 2, EPI_GMIN = 3, EPI_HAMP = 4, EPI_HAMG = 5,
              EPI_CROSS = 6 };  // CROSS: the full distance matrix, out[row * ld + query] (compute_cross_distances)

template <int MODE, int KIND, int LDW>
__device__ __forceinline__ void run_epilogue(const Tf32Params &P, const EpiCtx &E) {
  constexpr bool NF = KIND == OP_F16N || KIND == OP_F8C || KIND == OP_F8H;  // |b|^2 folded into the contraction / constant
  const float c0 = (KIND == OP_F8C || KIND == OP_F8H) ? P.c0 : 0.f;
  // no |b|^2 tiles travel: folded norms, and the packed Hamming modes (integer epilogue)
  constexpr bool NOBN = NF || MODE == EPI_HAMP || MODE == EPI_HAMG;
  unsigned char *smem = E.smem;
  const uint32_t sbase = E.sbase, tmem_base = E.tmem_base;
  const int warp = E.warp, lane = E.lane;
  const int first_item = E.first_item, item_step = E.item_step, tq_div = E.tq_div;
  const uint32_t crank = E.crank;
  auto bar = [&](int i) { return sbase + Smem::bar_off + 8 * i; };
  {
    const int team = warp / TEAM_WARPS, w8 = warp % TEAM_WARPS;
    const int quarter = w8 & 3, half = w8 >> 2;
    const int t = quarter * 32 + lane;  // query row inside the tile == TMEM lane
    int *hist = (int *)(smem + Smem::hist_off) + warp * 256;
    float2 *mylist = P.scratch +
                     (((size_t)blockIdx.x * EPI_TEAMS + team) * (2 * TM) + half * TM + t) * P.cap;
    const float inf = __uint_as_float(0x7f800000u);
    const float asc = P.acc_scale ? __ldg(P.acc_scale) : -2.0f;
    const float inv_asc = 1.0f / asc;  // exact: asc is a power of two
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + half * HALF_N;
    uint32_t tcount = 0;
    uint32_t cross_buf = 0;  // EPI_CROSS: position in the staging ring
    const bool clk_on = (P.debug & 512) && warp == 0;
    long long ck_wait = 0, ck_work = 0, ck_back = 0, ck_t0 = clk_on ? clock64() : 0, ck_a = 0, ck_b = 0;
    for (int item = first_item; item < P.items; item += item_step) {
      const int sp = P.order ? item % P.splits : item / tq_div;
      const int qi = P.order ? item / P.splits : item - sp * tq_div;
      const int qt = E.pair ? qi * 2 + (int)crank : qi;
      const int jt0 = sp * P.range_tiles, jt1 = min(P.nbt, jt0 + P.range_tiles);
      const int q = qt * TM + t;
      const bool valid = q < P.nq;
      float thr = valid ? (P.thr_init ? P.thr_init[q] : inf) : -inf;
      float best = inf;
      constexpr bool k1 = MODE == EPI_NEAREST;
      const float margin = (k1 && valid) ? P.k1_margin[q] : 0.f;
      int cnt = 0;
      float thrp = __fmul_rn(thr - c0, inv_asc);  // NF: admit iff acc' > thrp (= (thr - c0) / asc, asc < 0)
      uint32_t tau3 = 0u;
      const uint32_t ham_mask = P.ham_slots == 3 ? 0x808080u : 0x8080u;
      if (MODE == EPI_HAMP) tau3 = ham_tau3(thr);
      // ---- software-pipelined drain (folded-norm / constant-norm kinds, top-k' and k = 1 modes):
      // the accumulators of tile t+1 are loaded WHILE tile t is being tested.  A half tile is 64
      // registers (va: columns 0-63, vb: 64-127 of this thread's half): after va of tile t has been
      // tested it is free, so the epilogue waits for tile t+1's accumulator (long complete by then),
      // issues its first load into va, tests vb of tile t in the load's shadow, then issues the second
      // load into vb, waits for both and hands the buffer back.  Per tile the warp no longer pays
      // "wait for the accumulator + two loads" (~570 of its ~1640 clk) in series with the tests.
      // ---- FP16 accumulators (OP_F8H, top-k' mode): a half tile is ONE packed load of 64 registers;
      // tile t+1 is loaded into the other register set while tile t is tested (two tiles per trip)
      constexpr bool HACC = KIND == OP_F8H;
      if constexpr (HACC && MODE == EPI_LISTS) {
        uint32_t va[64], vb[64];
        // 16-bit accumulators still occupy one 32-bit TMEM column each: same column addresses
        auto hand_back_h = [&](uint32_t b) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            const uint32_t te = b ? E.t_empty_addr1 : E.t_empty_addr0;
            if (E.t_empty_remote)
              mbar_arrive_cluster(te);
            else
              mbar_arrive(te);
          }
        };
        auto room_check = [&]() {
          unsigned need = __ballot_sync(0xffffffffu, cnt > P.cap - HALF_N);
          while (need) {
            const int owner = __ffs(need) - 1;
            need &= need - 1;
            float2 *l = (float2 *)__shfl_sync(0xffffffffu, (unsigned long long)mylist, owner);
            const int n = __shfl_sync(0xffffffffu, cnt, owner);
            __syncwarp();
            const float nt = warp_select_compact(l, n, P.kprime, hist);
            if (lane == owner) {
              thr = nt;
              cnt = P.kprime;
            }
          }
          thrp = __fmul_rn(thr - c0, inv_asc);
        };
        {
          const uint32_t b0 = tcount & 1;
          mbar_wait(bar(E.t_full0 + b0), (tcount >> 1) & 1);
          tc_fence_after();
          tc_ld128p(lane_addr + b0 * TN, va);
          tc_wait_ld();
          hand_back_h(b0);
        }
        for (int jt = jt0; jt < jt1;) {
          {  // tile jt sits in va
            const int n0 = jt * P.tile_stride * TN + half * HALF_N + P.id0;
            const bool more = jt + 1 < jt1;
            const uint32_t nbuf = (tcount + 1) & 1;
            if (more) {
              mbar_wait(bar(E.t_full0 + nbuf), ((tcount + 1) >> 1) & 1);
              tc_fence_after();
              tc_ld128p(lane_addr + nbuf * TN, vb);
            }
            if (!(P.debug & 1)) process_wide_lists_h(va, thr, thrp, mylist, cnt, n0, asc, c0);
            if (more) {
              tc_wait_ld();
              hand_back_h(nbuf);
            }
            room_check();
            jt++;
            tcount++;
            if (!more) break;
          }
          {  // tile jt sits in vb
            const int n0 = jt * P.tile_stride * TN + half * HALF_N + P.id0;
            const bool more = jt + 1 < jt1;
            const uint32_t nbuf = (tcount + 1) & 1;
            if (more) {
              mbar_wait(bar(E.t_full0 + nbuf), ((tcount + 1) >> 1) & 1);
              tc_fence_after();
              tc_ld128p(lane_addr + nbuf * TN, va);
            }
            if (!(P.debug & 1)) process_wide_lists_h(vb, thr, thrp, mylist, cnt, n0, asc, c0);
            if (more) {
              tc_wait_ld();
              hand_back_h(nbuf);
            }
            room_check();
            jt++;
            tcount++;
          }
        }
      }
      constexpr bool PIPE = NF && !HACC && LDW == 128 && EPI_TEAMS == 1 && (MODE == EPI_LISTS || MODE == EPI_NEAREST);
      const bool piped = PIPE && !(P.debug & (2048 | 512 | 1 | 256));  // (bit 2048: the serial drain, A/B)
#ifdef YB_PIPE4
      // experiment (-DYB_PIPE4, top-k' mode): the half tile as FOUR 32-column register sets; after a
      // set has been tested the same columns of the next tile are loaded into it, so three of the
      // four loads of a tile are long complete when the one tcgen05.wait::ld per tile is reached
      if (PIPE && piped && MODE == EPI_LISTS) {
        uint32_t r0[32], r1[32], r2[32], r3[32];
        auto hand_back4 = [&](uint32_t b) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            const uint32_t te = b ? E.t_empty_addr1 : E.t_empty_addr0;
            if (E.t_empty_remote)
              mbar_arrive_cluster(te);
            else
              mbar_arrive(te);
          }
        };
        {
          const uint32_t b0 = tcount & 1;
          mbar_wait(bar(E.t_full0 + b0), (tcount >> 1) & 1);
          tc_fence_after();
          const uint32_t ta = lane_addr + b0 * TN;
          tc_ld32(ta, r0);
          tc_ld32(ta + 32, r1);
          tc_ld32(ta + 64, r2);
          tc_ld32(ta + 96, r3);
          tc_wait_ld();
          hand_back4(b0);
        }
        for (int jt = jt0; jt < jt1; jt++, tcount++) {
          const int n0 = jt * P.tile_stride * TN + half * HALF_N + P.id0;
          const bool more = jt + 1 < jt1;
          const uint32_t nbuf = (tcount + 1) & 1;
          const uint32_t tan = lane_addr + nbuf * TN;
          process_wide_lists<32>(r0, thr, thrp, mylist, cnt, n0, asc, c0);
          if (more) {
            mbar_wait(bar(E.t_full0 + nbuf), ((tcount + 1) >> 1) & 1);
            tc_fence_after();
            tc_ld32(tan, r0);
          }
          process_wide_lists<32>(r1, thr, thrp, mylist, cnt, n0 + 32, asc, c0);
          if (more) tc_ld32(tan + 32, r1);
          process_wide_lists<32>(r2, thr, thrp, mylist, cnt, n0 + 64, asc, c0);
          if (more) tc_ld32(tan + 64, r2);
          process_wide_lists<32>(r3, thr, thrp, mylist, cnt, n0 + 96, asc, c0);
          if (more) {
            tc_ld32(tan + 96, r3);
            tc_wait_ld();
            hand_back4(nbuf);
          }
          unsigned need = __ballot_sync(0xffffffffu, cnt > P.cap - HALF_N);
          while (need) {
            const int owner = __ffs(need) - 1;
            need &= need - 1;
            float2 *l = (float2 *)__shfl_sync(0xffffffffu, (unsigned long long)mylist, owner);
            const int n = __shfl_sync(0xffffffffu, cnt, owner);
            __syncwarp();
            const float nt = warp_select_compact(l, n, P.kprime, hist);
            if (lane == owner) {
              thr = nt;
              cnt = P.kprime;
            }
          }
          thrp = __fmul_rn(thr - c0, inv_asc);
        }
      } else
#endif
      if (PIPE && piped) {
        constexpr bool K1W = MODE == EPI_NEAREST;
        uint32_t va[64], vb[64];
        auto hand_back = [&](uint32_t b) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            const uint32_t te = b ? E.t_empty_addr1 : E.t_empty_addr0;
            if (E.t_empty_remote)
              mbar_arrive_cluster(te);
            else
              mbar_arrive(te);
          }
        };
        {  // prologue: the first tile of the item
          const uint32_t b0 = tcount & 1;
          mbar_wait(bar(E.t_full0 + b0), (tcount >> 1) & 1);
          tc_fence_after();
          const uint32_t ta = lane_addr + b0 * TN;
          tc_ldw<64>(ta, va);
          tc_ldw<64>(ta + 64, vb);
          tc_wait_ld();
          hand_back(b0);
          // experiment (YAEL_B200_EPI_DEPHASE=<cycles>): the two epilogue warps of a scheduler drain
          // the two column halves of the SAME tile in lockstep -- both wait for TMEM loads at the same
          // time, both run their max trees at the same time; a one-time delay of the second half's
          // warps puts one's tree under the other's wait
          if (P.dephase > 0 && half == 1) {
            const long long t0 = clock64();
            while (clock64() - t0 < P.dephase) {}
          }
        }
        for (int jt = jt0; jt < jt1; jt++, tcount++) {
          const int n0 = jt * P.tile_stride * TN + half * HALF_N + P.id0;
          const bool more = jt + 1 < jt1;
          const uint32_t nbuf = (tcount + 1) & 1;
          const uint32_t tan = lane_addr + nbuf * TN;
          if (!K1W) process_wide_lists<64>(va, thr, thrp, mylist, cnt, n0, asc, c0);
          else if (P.debug & 1024) process_wide_nf<K1W, 64>(va, thr, thrp, best, margin, mylist, cnt, P.cap, n0, asc, inv_asc, c0);
          else process_wide_k1<64>(va, thr, thrp, best, margin, mylist, cnt, P.cap, n0, asc, inv_asc, c0);
          if (more) {
            mbar_wait(bar(E.t_full0 + nbuf), ((tcount + 1) >> 1) & 1);
            tc_fence_after();
            tc_ldw<64>(tan, va);  // in flight while vb is tested
          }
          if (!K1W) process_wide_lists<64>(vb, thr, thrp, mylist, cnt, n0 + 64, asc, c0);
          else if (P.debug & 1024) process_wide_nf<K1W, 64>(vb, thr, thrp, best, margin, mylist, cnt, P.cap, n0 + 64, asc, inv_asc, c0);
          else process_wide_k1<64>(vb, thr, thrp, best, margin, mylist, cnt, P.cap, n0 + 64, asc, inv_asc, c0);
          if (more) {
            tc_ldw<64>(tan + 64, vb);
            tc_wait_ld();
            hand_back(nbuf);
          }
          // keep room for a full half tile of appends in every list of the warp
          unsigned need = __ballot_sync(0xffffffffu, !K1W && cnt > P.cap - HALF_N);
          while (need) {
            const int owner = __ffs(need) - 1;
            need &= need - 1;
            float2 *l = (float2 *)__shfl_sync(0xffffffffu, (unsigned long long)mylist, owner);
            const int n = __shfl_sync(0xffffffffu, cnt, owner);
            __syncwarp();
            const float nt = warp_select_compact(l, n, P.kprime, hist);
            if (lane == owner) {
              thr = nt;
              cnt = P.kprime;
            }
          }
          thrp = __fmul_rn(thr - c0, inv_asc);
        }
      }
      for (int jt = jt0; !(HACC && MODE == EPI_LISTS) && !(PIPE && piped) && jt < jt1; jt++, tcount++) {
        const uint32_t buf = tcount & 1, slot = tcount % NBN;
        if (EPI_TEAMS > 1 && (int)buf != team) continue;  // the other team's accumulator buffer
        bool handed_back = false;
        if (clk_on) ck_a = clock64();
        if (!NOBN) mbar_wait(bar(E.n_full0 + slot), (tcount / NBN) & 1);  // (no |b|^2 tiles otherwise)
        if (P.debug & 256) mbar_wait_spin(bar(E.t_full0 + buf), (tcount >> 1) & 1);
        else mbar_wait(bar(E.t_full0 + buf), (tcount >> 1) & 1);
        tc_fence_after();
        if (clk_on) {
          ck_b = clock64();
          ck_wait += ck_b - ck_a;
        }
        const float *bn = (const float *)(smem + Smem::bn_off + slot * TN * 4) + half * HALF_N;
        const int n0 = jt * P.tile_stride * TN + half * HALF_N + P.id0;
        if (MODE == EPI_DUMP) {
#pragma unroll 1
          for (int g = 0; g < HALF_N / 32; g++) {
            uint32_t v[32];
            tc_ld32(lane_addr + buf * TN + g * 32, v);
            tc_wait_ld();
            if (valid) {
              const long col0 = (long)jt * TN + half * HALF_N + g * 32;  // logical column
              float *drow = P.dump + (size_t)q * P.dump_ld + col0;
              if (col0 + 32 <= P.dump_ld && (P.dump_ld & 3) == 0) {
#pragma unroll
                for (int c4 = 0; c4 < 8; c4++) {  // 128 contiguous bytes per thread
                  float4 o;
                  o.x = fmaf(__uint_as_float(v[c4 * 4 + 0]), asc, NF ? c0 : bn[g * 32 + c4 * 4 + 0]);
                  o.y = fmaf(__uint_as_float(v[c4 * 4 + 1]), asc, NF ? c0 : bn[g * 32 + c4 * 4 + 1]);
                  o.z = fmaf(__uint_as_float(v[c4 * 4 + 2]), asc, NF ? c0 : bn[g * 32 + c4 * 4 + 2]);
                  o.w = fmaf(__uint_as_float(v[c4 * 4 + 3]), asc, NF ? c0 : bn[g * 32 + c4 * 4 + 3]);
                  *reinterpret_cast<float4 *>(drow + c4 * 4) = o;
                }
              } else {
#pragma unroll
                for (int c = 0; c < 32; c++)
                  if (col0 + c < P.dump_ld)
                    drow[c] = fmaf(__uint_as_float(v[c]), asc, NF ? c0 : bn[g * 32 + c]);
              }
            }
          }
        } else if (MODE == EPI_CROSS && P.cross_tma) {
          // compute_cross_distances (yael/nn.c:100-129): dist2[query + ld * row].  The tile leaves
          // through shared memory: the four warps of a column half (128 queries) stage 32 database
          // rows at a time as [32][128] floats -- thread = query, so a warp writes 128 contiguous
          // bytes per row, no bank conflicts -- and one elected thread hands the 16 KB to a TMA
          // store: 512 contiguous bytes per output row and no LSU store instructions.  (Storing
          // straight from the TMEM-lane layout is 128 bytes per warp instruction, 1024 instructions
          // per tile: measured 4.2 TB/s for that pattern alone against 5.4-6.0 TB/s for 512-byte
          // pieces, and the tile's stores took 10.9 k of its 12.1 k cycles.)
          const uint32_t ta = lane_addr + buf * TN;
          const int row0 = jt * TN + half * HALF_N;
          const uint32_t stg_off = half ? P.cross_stage1 : P.cross_stage0;
          const bool el = quarter == 0 && lane == 0;
          uint32_t va[32], vb[32];
          // 8 database rows per step through a ring of three 4 KB buffers, ONE barrier per step: at
          // step j the elected thread waits until the stores up to j-2 have read their buffers
          // before it joins the barrier, so whoever passes it may write buffer (j+1) % 3 next
          auto stage_and_store = [&](const uint32_t (&v)[32], int g) {
#pragma unroll
            for (int s8 = 0; s8 < 4; s8++) {
              float *stg = (float *)(smem + stg_off + cross_buf * CROSS_STAGE_BYTES);
#pragma unroll
              for (int c = 0; c < 8; c++) stg[c * TM + t] = __fmul_rn(asc, __uint_as_float(v[s8 * 8 + c]));
              fence_proxy_async_smem();
              if (el) bulk_wait_read1();
              named_bar_sync(1 + half, 128);
              if (el && !(P.debug & 1)) {
                tma_store_2d(E.map_out, sbase + stg_off + cross_buf * CROSS_STAGE_BYTES, qt * TM,
                             row0 + g * 32 + s8 * 8);
                bulk_commit();
              }
              cross_buf = cross_buf == 2 ? 0 : cross_buf + 1;
            }
          };
          tc_ld32(ta, va);
#pragma unroll 1
          for (int g = 0; g < HALF_N / 32; g += 2) {
            tc_wait_ld();
            tc_ld32(ta + (g + 1) * 32, vb);
            stage_and_store(va, g);
            tc_wait_ld();
            if (g + 2 < HALF_N / 32) {
              tc_ld32(ta + (g + 2) * 32, va);
            } else {  // the accumulator is in registers / on its way out: hand the buffer back
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                const uint32_t te = buf ? E.t_empty_addr1 : E.t_empty_addr0;
                if (E.t_empty_remote)
                  mbar_arrive_cluster(te);
                else
                  mbar_arrive(te);
              }
              handed_back = true;
            }
            stage_and_store(vb, g + 1);
          }
        } else if (MODE == EPI_CROSS) {
          // the same without staging (output pitch or address not 16-byte aligned, streamed query
          // chunks): the 32 lanes of a warp (32 consecutive queries) write 128 contiguous bytes per
          // database row.  The operands carry both norms (cross_l2_tensor, yb_knn.cu), so the value
          // is asc * acc.
          const uint32_t ta = lane_addr + buf * TN;
          const long col0 = (long)jt * TN + half * HALF_N;
          float *ocol = P.dump + (size_t)col0 * P.dump_ld + (valid ? q : 0);
          uint32_t va[32], vb[32];
          tc_ld32(ta, va);
#pragma unroll 1
          for (int g = 0; g < HALF_N / 32; g += 2) {
            tc_wait_ld();
            tc_ld32(ta + (g + 1) * 32, vb);
            if (valid && !(P.debug & 1)) {
#pragma unroll
              for (int c = 0; c < 32; c++)
                if (col0 + g * 32 + c < P.nb)
                  ocol[(size_t)(g * 32 + c) * P.dump_ld] = __fmul_rn(asc, __uint_as_float(va[c]));
            }
            tc_wait_ld();
            if (g + 2 < HALF_N / 32) tc_ld32(ta + (g + 2) * 32, va);
            if (valid && !(P.debug & 1)) {
#pragma unroll
              for (int c = 0; c < 32; c++)
                if (col0 + (g + 1) * 32 + c < P.nb)
                  ocol[(size_t)((g + 1) * 32 + c) * P.dump_ld] = __fmul_rn(asc, __uint_as_float(vb[c]));
            }
          }
        } else if (MODE == EPI_GMIN && NF && LDW == 128) {
          // sampling pass, early hand-back: the half tile comes into registers with two 64-column
          // loads and one wait, the buffer goes back at once, then the 8 group maxima are folded to
          // the requested group size (16 / 32 / 64 / 128 columns) and stored
          uint32_t va[64], vb[64];
          const uint32_t ta = lane_addr + buf * TN;
          tc_ldw<64>(ta, va);
          tc_ldw<64>(ta + 64, vb);
          tc_wait_ld();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            const uint32_t te = buf ? E.t_empty_addr1 : E.t_empty_addr0;
            if (E.t_empty_remote)
              mbar_arrive_cluster(te);
            else
              mbar_arrive(te);
          }
          handed_back = true;
          float g[8];
#pragma unroll
          for (int s4 = 0; s4 < 4; s4++) {
            g[s4] = group_max16(reinterpret_cast<const uint32_t(&)[16]>(va[16 * s4]));
            g[4 + s4] = group_max16(reinterpret_cast<const uint32_t(&)[16]>(vb[16 * s4]));
          }
          if (valid) {
            const int per_half = HALF_N / P.gsize;
            float *grow = P.gmin + (size_t)q * P.gmin_ld + ((long)jt * 2 + half) * per_half;
            // (a maximum of raw accumulators is a minimum of scores: asc < 0; NaN never wins fmaxf)
            if (P.gsize == 16 && (P.gmin_ld & 3) == 0) {
              // two 16-byte stores per thread and tile: the thread's 8 minima are one 32-byte sector
              // (as 8 scalar stores every warp instruction touched 32 sectors with 4 bytes each: 39 M
              // partial-sector writes per sampling pass, which paced it -- ncu: tensor pipe 30 %)
              float4 o0, o1;
              o0.x = fmaf(asc, g[0], c0); o0.y = fmaf(asc, g[1], c0); o0.z = fmaf(asc, g[2], c0); o0.w = fmaf(asc, g[3], c0);
              o1.x = fmaf(asc, g[4], c0); o1.y = fmaf(asc, g[5], c0); o1.z = fmaf(asc, g[6], c0); o1.w = fmaf(asc, g[7], c0);
              reinterpret_cast<float4 *>(grow)[0] = o0;
              reinterpret_cast<float4 *>(grow)[1] = o1;
            } else if (P.gsize == 16) {
#pragma unroll
              for (int i = 0; i < 8; i++) grow[i] = fmaf(asc, g[i], c0);
            } else if (P.gsize == 32) {
              float o[4];
#pragma unroll
              for (int i = 0; i < 4; i++) o[i] = fmaf(asc, fmaxf(g[2 * i], g[2 * i + 1]), c0);
              if ((P.gmin_ld & 3) == 0) {
                *reinterpret_cast<float4 *>(grow) = make_float4(o[0], o[1], o[2], o[3]);
              } else {
#pragma unroll
                for (int i = 0; i < 4; i++) grow[i] = o[i];
              }
            } else if (P.gsize == 64) {
              float o[2];
#pragma unroll
              for (int i = 0; i < 2; i++)
                o[i] = fmaf(asc, fmaxf(fmaxf(g[4 * i], g[4 * i + 1]), fmaxf(g[4 * i + 2], g[4 * i + 3])), c0);
              if ((P.gmin_ld & 1) == 0) {
                *reinterpret_cast<float2 *>(grow) = make_float2(o[0], o[1]);
              } else {
                grow[0] = o[0];
                grow[1] = o[1];
              }
            } else {
              const float a = fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3]));
              const float b = fmaxf(fmaxf(g[4], g[5]), fmaxf(g[6], g[7]));
              grow[0] = fmaf(asc, fmaxf(a, b), c0);
            }
          }
        } else if (MODE == EPI_GMIN) {
          // sampling pass: only the minimum of every column group leaves the SM
          uint32_t va[16], vb[16];
          const uint32_t ta = lane_addr + buf * TN;
          const int per_half = HALF_N / P.gsize;  // values this thread emits for the tile
          float *grow = P.gmin + (size_t)(valid ? q : 0) * P.gmin_ld + ((long)jt * 2 + half) * per_half;
          const int fold = P.gsize >> 4;          // 16-column groups per emitted value
          float gm = inf;
          tc_ld16(ta, va);
#pragma unroll 1
          for (int gg = 0; gg < 4; gg++) {
            tc_wait_ld();
            tc_ld16(ta + gg * 32 + 16, vb);
            gm = fminf(gm, NF ? fmaf(asc, group_max16(va), c0) : group_min16(va, bn + gg * 32, asc));
            if (((2 * gg + 1) % fold) == 0) {
              if (valid) grow[(2 * gg + 1) / fold - 1] = gm;
              gm = inf;
            }
            tc_wait_ld();
            if (gg < 3) tc_ld16(ta + gg * 32 + 32, va);
            gm = fminf(gm, NF ? fmaf(asc, group_max16(vb), c0) : group_min16(vb, bn + gg * 32 + 16, asc));
            if (((2 * gg + 2) % fold) == 0) {
              if (valid) grow[(2 * gg + 2) / fold - 1] = gm;
              gm = inf;
            }
          }
        } else if (MODE == EPI_HAMG) {
          // packed Hamming sampling pass: the smallest distance of every group of packed columns
          uint32_t va[16], vb[16];
          const uint32_t ta = lane_addr + buf * TN;
          const int per_half = HALF_N / P.gsize;  // values this thread emits for the tile
          float *grow = P.gmin + (size_t)(valid ? q : 0) * P.gmin_ld + ((long)jt * 2 + half) * per_half;
          const int fold = P.gsize >> 4;          // 16-column groups per emitted value
          float gm = inf;
          tc_ld16(ta, va);
#pragma unroll 1
          for (int gg = 0; gg < 4; gg++) {
            tc_wait_ld();
            tc_ld16(ta + gg * 32 + 16, vb);
            gm = fminf(gm, ham_group_min16(va, P.ham_magic, P.ham_slots));
            if (((2 * gg + 1) % fold) == 0) {
              if (valid) grow[(2 * gg + 1) / fold - 1] = gm;
              gm = inf;
            }
            tc_wait_ld();
            if (gg < 3) tc_ld16(ta + gg * 32 + 32, va);
            gm = fminf(gm, ham_group_min16(vb, P.ham_magic, P.ham_slots));
            if (((2 * gg + 2) % fold) == 0) {
              if (valid) grow[(2 * gg + 2) / fold - 1] = gm;
              gm = inf;
            }
          }
        } else if (MODE == EPI_HAMP && LDW == 128) {
          // packed Hamming pass, early hand-back (as the folded-norm k-NN epilogue below): the whole
          // half tile -- 128 accumulators = up to 384 distances per thread -- comes into registers
          // with two 64-column loads and ONE wait (a tcgen05.ld + wait::ld round trip costs a warp
          // ~230 cycles however many columns it brings: 8 of them per tile WERE the tile time), the
          // buffer goes back to the MMA issuer at once, the byte tests run in its shadow
          uint32_t va[64], vb[64];
          const uint32_t ta = lane_addr + buf * TN;
          const int r0 = (jt * P.tile_stride * TN + half * HALF_N) * P.ham_slots + P.id0;
          tc_ldw<64>(ta, va);
          tc_ldw<64>(ta + 64, vb);
          tc_wait_ld();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            const uint32_t te = buf ? E.t_empty_addr1 : E.t_empty_addr0;
            if (E.t_empty_remote)
              mbar_arrive_cluster(te);
            else
              mbar_arrive(te);
          }
          handed_back = true;
          if (!(P.debug & 1)) {
            process_wide_ham<64>(va, P.ham_magic, tau3, ham_mask, P.ham_slots, mylist, cnt, r0, P.ham_nb);
            process_wide_ham<64>(vb, P.ham_magic, tau3, ham_mask, P.ham_slots, mylist, cnt,
                                 r0 + 64 * P.ham_slots, P.ham_nb);
          }
        } else if (MODE == EPI_HAMP) {
          // packed Hamming pass: 8 groups of 16 accumulators (16 * ham_slots database rows each)
          uint32_t va[16], vb[16];
          const uint32_t ta = lane_addr + buf * TN;
          const int r0 = (jt * P.tile_stride * TN + half * HALF_N) * P.ham_slots + P.id0;
          tc_ld16(ta, va);
#pragma unroll 1
          for (int gg = 0; gg < 4; gg++) {
            tc_wait_ld();
            tc_ld16(ta + gg * 32 + 16, vb);
            if (!(P.debug & 1))
              process_group_ham(va, P.ham_magic, tau3, ham_mask, P.ham_slots, mylist, cnt,
                                r0 + gg * 32 * P.ham_slots, P.ham_nb);
            tc_wait_ld();
            if (gg < 3) tc_ld16(ta + gg * 32 + 32, va);
            if (!(P.debug & 1))
              process_group_ham(vb, P.ham_magic, tau3, ham_mask, P.ham_slots, mylist, cnt,
                                r0 + (gg * 32 + 16) * P.ham_slots, P.ham_nb);
          }
        } else if (NF && LDW == 128 && (MODE == EPI_LISTS || MODE == EPI_NEAREST)) {
          // The whole half tile (128 accumulators per thread) is pulled into registers with two
          // 64-column loads and ONE wait, the accumulator buffer is handed back to the MMA issuer
          // AT ONCE, and only then are the values looked at: the threshold tests overlap the MMAs
          // of the tile after next instead of sitting inside the buffer's ping-pong cycle
          // (MMA -> commit -> drain -> hand back -> MMA: with two buffers a tile costs half of
          // that cycle, and the hand-shake latencies alone are ~1350 of its ~3900 cycles).
          constexpr bool K1W = MODE == EPI_NEAREST;
          uint32_t va[64], vb[64];
          const uint32_t ta = lane_addr + buf * TN;
          tc_ldw<64>(ta, va);
          tc_ldw<64>(ta + 64, vb);
          tc_wait_ld();
          if (clk_on) {
            ck_a = clock64();
            ck_work += ck_a - ck_b;
            ck_b = ck_a;
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            const uint32_t te = buf ? E.t_empty_addr1 : E.t_empty_addr0;
            if (E.t_empty_remote)
              mbar_arrive_cluster(te);
            else
              mbar_arrive(te);
          }
          handed_back = true;
          if (!(P.debug & 1)) {
            if (!K1W && !(P.debug & 1024)) {  // (bit 1024: the out-of-line admission path only, A/B)
              process_wide_lists<64>(va, thr, thrp, mylist, cnt, n0, asc, c0);
              process_wide_lists<64>(vb, thr, thrp, mylist, cnt, n0 + 64, asc, c0);
            } else {
              process_wide_nf<K1W, 64>(va, thr, thrp, best, margin, mylist, cnt, P.cap, n0, asc, inv_asc, c0);
              process_wide_nf<K1W, 64>(vb, thr, thrp, best, margin, mylist, cnt, P.cap, n0 + 64, asc, inv_asc, c0);
            }
          }
        } else if (NF && LDW > 16 && (MODE == EPI_LISTS || MODE == EPI_NEAREST)) {
          // HALF_N / LDW wide loads, load g+1 in flight while g is processed
          constexpr int LW = (LDW == 32 || LDW == 64) ? LDW : 32;  // (LDW = 128 never gets here)
          if (P.debug & 4) {  // bring-up: TMEM traffic only, LDW columns per load
            uint32_t va[LW];
            const uint32_t ta = lane_addr + buf * TN;
#pragma unroll 1
            for (int g = 0; g < HALF_N / LW; g++) {
              tc_ldw<LW>(ta + g * LW, va);
              tc_wait_ld();
            }
          } else if (!(P.debug & 1)) {
            constexpr int NG = HALF_N / LW;
            constexpr bool K1W = MODE == EPI_NEAREST;
            uint32_t va[LW], vb[LW];
            const uint32_t ta = lane_addr + buf * TN;
            tc_ldw<LW>(ta, va);
#pragma unroll 1
            for (int g = 0; g < NG; g += 2) {
              tc_wait_ld();
              tc_ldw<LW>(ta + (g + 1) * LW, vb);
              process_wide_nf<K1W, LW>(va, thr, thrp, best, margin, mylist, cnt, P.cap, n0 + g * LW, asc,
                                        inv_asc, c0);
              tc_wait_ld();
              if (g + 2 < NG) tc_ldw<LW>(ta + (g + 2) * LW, va);
              process_wide_nf<K1W, LW>(vb, thr, thrp, best, margin, mylist, cnt, P.cap, n0 + (g + 1) * LW,
                                        asc, inv_asc, c0);
            }
          }
        } else if (!(P.debug & 1)) {
          // 8 groups of 16 columns, the TMEM load of group g+1 in flight while g is processed
          uint32_t va[16], vb[16];
          const uint32_t ta = lane_addr + buf * TN;
          if (P.debug & 8) {  // bring-up: math on whatever the registers hold, no TMEM traffic
#pragma unroll
            for (int c = 0; c < 16; c++) va[c] = vb[c] = 0x3f800000u + c + jt;
            process_group<false>(va, bn, thr, best, margin, mylist, cnt, P.cap, n0, asc);
            process_group<false>(vb, bn + 16, thr, best, margin, mylist, cnt, P.cap, n0 + 16, asc);
          } else if (P.debug & 4) {  // bring-up: TMEM traffic only
#pragma unroll 1
            for (int g = 0; g < 8; g++) {
              tc_ld16(ta + g * 16, va);
              tc_wait_ld();
            }
          } else {
#define YB_TILE_GROUPS(K1FLAG)                                                                  \
  tc_ld16(ta, va);                                                                              \
  _Pragma("unroll 1") for (int gg = 0; gg < 4; gg++) {                                          \
    tc_wait_ld();                                                                               \
    tc_ld16(ta + gg * 32 + 16, vb);                                                             \
    if (NF)                                                                                     \
      process_group_nf<K1FLAG>(va, thr, thrp, best, margin, mylist, cnt, P.cap, n0 + gg * 32,   \
                               asc, inv_asc, c0);                                               \
    else                                                                                        \
      process_group<K1FLAG>(va, bn + gg * 32, thr, best, margin, mylist, cnt, P.cap,            \
                            n0 + gg * 32, asc);                                                 \
    tc_wait_ld();                                                                               \
    if (gg < 3) tc_ld16(ta + gg * 32 + 32, va);                                                 \
    if (NF)                                                                                     \
      process_group_nf<K1FLAG>(vb, thr, thrp, best, margin, mylist, cnt, P.cap,                 \
                               n0 + gg * 32 + 16, asc, inv_asc, c0);                            \
    else                                                                                        \
      process_group<K1FLAG>(vb, bn + gg * 32 + 16, thr, best, margin, mylist, cnt, P.cap,       \
                            n0 + gg * 32 + 16, asc);                                            \
  }
            if (k1) {
              YB_TILE_GROUPS(true)
            } else {
              YB_TILE_GROUPS(false)
            }
#undef YB_TILE_GROUPS
          }
        }
        // accumulator buffer and |b|^2 slot are free again: ONE arrival per warp (256 per-thread
        // arrivals on the same barrier word serialise and cost more than the tile's math)
        if (clk_on) {
          ck_a = clock64();
          ck_work += ck_a - ck_b;
        }
        if (!handed_back) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            const uint32_t te = buf ? E.t_empty_addr1 : E.t_empty_addr0;
            if (E.t_empty_remote)
              mbar_arrive_cluster(te);
            else
              mbar_arrive(te);
            if (!NOBN) mbar_arrive(bar(E.n_empty0 + slot));
          }
        }
        // keep room for a full half tile of appends in every list of the warp
        // (a packed Hamming tile can append ham_slots entries per column)
        const int room = MODE == EPI_HAMP ? HALF_N * P.ham_slots : HALF_N;
        unsigned need = __ballot_sync(0xffffffffu, !k1 && cnt > P.cap - room);
        while (need) {
          const int owner = __ffs(need) - 1;
          need &= need - 1;
          float2 *l = (float2 *)__shfl_sync(0xffffffffu, (unsigned long long)mylist, owner);
          const int n = __shfl_sync(0xffffffffu, cnt, owner);
          __syncwarp();
          const float nt = warp_select_compact(l, n, P.kprime, hist);
          if (lane == owner) {
            thr = nt;
            cnt = P.kprime;
          }
        }
        if (MODE == EPI_HAMP) tau3 = ham_tau3(thr);
        if (NF) thrp = __fmul_rn(thr - c0, inv_asc);
        if (clk_on) ck_back += clock64() - ck_a;
      }
      if (MODE == EPI_NEAREST) {
        // k = 1: publish the candidates within the margin of the final best score
        if (valid) {
          const size_t l = (size_t)q * P.lists_ld + P.list0 + (sp * 2 + half) * EPI_TEAMS + team;
          const size_t o = l * P.kprime;
          int nout = 0;
          bool over = best != best;  // NaN: the list overflowed
          for (int e = 0; e < cnt; e++) {
            const float2 x = mylist[e];
            if (x.x < thr) {
              if (nout < P.kprime) {
                P.out_score[o + nout] = x.x;
                P.out_id[o + nout] = __float_as_int(x.y);
                nout++;
              } else {
                over = true;
              }
            }
          }
          for (int e = nout; e < P.kprime; e++) {
            P.out_score[o + e] = inf;
            P.out_id[o + e] = -1;
          }
          // out_thr: every unlisted row of this list's range scores >= thr; NaN = overflow
          P.out_thr[l] = over ? __uint_as_float(0x7fc00000u) : thr;
          if (P.out_cnt) P.out_cnt[l] = nout;
        }
      } else if (MODE == EPI_LISTS || MODE == EPI_HAMP) {
        // final compaction of over-full lists, then publish the shortlist of this item
        unsigned need = __ballot_sync(0xffffffffu, cnt > P.kprime);
        while (need) {
          const int owner = __ffs(need) - 1;
          need &= need - 1;
          float2 *l = (float2 *)__shfl_sync(0xffffffffu, (unsigned long long)mylist, owner);
          const int n = __shfl_sync(0xffffffffu, cnt, owner);
          __syncwarp();
          const float nt = warp_select_compact(l, n, P.kprime, hist);
          if (lane == owner) {
            thr = nt;
            cnt = P.kprime;
          }
        }
        __syncwarp();
        if (valid) {
          const size_t l = (size_t)q * P.lists_ld + P.list0 + (sp * 2 + half) * EPI_TEAMS + team;
          P.out_thr[l] = thr;
          if (P.out_cnt) P.out_cnt[l] = cnt;
        }
        // publish: the warp copies its 32 lists one after the other with coalesced accesses.
        // Only the valid entries are written -- the caller pre-fills the output with
        // (+inf, -1), and with tight thresholds a list holds a few dozen entries, not k'.
        for (int owner = 0; owner < 32; owner++) {
          const int n = __shfl_sync(0xffffffffu, valid ? cnt : 0, owner);
          if (n == 0) continue;
          const float2 *l = (const float2 *)__shfl_sync(0xffffffffu, (unsigned long long)mylist, owner);
          const int qo = __shfl_sync(0xffffffffu, q, owner);
          const size_t o = ((size_t)qo * P.lists_ld + P.list0 + (sp * 2 + half) * EPI_TEAMS + team) * P.kprime;
          for (int e = lane; e < n; e += 32) {
            const float2 x = l[e];
            P.out_score[o + e] = x.x;
            P.out_id[o + e] = __float_as_int(x.y);
          }
        }
      }
    }
    if (MODE == EPI_CROSS && P.cross_tma) bulk_wait0();  // staged tiles have left before the CTA does
    if (clk_on && lane == 0 && blockIdx.x < 160) {
      g_tf32_clk[blockIdx.x][4] = ck_wait;
      g_tf32_clk[blockIdx.x][5] = ck_work;
      g_tf32_clk[blockIdx.x][6] = ck_back;
      g_tf32_clk[blockIdx.x][7] = clock64() - ck_t0;
      g_tf32_clk[blockIdx.x][8] = tcount;
    }
  }
}

// ------------------------------------------------------------------ the kernel
template <int MODE, int KIND, int LDW>
__global__ void __launch_bounds__(TF32_THREADS, 1)
k_knn_tf32(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_b,
           const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_qx,
           const __grid_constant__ CUtensorMap map_bx, const Tf32Params P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto bar = [&](int i) { return sbase + Smem::bar_off + 8 * i; };
  // the extras-chunk machinery exists only in the folded-norm instantiations: the producer and
  // issuer loops of every other kind compile exactly as they did without it (the packed Hamming
  // pass lost 14 % to the mere presence of the run-time checks)
  constexpr bool NFK = KIND == OP_F16N;
  const int xk = NFK ? P.xk : -1;
  const bool xring = NFK && P.xring != 0;
  volatile uint32_t *tmem_ptr_smem = (volatile uint32_t *)(smem + Smem::tmem_ptr_off);

  if (threadIdx.x == 0) {
    mbar_init(bar(Smem::a_full), 1);
    mbar_init(bar(Smem::a_empty), 1);
    for (int i = 0; i < STAGES; i++) {
      mbar_init(bar(Smem::b_full + i), 1);
      mbar_init(bar(Smem::b_empty + i), P.pair ? 2 : 1);  // paired: both CTAs' MMAs must be done
    }
    for (int i = 0; i < NBN; i++) {
      mbar_init(bar(Smem::n_full + i), 1);
      mbar_init(bar(Smem::n_empty + i), TEAM_WARPS);  // a |b|^2 slot belongs to one tile, i.e. one team
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(bar(Smem::x_full + i), 1);
      mbar_init(bar(Smem::x_empty + i), 1);
      mbar_init(bar(Smem::t_full + i), 1);
      mbar_init(bar(Smem::t_empty + i), TEAM_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == EPI_WARPS + 1) {  // TMEM: all 512 columns (2 accumulator buffers of 256)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     sbase + Smem::tmem_ptr_off),
                 "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  uint32_t crank = 0;
  if (P.pair) {
    crank = cluster_ctarank();
    cluster_sync_all();  // the peer's barriers are initialised before anything is multicast
  }

  // work items: (range, query tile) -- or (range, query-tile pair) for paired CTAs, the two CTAs
  // of a cluster taking the two tiles of the pair and walking the same database range
  const int first_item = P.pair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int item_step = P.pair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int tq_div = P.pair ? P.tiles_q2 : P.tiles_q;
  EpiCtx ectx;
  ectx.smem = smem; ectx.sbase = sbase; ectx.tmem_base = tmem_base; ectx.warp = warp; ectx.lane = lane;
  ectx.first_item = first_item; ectx.item_step = item_step; ectx.tq_div = tq_div;
  ectx.pair = P.pair; ectx.crank = crank;
  ectx.t_empty_addr0 = bar(Smem::t_empty + 0); ectx.t_empty_addr1 = bar(Smem::t_empty + 1);
  ectx.t_empty_remote = 0;
  ectx.n_full0 = Smem::n_full; ectx.n_empty0 = Smem::n_empty; ectx.t_full0 = Smem::t_full;
  ectx.map_out = nullptr;

  if (warp == EPI_WARPS) {
    // ======================================================================== TMA producer
    regs_aux();
    constexpr int KCE = (KIND == OP_F8 || KIND == OP_F8C || KIND == OP_F8H) ? KC * 4 : ((KIND == OP_F16 || KIND == OP_F16N) ? KC * 2 : KC);  // elements per 128-byte K chunk (TMA coordinates)
    if (lane == 0) {
      uint32_t icount = 0, ccount = 0, tcount = 0;
      for (int item = first_item; item < P.items; item += item_step, icount++) {
        const int sp = P.order ? item % P.splits : item / tq_div;
        const int qi = P.order ? item / P.splits : item - sp * tq_div;
        const int qt = P.pair ? qi * 2 + (int)crank : qi;
        const int jt0 = sp * P.range_tiles, jt1 = min(P.nbt, jt0 + P.range_tiles);
        // query tile: wait until the MMAs of the previous item have drained A
        mbar_wait(bar(Smem::a_empty), (icount & 1) ^ 1);
        const int nkd = xk >= 0 ? xk : P.nkc;  // 128-byte data chunks
        mbar_expect_tx(bar(Smem::a_full),
                       (uint32_t)(nkd * A_CHUNK_BYTES + (xk >= 0 ? TM * 32 : 0)));
        for (int kc = 0; kc < nkd; kc++)
          tma_load_2d(sbase + Smem::a_off + kc * A_CHUNK_BYTES, &map_q, bar(Smem::a_full), kc * KCE,
                      qt * TM);
        if (xk >= 0)
          tma_load_2d(sbase + Smem::a_off + (xk < 0 ? 0 : xk) * A_CHUNK_BYTES, &map_qx, bar(Smem::a_full), P.xcol,
                      qt * TM);
        for (int jt = jt0; jt < jt1; jt++, tcount++) {
          const int jta = jt * P.tile_stride;  // actual database tile
          if (!(NFK || KIND == OP_F8C || KIND == OP_F8H || MODE == EPI_HAMP || MODE == EPI_HAMG)) {  // |b|^2 of the tile (the folded-norm kinds carry it in the operands, the packed Hamming modes do not use it)
            const uint32_t slot = tcount % NBN;
            mbar_wait(bar(Smem::n_empty + slot), ((tcount / NBN) & 1) ^ 1);
            mbar_expect_tx(bar(Smem::n_full + slot), TN * 4);
            bulk_load_1d(sbase + Smem::bn_off + slot * TN * 4, P.bnorm + (size_t)jta * TN, TN * 4,
                         bar(Smem::n_full + slot));
          }
          if (P.debug & 64) continue;  // bring-up: no operand rings at all (with 16: MMAs on stale data)
          if (xring) {  // extras of this tile: own 2-slot ring
            const uint32_t xs = tcount & 1;
            mbar_wait(bar(Smem::x_empty + xs), ((tcount >> 1) & 1) ^ 1);
            mbar_expect_tx(bar(Smem::x_full + xs), Smem::XB_BYTES);
            tma_load_2d(sbase + Smem::xb_off + xs * Smem::XB_BYTES, &map_bx, bar(Smem::x_full + xs),
                        P.xcol, jta * TN);
          }
          const int nring = xring ? xk : P.nkc;  // chunks of this tile that use the B ring
          for (int kc = 0; kc < nring; kc++, ccount++) {
            const uint32_t st = ccount % STAGES;
            mbar_wait(bar(Smem::b_empty + st), ((ccount / STAGES) & 1) ^ 1);
            if ((P.debug & 16) && jt > jt0) {  // bring-up: no TMA traffic after the first tile
              mbar_arrive(bar(Smem::b_full + st));
              continue;
            }
            if (kc == xk) {  // the 16 extra K elements: a 32-byte-wide box
              mbar_expect_tx(bar(Smem::b_full + st), TN * 32);
              tma_load_2d(sbase + Smem::b_off + st * B_CHUNK_BYTES, &map_bx, bar(Smem::b_full + st),
                          P.xcol, jta * TN);
              continue;
            }
            mbar_expect_tx(bar(Smem::b_full + st), B_CHUNK_BYTES);
            if (P.pair) {
              // my half of the chunk (128 rows), delivered to both CTAs of the cluster
              tma_load_2d_mc(sbase + Smem::b_off + st * B_CHUNK_BYTES + crank * (B_CHUNK_BYTES / 2),
                             &map_bh, bar(Smem::b_full + st), kc * KCE,
                             jta * TN + (int)crank * (TN / 2), (uint16_t)3);
            } else {
              tma_load_2d(sbase + Smem::b_off + st * B_CHUNK_BYTES, &map_b, bar(Smem::b_full + st),
                          kc * KCE, jta * TN);
            }
          }
        }
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    // ======================================================================== MMA issuer
    // every lane walks the loop; elect.sync inside the tcgen05 wrappers picks the issuing lane
    regs_aux();
    {
      uint32_t icount = 0, ccount = 0, tcount = 0;
      const bool skip_mma = (P.debug & 2) != 0;
      const bool ring_aligned = P.nkc == STAGES && P.last_k8 == 4 && xk < 0;
      const bool clk_on = (P.debug & 512) != 0;
      long long ck_acc = 0, ck_ops = 0, ck_x = 0, ck_t0 = clk_on ? clock64() : 0, ck_a = 0;
      for (int item = first_item; item < P.items; item += item_step, icount++) {
        const int sp = P.order ? item % P.splits : item / tq_div;
        const int jt0 = sp * P.range_tiles, jt1 = min(P.nbt, jt0 + P.range_tiles);
        mbar_wait(bar(Smem::a_full), icount & 1);
        for (int jt = jt0; jt < jt1; jt++, tcount++) {
          const uint32_t buf = tcount & 1;
          if (clk_on) ck_a = clock64();
          if (P.debug & 256) mbar_wait_spin(bar(Smem::t_empty + buf), ((tcount >> 1) & 1) ^ 1);
          else mbar_wait(bar(Smem::t_empty + buf), ((tcount >> 1) & 1) ^ 1);
          tc_fence_after();
          if (clk_on) ck_acc += clock64() - ck_a;
          const uint32_t d_tmem = tmem_base + buf * TN;
          if (ring_aligned && !skip_mma) {
            // d = 128 (K chunks == ring depth): chunk kc always sits in stage kc, so every
            // descriptor below is loop invariant and the body is wait, 4 MMAs, commit
#pragma unroll
            for (int kc = 0; kc < STAGES; kc++) {
              mbar_wait(bar(Smem::b_full + kc), tcount & 1);
              tc_fence_after();
              const uint64_t adesc = smem_desc_sw128(sbase + Smem::a_off + kc * A_CHUNK_BYTES);
              const uint64_t bdesc = smem_desc_sw128(sbase + Smem::b_off + kc * B_CHUNK_BYTES);
              tc_mma_elect<KIND>(d_tmem, adesc, bdesc, kc != 0);
              tc_mma_elect<KIND>(d_tmem, adesc + 2, bdesc + 2, 1);
              tc_mma_elect<KIND>(d_tmem, adesc + 4, bdesc + 4, 1);
              tc_mma_elect<KIND>(d_tmem, adesc + 6, bdesc + 6, 1);
              if (P.pair)
                tc_commit_mc_elect(bar(Smem::b_empty + kc), (uint16_t)3);
              else
                tc_commit_elect(bar(Smem::b_empty + kc));
            }
            ccount += STAGES;
          } else
          for (int kc = 0; kc < (xring ? xk : P.nkc); kc++, ccount++) {
            const uint32_t st = ccount % STAGES;
            if (clk_on) ck_a = clock64();
            if (!(P.debug & 64)) mbar_wait(bar(Smem::b_full + st), (ccount / STAGES) & 1);
            tc_fence_after();
            if (clk_on) ck_ops += clock64() - ck_a;
            const uint64_t adesc = smem_desc_sw128(sbase + Smem::a_off + kc * A_CHUNK_BYTES);
            const uint64_t bdesc = smem_desc_sw128(sbase + Smem::b_off + st * B_CHUNK_BYTES);
            // advancing K inside the 128-byte swizzle span: +32 bytes = +2 in 16-byte units
            if (kc == xk) {
              if (!skip_mma)
                tc_mma_elect<KIND>(d_tmem, smem_desc_sw32(sbase + Smem::a_off + kc * A_CHUNK_BYTES),
                                   smem_desc_sw32(sbase + Smem::b_off + st * B_CHUNK_BYTES), 1);
            } else if (!skip_mma) {
              const int last_data = (xk >= 0 ? xk : P.nkc) - 1;
              if (kc != last_data || P.last_k8 == 4) {
                const int reps = (P.debug & 32) ? 4 : 1;  // bring-up: marginal cost of an MMA
                for (int rep = 0; rep < reps; rep++) {
                  tc_mma_elect<KIND>(d_tmem, adesc, bdesc, kc != 0);
                  tc_mma_elect<KIND>(d_tmem, adesc + 2, bdesc + 2, 1);
                  tc_mma_elect<KIND>(d_tmem, adesc + 4, bdesc + 4, 1);
                  tc_mma_elect<KIND>(d_tmem, adesc + 6, bdesc + 6, 1);
                }
              } else {
                for (int k8 = 0; k8 < P.last_k8; k8++)
                  tc_mma_elect<KIND>(d_tmem, adesc + (uint64_t)(2 * k8), bdesc + (uint64_t)(2 * k8),
                                   (kc | k8) != 0);
              }
            }
            // smem slot reusable once these MMAs retire (paired: tell both CTAs, either may
            // multicast into the slot next)
            if (P.debug & 64) continue;
            if (P.pair)
              tc_commit_mc_elect(bar(Smem::b_empty + st), (uint16_t)3);
            else
              tc_commit_elect(bar(Smem::b_empty + st));
          }
          if (xring) {  // the norm term: one MMA over the 16 extra K elements
            const uint32_t xs = tcount & 1;
            if (clk_on) ck_a = clock64();
            if (!(P.debug & 64)) mbar_wait(bar(Smem::x_full + xs), (tcount >> 1) & 1);
            tc_fence_after();
            if (clk_on) ck_x += clock64() - ck_a;
            if (!skip_mma)
              tc_mma_elect<KIND>(d_tmem, smem_desc_sw32(sbase + Smem::a_off + (xk < 0 ? 0 : xk) * A_CHUNK_BYTES),
                                 smem_desc_sw32(sbase + Smem::xb_off + xs * Smem::XB_BYTES), 1);
            if (!(P.debug & 64)) tc_commit_elect(bar(Smem::x_empty + xs));
          }
          if (P.debug & 128) {  // bring-up (skeleton only): plain arrive instead of the commit
            if (lane == 0) mbar_arrive(bar(Smem::t_full + buf));
            __syncwarp();
          } else {
            tc_commit_elect(bar(Smem::t_full + buf));  // accumulator complete
          }
        }
        tc_commit_elect(bar(Smem::a_empty));  // query tile no longer needed
      }
      if (clk_on && lane == 0 && blockIdx.x < 160) {
        g_tf32_clk[blockIdx.x][0] = ck_acc;
        g_tf32_clk[blockIdx.x][1] = ck_ops;
        g_tf32_clk[blockIdx.x][2] = ck_x;
        g_tf32_clk[blockIdx.x][3] = clock64() - ck_t0;
      }
    }
  } else if (warp < EPI_WARPS) {
    regs_epilogue();
    run_epilogue<MODE, KIND, LDW>(P, ectx);
  } else {
    regs_aux();  // the two idle warps of the third warpgroup
  }

  tc_fence_before();
  __syncthreads();
  if (P.pair) cluster_sync_all();  // nobody leaves while the peer can still write into it
  if (warp == EPI_WARPS + 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------ the 2-SM kernel (folded-norm FP16)
// Same roles and the same epilogue as k_knn_tf32<MODE, OP_F16N>, but the two CTAs of a cluster form
// ONE MMA unit: tcgen05.mma.cta_group::2, M = 256 (each CTA's 128 queries) x N = 256, with HALF of
// every database chunk (128 rows) in each CTA's shared memory.  Why: a single-CTA SS-mode MMA reads
// 4 KB of A and 8 KB of B from shared memory per 128-cycle instruction while TMA writes the next
// 8 KB -- every database byte is written once and read once per tile, 144 KB + 36 KB of A per
// 1152-cycle tile, i.e. MORE than the 128 B/clk a shared memory delivers: with the folded-norm
// epilogue out of the way the 1-SM pass is paced by exactly that (bring-up switch 1, "TMA + MMA
// only": 2.23 of the 2.60 ms).  As a pair each SM holds, writes and reads only its half of B:
// 36 (A) + 36 (B reads) + 36 (B writes) KB per tile, 94 B/clk.
//   * only the leader CTA (cluster rank 0) issues MMAs; it waits on ITS full barriers, which both
//     CTAs' TMA loads complete (cta_group::2 loads can signal the leader's barrier)
//   * tcgen05.commit.cta_group::2 ... multicast releases the smem slots / publishes the
//     accumulators in both CTAs
//   * the peer's epilogue warps signal "accumulator drained" on the leader's barrier
constexpr int STAGES2 = 8;
// Wide resident layout: a query tile of up to MAX_NKC_WIDE chunks stays in shared memory and the
// database ring shrinks to what is left of the same 192 KB (12 - nka stages of 16 KB).  K = 3 d of
// the tensor-core compute_cross_distances at d = 128 is 7 chunks: resident, the queries are read
// once per work item instead of once per database tile (half the L2 -> SM operand traffic of the
// streamed variant, which is what bounded that kernel together with its 4 GB of output).
constexpr int MAX_NKC_WIDE = 7;
constexpr int B2_CHUNK_BYTES = (TN / 2) * KC * 4;  // 16 KB: this CTA's half of a 128-byte-wide chunk
constexpr int XB2_BYTES = (TN / 2) * 32;           // this CTA's half of the extras chunk
struct Smem2 {
  static constexpr int a_off = 0;
  static constexpr int b_off = MAX_NKC * A_CHUNK_BYTES;
  static constexpr int bn_off = b_off + STAGES2 * B2_CHUNK_BYTES;
  static constexpr int hist_off = bn_off + NBN * TN * 4;
  static constexpr int bar_off = hist_off + EPI_WARPS * 256 * 4;
  static constexpr int a_full = 0, a_empty = 1, b_full = 2, b_empty = b_full + STAGES2,
                       n_full = b_empty + STAGES2, n_empty = n_full + NBN,
                       t_full = n_empty + NBN, t_empty = t_full + 2, x_full = t_empty + 2,
                       x_empty = x_full + 2, nbar = x_empty + 2;
  static constexpr int xb_off = a_off + 3 * A_CHUNK_BYTES;  // 2-slot extras ring (as Smem::xb_off)
  static constexpr int tmem_ptr_off = bar_off + nbar * 8;
  static constexpr int total = tmem_ptr_off + 16;
};
static_assert(Smem2::bn_off == Smem::bn_off && Smem2::hist_off == Smem::hist_off,
              "run_epilogue() addresses |b|^2 tiles and histograms through Smem::");
static_assert(Smem2::bar_off == Smem::bar_off, "barrier block must sit at the same offset");
constexpr int TF32_SMEM2_BYTES = Smem2::total;
constexpr int CROSS_STAGE0_OFF = Smem2::bn_off;
constexpr int CROSS_STAGE1_OFF = (Smem2::total + 1023) & ~1023;
constexpr int TF32_SMEM2_CROSS_BYTES = CROSS_STAGE1_OFF + 3 * CROSS_STAGE_BYTES;
static_assert(3 * CROSS_STAGE_BYTES <= Smem2::bar_off - Smem2::bn_off, "half 0 staging overlays |b|^2 + histograms");
static_assert(TF32_SMEM2_CROSS_BYTES <= 232448, "shared memory per CTA");
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address

__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap *map,
                                                uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2sm_elect(uint32_t bar, uint16_t mask) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t"
      "}" ::"r"(bar),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tc_mma_f8_2sm_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                    uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma_f16_2sm_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                     uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D=F32, A=B=F16, K-major, N=256, M=256 (two CTAs x 128)
constexpr uint32_t IDESC_F16_2SM = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

// KIND: OP_F16N (k-NN / k-means) or OP_F8 in the packed Hamming modes (neither reads |b|^2 tiles)
template <int KIND>
__device__ __forceinline__ void tc_mma_2sm_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                 uint32_t accumulate) {
  if (KIND == OP_F8H)   // D = F16: c_format (bits 4-5 of the instruction descriptor) = 0
    tc_mma_f8_2sm_elect(d_tmem, a_desc, b_desc, IDESC_F16_2SM & ~(3u << 4), accumulate);
  else if (KIND == OP_F8 || KIND == OP_F8C)
    tc_mma_f8_2sm_elect(d_tmem, a_desc, b_desc, IDESC_F16_2SM, accumulate);
  else
    tc_mma_f16_2sm_elect(d_tmem, a_desc, b_desc, IDESC_F16_2SM, accumulate);
}

// STREAM: the query tile is NOT resident -- every ring stage carries a query chunk (16 KB) next to
// this CTA's half of the database chunk (16 KB), SST stages in the space of the resident query tile
// plus the B ring.  Lifts the d <= 240 limit of the resident layout (any d; the price is the query
// chunks re-read from L2 for every database tile, +50 % operand traffic).
constexpr int SST = 6;
constexpr int SSTAGE_BYTES = A_CHUNK_BYTES + B2_CHUNK_BYTES;  // 32 KB
static_assert(SST * SSTAGE_BYTES <= Smem2::bn_off && SST <= STAGES2, "streamed stages overlay a_off..bn_off");

template <int MODE, int LDW, int KIND = OP_F16N, bool STREAM = false>
__global__ void __launch_bounds__(TF32_THREADS, 1)
k_knn_2sm(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_bh,
          const __grid_constant__ CUtensorMap map_qx, const __grid_constant__ CUtensorMap map_bxh,
          const __grid_constant__ CUtensorMap map_out, const Tf32Params P) {
  static_assert(KIND == OP_F16N || KIND == OP_F8C || KIND == OP_F8H || (KIND == OP_F8 && (MODE == EPI_HAMP || MODE == EPI_HAMG)),
                "the 2-SM kernel carries no |b|^2 ring");
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto bar = [&](int i) { return sbase + Smem2::bar_off + 8 * i; };
  volatile uint32_t *tmem_ptr_smem = (volatile uint32_t *)(smem + Smem2::tmem_ptr_off);
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const int xk = KIND == OP_F16N ? P.xk : -1;  // index of the extras chunk (-1: the extras sit inside the last data chunk)
  const bool xring = KIND == OP_F16N && !STREAM && P.xring != 0;  // extras travel through their own 2-slot ring
  constexpr int KCE = (KIND == OP_F8 || KIND == OP_F8C || KIND == OP_F8H) ? KC * 4 : KC * 2;  // elements per 128-byte K chunk

  if (threadIdx.x == 0) {
    mbar_init(bar(Smem2::a_full), 1);
    mbar_init(bar(Smem2::a_empty), 1);
    for (int i = 0; i < STAGES2; i++) {
      mbar_init(bar(Smem2::b_full + i), 1);
      mbar_init(bar(Smem2::b_empty + i), 1);
    }
    for (int i = 0; i < NBN; i++) {
      mbar_init(bar(Smem2::n_full + i), 1);
      mbar_init(bar(Smem2::n_empty + i), TEAM_WARPS);
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(bar(Smem2::x_full + i), 1);
      mbar_init(bar(Smem2::x_empty + i), 1);
      mbar_init(bar(Smem2::t_full + i), 1);
      mbar_init(bar(Smem2::t_empty + i), 2 * TEAM_WARPS);  // both CTAs' epilogue warps (leader's copy is used)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == EPI_WARPS + 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     sbase + Smem2::tmem_ptr_off),
                 "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // barriers of both CTAs initialised, TMEM allocated in both
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int first_item = (int)(blockIdx.x >> 1), item_step = (int)(gridDim.x >> 1);
  const int tq_div = P.tiles_q2;
  const int nkd = xk >= 0 ? xk : P.nkc;        // 128-byte data chunks per row
  const int nring = xring ? xk : P.nkc;        // chunks of a tile that travel through the B ring
  const uint32_t b_off = (uint32_t)P.nka * A_CHUNK_BYTES;   // resident layout: the ring starts behind the query tile
  const uint32_t nst = (uint32_t)P.stages2;                 // ... and has this many stages (8 for nka = 4)

  if (warp == EPI_WARPS) {
    // ======================================================================== TMA producer
    regs_aux();
    if (lane == 0) {
      uint32_t icount = 0, ccount = 0, tcount = 0;
      for (int item = first_item; item < P.items; item += item_step, icount++) {
        const int sp = item / tq_div;
        const int qt = (item - sp * tq_div) * 2 + (int)crank;
        const int jt0 = sp * P.range_tiles, jt1 = min(P.nbt, jt0 + P.range_tiles);
        if (STREAM) {
          for (int jt = jt0; jt < jt1; jt++) {
            const int row0 = jt * P.tile_stride * TN + (int)crank * (TN / 2);
            for (int kc = 0; kc < P.nkc; kc++, ccount++) {
              const uint32_t st = ccount % SST;
              mbar_wait(bar(Smem2::b_empty + st), ((ccount / SST) & 1) ^ 1);
              const uint32_t a_dst = sbase + st * SSTAGE_BYTES, b_dst = a_dst + A_CHUNK_BYTES;
              const uint32_t fb = bar(Smem2::b_full + st) & PEER_BIT_MASK;
              if (kc == xk) {  // the 16 extra K elements of both operands: 32-byte-wide boxes
                if (leader) mbar_expect_tx(bar(Smem2::b_full + st), 2 * (TM * 32 + XB2_BYTES));
                tma_load_2d_2sm(a_dst, &map_qx, fb, P.xcol, qt * TM);
                tma_load_2d_2sm(b_dst, &map_bxh, fb, P.xcol, row0);
              } else {
                if (leader) mbar_expect_tx(bar(Smem2::b_full + st), 2 * (A_CHUNK_BYTES + B2_CHUNK_BYTES));
                tma_load_2d_2sm(a_dst, &map_q, fb, kc * KCE, qt * TM);
                tma_load_2d_2sm(b_dst, &map_bh, fb, kc * KCE, row0);
              }
            }
          }
          continue;
        }
        mbar_wait(bar(Smem2::a_empty), (icount & 1) ^ 1);
        if (leader)  // the leader's barrier collects the bytes of BOTH CTAs
          mbar_expect_tx(bar(Smem2::a_full), (uint32_t)(2 * (nkd * A_CHUNK_BYTES + (xk >= 0 ? TM * 32 : 0))));
        for (int kc = 0; kc < nkd; kc++)
          tma_load_2d_2sm(sbase + Smem2::a_off + kc * A_CHUNK_BYTES, &map_q,
                          bar(Smem2::a_full) & PEER_BIT_MASK, kc * KCE, qt * TM);
        if (xk >= 0)
          tma_load_2d_2sm(sbase + Smem2::a_off + xk * A_CHUNK_BYTES, &map_qx,
                          bar(Smem2::a_full) & PEER_BIT_MASK, P.xcol, qt * TM);
        for (int jt = jt0; jt < jt1; jt++, tcount++) {
          const int jta = jt * P.tile_stride;  // (no |b|^2 tiles: the operands carry the norms)
          const int row0 = jta * TN + (int)crank * (TN / 2);  // this CTA's half of the tile
          if (xring) {
            const uint32_t xs = tcount & 1;
            mbar_wait(bar(Smem2::x_empty + xs), ((tcount >> 1) & 1) ^ 1);
            if (leader) mbar_expect_tx(bar(Smem2::x_full + xs), 2 * XB2_BYTES);
            tma_load_2d_2sm(sbase + Smem2::xb_off + xs * XB2_BYTES, &map_bxh,
                            bar(Smem2::x_full + xs) & PEER_BIT_MASK, P.xcol, row0);
          }
          for (int kc = 0; kc < nring; kc++, ccount++) {
            const uint32_t st = ccount % nst;
            mbar_wait(bar(Smem2::b_empty + st), ((ccount / nst) & 1) ^ 1);
            if (kc == xk) {  // the 16 extra K elements as a 32-byte-wide box in a ring stage
              if (leader) mbar_expect_tx(bar(Smem2::b_full + st), 2 * XB2_BYTES);
              tma_load_2d_2sm(sbase + b_off + st * B2_CHUNK_BYTES, &map_bxh,
                              bar(Smem2::b_full + st) & PEER_BIT_MASK, P.xcol, row0);
              continue;
            }
            if (leader) mbar_expect_tx(bar(Smem2::b_full + st), 2 * B2_CHUNK_BYTES);
            tma_load_2d_2sm(sbase + b_off + st * B2_CHUNK_BYTES, &map_bh,
                            bar(Smem2::b_full + st) & PEER_BIT_MASK, kc * KCE, row0);
          }
        }
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    // ======================================================================== MMA issuer (leader CTA)
    regs_aux();
    if (leader) {
      uint32_t icount = 0, ccount = 0, tcount = 0;
      const int last_data = nkd - 1;
      const bool clk_on = (P.debug & 512) != 0;
      long long ck_acc = 0, ck_ops = 0, ck_x = 0, ck_t0 = clk_on ? clock64() : 0, ck_a = 0;
      for (int item = first_item; item < P.items; item += item_step, icount++) {
        const int sp = item / tq_div;
        const int jt0 = sp * P.range_tiles, jt1 = min(P.nbt, jt0 + P.range_tiles);
        if (!STREAM) mbar_wait(bar(Smem2::a_full), icount & 1);
        for (int jt = jt0; jt < jt1; jt++, tcount++) {
          const uint32_t buf = tcount & 1;
          if (clk_on) ck_a = clock64();
          mbar_wait(bar(Smem2::t_empty + buf), ((tcount >> 1) & 1) ^ 1);
          tc_fence_after();
          if (clk_on) ck_acc += clock64() - ck_a;
          const uint32_t d_tmem = tmem_base + buf * TN;
          if (STREAM) {
            for (int kc = 0; kc < P.nkc; kc++, ccount++) {
              const uint32_t st = ccount % SST;
              if (clk_on) ck_a = clock64();
              mbar_wait(bar(Smem2::b_full + st), (ccount / SST) & 1);
              tc_fence_after();
              if (clk_on) ck_ops += clock64() - ck_a;
              const uint32_t a_src = sbase + st * SSTAGE_BYTES, b_src = a_src + A_CHUNK_BYTES;
              if (kc == xk) {
                tc_mma_2sm_elect<KIND>(d_tmem, smem_desc_sw32(a_src), smem_desc_sw32(b_src), 1);
              } else {
                const uint64_t adesc = smem_desc_sw128(a_src), bdesc = smem_desc_sw128(b_src);
                const int steps = (kc != last_data || P.last_k8 == 4) ? 4 : P.last_k8;
                for (int k8 = 0; k8 < steps; k8++)
                  tc_mma_2sm_elect<KIND>(d_tmem, adesc + (uint64_t)(2 * k8), bdesc + (uint64_t)(2 * k8),
                                         (kc | k8) != 0);
              }
              tc_commit_2sm_elect(bar(Smem2::b_empty + st), (uint16_t)3);
            }
            tc_commit_2sm_elect(bar(Smem2::t_full + buf), (uint16_t)3);
            continue;
          }
          for (int kc = 0; kc < nring; kc++, ccount++) {
            const uint32_t st = ccount % nst;
            if (clk_on) ck_a = clock64();
            mbar_wait(bar(Smem2::b_full + st), (ccount / nst) & 1);
            tc_fence_after();
            if (clk_on) ck_ops += clock64() - ck_a;
            const uint64_t adesc = smem_desc_sw128(sbase + Smem2::a_off + kc * A_CHUNK_BYTES);
            const uint64_t bdesc = smem_desc_sw128(sbase + b_off + st * B2_CHUNK_BYTES);
            if (kc == xk) {
              tc_mma_2sm_elect<KIND>(d_tmem, smem_desc_sw32(sbase + Smem2::a_off + kc * A_CHUNK_BYTES),
                                     smem_desc_sw32(sbase + b_off + st * B2_CHUNK_BYTES), 1);
            } else if (kc != last_data || P.last_k8 == 4) {
              tc_mma_2sm_elect<KIND>(d_tmem, adesc, bdesc, kc != 0);
              tc_mma_2sm_elect<KIND>(d_tmem, adesc + 2, bdesc + 2, 1);
              tc_mma_2sm_elect<KIND>(d_tmem, adesc + 4, bdesc + 4, 1);
              tc_mma_2sm_elect<KIND>(d_tmem, adesc + 6, bdesc + 6, 1);
            } else {
              for (int k8 = 0; k8 < P.last_k8; k8++)
                tc_mma_2sm_elect<KIND>(d_tmem, adesc + (uint64_t)(2 * k8), bdesc + (uint64_t)(2 * k8),
                                       (kc | k8) != 0);
            }
            tc_commit_2sm_elect(bar(Smem2::b_empty + st), (uint16_t)3);
          }
          if (xring) {  // the norm term: one MMA over the 16 extra K elements
            const uint32_t xs = tcount & 1;
            if (clk_on) ck_a = clock64();
            mbar_wait(bar(Smem2::x_full + xs), (tcount >> 1) & 1);
            tc_fence_after();
            if (clk_on) ck_x += clock64() - ck_a;
            tc_mma_2sm_elect<KIND>(d_tmem, smem_desc_sw32(sbase + Smem2::a_off + xk * A_CHUNK_BYTES),
                                   smem_desc_sw32(sbase + Smem2::xb_off + xs * XB2_BYTES), 1);
            tc_commit_2sm_elect(bar(Smem2::x_empty + xs), (uint16_t)3);
          }
          tc_commit_2sm_elect(bar(Smem2::t_full + buf), (uint16_t)3);  // accumulators complete in both CTAs
        }
        if (!STREAM) tc_commit_2sm_elect(bar(Smem2::a_empty), (uint16_t)3);  // query tiles no longer needed
      }
      if (clk_on && lane == 0 && blockIdx.x < 160) {
        g_tf32_clk[blockIdx.x][0] = ck_acc;
        g_tf32_clk[blockIdx.x][1] = ck_ops;
        g_tf32_clk[blockIdx.x][2] = ck_x;
        g_tf32_clk[blockIdx.x][3] = clock64() - ck_t0;
      }
    }
  } else if (warp < EPI_WARPS) {
    regs_epilogue();
    EpiCtx ectx;
    ectx.smem = smem; ectx.sbase = sbase; ectx.tmem_base = tmem_base; ectx.warp = warp; ectx.lane = lane;
    ectx.first_item = first_item; ectx.item_step = item_step; ectx.tq_div = tq_div;
    ectx.pair = 2; ectx.crank = crank;
    // "accumulator drained" goes to the LEADER's barrier (the only MMA issuer)
    uint32_t a0, a1;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(a0) : "r"(bar(Smem2::t_empty + 0)));
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(a1) : "r"(bar(Smem2::t_empty + 1)));
    ectx.t_empty_addr0 = leader ? bar(Smem2::t_empty + 0) : a0;
    ectx.t_empty_addr1 = leader ? bar(Smem2::t_empty + 1) : a1;
    ectx.t_empty_remote = leader ? 0 : 1;   // the leader's own warps arrive locally
    ectx.n_full0 = Smem2::n_full; ectx.n_empty0 = Smem2::n_empty; ectx.t_full0 = Smem2::t_full;
    ectx.map_out = &map_out;
    run_epilogue<MODE, KIND, LDW>(P, ectx);
  } else {
    regs_aux();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody leaves (or frees TMEM) while the pair is still working
  if (warp == EPI_WARPS + 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

__global__ void k_fill_f32(float *p, long n, float v) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D map over a row-major [rows][d] float matrix, box = 32 floats x box_rows, 128-byte swizzle,
// out-of-bounds elements read as zero
static int make_map(CUtensorMap *m, const float *ptr, long rows, int d, int box_rows) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return fail(6, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)d * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)ptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(6, "cuTensorMapEncodeTiled failed with code %d", (int)r);
  return 0;
}

// the same over a row-major [rows][pitch] byte matrix (E4M3 operands of the Hamming path):
// box = 128 bytes x box_rows; `pitch` (bytes per row, a multiple of 16) may be smaller than the
// box, the rest of the 128-byte span then reads as zero
static int make_map_u8(CUtensorMap *m, const void *ptr, long rows, int pitch, int box_rows) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return fail(6, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch};
  cuuint32_t box[2] = {(cuuint32_t)(KC * 4), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)ptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(6, "cuTensorMapEncodeTiled (u8) failed with code %d", (int)r);
  return 0;
}

// the same over a row-major [rows][d] FP16 matrix (d a multiple of 8: 16-byte row pitch):
// box = 64 halfs (128 bytes) x box_rows
static int make_map_f16(CUtensorMap *m, const void *ptr, long rows, int d, int box_rows) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return fail(6, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)d * 2};
  cuuint32_t box[2] = {(cuuint32_t)(KC * 2), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void *)ptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(6, "cuTensorMapEncodeTiled (f16) failed with code %d", (int)r);
  return 0;
}

// the 16 extra K elements (columns [col0, col0 + 16)) of an FP16 operand matrix with row pitch d
// halfs: box = 16 halfs (32 bytes) x box_rows, 32-byte swizzle (see smem_desc_sw32)
static int make_map_f16_extras(CUtensorMap *m, const void *ptr, long rows, int d, int box_rows) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return fail(6, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)d * 2};
  cuuint32_t box[2] = {16u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void *)ptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(6, "cuTensorMapEncodeTiled (f16 extras) failed with code %d", (int)r);
  return 0;
}

// the output of compute_cross_distances: dist2[row * ld + query], box = 128 queries x 8 rows of
// floats, no swizzle (the staging tile is plain row-major); stores beyond na / nb are clipped
static int make_map_out(CUtensorMap *m, float *ptr, long na, long nb, long ld) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return fail(6, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {(cuuint64_t)na, (cuuint64_t)nb};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)TM, 8u};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)ptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(6, "cuTensorMapEncodeTiled (output) failed with code %d", (int)r);
  return 0;
}

// How the CTAs of the tensor pass are organised (Tf32Plan::pair):
//   0  independent CTAs
//   1  clusters of 2 that multicast every database chunk (halves the L2 -> SM traffic; measured
//      without effect on the pass: the L2 feed is not what bounds it), opt-in
//   2  clusters of 2 that form ONE cta_group::2 MMA unit (k_knn_2sm): the default for the
//      folded-norm FP16 kind whenever there are at least two query tiles
// YAEL_B200_PAIR=0|1|2 overrides (2 only applies to the folded-norm FP16 kind).
int tf32_pair_mode(int kind, int tiles_q) {
  const bool ham = kind == OP_F8P || kind == OP_F8C;
  const bool can2 = (kind == OP_F16N || ham) && tiles_q >= 2;
  int mode = can2 ? 2 : 0;
  if (const char *e = getenv(ham ? "YAEL_B200_HAM_PAIR" : "YAEL_B200_PAIR")) mode = atoi(e);
  if (mode == 2 && !can2) mode = 0;
  if (mode == 1 && (kind == OP_F8 || ham)) mode = 0;
  if (mode != 0 && (sm_count() & 1)) mode = 0;
  return mode;
}

int tf32_kprime_for(int k) {
  int pct = 200;  // k' = 2k; YAEL_B200_KPRIME_PCT: experiment knob
  if (const char *e = getenv("YAEL_B200_KPRIME_PCT")) pct = atoi(e) >= 100 ? atoi(e) : pct;
  int a = (int)((long)k * pct / 100), b = k + 32 < 8 * k ? k + 32 : 8 * k;
  return a > b ? a : b;
}

// Plan a pass over `nbt_logical` database tiles (a sampling pass sees every tile_stride-th tile)
// that keeps kp candidates per list.
Tf32Plan tf32_plan_tiles(int nq, int nbt_logical, int d, int kp, int kind) {
  Tf32Plan p = {};
  // TMA: 16-byte row pitch; the query tile (A) is resident: at most MAX_NKC chunks of 128 bytes
  const bool h = kind == OP_F16 || kind == OP_F16N;
  const int per_chunk = h ? 2 * KC : KC, mult = h ? 8 : 4;
  if (d < 1 || (d % mult) != 0) return p;
  // chunks a resident query tile would need (folded norms: the 16 extras are a chunk of their own
  // when the data fill whole chunks); more than MAX_NKC: only the streamed 2-SM kernel can do it
  int chunks = (d + per_chunk - 1) / per_chunk;
  if (kind == OP_F16N && d >= 16 && ((d - 16) % per_chunk) == 0) chunks = (d - 16) / per_chunk + 1;
  // (5 .. MAX_NKC_WIDE chunks: the 2-SM kernel's wide resident layout; YAEL_B200_WIDE=0: stream them)
  const bool wide_ok = kind == OP_F16N && !(getenv("YAEL_B200_WIDE") && atoi(getenv("YAEL_B200_WIDE")) == 0);
  bool stream = chunks > (wide_ok ? MAX_NKC_WIDE : MAX_NKC);
  if (kind == OP_F16N && getenv("YAEL_B200_STREAM")) stream = atoi(getenv("YAEL_B200_STREAM")) != 0 || stream;
  const bool wide = !stream && chunks > MAX_NKC;
  if ((stream || wide) && kind != OP_F16N) return p;
  if (nq < 1 || nbt_logical < 1 || kp < 1) return p;
  if (kp + 2 * HALF_N > MAXL) return p;  // the in-register compaction handles MAXL entries
  int cap = pow2_ceil(8 * kp);
  if (cap < 512) cap = 512;
  if (cap > MAXL) cap = MAXL;
  const int pair = tf32_pair_mode(kind, (nq + TM - 1) / TM);
  if ((stream || wide) && pair != 2) return p;  // (fewer than two query tiles, odd SM count: exact engine)
  p.stream = stream ? 1 : 0;
  p.nka = stream ? 0 : (chunks > MAX_NKC ? chunks : MAX_NKC);
  const int G = pair ? sm_count() / 2 : sm_count();          // schedulable units (CTAs or pairs)
  const int tiles_q = pair ? ((nq + TM - 1) / TM + 1) / 2 : (nq + TM - 1) / TM;  // tiles or pairs
  const int nbt = nbt_logical;
  // database ranges: the smallest split count whose last wave is at least 97 % full; every
  // range at least 8 tiles long
  int best_s = 1;
  double best_eff = 0.0;
  for (int s = 1; s <= 160; s++) {  // up to one range per SM for a single query tile
    if (s > 1 && nbt / s < 8) break;
    int range = (nbt + s - 1) / s;
    int s_eff = (nbt + range - 1) / range;
    long items = (long)tiles_q * s_eff;
    long waves = (items + G - 1) / G;
    double eff = (double)items / (double)(waves * G);
    if (eff > best_eff + 0.02) {
      best_eff = eff;
      best_s = s;
    }
    if (eff >= 0.97) break;
  }
  int range = (nbt + best_s - 1) / best_s;
  p.splits = (nbt + range - 1) / range;
  p.lists = 2 * EPI_TEAMS * p.splits;
  p.kprime = kp;
  p.cap = cap;
  long items = (long)tiles_q * p.splits;
  p.ctas = (int)(items < G ? items : G) * (pair ? 2 : 1);
  p.pair = pair;
  p.kind = kind;
  p.ws_bytes = Carver::need(sizeof(float2) * (size_t)p.ctas * EPI_TEAMS * 2 * TM * cap) + 256;
  p.ok = 1;
  return p;
}

Tf32Plan tf32_plan(int nq, int nb, int d, int k, int kind) {
  if (nb < 1) {
    Tf32Plan p = {};
    return p;
  }
  return tf32_plan_tiles(nq, (nb + TN - 1) / TN, d, tf32_kprime_for(k), kind);
}

template <int MODE, int KIND = OP_TF32, int LDW = 16>
static int launch_mode(const Tf32Plan &plan, const CUtensorMap &mq, const CUtensorMap &mb,
                       const CUtensorMap &mbh, const CUtensorMap &mqx, const CUtensorMap &mbx,
                       const Tf32Params &P, cudaStream_t st) {
  static bool attr[64] = {};  // per device: function attributes belong to the current device
  cudaError_t ae = cudaSuccess;
  once_per_device(attr, [&ae] {
    ae = cudaFuncSetAttribute(k_knn_tf32<MODE, KIND, LDW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              TF32_SMEM_BYTES);
  });
  if (ae != cudaSuccess) {
    attr[dev_index()] = false;
    return fail(6, "cannot reserve %d bytes of shared memory: %s", TF32_SMEM_BYTES, cudaGetErrorString(ae));
  }
  if (plan.pair) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(plan.ctas);
    cfg.blockDim = dim3(TF32_THREADS);
    cfg.dynamicSmemBytes = TF32_SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_knn_tf32<MODE, KIND, LDW>, mq, mb, mbh, mqx, mbx, P);
    if (e != cudaSuccess) return fail(2, "k_knn_tf32 cluster launch: %s", cudaGetErrorString(e));
    count_launch();
  } else {
    k_knn_tf32<MODE, KIND, LDW><<<plan.ctas, TF32_THREADS, TF32_SMEM_BYTES, st>>>(mq, mb, mbh, mqx, mbx, P);
    YB_LAUNCH_CHECK();
  }
  return 0;
}

template <int MODE, int LDW, int KIND = OP_F16N, bool STREAM = false>
static int launch_2sm(const Tf32Plan &plan, const CUtensorMap &mq, const CUtensorMap &mbh,
                      const CUtensorMap &mqx, const CUtensorMap &mbxh, const Tf32Params &P, cudaStream_t st,
                      const CUtensorMap *mout = nullptr) {
  static bool attr[64] = {};
  cudaError_t ae = cudaSuccess;
  constexpr int smem_bytes = MODE == EPI_CROSS ? TF32_SMEM2_CROSS_BYTES : TF32_SMEM2_BYTES;
  once_per_device(attr, [&ae] {
    ae = cudaFuncSetAttribute(k_knn_2sm<MODE, LDW, KIND, STREAM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  });
  if (ae != cudaSuccess) {
    attr[dev_index()] = false;
    return fail(6, "cannot reserve %d bytes of shared memory: %s", smem_bytes, cudaGetErrorString(ae));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(plan.ctas);
  cfg.blockDim = dim3(TF32_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k_knn_2sm<MODE, LDW, KIND, STREAM>, mq, mbh, mqx, mbxh,
                                     mout ? *mout : mq, P);
  if (e != cudaSuccess) return fail(2, "k_knn_2sm cluster launch: %s", cudaGetErrorString(e));
  count_launch();
  return 0;
}

static int launch_tf32(const Tf32Plan &plan, int nq, int nb, int d, int nbt_logical,
                       int tile_stride, const float *base, const float *query,
                       const float *bnorm_padded, const float *thr_init, const float *k1_margin,
                       float *out_score, int *out_id, float *out_thr, float *dump, long dump_ld,
                       void *ws, cudaStream_t st, const Tf32Out *oo = nullptr) {
  if ((((uintptr_t)base) & 15) || (((uintptr_t)query) & 15))
    return fail(6, "tf32 path needs 16-byte aligned matrices");
  CUtensorMap mq, mb, mbh, mqx, mbx, mbxh;
  int rc;
  Tf32Params P = {};
  P.nq = nq; P.nb = nb; P.d = d;
  P.xk = -1;
  if (plan.kind == OP_F8 || plan.kind == OP_F8C) {
    // E4M3 operands: d floats of pitch = 4*d bytes = 4*d elements per row; a K chunk is the same
    // 128-byte swizzle span (128 elements), an MMA covers 32 of them.  OP_F8C: the database copy is
    // padded to whole tiles with NaN rows (as the folded-norm FP16 copy)
    const int pitch = 4 * d;
    const long nbm = plan.kind == OP_F8C ? tf32_padded_rows(nb) : nb;
    if ((rc = make_map_u8(&mq, query, nq, pitch, TM))) return rc;
    if ((rc = make_map_u8(&mb, base, nbm, pitch, TN))) return rc;
    if ((rc = make_map_u8(&mbh, base, nbm, pitch, TN / 2))) return rc;
    P.nkc = (pitch + 127) / 128;
    P.last_k8 = (pitch - (P.nkc - 1) * 128 + 31) / 32;
  } else if (plan.kind == OP_F16 || plan.kind == OP_F16N) {
    // FP16 operands: d halfs per row (d % 8 == 0); a K chunk holds 64 of them, an MMA 16
    if (d % 8) return fail(6, "FP16 operands need a row pitch that is a multiple of 8 elements");
    const int pitch = 2 * d;
    // folded norms: the FP16 database copy is padded to whole tiles with rows whose |b|^2 element
    // is NaN (never admitted): the map covers them, so tail tiles need no masks
    const long nb_pad = plan.kind == OP_F16N ? tf32_padded_rows(nb) : nb;
    if ((rc = make_map_f16(&mq, query, nq, d, TM))) return rc;
    if ((rc = make_map_f16(&mb, base, nb_pad, d, TN))) return rc;
    if ((rc = make_map_f16(&mbh, base, nb_pad, d, TN / 2))) return rc;
    if (plan.kind == OP_F16N && ((d - 16) % 64) != 0) {
      // the 16 extras fit behind the data inside the last 128-byte chunk: nothing special
      // (d = 96: 2.49 ms this way, 2.58 ms with a separate extras chunk)
      P.nkc = (pitch + 127) / 128;
      P.last_k8 = (pitch - (P.nkc - 1) * 128 + 31) / 32;
    } else if (plan.kind == OP_F16N) {
      // data columns [0, d - 16) fill whole 128-byte chunks; the 16 extras travel as one
      // 32-byte-wide chunk (a zero-filled 128-byte one costs 2.74 ms instead of 2.62 at d = 128)
      const int dd = d - 16, dbytes = 2 * dd;
      const int nkd = (dbytes + 127) / 128;
      if (nkd + 1 > (plan.nka > MAX_NKC ? plan.nka : MAX_NKC) && !plan.stream)
        return fail(6, "folded-norm FP16 operands: d = %d needs too many chunks", dd);
      P.xk = nkd;
      P.xcol = dd;
      P.xring = nkd <= 2 && !plan.stream && !getenv("YAEL_B200_NO_XRING");
      P.nkc = nkd + 1;
      P.last_k8 = (dbytes - (nkd - 1) * 128 + 31) / 32;
      if ((rc = make_map_f16_extras(&mqx, query, nq, d, TM))) return rc;
      if ((rc = make_map_f16_extras(&mbx, base, nb_pad, d, TN))) return rc;
      if ((rc = make_map_f16_extras(&mbxh, base, nb_pad, d, TN / 2))) return rc;
    } else {
      P.nkc = (pitch + 127) / 128;
      P.last_k8 = (pitch - (P.nkc - 1) * 128 + 31) / 32;
    }
  } else {
    if ((rc = make_map(&mq, query, nq, d, TM))) return rc;
    if ((rc = make_map(&mb, base, nb, d, TN))) return rc;
    if ((rc = make_map(&mbh, base, nb, d, TN / 2))) return rc;
    P.nkc = (d + KC - 1) / KC;
    P.last_k8 = (d - (P.nkc - 1) * KC + 7) / 8;
  }
  if (P.xk < 0) {  // unused by the other kinds, but kernel parameters all the same
    mqx = mq;
    mbx = mb;
    mbxh = mbh;
  }
  P.tiles_q = (nq + TM - 1) / TM;
  P.nbt = nbt_logical;
  P.range_tiles = (P.nbt + plan.splits - 1) / plan.splits;
  P.splits = plan.splits;
  P.lists = plan.lists;
  P.pair = plan.pair;
  P.tiles_q2 = (P.tiles_q + 1) / 2;
  P.items = (plan.pair ? P.tiles_q2 : P.tiles_q) * P.splits;
  P.kprime = plan.kprime;
  P.cap = plan.cap;
  P.bnorm = bnorm_padded;
  P.scratch = (float2 *)ws;
  P.out_score = out_score;
  P.out_id = out_id;
  P.out_thr = out_thr;
  P.thr_init = thr_init;
  P.k1_margin = k1_margin;
  P.tile_stride = tile_stride;
  P.gmin = oo ? oo->gmin : nullptr;
  P.gmin_ld = oo ? oo->gmin_ld : 0;
  P.gsize = oo ? oo->gsize : 16;
  P.out_cnt = oo ? oo->out_cnt : nullptr;
  P.lists_ld = (oo && oo->lists_ld > 0) ? oo->lists_ld : plan.lists;
  P.list0 = oo ? oo->list0 : 0;
  P.id0 = oo ? oo->id0 : 0;
  {
    const char *e = getenv("YAEL_B200_TF32_DEBUG");
    P.debug = e ? atoi(e) : 0;
  }
  P.dump = dump;
  P.dump_ld = dump_ld;
  P.nka = plan.nka > MAX_NKC ? plan.nka : MAX_NKC;
  P.stages2 = STAGES2 + MAX_NKC - P.nka;   // the same 192 KB: query tile + ring
  if (P.nkc > P.nka && !plan.stream) return fail(6, "resident query tile: %d chunks do not fit", P.nkc);
  CUtensorMap mout;
  const bool cross = dump && oo && oo->cross;
  // staged TMA stores need a 16-byte aligned output and row pitch (YAEL_B200_CROSS_TMA=0: the
  // direct stores, A/B)
  if (cross && (dump_ld % 4) == 0 && (((uintptr_t)dump) & 15) == 0 &&
      !(getenv("YAEL_B200_CROSS_TMA") && atoi(getenv("YAEL_B200_CROSS_TMA")) == 0)) {
    if ((rc = make_map_out(&mout, dump, nq, nb, dump_ld))) return rc;
    P.cross_tma = 1;
    P.cross_stage0 = CROSS_STAGE0_OFF;
    P.cross_stage1 = CROSS_STAGE1_OFF;
  }
  const int mode = (dump && oo && oo->cross) ? EPI_CROSS
                   : (dump ? EPI_DUMP : (P.gmin ? EPI_GMIN : (k1_margin ? EPI_NEAREST : EPI_LISTS)));
  if (mode == EPI_CROSS && !(plan.kind == OP_F16N && plan.pair == 2))
    return fail(6, "the cross-distance mode needs the 2-SM folded-norm kernel");
  P.acc_scale = plan.acc_scale;
  {
    const char *e = getenv("YAEL_B200_EPI_DEPHASE");
    P.dephase = e ? atoi(e) : 0;
  }
  {
    const char *e = getenv("YAEL_B200_TF32_ORDER");
    P.order = e ? atoi(e) : 0;
  }
  P.ham_slots = plan.ham_slots;
  P.ham_nb = plan.ham_nb;
  P.ham_magic = plan.ham_magic;
  P.c0 = plan.score_c0;
  if (plan.kind == OP_F8C) {  // E4M3 with a constant norm: the folded-norm epilogues, no |b|^2 ring
    if (plan.pair == 1) return fail(6, "the E4M3 operand kind has no multicast-pair variant");
    if (plan.pair == 2) {
      switch (mode) {
        case EPI_DUMP: return launch_2sm<EPI_DUMP, 16, OP_F8C>(plan, mq, mbh, mqx, mbxh, P, st);
        case EPI_GMIN: return launch_2sm<EPI_GMIN, 128, OP_F8C>(plan, mq, mbh, mqx, mbxh, P, st);
        case EPI_LISTS: {
          // FP16 accumulators for the full pass (YAEL_B200_HAM_F16ACC=0: FP32, A/B)
          static const bool hacc = getenv("YAEL_B200_HAM_F16ACC") && atoi(getenv("YAEL_B200_HAM_F16ACC")) != 0;
          if (hacc) return launch_2sm<EPI_LISTS, 128, OP_F8H>(plan, mq, mbh, mqx, mbxh, P, st);
          return launch_2sm<EPI_LISTS, 128, OP_F8C>(plan, mq, mbh, mqx, mbxh, P, st);
        }
        default: return fail(6, "the E4M3 operand kind has no k = 1 margin mode");
      }
    }
    switch (mode) {
      case EPI_DUMP: return launch_mode<EPI_DUMP, OP_F8C>(plan, mq, mb, mbh, mqx, mbx, P, st);
      case EPI_GMIN: return launch_mode<EPI_GMIN, OP_F8C, 128>(plan, mq, mb, mbh, mqx, mbx, P, st);
      case EPI_LISTS: return launch_mode<EPI_LISTS, OP_F8C, 128>(plan, mq, mb, mbh, mqx, mbx, P, st);
      default: return fail(6, "the E4M3 operand kind has no k = 1 margin mode");
    }
  }
  if (plan.kind == OP_F8 && plan.ham_slots > 1 && !dump) {
    // packed Hamming passes: ham_slots database rows per accumulator, integer epilogue
    if (plan.pair == 1) return fail(6, "the E4M3 operand kind has no multicast-pair variant");
    if (plan.kprime + 2 * HALF_N * plan.ham_slots > MAXL)
      return fail(6, "packed Hamming pass: k' = %d leaves no room for a tile of appends", plan.kprime);
    // YAEL_B200_HAM_LDW=16: the round-1 epilogue (8 x 16-column loads per tile), A/B knob
    int hldw = 128;
    if (const char *e = getenv("YAEL_B200_HAM_LDW")) hldw = atoi(e);
    if (plan.pair == 2) {  // cta_group::2 pairs: each SM holds half of every database chunk
      if (P.gmin) return launch_2sm<EPI_HAMG, 16, OP_F8>(plan, mq, mbh, mqx, mbxh, P, st);
      if (hldw == 128) return launch_2sm<EPI_HAMP, 128, OP_F8>(plan, mq, mbh, mqx, mbxh, P, st);
      return launch_2sm<EPI_HAMP, 16, OP_F8>(plan, mq, mbh, mqx, mbxh, P, st);
    }
    if (P.gmin) return launch_mode<EPI_HAMG, OP_F8>(plan, mq, mb, mbh, mqx, mbx, P, st);
    if (hldw == 128) return launch_mode<EPI_HAMP, OP_F8, 128>(plan, mq, mb, mbh, mqx, mbx, P, st);
    return launch_mode<EPI_HAMP, OP_F8>(plan, mq, mb, mbh, mqx, mbx, P, st);
  }
  if (plan.kind == OP_F8) {
    if (plan.pair) return fail(6, "the E4M3 operand kind has no paired-CTA variant");
    switch (mode) {
      case EPI_DUMP: return launch_mode<EPI_DUMP, OP_F8>(plan, mq, mb, mbh, mqx, mbx, P, st);
      case EPI_GMIN: return launch_mode<EPI_GMIN, OP_F8>(plan, mq, mb, mbh, mqx, mbx, P, st);
      case EPI_LISTS: return launch_mode<EPI_LISTS, OP_F8>(plan, mq, mb, mbh, mqx, mbx, P, st);
      default: return fail(6, "the E4M3 operand kind has no k = 1 margin mode");
    }
  }
  if (plan.kind == OP_F16) {  // plain FP16 operands: the k = 1 margin mode (k-means)
    switch (mode) {
      case EPI_NEAREST: return launch_mode<EPI_NEAREST, OP_F16>(plan, mq, mb, mbh, mqx, mbx, P, st);
      default: return fail(6, "plain FP16 operands are only instantiated for the k = 1 margin mode");
    }
  }
  // columns per tcgen05.ld of the folded-norm list / nearest epilogues (YAEL_B200_LDW: A/B knob)
  // 128 = the early hand-back epilogue (whole half tile in registers, buffer released before the
  // threshold tests).  Measured at the bench shape with the 2-SM kernel: 16 -> 2.51 ms, 32 -> 2.43,
  // 64 -> 2.50, 128 -> 2.33 (1-SM kernel: 2.52 / 2.47 / 2.50 / 2.40); k-means config 4: 176 -> 150 ms.
  int ldw = 128;
  if (const char *e = getenv("YAEL_B200_LDW")) ldw = atoi(e);
  if (plan.kind == OP_F16N && plan.pair == 2 && plan.stream) {  // streamed query chunks: any d
    switch (mode) {
      case EPI_CROSS: return launch_2sm<EPI_CROSS, 16, OP_F16N, true>(plan, mq, mbh, mqx, mbxh, P, st, P.cross_tma ? &mout : nullptr);
      case EPI_DUMP: return launch_2sm<EPI_DUMP, 16, OP_F16N, true>(plan, mq, mbh, mqx, mbxh, P, st);
      case EPI_GMIN: return launch_2sm<EPI_GMIN, 128, OP_F16N, true>(plan, mq, mbh, mqx, mbxh, P, st);
      case EPI_LISTS: return launch_2sm<EPI_LISTS, 128, OP_F16N, true>(plan, mq, mbh, mqx, mbxh, P, st);
      default: return launch_2sm<EPI_NEAREST, 128, OP_F16N, true>(plan, mq, mbh, mqx, mbxh, P, st);
    }
  }
  if (plan.stream) return fail(6, "streamed query chunks need the 2-SM folded-norm kernel");
  if (plan.kind == OP_F16N && plan.pair == 2) {  // cta_group::2 pairs (k_knn_2sm)
    switch (mode) {
      case EPI_CROSS: return launch_2sm<EPI_CROSS, 16>(plan, mq, mbh, mqx, mbxh, P, st, P.cross_tma ? &mout : nullptr);
      case EPI_DUMP: return launch_2sm<EPI_DUMP, 16>(plan, mq, mbh, mqx, mbxh, P, st);
      case EPI_GMIN:
        if (ldw == 16) return launch_2sm<EPI_GMIN, 16>(plan, mq, mbh, mqx, mbxh, P, st);
        return launch_2sm<EPI_GMIN, 128>(plan, mq, mbh, mqx, mbxh, P, st);
      case EPI_LISTS:
        if (ldw == 128) return launch_2sm<EPI_LISTS, 128>(plan, mq, mbh, mqx, mbxh, P, st);
        if (ldw == 64) return launch_2sm<EPI_LISTS, 64>(plan, mq, mbh, mqx, mbxh, P, st);
        if (ldw == 32) return launch_2sm<EPI_LISTS, 32>(plan, mq, mbh, mqx, mbxh, P, st);
        return launch_2sm<EPI_LISTS, 16>(plan, mq, mbh, mqx, mbxh, P, st);
      default:
        if (ldw == 128) return launch_2sm<EPI_NEAREST, 128>(plan, mq, mbh, mqx, mbxh, P, st);
        if (ldw == 32) return launch_2sm<EPI_NEAREST, 32>(plan, mq, mbh, mqx, mbxh, P, st);
        return launch_2sm<EPI_NEAREST, 16>(plan, mq, mbh, mqx, mbxh, P, st);
    }
  }
  if (plan.kind == OP_F16N) {  // FP16 operands with folded norms: top-k', sampling, dump
    switch (mode) {
      case EPI_DUMP: return launch_mode<EPI_DUMP, OP_F16N>(plan, mq, mb, mbh, mqx, mbx, P, st);
      case EPI_GMIN:
        if (ldw == 16) return launch_mode<EPI_GMIN, OP_F16N>(plan, mq, mb, mbh, mqx, mbx, P, st);
        return launch_mode<EPI_GMIN, OP_F16N, 128>(plan, mq, mb, mbh, mqx, mbx, P, st);
      case EPI_LISTS:
        if (ldw == 128) return launch_mode<EPI_LISTS, OP_F16N, 128>(plan, mq, mb, mbh, mqx, mbx, P, st);
        if (ldw == 64) return launch_mode<EPI_LISTS, OP_F16N, 64>(plan, mq, mb, mbh, mqx, mbx, P, st);
        if (ldw == 32) return launch_mode<EPI_LISTS, OP_F16N, 32>(plan, mq, mb, mbh, mqx, mbx, P, st);
        return launch_mode<EPI_LISTS, OP_F16N>(plan, mq, mb, mbh, mqx, mbx, P, st);
      default:
        if (ldw == 128) return launch_mode<EPI_NEAREST, OP_F16N, 128>(plan, mq, mb, mbh, mqx, mbx, P, st);
        if (ldw == 32) return launch_mode<EPI_NEAREST, OP_F16N, 32>(plan, mq, mb, mbh, mqx, mbx, P, st);
        return launch_mode<EPI_NEAREST, OP_F16N>(plan, mq, mb, mbh, mqx, mbx, P, st);
    }
  }
  switch (mode) {
    case EPI_DUMP: return launch_mode<EPI_DUMP>(plan, mq, mb, mbh, mqx, mbx, P, st);
    case EPI_GMIN: return launch_mode<EPI_GMIN>(plan, mq, mb, mbh, mqx, mbx, P, st);
    case EPI_NEAREST: return launch_mode<EPI_NEAREST>(plan, mq, mb, mbh, mqx, mbx, P, st);
    default: return launch_mode<EPI_LISTS>(plan, mq, mb, mbh, mqx, mbx, P, st);
  }
}

int tf32_shortlist(const Tf32Plan &plan, int nq, int nb, int d, int nbt_logical, int tile_stride,
                   const float *base, const float *query, const float *bnorm_padded,
                   const float *thr_init, float *out_score, int *out_id, float *out_thr, void *ws,
                   cudaStream_t st, const Tf32Out *oo) {
  return launch_tf32(plan, nq, nb, d, nbt_logical, tile_stride, base, query, bnorm_padded,
                     thr_init, nullptr, out_score, out_id, out_thr, nullptr, 0, ws, st, oo);
}

// k = 1 mode: per list the (at most plan.kprime) rows whose TF32 score is within k1_margin[q]
// of the list's best score; out_thr[q][l] = best + margin, or NaN when more rows than that
// qualified (the caller re-does such queries exactly).
int tf32_nearest(const Tf32Plan &plan, int nq, int nb, int d, const float *base, const float *query,
                 const float *bnorm_padded, const float *k1_margin, float *out_score, int *out_id,
                 float *out_thr, void *ws, cudaStream_t st) {
  return launch_tf32(plan, nq, nb, d, tf32_tiles(nb), 1, base, query, bnorm_padded, nullptr,
                     k1_margin, out_score, out_id, out_thr, nullptr, 0, ws, st);
}

Tf32Plan tf32_plan_nearest(int nq, int nb, int d, int kind) {
  Tf32Plan p = tf32_plan_tiles(nq, tf32_tiles(nb), d, 8, kind);
  if (p.ok) {
    p.cap = 64;
    p.ws_bytes = Carver::need(sizeof(float2) * (size_t)p.ctas * EPI_TEAMS * 2 * TM * p.cap) + 256;
  }
  return p;
}

// raw TF32 scores of the logical tiles (every tile_stride-th database tile): scores[q][ld]
int tf32_scores(const Tf32Plan &plan, int nq, int nb, int d, int nbt_logical, int tile_stride,
                const float *base, const float *query, const float *bnorm_padded, float *scores,
                long ld, void *ws, cudaStream_t st) {
  return launch_tf32(plan, nq, nb, d, nbt_logical, tile_stride, base, query, bnorm_padded, nullptr,
                     nullptr, nullptr, nullptr, nullptr, scores, ld, ws, st);
}

// group minima of the logical tiles: gmin[q][(2*j + h) * (128/gsize) + g] = the smallest score
// among columns [g*gsize, (g+1)*gsize) of half h of logical tile j
int tf32_group_min(const Tf32Plan &plan, int nq, int nb, int d, int nbt_logical, int tile_stride,
                   const float *base, const float *query, const float *bnorm_padded, float *gmin,
                   long ld, int gsize, void *ws, cudaStream_t st) {
  Tf32Out oo = {};
  oo.gmin = gmin;
  oo.gmin_ld = ld;
  oo.gsize = gsize;
  return launch_tf32(plan, nq, nb, d, nbt_logical, tile_stride, base, query, bnorm_padded, nullptr,
                     nullptr, nullptr, nullptr, nullptr, nullptr, 0, ws, st, &oo);
}

// the full distance matrix out[row * ld + query] = asc * acc (operands carry both norms)
int tf32_cross(const Tf32Plan &plan, int nq, int nb, int d, const float *base, const float *query,
               float *out, long ld, void *ws, cudaStream_t st) {
  Tf32Out oo = {};
  oo.cross = 1;
  return launch_tf32(plan, nq, nb, d, tf32_tiles(nb), 1, base, query, nullptr, nullptr, nullptr, nullptr,
                     nullptr, nullptr, out, ld, ws, st, &oo);
}

long tf32_padded_rows(int nb) { return (long)((nb + TN - 1) / TN) * TN; }
int tf32_tiles(int nb) { return (nb + TN - 1) / TN; }

int fill_f32(float *p, long n, float v, cudaStream_t st) {
  if (n <= 0) return 0;
  k_fill_f32<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, n, v);
  YB_LAUNCH_CHECK();
  return 0;
}

}  // namespace yb

using namespace yb;

// Debug / test entry: raw TF32 scores s[q][n] = |b_n|^2 - 2 <q, b_n> for every pair (the tensor
// path with the top-k switched off).  Used by tests to bound the TF32 error against the
// certificate's model, and for bring-up.
// bring-up: per-CTA clock attribution of the last instrumented tensor pass (see g_tf32_clk)
extern "C" int yb_debug_tf32_clocks(long long *out, int n_cta) {
  if (n_cta > 160) n_cta = 160;
  YB_CUDA(cudaDeviceSynchronize());
  YB_CUDA(cudaMemcpyFromSymbol(out, g_tf32_clk, sizeof(long long) * 16 * (size_t)n_cta));
  return 0;
}

extern "C" int yb_debug_tf32_scores(int nq, int nb, int d, const float *base, const float *query,
                                    float *scores, yb_stream_t s) {
  Guard g;
  cudaStream_t st = stream_of(s);
  Tf32Plan plan = tf32_plan(nq, nb, d, 1);
  if (!plan.ok) return fail(3, "tf32 path does not support this shape (d=%d)", d);
  long padded = tf32_padded_rows(nb);
  ScratchScope ws(Carver::need(4ull * padded) + Carver::need(plan.ws_bytes), st);
  Carver c(ws.p);
  float *bn = c.take<float>(padded);
  void *tws = c.take<char>(plan.ws_bytes);
  int rc;
  if ((rc = row_norms_seq(base, nb, d, d, bn, nullptr, st))) return rc;
  if ((rc = fill_f32(bn + nb, padded - nb, __builtin_inff(), st))) return rc;
  return launch_tf32(plan, nq, nb, d, tf32_tiles(nb), 1, base, query, bn, nullptr, nullptr, nullptr,
                     nullptr, nullptr, scores, nb, tws, st);
}
