// yb_knn.cu -- exact k-NN orchestration (knn_full, yael/nn.c:451-525) on device pointers.
//
// Two engines produce the same values:
//   engine 1  tcgen05 TF32 shortlist (yb_knn_tf32.cu) -> exact FP32 re-rank (k_rerank below)
//             with a per-query certificate; queries that fail it are re-done by engine 0.
//   engine 0  exact FP32 distance slab (k_l2_simt) -> per-row select (k_kmin_rows).
// Both end in the reference's distance formula with the dot product as a sequential FP32
// FMA chain, and both order results by (distance, id).
#include <stdlib.h>

#include <cuda_fp16.h>

#include "yb_common.cuh"
#include "yb_internal.cuh"

namespace yb {

static int g_engine_force = -1;
static thread_local int g_last_engine = 0;
static thread_local int g_last_operands = 0;  // operand kind of the last tensor pass (0 TF32, 2 FP16)
static thread_local long g_last_uncert = 0;

// Absolute part of the tensor-score error bound, as a fraction of max|b - mu|^2: the squared norm of
// a centred row is an FP32 sum (<= ~10 roundings of 2^-24 for d <= 144 as the conversion kernel
// sums it) and the accumulator acc' = <q,b> - |b|^2/2 is rounded (towards zero, 2^-23, plus the
// alignment loss of the 16 products of an instruction, <= 2^-20 of the largest one) once per MMA,
// 9 to 18 MMAs per tile, at a magnitude of up to |b|^2/2 + |q||b|.  The |q||b| share of that is
// inside the 5 % head room of the Cauchy-Schwarz term; the |b|^2 share is NOT proportional to |q|
// and needs its own term: 2 * 18 * (2^-23 + 2^-20) / 2 + 10 * 2^-24 < 2.3e-5.
constexpr float kAccAbs = 2.3e-5f;
// Rounding of the reference's own FP32 distance (sequential float norm, sequential FMA dot product,
// two float additions; yael/nn.c:100-129) relative to |q|^2 + |b|^2 of the UNCENTRED rows: what two
// exact near-ties may be swapped by in the reference's ranking.  2e-6 = 33 x 2^-24 covers the
// random-walk growth of a d <= 1000 chain with a wide margin.
constexpr float kRefRound = 2e-6f;

// ------------------------------------------------------------------ exact re-rank
// One CTA (4 warps) per query.  Candidates come as ids; each warp takes 32 candidates at a
// time and every lane walks its candidate's row sequentially (warp_rows_seq pattern,
// restated here because rows are gathered).
//
// MODE 0: knn_reorder_shortlist (yael/nn.c:528-580): ids = idx[q][0..ki) up to the first
//         negative id; distance = compute_distances_1 (both norms double, nn.c:132-154);
//         order (distance, position).
// MODE 1: shortlist re-rank for knn_full: ids = lists[q][0..m) (id < 0 = empty slot);
//         distance = nn.c:100-129 with the base row as the a-operand (float norm) and the
//         query as the b-operand (double norm); order (distance, id); emits the k best,
//         padding, and the certificate flag.
struct RerankArgs {
  int nq, nb, d, k;
  const float *base;
  const float *query;
  // MODE 0
  int *idx;
  float *dis;
  // MODE 1: candidates = cand_id[q][pos] for pos = sel[q][j] (or pos = j when sel == NULL),
  // j < m; cand_score carries their TF32 scores
  const int *cand_id;      // [nq][cand_stride]
  const float *cand_score; // [nq][cand_stride]
  const int *sel;          // [nq][m] positions (ascending TF32 score), -1 = none
  int cand_stride;
  int m;
  int all_listed;          // 1: every database row is a candidate (nothing was ever dropped)
  const float *cand_thr;   // [nq][lists] final admission threshold of every shortlist
  int lists;
  const float *query_c;    // centred queries the tensor pass saw (row pitch qc_ld)
  int qc_ld;
  float err_scale;         // certificate: E_q = err_scale * |q| * max|b|
  const float *bmax;       // device scalar: max |b| (sqrt of max squared norm)
  const float *err_abs;    // device scalar (or NULL): E_q += err_abs * (|q| + max|b|) (FP16 operands)
  const float *mu_norm;    // device scalar: |mu|, the norm of the centring vector (reference-rounding term)
  int *assign;
  int id_offset;
  int *uncert_flags;       // [nq] 1 = certificate failed
  unsigned long long *gsort; // global sort slab when m_pad > 4096
  int m_pad;
  int k1;                  // k == 1: nn_single_full start value (-1, 1e30f), nn.c:404-407
  int prefetch;            // issue L2 prefetches for every candidate row before the chains start
  int cpa;                 // rows staged with cp.async into [32][RR_PITCH] tiles at tile_off (d % 4 == 0,
  unsigned tile_off;       // 16-byte aligned base): a row is one instruction, the chains read float4
};

constexpr int RR_T = 128;  // threads per query in the re-rank kernels
constexpr int RR_PITCH = 132;  // floats per staged row (cp.async variant): 16-byte aligned, conflict-free float4 reads
constexpr int RR_TILE_BYTES = (RR_T / 32) * 32 * RR_PITCH * 4;

__device__ __forceinline__ void rr_cp_async16(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc)
               : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(RR_T) k_rerank(RerankArgs A) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ double qn_sh;
  // transposition tiles of the scalar path (the cp.async path has its own layout at the same offset)
  float (*tile)[32][33] = reinterpret_cast<float (*)[32][33]>(smem_raw + A.tile_off);
  const int q = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int d = A.d;
  const int m = MODE == 0 ? A.k : A.m;
  const int m_pad = A.m_pad;
  float *qs = reinterpret_cast<float *>(smem_raw);
  size_t off = ((size_t)d * sizeof(float) + 15) & ~(size_t)15;
  int *ids = reinterpret_cast<int *>(smem_raw + off);
  off += ((size_t)m * sizeof(int) + 15) & ~(size_t)15;
  float *dv = reinterpret_cast<float *>(smem_raw + off);
  off += ((size_t)m * sizeof(float) + 15) & ~(size_t)15;
  unsigned long long *sortbuf = (m_pad <= 4096)
                                    ? reinterpret_cast<unsigned long long *>(smem_raw + off)
                                    : A.gsort + (size_t)q * m_pad;

  for (int t = tid; t < d; t += RR_T) qs[t] = A.query[(size_t)q * d + t];
  int ki = m;
  if (MODE == 0) {
    for (int j = tid; j < m; j += RR_T) ids[j] = A.idx[(size_t)q * m + j];
  } else {
    for (int j = tid; j < m; j += RR_T) {
      int pos = A.sel ? A.sel[(size_t)q * m + j] : j;
      ids[j] = pos >= 0 ? A.cand_id[(size_t)q * A.cand_stride + pos] : -1;
    }
  }
  __syncthreads();
  if (MODE == 0) {  // stop at the first negative id (nn.c:546-551)
    ki = m;
    for (int j = 0; j < m; j++)
      if (ids[j] < 0) {
        ki = j;
        break;
      }
  }
  if (A.prefetch) {  // the candidate rows are a gather from HBM: start every line towards L2 now
    const int lines = (d * 4 + 127) >> 7;
    for (int e = tid; e < ki * lines; e += RR_T) {
      const int id = ids[e / lines];
      if (id >= 0)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(A.base + (size_t)id * d + (size_t)(e % lines) * 32));
    }
  }
  if (tid == 0) {
    double s = 0.0;
    for (int t = 0; t < d; t++) s += (double)__fmul_rn(qs[t], qs[t]);
    qn_sh = s;
  }
  __syncthreads();
  const double qn = qn_sh;

  // cp.async variant (the kernel was issue-bound: 32 scalar loads + 32 stores + 32 x 5 chain
  // instructions per 32 x 32 block): every candidate row of the block comes in with ONE 16-byte
  // cp.async per lane (128 coordinates at a time), and lane c walks row c reading float4 -- the same
  // operations in the same order
  for (int c0 = warp * 32; A.cpa && c0 < ki; c0 += RR_T) {
    const int c = c0 + lane;
    const int id = c < ki ? ids[c] : -1;
    float *wt = reinterpret_cast<float *>(smem_raw + A.tile_off) + (size_t)warp * 32 * RR_PITCH;
    float nf = 0.f, dot = 0.f;
    double nd = 0.0;
    for (int t0 = 0; t0 < d; t0 += 128) {
      const int w = min(128, d - t0);
      const bool col = lane * 4 < w;
#pragma unroll 8
      for (int r = 0; r < 32; r++) {
        const int idr = __shfl_sync(0xffffffffu, id, r);
        if (idr >= 0 && col) rr_cp_async16(wt + r * RR_PITCH + lane * 4, A.base + (size_t)idr * d + t0 + lane * 4);
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncwarp();
      if (id >= 0) {
        const float4 *cr = reinterpret_cast<const float4 *>(wt + lane * RR_PITCH);
        const float4 *qr = reinterpret_cast<const float4 *>(qs + t0);
#pragma unroll 8
        for (int t4 = 0; t4 < (w >> 2); t4++) {
          const float4 v = cr[t4], qv = qr[t4];
          float sq;
          sq = __fmul_rn(v.x, v.x); if (MODE == 0) nd += (double)sq; else nf = __fadd_rn(nf, sq); dot = fmaf(v.x, qv.x, dot);
          sq = __fmul_rn(v.y, v.y); if (MODE == 0) nd += (double)sq; else nf = __fadd_rn(nf, sq); dot = fmaf(v.y, qv.y, dot);
          sq = __fmul_rn(v.z, v.z); if (MODE == 0) nd += (double)sq; else nf = __fadd_rn(nf, sq); dot = fmaf(v.z, qv.z, dot);
          sq = __fmul_rn(v.w, v.w); if (MODE == 0) nd += (double)sq; else nf = __fadd_rn(nf, sq); dot = fmaf(v.w, qv.w, dot);
        }
      }
      __syncwarp();
    }
    if (c < ki) {
      float base = MODE == 0 ? (float)(nd + qn) : (float)(qn + (double)nf);
      dv[c] = id >= 0 ? __fadd_rn(base, __fmul_rn(-2.0f, dot)) : __uint_as_float(0x7fc00000u);
    }
  }
  for (int c0 = warp * 32; !A.cpa && c0 < ki; c0 += RR_T) {
    const int c = c0 + lane;
    const int id = c < ki ? ids[c] : -1;
    const float *rowp = id >= 0 ? A.base + (size_t)id * d : nullptr;
    float nf = 0.f, dot = 0.f;
    double nd = 0.0;
    for (int t0 = 0; t0 < d; t0 += 32) {
      const int w = min(32, d - t0);
      // all 32 row loads of the chunk in flight before the first one is consumed
      float gv[32];
#pragma unroll
      for (int r = 0; r < 32; r++) {
        const float *p = (const float *)__shfl_sync(0xffffffffu, (unsigned long long)rowp, r);
        gv[r] = (p != nullptr && lane < w) ? __ldg(p + t0 + lane) : 0.f;
      }
#pragma unroll
      for (int r = 0; r < 32; r++) tile[warp][r][lane] = gv[r];
      __syncwarp();
      if (rowp != nullptr) {
        for (int t = 0; t < w; t++) {
          float v = tile[warp][lane][t];
          float sq = __fmul_rn(v, v);
          if (MODE == 0) nd += (double)sq; else nf = __fadd_rn(nf, sq);
          dot = fmaf(v, qs[t0 + t], dot);
        }
      }
      __syncwarp();
    }
    if (c < ki) {
      float base = MODE == 0 ? (float)(nd + qn) : (float)(qn + (double)nf);
      dv[c] = id >= 0 ? __fadd_rn(base, __fmul_rn(-2.0f, dot)) : __uint_as_float(0x7fc00000u);
    }
  }
  __syncthreads();
  for (int j = tid; j < m_pad; j += RR_T) {
    unsigned long long key = ~0ull;
    if (j < ki && ids[j] >= 0) {
      uint32_t fk = float_key(dv[j]);
      if (MODE == 0)
        key = ((unsigned long long)fk << 32) | (unsigned)j;
      else if (!is_nan_key(fk))
        key = ((unsigned long long)fk << 32) | (unsigned)j;
    }
    sortbuf[j] = key;
  }
  if (MODE == 1) {
    // order by (distance, id): positions are not ids, so fold the id in instead
    __syncthreads();
    for (int j = tid; j < m_pad; j += RR_T) {
      unsigned long long key = sortbuf[j];
      if (key != ~0ull) sortbuf[j] = (key & 0xffffffff00000000ull) | (unsigned)ids[(int)(uint32_t)key];
    }
  }
  bitonic_sort_u64(sortbuf, m_pad, tid, RR_T, [] { __syncthreads(); });

  if (MODE == 0) {
    for (int j = tid; j < ki; j += RR_T) {
      int pos = (int)(uint32_t)sortbuf[j];
      A.dis[(size_t)q * m + j] = dv[pos];
      A.idx[(size_t)q * m + j] = ids[pos];
    }
  } else {
    const int k = A.k;
    for (int j = tid; j < k; j += RR_T) {
      unsigned long long key = j < m_pad ? sortbuf[j] : ~0ull;
      uint32_t fk = (uint32_t)(key >> 32);
      uint32_t bits = (fk & 0x80000000u) ? (fk & 0x7fffffffu) : ~fk;
      if (A.k1 && !(key != ~0ull && __uint_as_float(bits) < 1e30f)) {
        A.assign[(size_t)q * k + j] = -1;
        A.dis[(size_t)q * k + j] = 1e30f;
      } else if (key != ~0ull) {
        A.assign[(size_t)q * k + j] = (int)(uint32_t)key + A.id_offset;
        A.dis[(size_t)q * k + j] = __uint_as_float(bits);
      } else {  // yael/nn.c:515-518
        A.assign[(size_t)q * k + j] = -1;
        A.dis[(size_t)q * k + j] = __uint_as_float(0xffffffffu);
      }
    }
    // Certificate.  A database row that is NOT among the m candidates was either refused by
    // a shortlist (TF32 score >= that list's final admission threshold) or dropped by the merge
    // (TF32 score >= the largest selected score).  With T the smallest of those bounds, its
    // exact score is >= T - E_q, so the k-th exact distance D_k (score S_k = D_k - |q|^2)
    // cannot be beaten by such a row when S_k + E_q < T.  T = +inf: nothing was ever refused.
    // (the reductions over lists / candidates / coordinates are done by warp 0 in parallel: a
    // single thread chasing 200 dependent global loads used to dominate this kernel)
    if (warp == 0) {
      const float inf = __uint_as_float(0x7f800000u);
      float T = inf;
      int nvalid = 0;
      float mx = -inf;
      double qcn = 0.0;
      if (!A.all_listed) {
        for (int l = lane; l < A.lists; l += 32) T = fminf(T, A.cand_thr[(size_t)q * A.lists + l]);
        for (int j = lane; j < m; j += 32) {
          int pos = A.sel ? A.sel[(size_t)q * m + j] : j;
          if (pos >= 0 && ids[j] >= 0) {
            nvalid++;
            mx = fmaxf(mx, A.cand_score[(size_t)q * A.cand_stride + pos]);
          }
        }
        const float *qc = A.query_c + (size_t)q * A.qc_ld;
        for (int t = lane; t < d; t += 32) qcn += (double)qc[t] * (double)qc[t];
#pragma unroll
        for (int o = 16; o; o >>= 1) {
          T = fminf(T, __shfl_xor_sync(0xffffffffu, T, o));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          nvalid += __shfl_xor_sync(0xffffffffu, nvalid, o);
          qcn += __shfl_xor_sync(0xffffffffu, qcn, o);
        }
        if (nvalid == m) T = fminf(T, mx);  // the merge may have dropped rows at or above mx
      }
      if (lane == 0) {
        int flag = 0;
        if (T < inf) {
          unsigned long long key = (k - 1) < m_pad ? sortbuf[k - 1] : ~0ull;
          if (key == ~0ull) {
            flag = 1;  // fewer than k candidates survived although rows were dropped
          } else {
            uint32_t fk = (uint32_t)(key >> 32);
            uint32_t bits = (fk & 0x80000000u) ? (fk & 0x7fffffffu) : ~fk;
            // the tensor pass scored CENTRED operands: S_c(b) = |q-b|^2 - |q-mu|^2
            double Dk = (double)__uint_as_float(bits);
            // E_q: operand rounding (Cauchy-Schwarz term) + an ABSOLUTE term for the FP32 steps
            // that do not shrink with |q - mu| (kAccAbs: rounding of |b - mu|^2 and of the tensor
            // core's FP32 accumulation, whose partial sums are of size |b - mu|^2 / 2 however small
            // the query is) + the rounding of the REFERENCE's own FP32 arithmetic, which works on
            // the uncentred rows (kRefRound * (|q|^2 + max|b|^2), |b| <= |b - mu| + |mu|).
            const double bmx = (double)(*A.bmax), mun = A.mu_norm ? (double)(*A.mu_norm) : 0.0;
            double E = (double)A.err_scale * sqrt(qcn) * bmx + (double)kAccAbs * bmx * bmx +
                       4e-5 * (fabs(Dk) + qcn) +
                       (double)kRefRound * (fabs(Dk) + qn + (bmx + mun) * (bmx + mun));
            if (A.err_abs) E += (double)(*A.err_abs) * (sqrt(qcn) + bmx);
            if (!((Dk - qcn) + E < (double)T)) flag = 1;
          }
        }
        A.uncert_flags[q] = flag;
      }
    }
  }
}

// ------------------------------------------------------------------ shard merge
// One CTA per query: G*k (distance, id) pairs -> k best by (distance, id).  Shard g's list for
// query q starts at ain / din + g * sstride + q * k.  The lists yb_knn_l2 produces are sorted, so
// the merge is a RANKING, not a sort: the merged position of an entry is its position in its own
// list plus, for every other list, the number of entries below it (a binary search in shared
// memory; keys are distinct because row ids are).  Unsorted input falls back to a bitonic sort.
__global__ void __launch_bounds__(128)
k_knn_merge(int k, int G, long nq, const int *__restrict__ ain, const float *__restrict__ din,
            long sstride, int *__restrict__ aout, float *__restrict__ dout,
            unsigned long long *gsort, int m_pad) {
  extern __shared__ unsigned long long ssort[];
  const long q = blockIdx.x;
  const int tid = threadIdx.x, m = G * k;
  unsigned long long *buf = m_pad <= 4096 ? ssort : gsort + q * m_pad;
  for (int j = tid; j < m_pad; j += 128) {
    unsigned long long key = ~0ull;
    if (j < m) {
      int g = j / k, r = j - g * k;
      size_t src = (size_t)g * sstride + (size_t)q * k + r;
      int id = ain[src];
      uint32_t fk = float_key(din[src]);
      if (id >= 0 && !is_nan_key(fk)) key = ((unsigned long long)fk << 32) | (unsigned)id;
    }
    buf[j] = key;
  }
  __syncthreads();
  int unsorted = 0;
  for (int j = tid; j < m; j += 128)
    if (j % k != 0 && buf[j] < buf[j - 1]) unsorted = 1;
  if (!__syncthreads_or(unsorted)) {
    for (int j = tid; j < k; j += 128) {  // padding first; the ranked writes below overwrite it
      aout[q * k + j] = -1;
      dout[q * k + j] = __uint_as_float(0xffffffffu);
    }
    __syncthreads();
    for (int j = tid; j < m; j += 128) {
      const unsigned long long key = buf[j];
      if (key == ~0ull) continue;
      const int g = j / k;
      int rank = j - g * k;
      for (int o = 0; o < G && rank < k; o++) {
        if (o == g) continue;
        const unsigned long long *l = buf + o * k;
        int lo = 0, hi = k;  // lower_bound: entries of list o below key
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (l[mid] < key) lo = mid + 1; else hi = mid;
        }
        rank += lo;
      }
      if (rank < k) {
        const uint32_t fk = (uint32_t)(key >> 32);
        const uint32_t bits = (fk & 0x80000000u) ? (fk & 0x7fffffffu) : ~fk;
        aout[q * k + rank] = (int)(uint32_t)key;
        dout[q * k + rank] = __uint_as_float(bits);
      }
    }
    return;
  }
  bitonic_sort_u64(buf, m_pad, tid, 128, [] { __syncthreads(); });
  for (int j = tid; j < k; j += 128) {
    unsigned long long key = buf[j];
    if (key != ~0ull) {
      uint32_t fk = (uint32_t)(key >> 32);
      uint32_t bits = (fk & 0x80000000u) ? (fk & 0x7fffffffu) : ~fk;
      aout[q * k + j] = (int)(uint32_t)key;
      dout[q * k + j] = __uint_as_float(bits);
    } else {
      aout[q * k + j] = -1;
      dout[q * k + j] = __uint_as_float(0xffffffffu);
    }
  }
}

static size_t rerank_smem_bytes(int d, int m, int m_pad) {
  size_t a16 = 15;
  size_t b = ((size_t)d * 4 + a16) & ~a16;
  b += 2 * (((size_t)m * 4 + a16) & ~a16);
  if (m_pad <= 4096) b += (size_t)m_pad * 8;
  return b;
}

// chooses the cp.async row staging when the shape allows it (YAEL_B200_RR_CPA=0: the scalar path)
// and returns the dynamic shared memory the launch needs
static size_t rerank_setup(RerankArgs &A, size_t smem) {
  // Measured at the bench shape (10 k queries x 200 candidates x 128): scalar path 0.431 ms, cp.async
  // staging 0.470 ms -- a third of the instructions, but 3 instead of 10 CTAs per SM hide less of the
  // gather latency.  Opt-in (YAEL_B200_RR_CPA=1).
  static const bool on = getenv("YAEL_B200_RR_CPA") && atoi(getenv("YAEL_B200_RR_CPA")) != 0;
  static const bool pf = !(getenv("YAEL_B200_RR_PREFETCH") && atoi(getenv("YAEL_B200_RR_PREFETCH")) == 0);
  A.prefetch = pf ? 1 : 0;
  A.cpa = 0;
  A.tile_off = 0;
  const size_t off = (smem + 15) & ~(size_t)15;
  A.tile_off = (unsigned)off;
  if (on && (A.d & 3) == 0 && (((uintptr_t)A.base) & 15) == 0 && off + RR_TILE_BYTES <= 160 * 1024) {
    A.cpa = 1;
    return off + RR_TILE_BYTES;
  }
  return off + (RR_T / 32) * 32 * 33 * 4;
}

static void rerank_attrs() {
  static bool done[64] = {};
  once_per_device(done, [] {
    cudaFuncSetAttribute(k_rerank<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_rerank<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  });
}

// ------------------------------------------------------------------ engine 0
static size_t exact_chunk_rows(int nq, int nb) {
  size_t budget = (size_t)1 << 30;  // distance slab budget (bytes)
  const char *e = getenv("YAEL_B200_SLAB_MB");
  if (e && atol(e) > 0) budget = (size_t)atol(e) << 20;
  size_t rows = budget / (sizeof(float) * (size_t)nb);
  if (rows < 1) rows = 1;
  if (rows > (size_t)nq) rows = nq;
  if (rows > 64) rows &= ~(size_t)63;
  return rows;
}

size_t knn_exact_ws_bytes(int nq, int nb, int k) {
  size_t rows = exact_chunk_rows(nq, nb);
  return l2_ws_bytes(nb, nq) + Carver::need(sizeof(float) * rows * (size_t)nb) +
         kmin_ws_bytes((long)rows, k) + 1024;
}

int knn_exact(int nq, int nb, int d, int k, const float *base, const float *query,
              const float *w, int *assign, float *dis, int id_offset, void *wsp,
              cudaStream_t st) {
  Carver c(wsp);
  size_t rows = exact_chunk_rows(nq, nb);
  float *an = c.take<float>(nb);
  double *bn = c.take<double>(nq);
  float *slab = c.take<float>(rows * (size_t)nb);
  void *kws = c.take<char>(kmin_ws_bytes((long)rows, k));
  int rc;
  if ((rc = row_norms_seq(base, nb, d, d, an, nullptr, st))) return rc;
  if ((rc = row_norms_seq(query, nq, d, d, nullptr, bn, st))) return rc;
  for (long q0 = 0; q0 < nq; q0 += (long)rows) {
    long nr = nq - q0 < (long)rows ? nq - q0 : (long)rows;
    {
      ProfScope ps(5, st);
      if ((rc = l2_matrix(d, nb, nr, base, d, query + q0 * d, d, an, bn + q0, w, slab, nb, st)))
        return rc;
    }
    {
      ProfScope ps(6, st);
      if ((rc = kmin_rows(slab, nb, nb, nr, k, +1, assign + q0 * k, dis + q0 * k, id_offset,
                          k == 1 ? 1 : 0, kws, st)))
        return rc;
    }
  }
  return 0;
}

// ------------------------------------------------------------------ engine 1
__global__ void k_sqrt_max(const float *__restrict__ sq, long n, float *out) {
  float m = 0.f;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float v = sq[i];
    if (v == v) m = fmaxf(m, v);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax((int *)out, __float_as_int(sqrtf(m)));  // m >= 0
}

__global__ void k_collect_flags(const int *__restrict__ flags, int nq, int *list, int *count) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nq && flags[q]) list[atomicAdd(count, 1)] = q;
}

__global__ void k_gather_rows(const float *__restrict__ src, const int *__restrict__ rows, int n,
                              int d, float *__restrict__ dst) {
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < (long)n * d) dst[t] = src[(size_t)rows[t / d] * d + t % d];
}

__global__ void k_scatter_results(const int *__restrict__ rows, int n, int k,
                                  const int *__restrict__ a_src, const float *__restrict__ d_src,
                                  int *__restrict__ a_dst, float *__restrict__ d_dst) {
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < (long)n * k) {
    size_t o = (size_t)rows[t / k] * k + t % k;
    a_dst[o] = a_src[t];
    d_dst[o] = d_src[t];
  }
}

// thr[q] = the j-th smallest sampled score (kmin_rows output), +inf when the sample held
// fewer than j rows
__global__ void k_threshold_from_kmin(const int *__restrict__ idx, const float *__restrict__ vals,
                                      int nq, int j, float *__restrict__ thr) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  thr[q] = idx[(size_t)q * j + (j - 1)] >= 0 ? vals[(size_t)q * j + (j - 1)]
                                              : __uint_as_float(0x7f800000u);
}

// ------------------------------------------------------------------ block radix select
// The key of 0-based rank `rank` among the keys the CTA's threads hold in registers (kreg;
// `absent` marks unused slots; rank < number of present keys).  MSB radix select with 8-bit
// digits that STARTS AT THE HIGHEST BIT IN WHICH THE KEYS DIFFER: scores of one query share
// sign, exponent and often the top mantissa bits, and a first pass over those bits would push
// every key through one or two shared-memory atomics addresses.  For 64-bit keys (score key
// << 32 | position) the all-zero bits between the position field (low_bits wide) and bit 32
// are skipped the same way.
struct SelectSmem {
  int hist[256];
  unsigned long long red[2 * 8];
  unsigned long long prefix;
  int rank;
};

template <int T, int PER, typename K>
__device__ __forceinline__ K block_select_rank(const K (&kreg)[PER], K absent, int rank, int low_bits,
                                               SelectSmem &sm) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr bool WIDE = sizeof(K) == 8;
  K mn = absent, mx = 0;
  bool any = false;
#pragma unroll
  for (int j = 0; j < PER; j++)
    if (kreg[j] != absent) {
      mn = any ? (kreg[j] < mn ? kreg[j] : mn) : kreg[j];
      mx = kreg[j] > mx ? kreg[j] : mx;
      any = true;
    }
  unsigned long long a = any ? (unsigned long long)mn : ~0ull, b = any ? (unsigned long long)mx : 0ull;
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    unsigned long long a2 = __shfl_xor_sync(0xffffffffu, a, o), b2 = __shfl_xor_sync(0xffffffffu, b, o);
    a = a2 < a ? a2 : a;
    b = b2 > b ? b2 : b;
  }
  __syncthreads();
  if (lane == 0) {
    sm.red[warp] = a;
    sm.red[8 + warp] = b;
  }
  __syncthreads();
  a = ~0ull;
  b = 0;
#pragma unroll
  for (int w = 0; w < T / 32; w++) {
    a = sm.red[w] < a ? sm.red[w] : a;
    b = sm.red[8 + w] > b ? sm.red[8 + w] : b;
  }
  const unsigned long long diff = a ^ b;
  if (diff == 0) return (K)a;
  const int top = 63 - __clzll((long long)diff);
  K decided = top >= (int)(8 * sizeof(K)) - 1 ? (K)0 : (K)(~(K)0 << (top + 1));
  K prefix = (K)b & decided;
  int floor_bit = (WIDE && top >= 32) ? 32 : 0;
  int shift = max(top - 7, floor_bit);
  if (WIDE && top < 32) shift = max(0, min(top + 1, low_bits) - 8);
  while (true) {
    for (int h = tid; h < 256; h += T) sm.hist[h] = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PER; j++)
      if (kreg[j] != absent && (kreg[j] & decided) == prefix)
        atomicAdd(&sm.hist[(int)(kreg[j] >> shift) & 255], 1);
    __syncthreads();
    if (warp == 0) {
      int h[8], ssum = 0;
#pragma unroll
      for (int c = 0; c < 8; c++) {
        h[c] = sm.hist[lane * 8 + c];
        ssum += h[c];
      }
      int inc = ssum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      int below = inc - ssum;
      if (below <= rank && rank < inc) {
        int c = 0;
        while (below + h[c] <= rank) below += h[c++];
        sm.prefix = (unsigned long long)(lane * 8 + c);
        sm.rank = rank - below;
      }
    }
    __syncthreads();
    const K dmask = (K)0xff << shift;
    prefix = (prefix & ~dmask) | ((K)sm.prefix << shift);
    decided |= dmask;
    rank = sm.rank;
    if (shift == floor_bit) {
      if (floor_bit == 0) break;
      floor_bit = 0;  // continue in the position field
      shift = max(0, low_bits - 8);
    } else {
      shift = max(shift - 8, floor_bit);
    }
  }
  return prefix;
}

// ------------------------------------------------------------------ shortlist merge
// One CTA per query: the union of the query's shortlists (cnt[q][l] entries each, a few dozen
// with sampled thresholds) -> positions of its kp smallest TF32 scores (ties by position),
// and/or the kp-th smallest score itself.  Replaces "pre-fill every slot with +inf, then
// radix-select over lists * k' slots": only the published entries are ever touched.
//   gather   lists are placed at the exclusive prefix sum of their counts (deterministic)
//   select   MSB radix select (8-bit digits) of the kp-th smallest 64-bit key (score key << 32
//            | position; keys are unique), every thread holding its keys in registers -- a
//            full bitonic sort of ~3k' keys per query is bound by shared-memory bandwidth and
//            costs as much as the radix select over all slots it was meant to replace
//   emit     ordered compaction of the keys <= that key
// Unions larger than the buffer (or more lists than ML_LISTS) are folded in list by list with
// a sort (keep the kp best, append): rare, adversarial inputs only.
constexpr int ML_T = 128;
constexpr int ML_CAP = 2048;     // needs kp + k' <= ML_CAP
constexpr int ML_PER = ML_CAP / ML_T;
constexpr int ML_LISTS = 1024;

__device__ __forceinline__ int ml_block_excl_scan(int v, int *wtot, int &total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // wtot may still be read from a previous call
  if (lane == 31) wtot[warp] = inc;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < ML_T / 32; w++) {
    if (w < warp) base += wtot[w];
    tot += wtot[w];
  }
  total = tot;
  return base + inc - v;
}

__global__ void __launch_bounds__(ML_T)
k_merge_lists(const int *__restrict__ cnt, const float *__restrict__ score, int lists, int kprime,
              int kp, int *__restrict__ sel, float *__restrict__ thr_out) {
  __shared__ unsigned long long buf[ML_CAP];
  __shared__ int loff[ML_LISTS];
  __shared__ SelectSmem ssm;
  __shared__ int wtot[ML_T / 32];
  const int q = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int *c = cnt + (size_t)q * lists;
  const float *sc = score + (size_t)q * lists * kprime;
  const float inf = __uint_as_float(0x7f800000u);

  // ---- gather
  const int per = (lists + ML_T - 1) / ML_T;  // contiguous lists per thread
  int mine = 0;
  for (int l = tid * per; l < min(lists, (tid + 1) * per); l++) mine += c[l];
  int total;
  int off = ml_block_excl_scan(mine, wtot, total);
  int n;
  if (total <= ML_CAP && lists <= ML_LISTS) {
    for (int l = tid * per; l < min(lists, (tid + 1) * per); l++) {
      loff[l] = off;
      off += c[l];
    }
    __syncthreads();
    for (int l = warp; l < lists; l += ML_T / 32) {
      const int m = c[l], b = loff[l];
      for (int e = lane; e < m; e += 32)
        buf[b + e] = ((unsigned long long)float_key(sc[(size_t)l * kprime + e]) << 32) |
                     (unsigned)(l * kprime + e);
    }
    n = total;
    __syncthreads();
  } else {
    n = 0;
    for (int l = 0; l < lists; l++) {
      const int m = c[l];
      if (m == 0) continue;
      if (n + m > ML_CAP) {  // keep the kp best so far
        const int n_pad = pow2_ceil(n);
        for (int j = n + tid; j < n_pad; j += ML_T) buf[j] = ~0ull;
        __syncthreads();
        bitonic_sort_u64(buf, n_pad, tid, ML_T, [] { __syncthreads(); });
        n = min(n, kp);
      }
      for (int e = tid; e < m; e += ML_T)
        buf[n + e] = ((unsigned long long)float_key(sc[(size_t)l * kprime + e]) << 32) |
                     (unsigned)(l * kprime + e);
      n += m;
      __syncthreads();
    }
  }

  // ---- select: kstar = the kp-th smallest key (only when more than kp keys are present)
  unsigned long long kreg[ML_PER];
#pragma unroll
  for (int j = 0; j < ML_PER; j++) {
    const int i = j * ML_T + tid;
    kreg[j] = i < n ? buf[i] : ~0ull;  // ~0 is never a key: its position field would be 2^32-1
  }
  unsigned long long kstar = ~0ull - 1;  // n <= kp: everything is selected
  if (n > kp) {
    int low_bits = 1;
    while (low_bits < 32 && ((long)1 << low_bits) < (long)lists * kprime) low_bits++;
    kstar = block_select_rank<ML_T, ML_PER, unsigned long long>(kreg, ~0ull, kp - 1, low_bits, ssm);
  }

  // ---- emit
  if (sel) {
    int mysel = 0;
#pragma unroll
    for (int j = 0; j < ML_PER; j++) mysel += (kreg[j] <= kstar);
    int tot2;
    int o = ml_block_excl_scan(mysel, wtot, tot2);
    int *out = sel + (size_t)q * kp;
#pragma unroll
    for (int j = 0; j < ML_PER; j++)
      if (kreg[j] <= kstar) out[o++] = (int)(uint32_t)kreg[j];
    for (int j = tot2 + tid; j < kp; j += ML_T) out[j] = -1;
  }
  if (thr_out && tid == 0) {
    float t = inf;
    if (n >= kp) {
      unsigned long long kk = kstar;
      if (n == kp) {  // the largest key present
        kk = 0;
        for (int i = 0; i < n; i++) kk = buf[i] > kk ? buf[i] : kk;
      }
      const uint32_t fk = (uint32_t)(kk >> 32);
      t = __uint_as_float((fk & 0x80000000u) ? (fk & 0x7fffffffu) : ~fk);
    }
    thr_out[q] = t;
  }
}

// thr[q] = the j-th smallest of vals[q][0..n) (n <= RK_T * RK_PER; +inf when n < j): MSB radix
// select on the order-preserving keys, every thread holding its values in registers.
constexpr int RK_T = 256;
constexpr int RK_PER = 16;

__global__ void __launch_bounds__(RK_T)
k_row_kth(const float *__restrict__ vals, long ld, int n, int j, float *__restrict__ thr) {
  __shared__ SelectSmem ssm;
  const int q = blockIdx.x, tid = threadIdx.x;
  const float *row = vals + (size_t)q * ld;
  uint32_t kreg[RK_PER];
#pragma unroll
  for (int i = 0; i < RK_PER; i++) {
    const int c = i * RK_T + tid;
    kreg[i] = c < n ? float_key(row[c]) : 0xffffffffu;  // absent = the NaN key (sorts last)
  }
  if (n < j) {
    if (tid == 0) thr[q] = __uint_as_float(0x7f800000u);
    return;
  }
  const uint32_t prefix = block_select_rank<RK_T, RK_PER, uint32_t>(kreg, 0xffffffffu, j - 1, 0, ssm);
  if (tid == 0) thr[q] = __uint_as_float((prefix & 0x80000000u) ? (prefix & 0x7fffffffu) : ~prefix);
}

int row_kth_max_n() { return RK_T * RK_PER; }
int row_kth(const float *vals, long ld, int nrow, int n, int j, float *thr, cudaStream_t st) {
  if (nrow <= 0) return 0;
  if (n > RK_T * RK_PER) return fail(3, "row_kth: %d values per row exceed %d", n, RK_T * RK_PER);
  k_row_kth<<<nrow, RK_T, 0, st>>>(vals, ld, n, j, thr);
  YB_LAUNCH_CHECK();
  return 0;
}

// TF32 operands keep 10 explicit mantissa bits; the hardware drops (or rounds) the rest, so
// each operand carries a relative error < 2^-10 and each product < 2^-9 (+2^-20); with
// Cauchy-Schwarz the score |b|^2 - 2<q,b> is off by at most 2 * 2^-9 * |q||b|.  5 % head
// room covers the FP32 accumulation inside the tensor core, the norm rounding and the 7 mantissa
// bits of a listed score that carry its column index (relative 2^-16).
static const float kTf32ErrScale = 1.05f / 256.0f;
// FP16 operands are rounded to nearest: relative error <= 2^-11 per operand, 2^-10 per product,
// 2 * 2^-10 |q||b| on the score (plus the absolute term for sub-normal values, scal[5])
static const float kF16ErrScale = 1.05f / 512.0f;

// ------------------------------------------------------------------ centring
// Squared L2 distances are translation invariant: |q-b|^2 = |(q-mu)-(b-mu)|^2 for ANY mu.  The
// tensor pass therefore runs on copies shifted by the column mean of the database: operand
// norms shrink, and with them the TF32 error bound err*|q-mu||b-mu| that sizes the shortlist
// margin (for k-means on unstructured data this is the difference between a margin that
// covers every centroid and one that covers one or two).  The exact re-rank always reads the
// caller's original rows.
constexpr int CM_ROWS = 256;    // rows per block in the column-mean reduction
constexpr int CM_BLOCKS = 128;  // the mean is taken over at most CM_BLOCKS * CM_ROWS sampled rows
                                // (any shift vector is valid; a sampled mean is as good)

__global__ void __launch_bounds__(256)
k_col_partial(const float *__restrict__ x, long n, int d, long block_step,
              float *__restrict__ psum, int *__restrict__ pcnt) {
  // thread t owns columns t, t+256, ...; sequential over the block's rows (deterministic)
  const long r0 = (long)blockIdx.x * block_step, r1 = min(n, r0 + CM_ROWS);
  for (int c = threadIdx.x; c < d; c += 256) {
    float s = 0.f;
    int m = 0;
    for (long r = r0; r < r1; r++) {
      float v = x[r * d + c];
      if (isfinite(v)) {
        s += v;
        m++;
      }
    }
    psum[(size_t)blockIdx.x * d + c] = s;
    pcnt[(size_t)blockIdx.x * d + c] = m;
  }
}

__global__ void k_col_final(const float *__restrict__ psum, const int *__restrict__ pcnt, int nblk,
                            int d, float *__restrict__ mu) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  double s = 0.0;
  long m = 0;
  for (int b = 0; b < nblk; b++) {
    s += (double)psum[(size_t)b * d + c];
    m += pcnt[(size_t)b * d + c];
  }
  mu[c] = m > 0 ? (float)(s / (double)m) : 0.f;
}

// Logical -> physical row index: identity, or "the first `rows` rows of every `stride` rows"
// (the sample tiles of a database that is still arriving from the host).
struct RowMap {
  long rows, stride;  // rows == 0: identity
  __host__ __device__ long operator()(long r) const {
    return rows ? (r / rows) * stride + (r % rows) : r;
  }
  __host__ __device__ long count(long n) const {  // logical rows that map below n (upper bound)
    return rows ? ((n + stride - 1) / stride) * rows : n;
  }
};

// one warp per row: out[r][0..dpad) = x[r][c] - mu[c] (zero padded), optionally norm[r] = |out[r]|^2
__global__ void __launch_bounds__(256)
k_center_rows(const float *__restrict__ x, long n, int d, int dpad, const float *__restrict__ mu,
              float *__restrict__ out, float *__restrict__ norm, RowMap rm) {
  const long rl = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (rl >= rm.count(n)) return;
  const long r = rm(rl);
  if (r >= n) return;
  float s = 0.f;
  for (int c = lane; c < dpad; c += 32) {
    float v = c < d ? __fsub_rn(x[r * d + c], __ldg(mu + c)) : 0.f;
    out[r * dpad + c] = v;
    s = fmaf(v, v, s);
  }
  if (norm) {
    s = warp_sum(s);
    if (lane == 0) norm[r] = s;
  }
}

// d % 4 == 0 (pitch == d): one warp per 4 rows, 16-byte accesses, the 4 row loads in flight
// together.  Same arithmetic per element as k_center_rows (the norm is a TF32-pass operand:
// only its order of summation differs, which the error model covers).
__global__ void __launch_bounds__(256)
k_center_rows_v4(const float *__restrict__ x, long n, int d, const float *__restrict__ mu,
                 float *__restrict__ out, float *__restrict__ norm, RowMap rm) {
  const long rl0 = ((long)blockIdx.x * 8 + (threadIdx.x >> 5)) * 4;
  if (rl0 >= rm.count(n)) return;
  const long r0 = rm(rl0);  // blocks are multiples of 4 rows: the 4 rows stay consecutive
  const int lane = threadIdx.x & 31;
  const int d4 = d >> 2;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = lane; c < d4; c += 32) {
    const float4 m = __ldg(reinterpret_cast<const float4 *>(mu) + c);
    float4 v[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (r0 + i < n) v[i] = __ldg(reinterpret_cast<const float4 *>(x + (r0 + i) * d) + c);
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (r0 + i < n) {
        float4 o;
        o.x = __fsub_rn(v[i].x, m.x);
        o.y = __fsub_rn(v[i].y, m.y);
        o.z = __fsub_rn(v[i].z, m.z);
        o.w = __fsub_rn(v[i].w, m.w);
        reinterpret_cast<float4 *>(out + (r0 + i) * d)[c] = o;
        s[i] = fmaf(o.x, o.x, s[i]);
        s[i] = fmaf(o.y, o.y, s[i]);
        s[i] = fmaf(o.z, o.z, s[i]);
        s[i] = fmaf(o.w, o.w, s[i]);
      }
  }
  if (norm) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      float t = warp_sum(s[i]);
      if (lane == 0 && r0 + i < n) norm[r0 + i] = t;
    }
  }
}

// scal[7] = |mu| (norm of the centring vector): enters the reference-rounding term of the
// certificate and of the k = 1 margin
__global__ void k_mu_norm(const float *__restrict__ mu, int d, float *__restrict__ out) {
  float s = 0.f;
  for (int c = threadIdx.x; c < d; c += 32) {
    const float v = mu[c];
    if (isfinite(v)) s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  if (threadIdx.x == 0) *out = sqrtf(s);
}

static void launch_center_rows(const float *x, long n, int d, int dpad, const float *mu, float *out,
                               float *norm, cudaStream_t st, RowMap rm = RowMap{0, 0}) {
  if (n <= 0) return;
  const long nl = rm.count(n);
  if (dpad == d && (((uintptr_t)x | (uintptr_t)out | (uintptr_t)mu) & 15) == 0 && (rm.rows % 4) == 0)
    k_center_rows_v4<<<(unsigned)((nl + 31) / 32), 256, 0, st>>>(x, n, d, mu, out, norm, rm);
  else
    k_center_rows<<<(unsigned)((nl + 7) / 8), 256, 0, st>>>(x, n, d, dpad, mu, out, norm, rm);
  count_launch();
}

static size_t center_ws_bytes(long nb, int d) {
  (void)nb;
  return 2 * Carver::need(sizeof(float) * (size_t)CM_BLOCKS * d) + Carver::need(sizeof(float) * d);
}

// base_c / query_c: centred copies with row pitch dpad (multiple of 4 floats); bnorm[nb] = |b-mu|^2
static int center_operands(int nq, int nb, int d, int dpad, const float *base, const float *query,
                           float *base_c, float *query_c, float *bnorm, float *mu_norm, void *ws,
                           cudaStream_t st) {
  Carver c(ws);
  int nblk = (int)(((long)nb + CM_ROWS - 1) / CM_ROWS);
  if (nblk > CM_BLOCKS) nblk = CM_BLOCKS;
  long step = nblk > 0 ? (long)nb / nblk : CM_ROWS;
  if (step < CM_ROWS) step = CM_ROWS;
  float *psum = c.take<float>((size_t)CM_BLOCKS * d);
  int *pcnt = c.take<int>((size_t)CM_BLOCKS * d);
  float *mu = c.take<float>(d);
  k_col_partial<<<nblk, 256, 0, st>>>(base, nb, d, step, psum, pcnt);
  YB_LAUNCH_CHECK();
  k_col_final<<<(d + 127) / 128, 128, 0, st>>>(psum, pcnt, nblk, d, mu);
  YB_LAUNCH_CHECK();
  k_mu_norm<<<1, 32, 0, st>>>(mu, d, mu_norm);
  YB_LAUNCH_CHECK();
  launch_center_rows(base, nb, d, dpad, mu, base_c, bnorm, st);
  launch_center_rows(query, nq, d, dpad, mu, query_c, nullptr, st);
  YB_CUDA(cudaGetLastError());
  return 0;
}


// ------------------------------------------------------------------ FP16 operands
// The tensor pass can read its operands as FP16 (kind::f16) instead of FP32-as-TF32: the same 10
// explicit mantissa bits (and round-to-nearest instead of truncation), half the operand bytes,
// i.e. half as many MMAs, shared-memory operand reads and accumulator passes per tile.  FP16's
// narrow exponent range is handled by an exact power-of-two scale 2^sigma chosen from a sample
// of the centred data (8x head room); the kernel's epilogue multiplies the accumulator by
// -2^(1-2 sigma) in the same FMA that adds |b|^2, so scores, thresholds and the certificate stay
// in the caller's units.  A finite value that still overflows FP16 raises a flag and the caller
// repeats the pass with TF32 operands.  scal[] layout (device floats): [0] max |b-mu|,
// [1] flag count (int), [2] 2^sigma, [3] -2^(1-2 sigma), [4] overflow flag (int),
// [5] absolute-error coefficient, [6] sampled max |x|.
__global__ void __launch_bounds__(256)
k_absmax_sample(const float *__restrict__ x, long n, int d, long block_step, float *__restrict__ out) {
  // max |x| over the block's CM_ROWS sampled rows (contiguous in memory: flat, coalesced, 8 loads
  // in flight per thread).  |x - mu| <= max|x| + max|mu|: k_pick_scale adds the second term.
  const long r0 = (long)blockIdx.x * block_step, r1 = min(n, r0 + CM_ROWS);
  const float *p = x + r0 * d;
  const long cnt = (r1 - r0) * d;
  float m = 0.f;
  for (long i0 = threadIdx.x; i0 < cnt; i0 += 256 * 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = i0 + j * 256 < cnt ? fabsf(p[i0 + j * 256]) : 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (isfinite(v[j])) m = fmaxf(m, v[j]);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax((int *)out, __float_as_int(m));  // m >= 0
}

__global__ void k_pick_scale(float *__restrict__ scal, int d, const float *__restrict__ mu) {
  float mm = 0.f;
  for (int c = 0; c < d; c++)
    if (isfinite(mu[c])) mm = fmaxf(mm, fabsf(mu[c]));
  const float m = scal[6] + mm;  // >= max |x - mu| over the sample
  int sigma = 0;
  if (m > 0.f && isfinite(m)) {
    const float x = 8188.0f / m;
    int e = 41;
    if (isfinite(x)) frexpf(x, &e);  // 8188 / m = f * 2^e, f in [0.5, 1): 2^(e-1) <= 8188 / m
    sigma = e - 1;
  }
  sigma = max(-40, min(40, sigma));
  scal[2] = ldexpf(1.0f, sigma);
  scal[3] = -ldexpf(1.0f, 1 - 2 * sigma);
  // values below FP16's normal range round with an ABSOLUTE error <= 2^-25 (scaled units); summed
  // over a dot product and brought back to the caller's units that is at most
  // coef * (|q-mu| + |b-mu|), coef = 2 sqrt(d) 2^-25 2^-sigma (2x for the -2<q,b> factor)
  scal[5] = 2.0f * sqrtf((float)d) * ldexpf(1.0f, -24 - sigma);
}

__device__ __forceinline__ unsigned pack_h2(float a, float b, float sc, bool &over) {
  const float x = a * sc, y = b * sc;
  over |= (fabsf(x) > 65504.0f && fabsf(a) < __int_as_float(0x7f800000)) ||
          (fabsf(y) > 65504.0f && fabsf(b) < __int_as_float(0x7f800000));
  const __half2 h = __floats2half2_rn(x, y);
  return *reinterpret_cast<const unsigned *>(&h);
}

// The 16 extra K elements behind the dh data elements of an FP16 operand row (one MMA K step):
// database rows (role 1) carry -beta split into three FP16 pieces, beta = 2^(2 sigma - 1) |b-mu|^2
// / 2^15, queries (role 2) carry 2^15 three times, so that the contraction itself yields
// 2^(2 sigma) (<q,b> - |b|^2 / 2) and the epilogue needs no |b|^2.  The split is exact (3 x 11
// bits cover the 24-bit float); beta > 65504 raises the overflow flag like any other value.
constexpr int kNfExtra = 16;
__device__ __forceinline__ void write_row_extras(__half *dst, int role, float norm, float sc,
                                                 int *oflag) {
  uint4 a = make_uint4(0u, 0u, 0u, 0u), b = a;
  if (role == 1) {
    const float w = norm * sc * sc * (0.5f / 32768.0f);
    if (w > 65504.0f && w < __int_as_float(0x7f800000)) *oflag = 1;
    const __half h0 = __float2half_rn(w);
    const float r1 = w - __half2float(h0);
    const __half h1 = __float2half_rn(r1);
    const float r2 = r1 - __half2float(h1);
    const __half h2 = __float2half_rn(r2);
    const unsigned short n0 = __half_as_ushort(__hneg(h0)), n1 = __half_as_ushort(__hneg(h1)),
                         n2 = __half_as_ushort(__hneg(h2));
    a.x = (unsigned)n0 | ((unsigned)n1 << 16);
    a.y = (unsigned)n2;
  } else if (role == 2) {
    a.x = 0x78007800u;  // 32768.0, 32768.0
    a.y = 0x00007800u;  // 32768.0, 0
  }
  reinterpret_cast<uint4 *>(dst)[0] = a;
  reinterpret_cast<uint4 *>(dst)[1] = b;
}

// rows [n, n_pad) of the FP16 database copy: zeros with a NaN |b|^2 element (never admitted)
__global__ void k_fill_pad_rows_h(__half *__restrict__ out_h, long n, long n_pad, int pitch, int dh) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (n_pad - n) * pitch) return;
  const int c = (int)(t % pitch);
  out_h[n * pitch + t] = __ushort_as_half(c == dh ? (unsigned short)0x7E00 : (unsigned short)0);
}

// one warp per row, any d: out_h[r][0..dh) = fp16((x - mu) * 2^sigma) zero padded, row pitch
// `pitch` halfs (dh, or dh + 16 with the extras above when role != 0); optional FP32 copy
// out_f[r][0..df) and norm[r] = |x - mu|^2 (FP32 values, before the conversion)
__global__ void __launch_bounds__(256)
k_center_rows_h(const float *__restrict__ x, long n, int d, int dh, int df, int pitch, int role,
                const float *__restrict__ mu, const float *__restrict__ scal,
                __half *__restrict__ out_h, float *__restrict__ out_f, float *__restrict__ norm,
                int *__restrict__ oflag) {
  const long r = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n) return;
  const float sc = scal[2];
  const int dm = dh > df ? dh : df;
  float s = 0.f;
  bool over = false;
  for (int c = lane; c < dm; c += 32) {
    const float v = c < d ? __fsub_rn(x[r * d + c], __ldg(mu + c)) : 0.f;
    if (c < dh) {
      const float w = v * sc;
      over |= fabsf(w) > 65504.0f && fabsf(v) < __int_as_float(0x7f800000);
      out_h[r * pitch + c] = __float2half_rn(w);
    }
    if (out_f && c < df) out_f[r * df + c] = v;
    s = fmaf(v, v, s);
  }
  if (over) *oflag = 1;
  s = warp_sum(s);
  if (lane == 0) {
    if (norm) norm[r] = s;
    if (role) write_row_extras(out_h + r * pitch + dh, role, s, sc, oflag);
  }
}

// d % 8 == 0, 16-byte aligned: one warp per 4 rows, 2 x 16-byte loads and one 16-byte store per
// 8 columns
__global__ void __launch_bounds__(256)
k_center_rows_h8(const float *__restrict__ x, long n, int d, int pitch, int role,
                 const float *__restrict__ mu, const float *__restrict__ scal,
                 __half *__restrict__ out_h, float *__restrict__ out_f, float *__restrict__ norm,
                 int *__restrict__ oflag) {
  const long r0 = ((long)blockIdx.x * 8 + (threadIdx.x >> 5)) * 4;
  if (r0 >= n) return;
  const int lane = threadIdx.x & 31;
  const int d8 = d >> 3;
  const float sc = scal[2];
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  bool over = false;
  for (int f = lane; f < 4 * d8; f += 32) {
    const int i = f / d8, c = f - i * d8;
    if (r0 + i >= n) continue;
    const float4 m0 = __ldg(reinterpret_cast<const float4 *>(mu) + 2 * c);
    const float4 m1 = __ldg(reinterpret_cast<const float4 *>(mu) + 2 * c + 1);
    const float4 a = __ldg(reinterpret_cast<const float4 *>(x + (r0 + i) * d) + 2 * c);
    const float4 b = __ldg(reinterpret_cast<const float4 *>(x + (r0 + i) * d) + 2 * c + 1);
    float4 u, w;
    u.x = __fsub_rn(a.x, m0.x); u.y = __fsub_rn(a.y, m0.y);
    u.z = __fsub_rn(a.z, m0.z); u.w = __fsub_rn(a.w, m0.w);
    w.x = __fsub_rn(b.x, m1.x); w.y = __fsub_rn(b.y, m1.y);
    w.z = __fsub_rn(b.z, m1.z); w.w = __fsub_rn(b.w, m1.w);
    if (out_f) {
      reinterpret_cast<float4 *>(out_f + (r0 + i) * d)[2 * c] = u;
      reinterpret_cast<float4 *>(out_f + (r0 + i) * d)[2 * c + 1] = w;
    }
    float t = 0.f;
    t = fmaf(u.x, u.x, t); t = fmaf(u.y, u.y, t); t = fmaf(u.z, u.z, t); t = fmaf(u.w, u.w, t);
    t = fmaf(w.x, w.x, t); t = fmaf(w.y, w.y, t); t = fmaf(w.z, w.z, t); t = fmaf(w.w, w.w, t);
#pragma unroll
    for (int ii = 0; ii < 4; ii++)
      if (ii == i) s[ii] += t;
    uint4 o;
    o.x = pack_h2(u.x, u.y, sc, over);
    o.y = pack_h2(u.z, u.w, sc, over);
    o.z = pack_h2(w.x, w.y, sc, over);
    o.w = pack_h2(w.z, w.w, sc, over);
    reinterpret_cast<uint4 *>(out_h + (r0 + i) * pitch)[c] = o;
  }
  if (over) *oflag = 1;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float t = warp_sum(s[i]);
    if (lane == 0 && r0 + i < n) {
      if (norm) norm[r0 + i] = t;
      if (role) write_row_extras(out_h + (r0 + i) * pitch + d, role, t, sc, oflag);
    }
  }
}

// role 0: plain rows of dh halfs; 1 / 2: database / query rows of dh + 16 halfs (extras)
static void launch_center_rows_h(const float *x, long n, int d, int dh, int df, int role,
                                 const float *mu, const float *scal, __half *out_h, float *out_f,
                                 float *norm, cudaStream_t st) {
  if (n <= 0) return;
  int *oflag = (int *)(scal + 4);
  const int pitch = role ? dh + kNfExtra : dh;
  const bool fast = dh == d && (d & 7) == 0 && (!out_f || df == d) &&
                    (((uintptr_t)x | (uintptr_t)out_h | (uintptr_t)mu | (uintptr_t)out_f) & 15) == 0;
  if (fast)
    k_center_rows_h8<<<(unsigned)((n + 31) / 32), 256, 0, st>>>(x, n, d, pitch, role, mu, scal, out_h,
                                                                out_f, norm, oflag);
  else
    k_center_rows_h<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(x, n, d, dh, df, pitch, role, mu, scal,
                                                              out_h, out_f, norm, oflag);
  count_launch();
}

// FP16 counterpart of center_operands: base_h / query_h with dh data halfs (multiple of 8) plus
// the 16 extra elements per row (pitch dh + 16); base_h holds tf32_padded_rows(nb) rows, the
// padding marked NaN; optional FP32 centred queries query_c (pitch df) and their squared norms
// qcnorm; bnorm[nb].  scal must have been zeroed by the caller.
static int center_operands_h(int nq, int nb, int d, int dh, int df, const float *base,
                             const float *query, __half *base_h, __half *query_h, float *query_c,
                             float *bnorm, float *qcnorm, float *scal, void *ws, cudaStream_t st,
                             bool fold) {
  Carver c(ws);
  int nblk = (int)(((long)nb + CM_ROWS - 1) / CM_ROWS);
  if (nblk > CM_BLOCKS) nblk = CM_BLOCKS;
  long step = nblk > 0 ? (long)nb / nblk : CM_ROWS;
  if (step < CM_ROWS) step = CM_ROWS;
  float *psum = c.take<float>((size_t)CM_BLOCKS * d);
  int *pcnt = c.take<int>((size_t)CM_BLOCKS * d);
  float *mu = c.take<float>(d);
  k_col_partial<<<nblk, 256, 0, st>>>(base, nb, d, step, psum, pcnt);
  YB_LAUNCH_CHECK();
  k_col_final<<<(d + 127) / 128, 128, 0, st>>>(psum, pcnt, nblk, d, mu);
  YB_LAUNCH_CHECK();
  k_mu_norm<<<1, 32, 0, st>>>(mu, d, scal + 7);
  YB_LAUNCH_CHECK();
  k_absmax_sample<<<nblk, 256, 0, st>>>(base, nb, d, step, scal + 6);
  YB_LAUNCH_CHECK();
  int qblk = (int)(((long)nq + CM_ROWS - 1) / CM_ROWS);
  if (qblk > CM_BLOCKS) qblk = CM_BLOCKS;
  long qstep = qblk > 0 ? (long)nq / qblk : CM_ROWS;
  if (qstep < CM_ROWS) qstep = CM_ROWS;
  k_absmax_sample<<<qblk, 256, 0, st>>>(query, nq, d, qstep, scal + 6);
  YB_LAUNCH_CHECK();
  k_pick_scale<<<1, 1, 0, st>>>(scal, d, mu);
  YB_LAUNCH_CHECK();
  launch_center_rows_h(base, nb, d, dh, dh, fold ? 1 : 0, mu, scal, base_h, nullptr, bnorm, st);
  launch_center_rows_h(query, nq, d, dh, df, fold ? 2 : 0, mu, scal, query_h, query_c, qcnorm, st);
  if (fold) {
    const long pad = tf32_padded_rows(nb) - nb, tot = pad * (dh + kNfExtra);
    if (tot > 0) {
      k_fill_pad_rows_h<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(base_h, nb, tf32_padded_rows(nb),
                                                                       dh + kNfExtra, dh);
      count_launch();
    }
  }
  YB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------- compute_cross_distances, tensor cores
// Split-precision FP16 operands for the FULL distance matrix (yael/nn.c:100-129 within the north
// star's 1e-5): a centred, scaled coordinate w = hi + lo (+ r, |r| <= 2^-22 |w|), both halves FP16, and
//     <a, b> ~ <a_hi, b_hi> + <a_hi, b_lo> + <a_lo, b_hi>
// is ONE contraction over K = 3 dh: query rows [hi | hi | lo], database rows [hi | lo | hi].  The 16
// extra K elements carry BOTH norms: database rows -beta pieces x query 2^15 (as write_row_extras)
// and query rows -alpha pieces x database 2^15, alpha = 2^(2 sigma - 1) |a-mu|^2 / 2^15.  The
// accumulator is then 2^(2 sigma) (<a,b> - |a|^2/2 - |b|^2/2) and the distance asc * acc: the
// epilogue only scales and stores.
__global__ void __launch_bounds__(256)
k_split_rows_h(const float *__restrict__ x, long n, int d, int dh, int role,
               const float *__restrict__ mu, const float *__restrict__ scal,
               __half *__restrict__ out_h, int *__restrict__ oflag) {
  const long r = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n) return;
  const float sc = scal[2];
  const int pitch = 3 * dh + kNfExtra;
  __half *row = out_h + r * pitch;
  float s = 0.f;
  bool over = false;
  for (int c = lane; c < dh; c += 32) {
    const float v = c < d ? __fsub_rn(x[r * d + c], __ldg(mu + c)) : 0.f;
    const float w = v * sc;
    over |= fabsf(w) > 65504.0f && fabsf(v) < __int_as_float(0x7f800000);
    const __half hi = __float2half_rn(w);
    const __half lo = __float2half_rn(w - __half2float(hi));
    row[c] = hi;
    row[dh + c] = role == 1 ? lo : hi;
    row[2 * dh + c] = role == 1 ? hi : lo;
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  if (lane == 0) {
    const float w = s * sc * sc * (0.5f / 32768.0f);
    if (w > 65504.0f && w < __int_as_float(0x7f800000)) over = true;
    const __half h0 = __float2half_rn(w);
    const float r1 = w - __half2float(h0);
    const __half h1 = __float2half_rn(r1);
    const __half h2 = __float2half_rn(r1 - __half2float(h1));
    const __half big = __ushort_as_half((unsigned short)0x7800);  // 32768.0
    __half *e = row + 3 * dh;
    const int own = role == 1 ? 0 : 3, other = role == 1 ? 3 : 0;
    e[own + 0] = __hneg(h0); e[own + 1] = __hneg(h1); e[own + 2] = __hneg(h2);
    e[other + 0] = big; e[other + 1] = big; e[other + 2] = big;
    for (int j = 6; j < kNfExtra; j++) e[j] = __ushort_as_half((unsigned short)0);
  }
  if (over) *oflag = 1;
}

int cross_l2_tensor(int d, int na, int nb, const float *a, const float *b, float *out, long ldd,
                    cudaStream_t st) {
  if (d < 1 || na <= 128 || nb < 1) return -1000;  // the 2-SM kernel pairs two query tiles
  const int dh = (d + 15) & ~15, dop = 3 * dh + kNfExtra;
  Tf32Plan plan = tf32_plan_tiles(na, tf32_tiles(nb), dop, 1, 3);
  if (!plan.ok || plan.pair != 2) return -1000;
  const long padded = tf32_padded_rows(nb);
  const size_t need = Carver::need(2 * (size_t)padded * dop) + Carver::need(2 * (size_t)na * dop) +
                      Carver::need(64) + center_ws_bytes(nb, d) + Carver::need(plan.ws_bytes) + 1024;
  ScratchScope ws(need, st);
  Carver c(ws.p);
  __half *b_h = (__half *)c.take<char>(2 * (size_t)padded * dop);
  __half *a_h = (__half *)c.take<char>(2 * (size_t)na * dop);
  float *scal = c.take<float>(16);
  void *cws = c.take<char>(center_ws_bytes(nb, d));
  void *tws = c.take<char>(plan.ws_bytes);
  YB_CUDA(cudaMemsetAsync(scal, 0, 64, st));
  {
    ProfScope ps(0, st);
    Carver cc(cws);  // column mean of the database rows and the power-of-two scale (as center_operands_h)
    int nblk = (int)(((long)nb + CM_ROWS - 1) / CM_ROWS);
    if (nblk > CM_BLOCKS) nblk = CM_BLOCKS;
    long step = nblk > 0 ? (long)nb / nblk : CM_ROWS;
    if (step < CM_ROWS) step = CM_ROWS;
    float *psum = cc.take<float>((size_t)CM_BLOCKS * d);
    int *pcnt = cc.take<int>((size_t)CM_BLOCKS * d);
    float *mu = cc.take<float>(d);
    k_col_partial<<<nblk, 256, 0, st>>>(b, nb, d, step, psum, pcnt);
    YB_LAUNCH_CHECK();
    k_col_final<<<(d + 127) / 128, 128, 0, st>>>(psum, pcnt, nblk, d, mu);
    YB_LAUNCH_CHECK();
    k_mu_norm<<<1, 32, 0, st>>>(mu, d, scal + 7);
    YB_LAUNCH_CHECK();
    // sampled max |x| with 8x head room; a value that still overflows raises the flag (-1001)
    k_absmax_sample<<<nblk, 256, 0, st>>>(b, nb, d, step, scal + 6);
    YB_LAUNCH_CHECK();
    int qblk = (int)(((long)na + CM_ROWS - 1) / CM_ROWS);
    if (qblk > CM_BLOCKS) qblk = CM_BLOCKS;
    long qstep = qblk > 0 ? (long)na / qblk : CM_ROWS;
    if (qstep < CM_ROWS) qstep = CM_ROWS;
    k_absmax_sample<<<qblk, 256, 0, st>>>(a, na, d, qstep, scal + 6);
    YB_LAUNCH_CHECK();
    k_pick_scale<<<1, 1, 0, st>>>(scal, d, mu);
    YB_LAUNCH_CHECK();
    int *oflag = (int *)(scal + 4);
    k_split_rows_h<<<(unsigned)((nb + 7) / 8), 256, 0, st>>>(b, nb, d, dh, 1, mu, scal, b_h, oflag);
    YB_LAUNCH_CHECK();
    k_split_rows_h<<<(unsigned)((na + 7) / 8), 256, 0, st>>>(a, na, d, dh, 2, mu, scal, a_h, oflag);
    YB_LAUNCH_CHECK();
    if (padded > nb)
      YB_CUDA(cudaMemsetAsync(b_h + (size_t)nb * dop, 0, 2 * (size_t)(padded - nb) * dop, st));
  }
  plan.acc_scale = scal + 3;
  int rc;
  {
    ProfScope ps(1, st);
    rc = tf32_cross(plan, na, nb, dop, (const float *)b_h, (const float *)a_h, out, ldd, tws, st);
  }
  if (rc) return rc;
  int overflow = 0;
  YB_CUDA(cudaMemcpyAsync(&overflow, scal + 4, sizeof(int), cudaMemcpyDeviceToHost, st));
  YB_CUDA(cudaStreamSynchronize(st));
  return overflow ? -1001 : 0;
}

// operand kind of the resident tensor passes: FP16 unless YAEL_B200_OPERANDS=tf32
static int tensor_operand_kind() {
  const char *e = getenv("YAEL_B200_OPERANDS");
  return (e && (e[0] == 't' || e[0] == 'T')) ? 0 : 2;
}

// Queries whose certificate failed are re-done by the exact engine (own allocations: rare
// path).  flag_list / flag_count live in the caller's scratch.
static int redo_flagged_exact(int nq, int nb, int d, int k, const float *base, const float *query,
                              int *assign, float *dis, int id_offset, const int *flag_list,
                              const int *flag_count_dev, int *n_flag_out, cudaStream_t st) {
  (void)nq;
  int n_flag = 0;
  YB_CUDA(cudaMemcpyAsync(&n_flag, flag_count_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
  YB_CUDA(cudaStreamSynchronize(st));
  *n_flag_out = n_flag;
  if (n_flag <= 0) return 0;
  // pooled blocks sized in steps of 4096 queries, so that a k-means loop (which lands here with
  // a slightly different count every iteration) reuses them instead of paying cudaMalloc/cudaFree
  const int n_cap = (n_flag + 4095) & ~4095;
  float *qsub = (float *)yb_malloc(sizeof(float) * (size_t)n_cap * d);
  float *dsub = (float *)yb_malloc(sizeof(float) * (size_t)n_cap * k);
  int *asub = (int *)yb_malloc(sizeof(int) * (size_t)n_cap * k);
  int *rows = (int *)yb_malloc(sizeof(int) * (size_t)n_cap);
  void *ews = yb_malloc(knn_exact_ws_bytes(n_cap, nb, k));
  YB_CUDA(cudaMemcpyAsync(rows, flag_list, sizeof(int) * (size_t)n_flag, cudaMemcpyDeviceToDevice, st));
  long tot = (long)n_flag * d;
  k_gather_rows<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(query, rows, n_flag, d, qsub);
  count_launch();
  int rc;
  {
    ProfScope ps(4, st);
    rc = knn_exact(n_flag, nb, d, k, base, qsub, nullptr, asub, dsub, id_offset, ews, st);
  }
  if (!rc) {
    tot = (long)n_flag * k;
    k_scatter_results<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(rows, n_flag, k, asub, dsub,
                                                                     assign, dis);
    count_launch();
    cudaStreamSynchronize(st);
  }
  yb_free(qsub); yb_free(dsub); yb_free(asub); yb_free(rows); yb_free(ews);
  return rc;
}

// Queries whose certificate failed, first resort: the tensor path once more for just those
// queries with 4x looser admission thresholds (a fraction of a millisecond for a handful of
// queries; the exact engine needs ~2 ms however few they are).  What still fails there -- or a
// shape the tensor path refuses -- goes to the exact engine.
int knn_tf32_path(int nq, int nb, int d, int k, const float *base, const float *query,
                  const float *w, int *assign, float *dis, int id_offset, int force,
                  int *engine_out, long *uncert_out, cudaStream_t st, int kind, int retry);

static int redo_flagged_tensor(int nq, int nb, int d, int k, const float *base, const float *query,
                               int *assign, float *dis, int id_offset, const int *flag_list,
                               const int *flag_count_dev, int *n_flag_out, cudaStream_t st,
                               int kind) {
  int n_flag = 0;
  YB_CUDA(cudaMemcpyAsync(&n_flag, flag_count_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
  YB_CUDA(cudaStreamSynchronize(st));
  *n_flag_out = n_flag;
  if (n_flag <= 0) return 0;
  if (n_flag > nq / 4 || k > nb)  // not a few stragglers: the exact engine directly
    return redo_flagged_exact(nq, nb, d, k, base, query, assign, dis, id_offset, flag_list,
                              flag_count_dev, n_flag_out, st);
  const int n_cap = (n_flag + 255) & ~255;
  float *qsub = (float *)yb_malloc(sizeof(float) * (size_t)n_cap * d);
  float *dsub = (float *)yb_malloc(sizeof(float) * (size_t)n_cap * k);
  int *asub = (int *)yb_malloc(sizeof(int) * (size_t)n_cap * k);
  int *rows = (int *)yb_malloc(sizeof(int) * (size_t)n_cap);
  YB_CUDA(cudaMemcpyAsync(rows, flag_list, sizeof(int) * (size_t)n_flag, cudaMemcpyDeviceToDevice, st));
  long tot = (long)n_flag * d;
  k_gather_rows<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(query, rows, n_flag, d, qsub);
  count_launch();
  int eng = 0;
  long unc = 0;
  int rc;
  {
    ProfScope ps(4, st);
    // NOTE: this reserves the device scratch again; the caller no longer reads its own
    rc = knn_tf32_path(n_flag, nb, d, k, base, qsub, nullptr, asub, dsub, id_offset, 1, &eng, &unc, st,
                       kind, 1);
    if (rc == -1000 || rc == -1001) {
      void *ews = yb_malloc(knn_exact_ws_bytes(n_cap, nb, k));
      rc = knn_exact(n_flag, nb, d, k, base, qsub, nullptr, asub, dsub, id_offset, ews, st);
      cudaStreamSynchronize(st);
      yb_free(ews);
    }
  }
  if (!rc) {
    tot = (long)n_flag * k;
    k_scatter_results<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(rows, n_flag, k, asub, dsub,
                                                                     assign, dis);
    count_launch();
    cudaStreamSynchronize(st);
  }
  yb_free(qsub); yb_free(dsub); yb_free(asub); yb_free(rows);
  return rc;
}

// ------------------------------------------------------------------ engine 1, k = 1
// margin[q] = 2.05 * E_q + 2 * delta_q.  E_q bounds the error of a tensor score (operand rounding:
// err_scale |q-mu| max|b-mu|, sub-normal FP16 values: err_abs, and the ABSOLUTE term kAccAbs
// max|b-mu|^2 for the FP32 roundings that do not shrink with |q-mu|: a query AT the centring
// vector still sees scores of size |b-mu|^2 rounded in FP32); delta_q = kRefRound ((|q-mu|+|mu|)^2
// + (max|b-mu|+|mu|)^2) bounds the rounding of the reference's own FP32 distance, so the row the
// reference's arithmetic ranks first is a candidate even when it is an exact near-tie.
// scal[0] = max|b-mu|, scal[7] = |mu|.
__device__ __forceinline__ float k1_margin_of(float qn, float bmax, float mun, float err_scale,
                                              float err_abs) {
  const float e = err_scale * qn * bmax + err_abs * (qn + bmax) + kAccAbs * bmax * bmax;
  const float dl = kRefRound * ((qn + mun) * (qn + mun) + (bmax + mun) * (bmax + mun));
  return 2.05f * e + 2.0f * dl + 1e-30f;
}

__global__ void k_k1_margin(const float *__restrict__ query, int nq, int d,
                            const float *__restrict__ scal, float err_scale,
                            float *__restrict__ margin) {
  const int q = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (q >= nq) return;
  float s = 0.f;
  for (int t = lane; t < d; t += 32) {
    float v = query[(size_t)q * d + t];
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  if (lane == 0) margin[q] = k1_margin_of(sqrtf(s), scal[0], scal[7], err_scale, 0.f);
}

// the same from the squared norms of the centred queries (FP16 operand path; scal[5] = err_abs)
__global__ void k_k1_margin_n(const float *__restrict__ qcnorm, int nq, const float *__restrict__ scal,
                              float err_scale, float *__restrict__ margin) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  margin[q] = k1_margin_of(sqrtf(qcnorm[q]), scal[0], scal[7], err_scale, scal[5]);
}

// exact re-rank of the few k = 1 candidates: one warp per query, one lane per candidate slot;
// every lane walks its candidate row with the reference's arithmetic (yael/nn.c:100-129: float
// norm of the base row, double norm of the query, sequential FP32 FMA dot product) and the warp
// takes the (distance, id) minimum.  nn_single_full semantics (yael/nn.c:404-440).
__global__ void __launch_bounds__(128)
k_rerank_k1(int nq, int d, int slots, int lists, const float *__restrict__ base,
            const float *__restrict__ query, const int *__restrict__ cand_id,
            const float *__restrict__ cand_thr, int *__restrict__ assign, float *__restrict__ dis,
            int id_offset, int *__restrict__ flags, const double *__restrict__ qnorm) {
  // One warp per query, a handful of candidates each (the rows within the TF32 margin of the
  // best score).  The candidate rows and the query are brought into shared memory with
  // coalesced loads, all in flight together; then lane c walks candidate c sequentially -- the
  // reference's order of operations (nn.c:100-129: float norm of the base row, double norm of
  // the query, one FMA chain for the dot product), so distances equal the oracle's bit for bit.
  constexpr int RB = 8;          // candidates per batch
  constexpr int RP = 128 + 1;    // row pitch (d <= 128 on this path), odd: no bank conflicts
  __shared__ float rows[4][RB][RP];
  __shared__ float qs[4][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 4 + warp;
  if (q >= nq) return;
  const float *qrow = query + (size_t)q * d;
  const bool one_chunk = d <= 128;  // longer rows go through shared memory 128 coordinates at a time
  if (one_chunk)
    for (int t = lane; t < d; t += 32) qs[warp][t] = qrow[t];
  unsigned long long best = ~0ull;
  const double qn = qnorm[q];  // sequential double sum of squares (k_row_norms_seq, nn.c:108-120)
  for (int s0 = 0; s0 < slots; s0 += 32) {
    const int slot = s0 + lane;
    const int id = slot < slots ? cand_id[(size_t)q * slots + slot] : -1;
    unsigned todo = __ballot_sync(0xffffffffu, id >= 0);
    while (todo) {
      // this batch: the first RB candidates of `todo`; lane c < nb_ takes candidate c
      int src[RB];
      int nb_ = 0;
#pragma unroll
      for (int c = 0; c < RB; c++) {
        src[c] = todo ? __ffs(todo) - 1 : -1;
        if (todo) {
          todo &= todo - 1;
          nb_++;
        }
      }
      int myid = -1;
      float nf = 0.f, dot = 0.f;
      for (int t0 = 0; t0 < d; t0 += 128) {  // the chains run on across the chunks: same order
        const int dc = d - t0 < 128 ? d - t0 : 128;
        __syncwarp();
        if (!one_chunk)
          for (int t = lane; t < dc; t += 32) qs[warp][t] = qrow[t0 + t];
#pragma unroll
        for (int c = 0; c < RB; c++) {
          if (src[c] < 0) break;
          const int idc = __shfl_sync(0xffffffffu, id, src[c]);
          if (c == lane) myid = idc;
          const float *brow = base + (size_t)idc * d + t0;
          for (int t = lane; t < dc; t += 32) rows[warp][c][t] = __ldg(brow + t);
        }
        __syncwarp();
        if (lane < nb_) {
#pragma unroll 16
          for (int t = 0; t < dc; t++) {  // (unrolled: the shared-memory loads run ahead of the chains)
            const float v = rows[warp][lane][t];
            nf = __fadd_rn(nf, __fmul_rn(v, v));
            dot = fmaf(v, qs[warp][t], dot);
          }
        }
      }
      if (lane < nb_) {
        const float dist = __fadd_rn((float)(qn + (double)nf), __fmul_rn(-2.0f, dot));
        const uint32_t fk = float_key(dist);
        if (!is_nan_key(fk)) {
          unsigned long long key = ((unsigned long long)fk << 32) | (unsigned)myid;
          best = key < best ? key : best;
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
    best = t < best ? t : best;
  }
  if (lane == 0) {
    int flag = 0;
    for (int l = 0; l < lists; l++) {
      float t = cand_thr[(size_t)q * lists + l];
      if (t != t) flag = 1;  // NaN: a list overflowed, the candidate set is incomplete
    }
    flags[q] = flag;
    const uint32_t fk = (uint32_t)(best >> 32);
    const uint32_t bits = (fk & 0x80000000u) ? (fk & 0x7fffffffu) : ~fk;
    const float dv = __uint_as_float(bits);
    if (best != ~0ull && dv < 1e30f) {
      assign[q] = (int)(uint32_t)best + id_offset;
      dis[q] = dv;
    } else {
      assign[q] = -1;
      dis[q] = 1e30f;
    }
  }
}

// ---- lane-per-query variant (d <= 128, d % 4 == 0, 16-byte aligned rows) ------------------------
// k_rerank_k1 above spends a whole warp on one query although only one or two lanes have a candidate
// to walk: ncu shows it ISSUE-bound (91 % issue slots busy, ~1060 warp instructions per query; 14 ms
// per iteration at BASELINE configs[3]).  Here a warp owns 32 consecutive queries, lane = query: the
// 32 query rows and, per round, the lanes' next candidate rows are copied into padded shared-memory
// tiles with cp.async (16 bytes per lane: one instruction moves a whole 512-byte row and needs no
// registers, so all 32 rows of a tile are in flight together), and every lane walks ITS OWN pair
// with the same sequence of operations as above (float norm of the base row, one FMA chain, the
// query norm from k_row_norms_seq) reading float4 from a conflict-free pitch.  Rounds repeat until
// no lane has a candidate left.  Candidates are first filtered by their tensor score: a row is
// walked only if its score is within the query's margin of the BEST published score -- the tensor
// pass admitted against the running best of one list (one column half of one range), so most of what
// the other lists publish cannot be the nearest row; the margin argument of k1_margin_of holds
// against the final best a fortiori.
constexpr int RL_PITCH = 132;                       // floats per staged row: 16-byte aligned, 4 * lane mod 32 banks
constexpr int RL_WARPS = 2;
constexpr int RL_CAP = 16;                          // candidates kept per query (more: the query is redone exactly)
constexpr int RL_WARP_FLOATS = 2 * 32 * RL_PITCH;   // query tile + candidate tile
constexpr int RL_WARP_BYTES = RL_WARP_FLOATS * 4 + 32 * RL_CAP * 4 + 32 * RL_CAP + 32 * 8;
constexpr int RL_SMEM_BYTES = RL_WARPS * RL_WARP_BYTES;
static_assert(RL_WARP_BYTES % 16 == 0, "per-warp block keeps the tiles 16-byte aligned");

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// The (query, candidate) pairs of the warp's 32 queries are FLATTENED before they are walked: lane p
// of round r takes pair 32 r + p, so a warp needs ceil(pairs / 32) rounds instead of as many as its
// busiest query has candidates (a lane per query with 2-3 candidates on average but 8 in the worst
// lane of a warp: 1.9 ms against 1.1 ms for 2.5 M points).
__global__ void __launch_bounds__(32 * RL_WARPS)
k_rerank_k1_lanes(int nq, int d, int slots, int lists, const float *__restrict__ base,
                  const float *__restrict__ query, const int *__restrict__ cand_id,
                  const float *__restrict__ cand_score, const float *__restrict__ cand_thr,
                  const float *__restrict__ margin, int *__restrict__ assign, float *__restrict__ dis,
                  int id_offset, int *__restrict__ flags, const double *__restrict__ qnorm) {
  extern __shared__ __align__(16) unsigned char rl_sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char *mine = rl_sm + (size_t)warp * RL_WARP_BYTES;
  float *qt = reinterpret_cast<float *>(mine), *ct = qt + 32 * RL_PITCH;
  int *flat_id = reinterpret_cast<int *>(mine + RL_WARP_FLOATS * 4);            // [32 * RL_CAP]
  unsigned char *flat_q = mine + RL_WARP_FLOATS * 4 + 32 * RL_CAP * 4;          // [32 * RL_CAP]
  unsigned long long *bestk = reinterpret_cast<unsigned long long *>(flat_q + 32 * RL_CAP);  // [32]
  const long q0 = ((long)blockIdx.x * RL_WARPS + warp) * 32;
  if (q0 >= nq) return;
  const long q = q0 + lane;
  const bool valid = q < nq;
  const bool col = lane * 4 < d;   // this lane's 16 bytes of a row
  for (int r = 0; r < 32; r++)
    if (col && q0 + r < nq) cp_async16(qt + r * RL_PITCH + lane * 4, query + (size_t)(q0 + r) * d + lane * 4);
  bestk[lane] = ~0ull;
  // my query's candidates: those whose tensor score is within the margin of the best published one
  const int *cid = cand_id + (size_t)(valid ? q : 0) * slots;
  const float *csc = cand_score ? cand_score + (size_t)(valid ? q : 0) * slots : nullptr;
  float lim = __uint_as_float(0x7f800000u);
  if (valid && csc) {
    float sbest = lim;
    for (int s = 0; s < slots; s++)
      if (cid[s] >= 0) sbest = fminf(sbest, csc[s]);
    lim = sbest + (margin ? margin[q] : 0.f);
  }
  int mycand[RL_CAP];
  int n = 0;
  bool overflow = false;
  if (valid)
    for (int s = 0; s < slots; s++) {
      const int c = cid[s];
      if (c >= 0 && !(csc && csc[s] > lim)) {
        if (n < RL_CAP) {
#pragma unroll
          for (int e = 0; e < RL_CAP; e++)
            if (e == n) mycand[e] = c;
          n++;
        } else {
          overflow = true;
        }
      }
    }
  int off = n;  // exclusive prefix sum of the counts
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, off, o);
    if (lane >= o) off += t;
  }
  const int total = __shfl_sync(0xffffffffu, off, 31);
  off -= n;
#pragma unroll
  for (int e = 0; e < RL_CAP; e++)
    if (e < n) {
      flat_id[off + e] = mycand[e];
      flat_q[off + e] = (unsigned char)lane;
    }
  __syncwarp();
  for (int p0 = 0; p0 < total; p0 += 32) {
    const int p = p0 + lane;
    const int id = p < total ? flat_id[p] : -1;
    const int j = p < total ? flat_q[p] : 0;
#pragma unroll 8
    for (int r = 0; r < 32; r++) {
      const int idr = __shfl_sync(0xffffffffu, id, r);
      if (idr >= 0 && col) cp_async16(ct + r * RL_PITCH + lane * 4, base + (size_t)idr * d + lane * 4);
    }
    cp_async_wait_all();
    __syncwarp();
    if (id >= 0) {
      const float4 *cr = reinterpret_cast<const float4 *>(ct + lane * RL_PITCH);
      const float4 *qr = reinterpret_cast<const float4 *>(qt + j * RL_PITCH);
      float nf = 0.f, dot = 0.f;
#pragma unroll 8
      for (int t4 = 0; t4 < (d >> 2); t4++) {   // coordinate order, as nn.c:100-129
        const float4 v = cr[t4], w = qr[t4];
        nf = __fadd_rn(nf, __fmul_rn(v.x, v.x)); dot = fmaf(v.x, w.x, dot);
        nf = __fadd_rn(nf, __fmul_rn(v.y, v.y)); dot = fmaf(v.y, w.y, dot);
        nf = __fadd_rn(nf, __fmul_rn(v.z, v.z)); dot = fmaf(v.z, w.z, dot);
        nf = __fadd_rn(nf, __fmul_rn(v.w, v.w)); dot = fmaf(v.w, w.w, dot);
      }
      const double qn = qnorm[q0 + j];
      const float dist = __fadd_rn((float)(qn + (double)nf), __fmul_rn(-2.0f, dot));
      const uint32_t fk = float_key(dist);
      if (!is_nan_key(fk)) atomicMin(&bestk[j], ((unsigned long long)fk << 32) | (unsigned)id);
    }
    __syncwarp();  // the candidate tile is rewritten by the next round; bestk is read after the last
  }
  if (!valid) return;
  int flag = overflow ? 1 : 0;  // more candidates than RL_CAP: the exact engine redoes the query
  for (int l = 0; l < lists; l++) {
    const float t = cand_thr[(size_t)q * lists + l];
    if (t != t) flag = 1;  // NaN: a list overflowed, the candidate set is incomplete
  }
  flags[q] = flag;
  const unsigned long long best = bestk[lane];
  const uint32_t fk = (uint32_t)(best >> 32);
  const uint32_t bits = (fk & 0x80000000u) ? (fk & 0x7fffffffu) : ~fk;
  const float dv = __uint_as_float(bits);
  if (best != ~0ull && dv < 1e30f) {
    assign[q] = (int)(uint32_t)best + id_offset;
    dis[q] = dv;
  } else {
    assign[q] = -1;
    dis[q] = 1e30f;
  }
}

// launches the lane-per-query kernel when the shape allows it, else the warp-per-query one
static int rerank_k1(int nq, int d, int slots, int lists, const float *base, const float *query,
                     const int *cand_id, const float *cand_score, const float *cand_thr,
                     const float *margin, int *assign, float *dis, int id_offset, int *flags,
                     const double *qnorm, cudaStream_t st) {
  static const bool lanes_on = !(getenv("YAEL_B200_K1_LANES") && atoi(getenv("YAEL_B200_K1_LANES")) == 0);
  // (without tensor scores nothing filters the candidates and nobody reads the overflow flag: hkm)
  if (lanes_on && d <= 128 && (d & 3) == 0 && ((((uintptr_t)base) | ((uintptr_t)query)) & 15) == 0 &&
      (cand_score != nullptr || slots <= RL_CAP)) {
    static bool attr[64] = {};
    cudaError_t ae = cudaSuccess;
    once_per_device(attr, [&ae] {
      ae = cudaFuncSetAttribute(k_rerank_k1_lanes, cudaFuncAttributeMaxDynamicSharedMemorySize, RL_SMEM_BYTES);
    });
    if (ae != cudaSuccess) {
      attr[dev_index()] = false;
      return fail(6, "cannot reserve %d bytes of shared memory: %s", RL_SMEM_BYTES, cudaGetErrorString(ae));
    }
    const long per_cta = 32L * RL_WARPS;
    k_rerank_k1_lanes<<<(unsigned)((nq + per_cta - 1) / per_cta), 32 * RL_WARPS, RL_SMEM_BYTES, st>>>(
        nq, d, slots, lists, base, query, cand_id, cand_score, cand_thr, margin, assign, dis, id_offset, flags, qnorm);
  } else {
    k_rerank_k1<<<(nq + 3) / 4, 128, 0, st>>>(nq, d, slots, lists, base, query, cand_id, cand_thr, assign, dis,
                                              id_offset, flags, qnorm);
  }
  YB_LAUNCH_CHECK();
  return 0;
}

// (Round 2 tried a LANE per query -- 32 queries per warp, candidate and query rows staged through
// padded shared tiles, the query norm fused in: bit-identical, but 24.1 ms against 14.5 ms at
// BASELINE configs[3]: the staging loop's loads are only four deep.  Not kept; profiles/README.md.)
static int knn_tf32_nearest(int nq, int nb, int d, const float *base, const float *query,
                            int *assign, float *dis, int id_offset, long *uncert_out,
                            cudaStream_t st, int kind) {
  const bool f16 = kind == 2;
  // FP16 operands with folded norms (kind 3; the extras travel as a 32-byte chunk through their own
  // ring: BASELINE config 4 pass 186.8 -> 176.4 ms) unless YAEL_B200_K1_FOLD=0 (plain kind 2, FMA
  // epilogue)
  const char *kf = getenv("YAEL_B200_K1_FOLD");
  const bool fold = f16 && !(kf && atoi(kf) == 0);
  if (fold) kind = 3;
  const int dh = fold ? (d + 15) & ~15 : (d + 7) & ~7;     // FP16: data elements per row
  const int dpad = f16 ? dh + (fold ? kNfExtra : 0) : (d + 3) & ~3;  // operand row pitch in elements
  Tf32Plan plan = tf32_plan_nearest(nq, nb, dpad, kind);
  if (!plan.ok) return -1000;
  const int kp = plan.kprime, slots = plan.lists * kp;
  const long padded = tf32_padded_rows(nb);
  size_t need = Carver::need(sizeof(float) * (size_t)padded) + Carver::need(64) +
                Carver::need((f16 ? 2 : 4) * (size_t)padded * dpad) +
                Carver::need((f16 ? 2 : 4) * (size_t)nq * dpad) + center_ws_bytes(nb, d) +
                Carver::need(sizeof(float) * (size_t)nq) +
                Carver::need(sizeof(float) * (size_t)nq) +
                Carver::need(sizeof(float) * (size_t)nq * slots) +
                Carver::need(sizeof(int) * (size_t)nq * slots) +
                Carver::need(sizeof(float) * (size_t)nq * plan.lists) +
                2 * Carver::need(sizeof(int) * (size_t)nq) + Carver::need(plan.ws_bytes) + 1024 +
                Carver::need(sizeof(double) * (size_t)nq);
  int n_flag = 0;
  {
    ScratchScope ws(need, st);
    Carver c(ws.p);
    float *an = c.take<float>(padded);
    float *scal = c.take<float>(16);
    double *qnorm = c.take<double>(nq);
    float *margin = c.take<float>(nq);
    float *cscore = c.take<float>((size_t)nq * slots);
    int *cid = c.take<int>((size_t)nq * slots);
    float *cthr = c.take<float>((size_t)nq * plan.lists);
    int *flags = c.take<int>(nq);
    int *flag_list = c.take<int>(nq);
    void *tfws = c.take<char>(plan.ws_bytes);
    float *base_c = (float *)c.take<char>((f16 ? 2 : 4) * (size_t)padded * dpad);
    float *query_c = (float *)c.take<char>((f16 ? 2 : 4) * (size_t)nq * dpad);
    float *qcnorm = c.take<float>(nq);
    void *cws = c.take<char>(center_ws_bytes(nb, d));
    int rc;
    {
      ProfScope ps(0, st);
      YB_CUDA(cudaMemsetAsync(scal, 0, 64, st));
      if (f16) {
        // no FP32 copy of the centred queries (for k-means they are the 10^7 points): the margin
        // comes from the norms the conversion kernel emits
        if ((rc = center_operands_h(nq, nb, d, dh, dh, base, query, (__half *)base_c,
                                    (__half *)query_c, nullptr, an, qcnorm, scal, cws, st, fold)))
          return rc;
        plan.acc_scale = scal + 3;
      } else {
        if ((rc = center_operands(nq, nb, d, dpad, base, query, base_c, query_c, an, scal + 7, cws, st))) return rc;
      }
      if ((rc = fill_f32(an + nb, padded - nb, __builtin_inff(), st))) return rc;
      k_sqrt_max<<<2 * sm_count(), 256, 0, st>>>(an, nb, scal);
      YB_LAUNCH_CHECK();
      if (f16)
        k_k1_margin_n<<<(nq + 255) / 256, 256, 0, st>>>(qcnorm, nq, scal, kF16ErrScale, margin);
      else
        k_k1_margin<<<(nq + 3) / 4, 128, 0, st>>>(query_c, nq, dpad, scal, kTf32ErrScale, margin);
      YB_LAUNCH_CHECK();
    }
    {
      ProfScope ps(1, st);
      if ((rc = tf32_nearest(plan, nq, nb, dpad, base_c, query_c, an, margin, cscore, cid, cthr,
                             tfws, st)))
        return rc;
    }
    {
      ProfScope ps(3, st);
      if ((rc = row_norms_seq(query, nq, d, d, nullptr, qnorm, st))) return rc;
      if ((rc = rerank_k1(nq, d, slots, plan.lists, base, query, cid, cscore, cthr, margin, assign, dis,
                          id_offset, flags, qnorm, st)))
        return rc;
    }
    k_collect_flags<<<(nq + 255) / 256, 256, 0, st>>>(flags, nq, flag_list, (int *)(scal + 1));
    YB_LAUNCH_CHECK();
    int overflow = 0;
    if (f16) {  // read with the flag count: redo_flagged_exact synchronises the stream
      YB_CUDA(cudaMemcpyAsync(&overflow, scal + 4, sizeof(int), cudaMemcpyDeviceToHost, st));
      YB_CUDA(cudaStreamSynchronize(st));
      if (overflow) return -1001;  // a value outside FP16's range: the caller repeats with TF32
    }
    if ((rc = redo_flagged_exact(nq, nb, d, 1, base, query, assign, dis, id_offset, flag_list,
                                 (int *)(scal + 1), &n_flag, st)))
      return rc;
  }
  *uncert_out = n_flag;
  return 0;
}

// returns -1000 when the tensor-core path does not apply (caller falls through to engine 0)
int knn_tf32_path(int nq, int nb, int d, int k, const float *base, const float *query,
                  const float *w, int *assign, float *dis, int id_offset, int force,
                  int *engine_out, long *uncert_out, cudaStream_t st, int kind, int retry) {
  if (force == 0 || w != nullptr) return -1000;
  const bool f16 = kind == 2;
  // the tensor pass runs on centred, pitch-padded copies (16-byte row pitch: 4 floats / 8 halfs)
  // FP16: data elements per row, a multiple of 16 (one MMA K step) so that the K steps over the
  // data never touch the extras behind them
  const int dh = (d + 15) & ~15;
  const int dpad = f16 ? dh + kNfExtra : (d + 3) & ~3;     // operand row pitch in elements
  const int dqc = (d + 3) & ~3;   // pitch of the FP32 centred queries (certificate)
  const int opkind = f16 ? 3 : kind;  // FP16 passes of this path fold |b|^2 into the contraction
  Tf32Plan plan = tf32_plan(nq, nb, dpad, k, opkind);
  if (!plan.ok) return -1000;
  if (force < 0 && (double)nq * nb < 1e6) return -1000;  // tiny problems: not worth a TMA setup
  if (k == 1) {
    (void)retry;
    int rc1 = knn_tf32_nearest(nq, nb, d, base, query, assign, dis, id_offset, uncert_out, st, kind);
    if (rc1 == 0) *engine_out = 1;
    return rc1;
  }

  const int kp = plan.kprime;
  const int stride = plan.lists * kp;  // candidates per query produced by the tensor pass
  const int m = kp;                     // candidates per query that are re-ranked
  const int m_pad = pow2_ceil(m < 2 ? 2 : m);
  size_t smem = rerank_smem_bytes(d, m, m_pad);
  if (smem > 160 * 1024) return -1000;
  const long padded = tf32_padded_rows(nb);
  const int nbt = tf32_tiles(nb);
  const bool need_sel = plan.lists > 1;
  // merge by list counts (k_merge_lists) instead of pre-fill + radix select over every slot
  const bool fast_merge = need_sel && 2 * kp <= ML_CAP;

  // Sampling pre-passes (large databases).  A threshold tau_q with "about 3x the wanted number
  // of rows below it" turns the streaming top-k into a plain filter: lists hardly ever fill up,
  // so the epilogue never stops to compact.  Level 2 scans every 16th tile and keeps the j2
  // smallest scores; its own threshold comes from level 1, a raw score slab over a few tiles.
  // Any threshold is SOUND as long as enough rows pass it, which the re-rank verifies per
  // query (too few candidates -> exact engine); the levels only affect speed.
  const int kSampleStride = 16;
  const bool use_sample = nbt >= 8 * kSampleStride;
  const int nbt_s = (nbt + kSampleStride - 1) / kSampleStride;
  // j2-th smallest sampled group minimum: about 16 * j2 rows of the database pass the threshold
  // (relative spread 1 / sqrt(j2)).  2k' / 16 (25 for k = 100: ~400 rows for the 100 wanted)
  // leaves a query short of candidates about once in 10^4 (measured with FP16 operands, 10 k
  // queries: j2 = 38 / 24 / 20 -> 0 uncertified, pass 3.22 / 2.94 / 2.86 ms; j2 = 16 -> 9); such
  // queries are retried with 4x the threshold rank (redo_flagged_tensor), which costs little.
  int j2 = (2 * kp + kSampleStride - 1) / kSampleStride;
  if (j2 < 24) j2 = 24;
  if (const char *e = getenv("YAEL_B200_J2")) j2 = atoi(e) > 0 ? atoi(e) : j2;  // experiment knob
  if (retry) j2 *= 4;
  Tf32Plan splan = {};
  if (use_sample) splan = tf32_plan_tiles(nq, nbt_s, dpad, j2, opkind);
  const bool sample_ok = use_sample && splan.ok;
  const int sstride = sample_ok ? splan.lists * j2 : 1;
  // level 1: t1 tiles spread over the database, j1-th smallest -> about 3*j2 rows of level 2
  int t1 = nbt_s / 8;
  if (t1 > 16) t1 = 16;
  const int stride1 = t1 > 0 ? nbt / t1 : 1;
  const long rows1 = (long)t1 * 256;
  int j1 = t1 > 0 ? (int)((3L * j2 * t1 + nbt_s - 1) / nbt_s) : 0;
  if (j1 < 12) j1 = 12;
  Tf32Plan l1plan = {};
  const bool level1_ok = sample_ok && t1 >= 8 && (size_t)nq * rows1 * 4 <= ((size_t)1 << 30) &&
                         (l1plan = tf32_plan_tiles(nq, t1, dpad, 8, opkind)).ok;

  // single-level sampling: the sample pass emits the minimum of every group of gsize columns
  // and the threshold is an order statistic of those minima (with j2 << groups the j2 smallest
  // rows sit in distinct groups, so this is the j2-th smallest sampled score give or take a
  // rank or two) -- no lists, no second pass.  Falls back to the two-level scheme when the
  // minima of one query do not fit k_row_kth's registers.
  const long srows = (long)nbt_s * 256;
  // 32 columns per emitted minimum: half the values for k_row_kth and half the stores of 16, same
  // thresholds for all practical purposes (the j2 smallest sampled rows still sit in distinct
  // groups: 25 of 1960); measured 0.336 -> 0.282 ms for the sampling phase, 0 uncertified queries,
  // pass unchanged (64: 0.265 ms)
  int gsize = 32;
  if (const char *e = getenv("YAEL_B200_GSIZE")) gsize = atoi(e) >= 16 ? atoi(e) : gsize;  // experiment knob
  while (gsize < 128 && (srows / gsize > RK_T * RK_PER || (size_t)nq * (srows / gsize) * 4 > ((size_t)256 << 20)))
    gsize *= 2;
  const long gcols = srows / gsize;
  const bool gmin_ok = sample_ok && gcols <= RK_T * RK_PER && (long)j2 * 4 <= gcols &&
                       (size_t)nq * gcols * 4 <= ((size_t)512 << 20) && !getenv("YAEL_B200_TWO_LEVEL");
  size_t need = Carver::need(sizeof(float) * (size_t)padded) + Carver::need(64) +
                Carver::need(sizeof(float) * (size_t)nq * stride) +
                Carver::need(sizeof(int) * (size_t)nq * stride) +
                Carver::need(sizeof(float) * (size_t)nq * plan.lists) +
                Carver::need(sizeof(int) * (size_t)nq * kp) +
                2 * Carver::need(sizeof(int) * (size_t)nq) + kmin_ws_bytes(nq, kp) +
                Carver::need(sizeof(int) * (size_t)nq * plan.lists) +
                Carver::need(plan.ws_bytes) + 1024 +
                Carver::need((f16 ? 2 : 4) * (size_t)padded * dpad) +
                Carver::need((f16 ? 2 : 4) * (size_t)nq * dpad) +
                Carver::need(sizeof(float) * (size_t)nq * dqc) + center_ws_bytes(nb, d);
  if (sample_ok)
    need += Carver::need(sizeof(float) * (size_t)nq * sstride) +
            Carver::need(sizeof(int) * (size_t)nq * sstride) +
            2 * Carver::need(sizeof(float) * (size_t)nq * splan.lists) +
            2 * Carver::need(sizeof(int) * (size_t)nq * j2) + 2 * Carver::need(sizeof(float) * (size_t)nq) +
            kmin_ws_bytes(nq, j2) + Carver::need(splan.ws_bytes);
  if (gmin_ok) need += Carver::need(sizeof(float) * (size_t)nq * gcols);
  if (level1_ok)
    need += Carver::need(sizeof(float) * (size_t)nq * rows1) +
            2 * Carver::need(sizeof(int) * (size_t)nq * j1) + kmin_ws_bytes(nq, j1) +
            Carver::need(l1plan.ws_bytes);
  int n_flag = 0;
  {
    ScratchScope ws(need, st);
    Carver c(ws.p);
    float *an = c.take<float>(padded);
    float *scal = c.take<float>(16);  // [0] = max |b|, [1] = flag count (int)
    float *cscore = c.take<float>((size_t)nq * stride);
    int *cid = c.take<int>((size_t)nq * stride);
    float *cthr = c.take<float>((size_t)nq * plan.lists);
    int *sel = c.take<int>((size_t)nq * kp);
    int *ccnt = c.take<int>((size_t)nq * plan.lists);
    int *flags = c.take<int>(nq);
    int *flag_list = c.take<int>(nq);
    void *kws = c.take<char>(kmin_ws_bytes(nq, kp));
    void *tfws = c.take<char>(plan.ws_bytes);
    // operands of the tensor passes (FP32 rows read as TF32, or FP16) and, for FP16, a separate
    // FP32 copy of the centred queries for the certificate
    float *base_c = (float *)c.take<char>((f16 ? 2 : 4) * (size_t)padded * dpad);
    float *query_c = (float *)c.take<char>((f16 ? 2 : 4) * (size_t)nq * dpad);
    float *query_cf = f16 ? c.take<float>((size_t)nq * dqc) : query_c;
    void *cws = c.take<char>(center_ws_bytes(nb, d));
    float *thr_init = nullptr;
    int rc;
    {
      ProfScope ps(0, st);
      YB_CUDA(cudaMemsetAsync(scal, 0, 64, st));
      if (f16) {
        if ((rc = center_operands_h(nq, nb, d, dh, dqc, base, query, (__half *)base_c,
                                    (__half *)query_c, query_cf, an, nullptr, scal, cws, st, true)))
          return rc;
        plan.acc_scale = splan.acc_scale = l1plan.acc_scale = scal + 3;
      } else {
        if ((rc = center_operands(nq, nb, d, dpad, base, query, base_c, query_c, an, scal + 7, cws, st))) return rc;
      }
      if ((rc = fill_f32(an + nb, padded - nb, __builtin_inff(), st))) return rc;
      k_sqrt_max<<<2 * sm_count(), 256, 0, st>>>(an, nb, scal);
      YB_LAUNCH_CHECK();
    }
    if (sample_ok) {
      ProfScope ps(10, st);
      float *sscore = c.take<float>((size_t)nq * sstride);
      int *sid = c.take<int>((size_t)nq * sstride);
      float *sthr = c.take<float>((size_t)nq * splan.lists);
      int *scnt = c.take<int>((size_t)nq * splan.lists);
      int *ssel = c.take<int>((size_t)nq * j2);
      float *svals = (float *)c.take<int>((size_t)nq * j2);
      thr_init = c.take<float>(nq);
      float *thr1 = c.take<float>(nq);
      void *skws = c.take<char>(kmin_ws_bytes(nq, j2));
      void *stfws = c.take<char>(splan.ws_bytes);
      const float *thr_l1 = nullptr;
      if (gmin_ok) {
        float *gm = c.take<float>((size_t)nq * gcols);
        if ((rc = tf32_group_min(splan, nq, nb, dpad, nbt_s, kSampleStride, base_c, query_c, an, gm,
                                 gcols, gsize, stfws, st)))
          return rc;
        k_row_kth<<<nq, RK_T, 0, st>>>(gm, gcols, (int)gcols, j2, thr_init);
        YB_LAUNCH_CHECK();
      } else {
      if (level1_ok) {
        ProfScope ps1(11, st);
        float *slab = c.take<float>((size_t)nq * rows1);
        int *sel1 = c.take<int>((size_t)nq * j1);
        float *vals1 = (float *)c.take<int>((size_t)nq * j1);
        void *kws1 = c.take<char>(kmin_ws_bytes(nq, j1));
        void *tfws1 = c.take<char>(l1plan.ws_bytes);
        if ((rc = tf32_scores(l1plan, nq, nb, dpad, t1, stride1, base_c, query_c, an, slab, rows1, tfws1, st)))
          return rc;
        if ((rc = kmin_rows(slab, rows1, rows1, nq, j1, +1, sel1, vals1, 0, 0, kws1, st))) return rc;
        k_threshold_from_kmin<<<(nq + 255) / 256, 256, 0, st>>>(sel1, vals1, nq, j1, thr1);
        YB_LAUNCH_CHECK();
        thr_l1 = thr1;
      }
      if (fast_merge) {
        Tf32Out so = {scnt, 0, 0, 0};
        if ((rc = tf32_shortlist(splan, nq, nb, dpad, nbt_s, kSampleStride, base_c, query_c, an, thr_l1,
                                 sscore, sid, sthr, stfws, st, &so)))
          return rc;
        k_merge_lists<<<nq, ML_T, 0, st>>>(scnt, sscore, splan.lists, j2, j2, nullptr, thr_init);
        YB_LAUNCH_CHECK();
      } else {
        if ((rc = fill_f32(sscore, (long)nq * sstride, __builtin_inff(), st))) return rc;
        YB_CUDA(cudaMemsetAsync(sid, 0xff, sizeof(int) * (size_t)nq * sstride, st));
        if ((rc = tf32_shortlist(splan, nq, nb, dpad, nbt_s, kSampleStride, base_c, query_c, an, thr_l1,
                                 sscore, sid, sthr, stfws, st)))
          return rc;
        if ((rc = kmin_rows(sscore, sstride, sstride, nq, j2, +1, ssel, svals, 0, 0, skws, st)))
          return rc;
        k_threshold_from_kmin<<<(nq + 255) / 256, 256, 0, st>>>(ssel, svals, nq, j2, thr_init);
        YB_LAUNCH_CHECK();
      }
      }  // two-level scheme
    }
    if (!fast_merge) {
      if ((rc = fill_f32(cscore, (long)nq * stride, __builtin_inff(), st))) return rc;
      YB_CUDA(cudaMemsetAsync(cid, 0xff, sizeof(int) * (size_t)nq * stride, st));
    }
    {
      ProfScope ps(1, st);
      Tf32Out mo = {ccnt, 0, 0, 0};
      if ((rc = tf32_shortlist(plan, nq, nb, dpad, nbt, 1, base_c, query_c, an, thr_init, cscore, cid,
                               cthr, tfws, st, fast_merge ? &mo : nullptr)))
        return rc;
    }
    if (fast_merge) {
      ProfScope ps(2, st);
      k_merge_lists<<<nq, ML_T, 0, st>>>(ccnt, cscore, plan.lists, kp, kp, sel, nullptr);
      YB_LAUNCH_CHECK();
    } else if (need_sel) {
      // merge the per-list shortlists: the kp smallest TF32 scores of the union
      ProfScope ps(2, st);
      if ((rc = kmin_rows(cscore, stride, stride, nq, kp, +1, sel, nullptr, 0, 0, kws, st)))
        return rc;
    }

    RerankArgs A = {};
    A.nq = nq; A.nb = nb; A.d = d; A.k = k; A.base = base; A.query = query;
    A.dis = dis; A.assign = assign; A.id_offset = id_offset;
    A.cand_id = cid; A.cand_score = cscore; A.sel = need_sel ? sel : nullptr;
    A.cand_stride = stride; A.m = m; A.all_listed = (nb <= kp);
    A.cand_thr = cthr; A.lists = plan.lists; A.query_c = query_cf; A.qc_ld = f16 ? dqc : dpad;
    A.err_scale = f16 ? kF16ErrScale : kTf32ErrScale; A.bmax = scal; A.uncert_flags = flags;
    A.err_abs = f16 ? scal + 5 : nullptr;
    A.mu_norm = scal + 7;
    A.gsort = nullptr; A.m_pad = m_pad; A.k1 = (k == 1);
    rerank_attrs();
    {
      ProfScope ps(3, st);
      k_rerank<1><<<nq, RR_T, rerank_setup(A, smem), st>>>(A);
      YB_LAUNCH_CHECK();
    }
    k_collect_flags<<<(nq + 255) / 256, 256, 0, st>>>(flags, nq, flag_list, (int *)(scal + 1));
    YB_LAUNCH_CHECK();
    if (f16) {
      int overflow = 0;
      YB_CUDA(cudaMemcpyAsync(&overflow, scal + 4, sizeof(int), cudaMemcpyDeviceToHost, st));
      YB_CUDA(cudaStreamSynchronize(st));
      if (overflow) return -1001;  // a value outside FP16's range: the caller repeats with TF32
    }
    if (retry)
      rc = redo_flagged_exact(nq, nb, d, k, base, query, assign, dis, id_offset, flag_list,
                              (int *)(scal + 1), &n_flag, st);
    else
      rc = redo_flagged_tensor(nq, nb, d, k, base, query, assign, dis, id_offset, flag_list,
                               (int *)(scal + 1), &n_flag, st, kind);
    if (rc) return rc;
  }
  *engine_out = 1;
  *uncert_out = n_flag;
  return 0;
}

// ------------------------------------------------------------------ engine 1, database in host memory
// knn_full() on a host-resident database is a PCIe transfer (517 MB, 9.5 ms for the bench
// shape) followed by 5 ms of kernels.  Here the two overlap: the SAMPLE tiles (tile 0 of every
// group of 16 tiles: a 2-D copy, 1/16 of the bytes) go first, the centring vector, the centred
// queries and the admission thresholds are computed from them exactly as in the resident path,
// and the rest of the database follows in a few large 2-D copies; each chunk is centred and
// scanned by its own tensor pass as soon as its copy has landed, all passes publishing into one
// set of shortlists.  Merge, exact re-rank, certificate and fallback are the resident path's.
// Returns -1000 when the shape does not qualify (caller copies the database and calls the
// resident path).
int knn_tf32_streamed(int nq, int nb, int d, int k, const float *base_host, float *base,
                      const float *query, int *assign, float *dis, int id_offset, int force,
                      int *engine_out, long *uncert_out, cudaStream_t st) {
  if (force == 0 || k < 2) return -1000;
  const int dpad = (d + 3) & ~3;
  Tf32Plan plan0 = tf32_plan(nq, nb, dpad, k);
  if (!plan0.ok) return -1000;
  const int kp = plan0.kprime, m = kp;
  const int m_pad = pow2_ceil(m < 2 ? 2 : m);
  const size_t smem = rerank_smem_bytes(d, m, m_pad);
  if (smem > 160 * 1024 || 2 * kp > ML_CAP) return -1000;
  const size_t row_bytes = sizeof(float) * (size_t)d;
  const size_t total_bytes = row_bytes * (size_t)nb;
  const long kGroupRows = 16L * 256;
  const size_t pitch = (size_t)kGroupRows * row_bytes;
  if (total_bytes < ((size_t)96 << 20) || pitch > ((size_t)1 << 30)) return -1000;
  const long padded = tf32_padded_rows(nb);
  const int nbt = tf32_tiles(nb);
  const int kSampleStride = 16;
  const int nbt_s = (nbt + kSampleStride - 1) / kSampleStride;
  int j2 = (3 * kp + kSampleStride - 1) / kSampleStride;
  if (j2 < 32) j2 = 32;
  Tf32Plan splan = tf32_plan_tiles(nq, nbt_s, dpad, j2);
  if (!splan.ok) return -1000;
  const long srows = (long)nbt_s * 256;
  // 32 columns per emitted minimum: half the values for k_row_kth and half the stores of 16, same
  // thresholds for all practical purposes (the j2 smallest sampled rows still sit in distinct
  // groups: 25 of 1960); measured 0.336 -> 0.282 ms for the sampling phase, 0 uncertified queries,
  // pass unchanged (64: 0.265 ms)
  int gsize = 32;
  if (const char *e = getenv("YAEL_B200_GSIZE")) gsize = atoi(e) >= 16 ? atoi(e) : gsize;  // experiment knob
  while (gsize < 128 && (srows / gsize > RK_T * RK_PER || (size_t)nq * (srows / gsize) * 4 > ((size_t)256 << 20)))
    gsize *= 2;
  const long gcols = srows / gsize;
  if (gcols > RK_T * RK_PER || (long)j2 * 4 > gcols || (size_t)nq * gcols * 4 > ((size_t)512 << 20))
    return -1000;

  // chunks of whole 16-tile groups
  const long ngroups = ((long)nb + kGroupRows - 1) / kGroupRows;
  const long nfg = (long)nb / kGroupRows, rem = (long)nb % kGroupRows;
  constexpr int kMaxChunks = 16;
  int C = (int)(total_bytes / ((size_t)64 << 20));
  if (const char *e = getenv("YAEL_B200_H2D_CHUNKS")) C = atoi(e) > 0 ? atoi(e) : C;
  if (C < 2) C = 2;
  if (C > kMaxChunks) C = kMaxChunks;
  if (C > ngroups) C = (int)ngroups;
  Tf32Plan cplan[kMaxChunks];
  long g0[kMaxChunks + 1];
  int list0[kMaxChunks + 1];
  size_t tf_ws = splan.ws_bytes;
  list0[0] = 0;
  // The scan of a chunk overlaps the transfer of the next one; only the LAST chunk's scan (and the
  // merge / re-rank behind it) is exposed after the transfer ends.  So the last two chunks are
  // small (1/16 and 1/32 of the database) and the others share the rest evenly.
  long bound[kMaxChunks + 1];
  {
    const bool taper = C >= 4 && ngroups >= 64 && !getenv("YAEL_B200_H2D_UNIFORM");
    const long tail1 = taper ? ngroups / 32 : 0, tail2 = taper ? ngroups / 16 : 0;
    const int cu = taper ? C - 2 : C;           // evenly sized chunks
    const long gu = ngroups - tail1 - tail2;    // groups they cover
    for (int c = 0; c <= cu; c++) bound[c] = gu * c / cu;
    if (taper) {
      bound[C - 1] = gu + tail2;
      bound[C] = ngroups;
    }
  }
  for (int c = 0; c < C; c++) {
    g0[c] = bound[c];
    const long g1 = bound[c + 1];
    const long r0 = g0[c] * kGroupRows, r1 = g1 * kGroupRows < nb ? g1 * kGroupRows : nb;
    cplan[c] = tf32_plan_tiles(nq, tf32_tiles((int)(r1 - r0)), dpad, kp);
    if (!cplan[c].ok) return -1000;
    if (cplan[c].ws_bytes > tf_ws) tf_ws = cplan[c].ws_bytes;
    list0[c + 1] = list0[c] + cplan[c].lists;
  }
  g0[C] = ngroups;
  const int lists = list0[C];
  if (lists > ML_LISTS) return -1000;
  const size_t stride = (size_t)lists * kp;

  size_t need = Carver::need(sizeof(float) * (size_t)padded) + Carver::need(64) +
                2 * Carver::need(sizeof(float) * (size_t)nq * stride) +
                2 * Carver::need(sizeof(float) * (size_t)nq * lists) +
                Carver::need(sizeof(int) * (size_t)nq * kp) + 2 * Carver::need(sizeof(int) * (size_t)nq) +
                Carver::need(tf_ws) + 1024 + Carver::need(sizeof(float) * (size_t)nb * dpad) +
                Carver::need(sizeof(float) * (size_t)nq * dpad) + center_ws_bytes(nb, d) +
                Carver::need(sizeof(float) * (size_t)nq * gcols) + Carver::need(sizeof(float) * (size_t)nq);
  int n_flag = 0;
  cudaStream_t cs = copy_stream();
  cudaEvent_t ev_begin = nullptr, ev_sample = nullptr, ev_chunk[kMaxChunks] = {};
  auto new_event = [](cudaEvent_t *e) { return cudaEventCreateWithFlags(e, cudaEventDisableTiming); };
  YB_CUDA(new_event(&ev_begin));
  YB_CUDA(new_event(&ev_sample));
  for (int c = 0; c < C; c++) YB_CUDA(new_event(&ev_chunk[c]));
  struct EventGuard {
    cudaEvent_t *a, *b, *c;
    int n;
    ~EventGuard() {
      cudaEventDestroy(*a);
      cudaEventDestroy(*b);
      for (int i = 0; i < n; i++) cudaEventDestroy(c[i]);
    }
  } eguard = {&ev_begin, &ev_sample, ev_chunk, C};
  {
    ScratchScope ws(need, st);
    Carver c(ws.p);
    float *an = c.take<float>(padded);
    float *scal = c.take<float>(16);
    float *cscore = c.take<float>((size_t)nq * stride);
    int *cid = c.take<int>((size_t)nq * stride);
    float *cthr = c.take<float>((size_t)nq * lists);
    int *ccnt = c.take<int>((size_t)nq * lists);
    int *sel = c.take<int>((size_t)nq * kp);
    int *flags = c.take<int>(nq);
    int *flag_list = c.take<int>(nq);
    void *tfws = c.take<char>(tf_ws);
    float *base_c = c.take<float>((size_t)nb * dpad);
    float *query_c = c.take<float>((size_t)nq * dpad);
    char *cws = c.take<char>(center_ws_bytes(nb, d));
    float *gm = c.take<float>((size_t)nq * gcols);
    float *thr_init = c.take<float>(nq);
    int rc;

    // ---- copy stream: sample tiles first (the staging buffer may still be read by earlier
    // work on the compute stream: order behind it)
    YB_CUDA(cudaEventRecord(ev_begin, st));
    YB_CUDA(cudaStreamWaitEvent(cs, ev_begin, 0));
    if (nfg > 0)
      YB_CUDA(cudaMemcpy2DAsync(base, pitch, base_host, pitch, 256 * row_bytes, (size_t)nfg,
                                cudaMemcpyHostToDevice, cs));
    if (rem > 0) {
      const long r = nfg * kGroupRows, n = rem < 256 ? rem : 256;
      YB_CUDA(cudaMemcpyAsync(base + r * d, base_host + r * d, (size_t)n * row_bytes,
                              cudaMemcpyHostToDevice, cs));
    }
    YB_CUDA(cudaEventRecord(ev_sample, cs));

    // ---- compute stream: centring vector, queries, sample tiles, thresholds
    YB_CUDA(cudaStreamWaitEvent(st, ev_sample, 0));
    Carver cc(cws);
    float *psum = cc.take<float>((size_t)CM_BLOCKS * d);
    int *pcnt = cc.take<int>((size_t)CM_BLOCKS * d);
    float *mu = cc.take<float>(d);
    {
      ProfScope ps(0, st);
      int nblk = ngroups < CM_BLOCKS ? (int)ngroups : CM_BLOCKS;
      const long step = (ngroups / nblk) * kGroupRows;  // blocks start on sample tiles
      k_col_partial<<<nblk, 256, 0, st>>>(base, nb, d, step, psum, pcnt);
      YB_LAUNCH_CHECK();
      k_col_final<<<(d + 127) / 128, 128, 0, st>>>(psum, pcnt, nblk, d, mu);
      YB_LAUNCH_CHECK();
      launch_center_rows(query, nq, d, dpad, mu, query_c, nullptr, st);
      launch_center_rows(base, nb, d, dpad, mu, base_c, an, st, RowMap{256, kGroupRows});
      YB_CUDA(cudaGetLastError());
      if ((rc = fill_f32(an + nb, padded - nb, __builtin_inff(), st))) return rc;
      YB_CUDA(cudaMemsetAsync(scal, 0, 64, st));
      k_mu_norm<<<1, 32, 0, st>>>(mu, d, scal + 7);
      YB_LAUNCH_CHECK();
    }
    {
      ProfScope ps(10, st);
      if ((rc = tf32_group_min(splan, nq, nb, dpad, nbt_s, kSampleStride, base_c, query_c, an, gm,
                               gcols, gsize, tfws, st)))
        return rc;
      k_row_kth<<<nq, RK_T, 0, st>>>(gm, gcols, (int)gcols, j2, thr_init);
      YB_LAUNCH_CHECK();
    }

    // ---- chunks: copy, centre, scan
    for (int ch = 0; ch < C; ch++) {
      const long ga = g0[ch], gb = g0[ch + 1];
      const long gfull = gb < nfg ? gb : nfg;
      if (gfull > ga) {
        const long r = ga * kGroupRows + 256;
        YB_CUDA(cudaMemcpy2DAsync(base + r * d, pitch, base_host + r * d, pitch,
                                  (size_t)(kGroupRows - 256) * row_bytes, (size_t)(gfull - ga),
                                  cudaMemcpyHostToDevice, cs));
      }
      if (gb > nfg && rem > 256) {  // the partial last group
        const long r = nfg * kGroupRows + 256;
        YB_CUDA(cudaMemcpyAsync(base + r * d, base_host + r * d, (size_t)(nb - r) * row_bytes,
                                cudaMemcpyHostToDevice, cs));
      }
      YB_CUDA(cudaEventRecord(ev_chunk[ch], cs));
      YB_CUDA(cudaStreamWaitEvent(st, ev_chunk[ch], 0));
      const long r0 = ga * kGroupRows, r1 = gb * kGroupRows < nb ? gb * kGroupRows : nb;
      {
        ProfScope ps(0, st);
        launch_center_rows(base + r0 * d, r1 - r0, d, dpad, mu, base_c + r0 * dpad, an + r0, st);
        YB_CUDA(cudaGetLastError());
      }
      {
        ProfScope ps(1, st);
        Tf32Out oo = {};
        oo.out_cnt = ccnt;
        oo.lists_ld = lists;
        oo.list0 = list0[ch];
        oo.id0 = (int)r0;
        if ((rc = tf32_shortlist(cplan[ch], nq, (int)(r1 - r0), dpad, tf32_tiles((int)(r1 - r0)), 1,
                                 base_c + r0 * dpad, query_c, an + r0, thr_init, cscore, cid, cthr,
                                 tfws, st, &oo)))
          return rc;
      }
    }
    k_sqrt_max<<<2 * sm_count(), 256, 0, st>>>(an, nb, scal);
    YB_LAUNCH_CHECK();
    {
      ProfScope ps(2, st);
      k_merge_lists<<<nq, ML_T, 0, st>>>(ccnt, cscore, lists, kp, kp, sel, nullptr);
      YB_LAUNCH_CHECK();
    }
    RerankArgs A = {};
    A.nq = nq; A.nb = nb; A.d = d; A.k = k; A.base = base; A.query = query;
    A.dis = dis; A.assign = assign; A.id_offset = id_offset;
    A.cand_id = cid; A.cand_score = cscore; A.sel = sel;
    A.cand_stride = (int)stride; A.m = m; A.all_listed = 0;
    A.cand_thr = cthr; A.lists = lists; A.query_c = query_c; A.qc_ld = dpad;
    A.err_scale = kTf32ErrScale; A.bmax = scal; A.uncert_flags = flags;
    A.mu_norm = scal + 7;
    A.gsort = nullptr; A.m_pad = m_pad; A.k1 = 0;
    rerank_attrs();
    {
      ProfScope ps(3, st);
      k_rerank<1><<<nq, RR_T, rerank_setup(A, smem), st>>>(A);
      YB_LAUNCH_CHECK();
    }
    k_collect_flags<<<(nq + 255) / 256, 256, 0, st>>>(flags, nq, flag_list, (int *)(scal + 1));
    YB_LAUNCH_CHECK();
    if ((rc = redo_flagged_exact(nq, nb, d, k, base, query, assign, dis, id_offset, flag_list,
                                 (int *)(scal + 1), &n_flag, st)))
      return rc;
  }
  *engine_out = 1;
  *uncert_out = n_flag;
  return 0;
}

}  // namespace yb

using namespace yb;

// Debug / test entry: raw scores s[q][n] = |b_n|^2 - 2 <q, b_n> of the FP16-operand tensor pass
// (no centring; scale chosen as in the real pipeline).  Returns 7 when a value overflowed FP16.
extern "C" int yb_debug_f16_scores(int nq, int nb, int d, const float *base, const float *query,
                                   float *scores, yb_stream_t s) {
  Guard g;
  cudaStream_t st = stream_of(s);
  const int dh = (d + 15) & ~15, dop = dh + kNfExtra;
  Tf32Plan plan = tf32_plan(nq, nb, dop, 1, 3);
  if (!plan.ok) return fail(3, "tensor path does not support this shape (d=%d)", d);
  const long padded = tf32_padded_rows(nb);
  ScratchScope ws(Carver::need(4ull * padded) + Carver::need(64) + Carver::need(4ull * d) +
                      Carver::need(2ull * padded * dop) + Carver::need(2ull * nq * dop) +
                      Carver::need(plan.ws_bytes),
                  st);
  Carver c(ws.p);
  float *bn = c.take<float>(padded);
  float *scal = c.take<float>(16);
  float *mu = c.take<float>(d);
  __half *bh = c.take<__half>((size_t)padded * dop);
  __half *qh = c.take<__half>((size_t)nq * dop);
  void *tws = c.take<char>(plan.ws_bytes);
  YB_CUDA(cudaMemsetAsync(scal, 0, 64, st));
  YB_CUDA(cudaMemsetAsync(mu, 0, 4ull * d, st));
  k_absmax_sample<<<(nb + CM_ROWS - 1) / CM_ROWS, 256, 0, st>>>(base, nb, d, CM_ROWS, scal + 6);
  YB_LAUNCH_CHECK();
  k_absmax_sample<<<(nq + CM_ROWS - 1) / CM_ROWS, 256, 0, st>>>(query, nq, d, CM_ROWS, scal + 6);
  YB_LAUNCH_CHECK();
  k_pick_scale<<<1, 1, 0, st>>>(scal, d, mu);
  YB_LAUNCH_CHECK();
  launch_center_rows_h(base, nb, d, dh, dh, 1, mu, scal, bh, nullptr, bn, st);
  launch_center_rows_h(query, nq, d, dh, dh, 2, mu, scal, qh, nullptr, nullptr, st);
  {
    const long tot = (padded - nb) * dop;
    if (tot > 0) {
      k_fill_pad_rows_h<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(bh, nb, padded, dop, dh);
      count_launch();
    }
  }
  int rc;
  if ((rc = fill_f32(bn + nb, padded - nb, __builtin_inff(), st))) return rc;
  plan.acc_scale = scal + 3;
  if ((rc = tf32_scores(plan, nq, nb, dop, tf32_tiles(nb), 1, (const float *)bh, (const float *)qh, bn,
                        scores, nb, tws, st)))
    return rc;
  int overflow = 0;
  YB_CUDA(cudaMemcpyAsync(&overflow, scal + 4, sizeof(int), cudaMemcpyDeviceToHost, st));
  YB_CUDA(cudaStreamSynchronize(st));
  return overflow ? fail(7, "a value overflowed FP16") : 0;
}

// dists[j][i] *= w[i] (yael/nn.c:497-500), j < nr queries, i < nb base vectors
__global__ void k_weight_slab(float *__restrict__ slab, long nb, long nr, const float *__restrict__ w) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nb * nr) slab[t] = __fmul_rn(slab[t], __ldg(w + t % nb));
}

// knn_full for the distance types that are not a contraction (yael/nn.c:451-525 with
// compute_cross_distances_alt, 280-350): distance slab -> optional per-base weights -> per-row
// select, in query chunks.  k == 1 follows nn_single_full (yael/nn.c:404-440): strict '<' from
// (-1, 1e30f) for every distance type.
extern "C" int yb_knn_alt(int distance_type, int nq, int nb, int d, int k, const float *base,
                          const float *query, const float *b_weights, int *assign, float *dis,
                          yb_stream_t s) {
  if (nq <= 0) return 0;
  if (k <= 0 || k > nb) return fail(3, "yb_knn_alt: need 0 < k <= nb (k=%d, nb=%d)", k, nb);
  Guard g;
  cudaStream_t st = stream_of(s);
  const size_t rows = exact_chunk_rows(nq, nb);
  // pooled blocks, not the device workspace: the distance call below reserves that itself
  float *slab = (float *)yb_malloc(sizeof(float) * rows * (size_t)nb);
  void *kws = yb_malloc(kmin_ws_bytes((long)rows, k));
  int rc = 0;
  for (long q0 = 0; q0 < nq && !rc; q0 += (long)rows) {
    const long nr = nq - q0 < (long)rows ? nq - q0 : (long)rows;
    rc = yb_cross_distances_alt(distance_type, d, nb, (int)nr, base, d, query + q0 * d, d, slab, nb, s);
    if (!rc && b_weights) {
      const long tot = nr * nb;
      k_weight_slab<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(slab, nb, nr, b_weights);
      count_launch();
    }
    if (!rc)
      rc = kmin_rows(slab, nb, nb, nr, k, +1, assign + q0 * k, dis + q0 * k, 0, k == 1 ? 1 : 0, kws, st);
  }
  yb_free(slab);
  yb_free(kws);
  return rc;
}

extern "C" void yb_set_knn_engine(int engine) { g_engine_force = engine; }
extern "C" int yb_last_knn_engine(void) { return g_last_engine; }
extern "C" int yb_last_knn_operands(void) { return g_last_operands; }
extern "C" long yb_last_knn_uncertified(void) { return g_last_uncert; }

extern "C" int yb_knn_l2(int nq, int nb, int d, int k, const float *base, const float *query,
                          const float *b_weights, int *assign, float *dis, int id_offset,
                          yb_stream_t s) {
  if (nq <= 0) return 0;
  if (k <= 0 || k > nb) return fail(3, "yb_knn_l2: need 0 < k <= nb (k=%d, nb=%d)", k, nb);
  Guard g;
  cudaStream_t st = stream_of(s);
  g_last_operands = tensor_operand_kind();
  int rc = knn_tf32_path(nq, nb, d, k, base, query, b_weights, assign, dis, id_offset,
                         g_engine_force, &g_last_engine, &g_last_uncert, st, g_last_operands, 0);
  if (rc == -1001)  // data outside FP16's range (even scaled): the same pass on TF32 operands
    g_last_operands = 0, rc = knn_tf32_path(nq, nb, d, k, base, query, b_weights, assign, dis, id_offset,
                       g_engine_force, &g_last_engine, &g_last_uncert, st, 0, 0);
  if (rc != -1000) return rc;  // -1000: tensor-core path not applicable
  g_last_engine = 0;
  g_last_uncert = 0;
  ScratchScope ws(knn_exact_ws_bytes(nq, nb, k), st);
  return knn_exact(nq, nb, d, k, base, query, b_weights, assign, dis, id_offset, ws.p, st);
}

// knn_full() with the database in HOST memory (pinned for full PCIe speed; pageable works):
// base_dev is device scratch for nb*d floats that holds the database when the call returns.
extern "C" int yb_knn_l2_hostbase(int nq, int nb, int d, int k, const float *base_host,
                                   float *base_dev, const float *query, int *assign, float *dis,
                                   int id_offset, yb_stream_t s) {
  if (nq <= 0) return 0;
  if (k <= 0 || k > nb) return fail(3, "yb_knn_l2_hostbase: need 0 < k <= nb (k=%d, nb=%d)", k, nb);
  Guard g;
  cudaStream_t st = stream_of(s);
  const char *off = getenv("YAEL_B200_NO_STREAMED_H2D");
  if (!(off && atoi(off))) {
    int rc = knn_tf32_streamed(nq, nb, d, k, base_host, base_dev, query, assign, dis, id_offset,
                               g_engine_force, &g_last_engine, &g_last_uncert, st);
    if (rc != -1000) return rc;
  }
  YB_CUDA(cudaMemcpyAsync(base_dev, base_host, sizeof(float) * (size_t)nb * d, cudaMemcpyHostToDevice, st));
  return yb_knn_l2(nq, nb, d, k, base_dev, query, nullptr, assign, dis, id_offset, s);
}

// shard_stride: elements between the lists of consecutive shards (nq * k when every shard's
// [nq][k] block follows the previous one; 2 * nq * k when ids and distances of a shard travel
// in one [2][nq][k] buffer, i.e. one collective)
extern "C" int yb_knn_merge_strided(int nq, int k, int G, const int *assign_in, const float *dis_in,
                                    long shard_stride, int *assign_out, float *dis_out,
                                    yb_stream_t s) {
  if (nq <= 0 || k <= 0 || G <= 0) return 0;
  Guard g;
  cudaStream_t st = stream_of(s);
  int m_pad = pow2_ceil(G * k < 2 ? 2 : G * k);
  size_t wsb = m_pad <= 4096 ? 256 : Carver::need(sizeof(unsigned long long) * (size_t)nq * m_pad);
  ScratchScope ws(wsb, st);
  size_t smem = m_pad <= 4096 ? sizeof(unsigned long long) * (size_t)m_pad : 0;
  k_knn_merge<<<nq, 128, smem, st>>>(k, G, nq, assign_in, dis_in, shard_stride, assign_out, dis_out,
                                     (unsigned long long *)ws.p, m_pad);
  YB_LAUNCH_CHECK();
  return 0;
}

extern "C" int yb_knn_merge(int nq, int k, int G, const int *assign_in, const float *dis_in,
                             int *assign_out, float *dis_out, yb_stream_t s) {
  return yb_knn_merge_strided(nq, k, G, assign_in, dis_in, (long)nq * k, assign_out, dis_out, s);
}

extern "C" int yb_knn_reorder_shortlist(int nq, int nb, int d, int k, const float *base,
                                         const float *query, int *idx, float *dis,
                                         yb_stream_t s) {
  (void)nb;
  if (nq <= 0 || k <= 0) return 0;
  Guard g;
  cudaStream_t st = stream_of(s);
  int m_pad = pow2_ceil(k < 2 ? 2 : k);
  size_t wsb = m_pad <= 4096 ? 256 : Carver::need(sizeof(unsigned long long) * (size_t)nq * m_pad);
  ScratchScope ws(wsb, st);
  RerankArgs A = {};
  A.nq = nq; A.nb = nb; A.d = d; A.k = k; A.base = base; A.query = query;
  A.idx = idx; A.dis = dis; A.m = k; A.m_pad = m_pad; A.gsort = (unsigned long long *)ws.p;
  size_t smem = rerank_smem_bytes(d, k, m_pad);
  if (smem > 160 * 1024) return fail(3, "knn_reorder_shortlist: d=%d k=%d too large for one CTA", d, k);
  rerank_attrs();
  k_rerank<0><<<nq, RR_T, rerank_setup(A, smem), st>>>(A);
  YB_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ hkm_quantize (yael/hkm.c:144-162)
// One level of the tree walk: the bf children of every point's current node are its candidates
// (ids in the level's table), k_rerank_k1 scores them with the reference's arithmetic and takes the
// (distance, id) minimum -- nn() with k = 1 (yael/nn.c:608-621, 383-446).
__global__ void k_hkm_children(const int *__restrict__ vw, long n, int bf, int *__restrict__ cand) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * bf) return;
  const int node = vw[e / bf];
  cand[e] = node >= 0 ? node * bf + (int)(e % bf) : -1;
}

// a point whose children all have NaN distances gets child -1, as in the reference (nn_single_full
// starts from (-1, 1e30): vw * bf + (-1))
__global__ void k_hkm_step(const int *__restrict__ vw, long n, int bf, int *__restrict__ next) {
  const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  if (next[q] < 0) next[q] = vw[q] >= 0 ? vw[q] * bf - 1 : -1;
}

// idx[i] = leaf of point i.  levels[l]: DEVICE pointer to level l's [bf^(l+1)][d] table (the array of
// pointers itself is host memory); v, idx: device.
extern "C" int yb_hkm_quantize(int nlevel, int bf, int d, const float *const *levels, long n,
                               const float *v, int *idx, yb_stream_t s) {
  if (n <= 0) return 0;
  if (nlevel == 0) {  // a tree without levels has one leaf (the reference's loop leaves vw = 0)
    Guard g0;
    YB_CUDA(cudaMemsetAsync(idx, 0, sizeof(int) * (size_t)n, stream_of(s)));
    return 0;
  }
  if (nlevel < 0 || bf <= 0 || d <= 0) return fail(3, "yb_hkm_quantize: nlevel=%d bf=%d d=%d", nlevel, bf, d);
  if (n > 0x7fffffffL / (bf > 4 ? bf : 4)) return fail(3, "yb_hkm_quantize: n = %ld x bf = %d is too large", n, bf);
  Guard g;
  cudaStream_t st = stream_of(s);
  size_t need = Carver::need(sizeof(int) * (size_t)n * bf) + 3 * Carver::need(sizeof(int) * (size_t)n) +
                Carver::need(sizeof(float) * (size_t)n) + Carver::need(sizeof(double) * (size_t)n);
  ScratchScope ws(need, st);
  Carver c(ws.p);
  int *cand = c.take<int>((size_t)n * bf);
  int *cur = c.take<int>(n);
  int *next = c.take<int>(n);
  int *flags = c.take<int>(n);
  float *dis = c.take<float>(n);
  double *qnorm = c.take<double>(n);
  YB_CUDA(cudaMemsetAsync(cur, 0, sizeof(int) * (size_t)n, st));
  int rc;
  if ((rc = row_norms_seq(v, n, d, d, nullptr, qnorm, st))) return rc;
  for (int l = 0; l < nlevel; l++) {
    const long tot = n * bf;
    k_hkm_children<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(cur, n, bf, cand);
    YB_LAUNCH_CHECK();
    if ((rc = rerank_k1((int)n, d, bf, 0, levels[l], v, cand, nullptr, nullptr, nullptr,
                        l + 1 == nlevel ? idx : next, dis, 0, flags, qnorm, st)))
      return rc;
    k_hkm_step<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(cur, n, bf, l + 1 == nlevel ? idx : next);
    YB_LAUNCH_CHECK();
    int *t = cur; cur = next; next = t;
  }
  return 0;
}

extern "C" int yb_gather_rows(const float *src, const int *rows, int n, int d, float *dst,
                               yb_stream_t s) {
  if (n <= 0 || d <= 0) return 0;
  Guard g;
  cudaStream_t st = stream_of(s);
  long tot = (long)n * d;
  k_gather_rows<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(src, rows, n, d, dst);
  YB_LAUNCH_CHECK();
  return 0;
}
