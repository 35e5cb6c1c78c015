// yb_hamming_tc.cu -- nn_hamming on the tensor cores (engine 1 of yb_nn_hamming).
//
// At the BASELINE shape (10^4 queries x 10^7 codes) the popcount scan of yb_hamming.cu is bound
// by the POPC pipe (16 / clk / SM), not by HBM: 10^11 pairs cost >= 43 ms however the codes are
// loaded.  The same distances are an exact dense contraction: write every bit as +-1,
//     <q, b> = (#equal bits) - (#different bits) = bits - 2 * ham(q, b),
// so with |b|^2 replaced by the constant 2 * bits the score of the distance GEMM kernel
// (yb_knn_tf32.cu: s = |b|^2 - 2 <q, b>) is  s = 4 * ham(q, b)  -- and with E4M3 operands
// (+1.0 = 0x38, -1.0 = 0xB8; tcgen05.mma kind::f8f6f4, FP32 accumulation) every product and
// every partial sum is a small integer, i.e. the result is EXACT, not approximate.  The pass
// therefore needs no re-rank: it reuses the fused threshold/top-k epilogue of k_knn_tf32 and a
// finishing kernel orders the admitted (distance, id) pairs and certifies them:
//
//   every row that is not in a list has score >= T (the smallest list threshold); the k best
//   listed pairs are the answer iff the k-th of them has score < T.  Scores are integers and
//   ties are ordered by id, so this is bit-exact against compute_hamming + stable selection
//   (yael/hamming.c:177-219; oracle orc_nn_hamming).  Queries that fail the test (lists that
//   had to drop ties, adversarial row order) are re-done by the popcount scan.
//
// Work: 2 MMAs (K = 32 each) per 128 x 256 tile for 64-bit codes, so the pass is bound by the
// epilogue (one FFMA + one FMNMX per pair instead of 2 POPC + xor/add/compare).
#include <stdlib.h>

#include "yb_common.cuh"
#include "yb_internal.cuh"

namespace yb {

// ------------------------------------------------------------------ operand expansion
// Codes -> E4M3 rows.  An output row holds `slots` copies ("slots") of 64*W elements; element j
// of slot i is +-scale_i (bit set: +), scale = 1, 16, 256 (0x38, 0x58, 0x78).  Queries
// (interleave = 0) repeat their own code in every slot; the database (interleave = 1) packs
// `slots` CONSECUTIVE rows into one combined row (row slots*c + i in slot i; absent rows are
// zeros).  With both sides scaled the accumulator of (query, combined row c) is
//     acc = dot_0 + 2^8 dot_1 + 2^16 dot_2,   dot_i = <q, b_{slots*c+i}> = bits - 2 ham_i,
// every MMA (K = 32 elements) lying inside one slot: one 32-bit TMEM word carries up to three
// distances.  One thread per 16 bits -> 16 bytes (one 128-bit store).
__global__ void __launch_bounds__(256)
k_ham_expand(const unsigned long long *__restrict__ codes, long n_src, int W, int slots,
             int interleave, long n_out, uint4 *__restrict__ out) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int units = slots * W * 4;  // 16-byte units per output row
  if (t >= n_out * units) return;
  const long row_out = t / units;
  const int u = (int)(t - row_out * units);
  const int slot = u / (W * 4), w16 = u - slot * (W * 4);
  const long src = interleave ? row_out * slots + slot : row_out;
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (src < n_src) {
    const unsigned bits = (unsigned)(codes[src * W + (w16 >> 2)] >> (16 * (w16 & 3))) & 0xffffu;
    const unsigned neg = (slot == 0 ? 0xB8B8B8B8u : (slot == 1 ? 0xD8D8D8D8u : 0xF8F8F8F8u));
    auto four = [neg](unsigned x) {  // 4 bits -> 4 bytes: -scale with the sign cleared where the bit is set
      const unsigned spread = (x & 1u) | ((x & 2u) << 7) | ((x & 4u) << 14) | ((x & 8u) << 21);
      return neg ^ (spread << 7);
    };
    o.x = four(bits & 15u);
    o.y = four((bits >> 4) & 15u);
    o.z = four((bits >> 8) & 15u);
    o.w = four((bits >> 12) & 15u);
  }
  out[t] = o;
}

__global__ void k_thr_bump(float *__restrict__ thr, int n, float add) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) thr[i] += add;  // +inf stays +inf
}

__global__ void k_ham_collect(const int *__restrict__ flags, int nq, int *list, int *count) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nq && flags[q]) list[atomicAdd(count, 1)] = q;
}

// ------------------------------------------------------------------ finish: order + certify
// One CTA per query: gather the entries of its lists as (distance << 32 | row) keys, sort them,
// emit the k smallest and flag the query when they are not provably the answer.
constexpr int HF_T = 256;
constexpr int HF_CAP = 8192;      // keys per query held in shared memory (64 KB)
constexpr int HF_LISTS = 1024;    // lists per query

__global__ void __launch_bounds__(HF_T)
k_ham_tc_finish(int nq, int nb, int k, int lists, int kp, const int *__restrict__ cnt,
                const float *__restrict__ score, const int *__restrict__ id,
                const float *__restrict__ lthr, int id_offset, int *__restrict__ assign,
                uint16_t *__restrict__ dis, int *__restrict__ flags) {
  extern __shared__ unsigned long long keys[];
  __shared__ int loff[HF_LISTS + 1];
  __shared__ float lthr_s[HF_LISTS];
  __shared__ float tmin_s;
  const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int *qcnt = cnt + (size_t)q * lists;
  for (int l = tid; l < lists; l += HF_T) {
    loff[l + 1] = qcnt[l];
    lthr_s[l] = lthr[(size_t)q * lists + l];
  }
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    float t = __uint_as_float(0x7f800000u);
    for (int l = 0; l < lists; l++) {
      const int c = loff[l + 1];
      loff[l] = acc;
      acc += c;
      t = fminf(t, lthr_s[l]);
    }
    loff[lists] = acc;
    tmin_s = t;
  }
  __syncthreads();
  const int total = loff[lists];
  if (total > HF_CAP || total < k) {  // uniform over the CTA
    if (tid == 0) flags[q] = 1;
    return;
  }
  int m_pad = 2;
  while (m_pad < total) m_pad <<= 1;
  for (int l = warp; l < lists; l += HF_T / 32) {
    const int o = loff[l], n = loff[l + 1] - o;
    const size_t src = ((size_t)q * lists + l) * kp;
    for (int e = lane; e < n; e += 32) {
      const unsigned dist = (unsigned)__float2int_rn(score[src + e] * 0.25f);
      const unsigned row = (unsigned)id[src + e];
      // (rows of the tile padding can never be listed -- NaN accumulators -- but cost nothing to refuse)
      keys[o + e] = row < (unsigned)nb ? (((unsigned long long)dist << 32) | row) : ~0ull;
    }
  }
  for (int e = total + tid; e < m_pad; e += HF_T) keys[e] = ~0ull;
  bitonic_sort_u64(keys, m_pad, tid, HF_T, [] { __syncthreads(); });
  // certificate: every unlisted row has score >= tmin, the k-th listed one must be below it
  const float sk = 4.0f * (float)(unsigned)(keys[k - 1] >> 32);
  const bool ok = keys[k - 1] != ~0ull && sk < tmin_s;
  if (tid == 0) flags[q] = ok ? 0 : 1;
  if (!ok) return;
  for (int j = tid; j < k; j += HF_T) {
    const unsigned long long v = keys[j];
    assign[(size_t)q * k + j] = (int)(unsigned)v + id_offset;
    dis[(size_t)q * k + j] = (uint16_t)(v >> 32);
  }
}

// ------------------------------------------------------------------ host side
static int expand_codes(const unsigned long long *codes, long n_src, int W, int slots,
                        int interleave, long n_out, void *out, cudaStream_t st) {
  const long threads = n_out * slots * W * 4;
  if (threads <= 0) return 0;
  k_ham_expand<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(codes, n_src, W, slots, interleave,
                                                                  n_out, (uint4 *)out);
  YB_LAUNCH_CHECK();
  return 0;
}

// sampling geometry: every `stride`-th tile is scored once, the minima of runs of `gsize` columns
// are kept (at most row_kth_max_n() of them per query), the j2-th smallest of those is the
// admission threshold -- about j2 * stride rows of the whole database are at or below it
struct HamSample {
  int ok, stride, nbt_s, gsize, j2;
  long gcols;
};
static HamSample ham_sample_geometry(int nbt, int k) {
  HamSample s = {};
  if (nbt < 128) return s;
  const int maxn = row_kth_max_n();
  s.stride = (nbt + 2047) / 2048;
  if (s.stride < 16) s.stride = 16;
  s.nbt_s = (nbt + s.stride - 1) / s.stride;
  const long srows = (long)s.nbt_s * 256;
  s.gsize = 16;
  while (s.gsize < 128 && srows / s.gsize > maxn) s.gsize *= 2;
  s.gcols = srows / s.gsize;
  s.j2 = (3 * k + s.stride - 1) / s.stride;
  // floor of the order-statistic rank: Hamming distances fall into coarse classes (x3 rows per unit of
  // distance at the BASELINE shape), so the threshold snaps to a class boundary and a small rank is
  // stable.  Measured (10 k queries, k = 100, pass + finish in ms, 0 scan fallbacks throughout):
  // 10 M rows: rank 32 -> 13.7 + 0.92, 20 -> 12.9 + 0.60, 16 -> 12.2 + 0.39, 12 -> 12.0 + 0.35;
  // 1.25 M rows: 3.15 + 0.48, 2.51 + 0.30, 2.46 + 0.28, 2.46 + 0.29
  int j2_floor = 16;
  if (const char *e = getenv("YAEL_B200_HAM_J2")) j2_floor = atoi(e) > 0 ? atoi(e) : j2_floor;  // A/B knob
  if (s.j2 < j2_floor) s.j2 = j2_floor;
  s.ok = s.gcols <= maxn && (long)s.j2 * 4 <= s.gcols;
  return s;
}

// database rows per accumulator.  Codes of up to 64 bits CAN pack 3 rows, 128 bits 2 (the packed
// fields are one byte each and the sum must stay below 2^23), longer ones 1; fewer when k' leaves
// no room for a tile of packed appends.  Round 1 packed from 8 M rows on (the float epilogue of the
// one-row pass cost 2117 clk per tile: 10 M rows 23.0 ms unpacked vs 21.6 packed).  Round 2: with the
// constant-norm MAX-tree epilogue (operand kind 5: wide TMEM loads, early hand-back, no |b|^2 ring)
// the one-row pass takes ~1000 clk per 256-row tile and wins at every size (10 M rows: 12.2 ms vs
// 19.7 packed; 1.25 M: 2.5 vs 6.8) -- the ncu source page of the packed pass showed why packing could
// not pay: ~57 instructions per 16 accumulators at 0.17 IPC per warp plus a 160-instruction decoder
// on 15 % of the groups.  Default: one row per accumulator; YAEL_B200_HAM_SLOTS=2|3 selects the packed
// pass (tests keep it covered, A/B measurements).
static int ham_slots_for(int W, int k, long nb) {
  const int smax = W == 1 ? 3 : (W == 2 ? 2 : 1);
  (void)nb;
  int S = 1;
  if (const char *e = getenv("YAEL_B200_HAM_SLOTS")) {
    const int want = atoi(e);
    if (want >= 1) S = want < smax ? want : smax;
  }
  while (S > 1 && tf32_kprime_for(k) + 256 * S > 1024) S--;
  return S;
}

static Tf32Plan ham_plan(int nq, int nb, int W, int k, int S) {
  const int nc = (nb + S - 1) / S;
  // packed passes: E4M3 with the integer epilogue (planning kind 4 = OP_F8P, then kind 1); one row
  // per accumulator: kind 5 (E4M3, constant |b|^2 = 2 * bits: s = 2 bits - 2 dot = 4 ham) with the
  // MAX-tree epilogue of the folded-norm k-NN pass.  Both may run as cta_group::2 pairs.
  Tf32Plan plan = tf32_plan(nq, nc, 16 * W * S, k, S > 1 ? 4 : 5);
  if (S > 1) {
    plan.kind = 1;
    plan.ham_slots = S;
    plan.ham_nb = nb;
    plan.ham_magic = 8388608.0f + (float)(32 * W) * (S == 3 ? 65793.0f : 257.0f);
  } else {
    plan.score_c0 = 128.0f * (float)W;
  }
  return plan;
}

bool hamming_tc_supported(int nq, int nb, int W, int k) {
  if (W < 1 || W > 8 || nq < 1 || nb < 1 || k < 1 || k > nb) return false;
  Tf32Plan plan = ham_plan(nq, nb, W, k, ham_slots_for(W, k, nb));
  return plan.ok && plan.pair != 1 && plan.lists <= HF_LISTS;
}

// pb / pq: codes packed as W 64-bit words per row.  Results for every query whose certificate
// holds are written to assign / dis; the others are listed in flag_list_out (device memory from
// yb_malloc, owned by the caller when *n_flag_out > 0).  Returns -1000 when the shape does not
// qualify.
int hamming_tc(int nq, int nb, int W, int k, const unsigned long long *pb,
               const unsigned long long *pq, int *assign, uint16_t *dis, int id_offset,
               int **flag_list_out, int *n_flag_out, cudaStream_t st) {
  *flag_list_out = nullptr;
  *n_flag_out = 0;
  if (!hamming_tc_supported(nq, nb, W, k)) return -1000;
  const int S = ham_slots_for(W, k, nb);
  const int nc = (nb + S - 1) / S;   // combined rows the tensor pass sees
  const int dfl = 16 * W * S;        // row pitch in floats (64 * W * S one-byte elements)
  const int bits = 64 * W;
  Tf32Plan plan = ham_plan(nq, nb, W, k, S);
  const int kp = plan.kprime;
  const int nbt = tf32_tiles(nc);
  const long padded = tf32_padded_rows(nc);
  const HamSample sg = ham_sample_geometry(nbt, k);
  Tf32Plan splan = {};
  if (sg.ok) {
    splan = tf32_plan_tiles(nq, sg.nbt_s, dfl, sg.j2, S > 1 ? 0 : 5);
    if (S > 1) splan.kind = 1;
    splan.ham_slots = plan.ham_slots;
    splan.ham_nb = plan.ham_nb;
    splan.ham_magic = plan.ham_magic;
    splan.score_c0 = plan.score_c0;
  }
  const bool sample = sg.ok && splan.ok && splan.pair != 1;
  const size_t stride = (size_t)plan.lists * kp;
  const size_t rowb = 64ull * W * S;
  size_t need = Carver::need(rowb * (size_t)padded) + Carver::need(rowb * nq) +
                Carver::need(sizeof(float) * (size_t)padded) +
                Carver::need(sizeof(float) * nq * stride) + Carver::need(sizeof(int) * nq * stride) +
                2 * Carver::need(sizeof(float) * (size_t)nq * plan.lists) +
                3 * Carver::need(sizeof(int) * (size_t)nq) + Carver::need(64) +
                Carver::need(plan.ws_bytes) + 1024;
  if (sample)
    need += Carver::need(sizeof(float) * (size_t)nq * sg.gcols) + Carver::need(splan.ws_bytes);
  int n_flag = 0;
  {
    ScratchScope ws(need, st);
    Carver c(ws.p);
    void *base8 = c.take<char>(rowb * (size_t)padded);
    void *query8 = c.take<char>(rowb * nq);
    float *an = c.take<float>(padded);
    float *cscore = c.take<float>(nq * stride);
    int *cid = c.take<int>(nq * stride);
    float *cthr = c.take<float>((size_t)nq * plan.lists);
    int *ccnt = c.take<int>((size_t)nq * plan.lists);
    float *thr_init = c.take<float>(nq);
    int *flags = c.take<int>(nq);
    int *flag_list = c.take<int>(nq);
    int *flag_count = c.take<int>(16);
    void *tfws = c.take<char>(plan.ws_bytes);
    int rc;
    {
      ProfScope ps(12, st);
      if ((rc = expand_codes(pb, nb, W, S, S > 1, nc, base8, st))) return rc;
      if ((rc = expand_codes(pq, nq, W, S, 0, nq, query8, st))) return rc;
      if (S == 1) {  // one row per accumulator: constant |b|^2; the tile padding is E4M3 NaN (0x7F)
        if (padded > nc)
          YB_CUDA(cudaMemsetAsync((char *)base8 + rowb * (size_t)nc, 0x7F, rowb * (size_t)(padded - nc), st));
      }
      (void)bits;
      YB_CUDA(cudaMemsetAsync(flag_count, 0, 64, st));
    }
    const float *thr0 = nullptr;
    if (sample) {
      ProfScope ps(13, st);
      float *gm = c.take<float>((size_t)nq * sg.gcols);
      void *stfws = c.take<char>(splan.ws_bytes);
      if ((rc = tf32_group_min(splan, nq, nc, dfl, sg.nbt_s, sg.stride, (const float *)base8,
                               (const float *)query8, an, gm, sg.gcols, sg.gsize, stfws, st)))
        return rc;
      if ((rc = row_kth(gm, sg.gcols, nq, (int)sg.gcols, sg.j2, thr_init, st))) return rc;
      // scores are multiples of 4: admit the ties AT the order statistic as well
      k_thr_bump<<<(nq + 255) / 256, 256, 0, st>>>(thr_init, nq, 2.0f);
      YB_LAUNCH_CHECK();
      thr0 = thr_init;
    }
    {
      ProfScope ps(14, st);
      Tf32Out mo = {ccnt, 0, 0, 0};
      if ((rc = tf32_shortlist(plan, nq, nc, dfl, nbt, 1, (const float *)base8, (const float *)query8,
                               an, thr0, cscore, cid, cthr, tfws, st, &mo)))
        return rc;
    }
    {
      ProfScope ps(15, st);
      static bool attr[64] = {};
      once_per_device(attr, [] {
        cudaFuncSetAttribute(k_ham_tc_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, HF_CAP * 8);
      });
      k_ham_tc_finish<<<nq, HF_T, HF_CAP * 8, st>>>(nq, nb, k, plan.lists, kp, ccnt, cscore, cid, cthr,
                                                    id_offset, assign, dis, flags);
      YB_LAUNCH_CHECK();
      k_ham_collect<<<(nq + 255) / 256, 256, 0, st>>>(flags, nq, flag_list, flag_count);
      YB_LAUNCH_CHECK();
    }
    YB_CUDA(cudaMemcpyAsync(&n_flag, flag_count, sizeof(int), cudaMemcpyDeviceToHost, st));
    YB_CUDA(cudaStreamSynchronize(st));
    if (n_flag > 0) {
      int *keep = (int *)yb_malloc(sizeof(int) * (size_t)n_flag);
      YB_CUDA(cudaMemcpyAsync(keep, flag_list, sizeof(int) * (size_t)n_flag, cudaMemcpyDeviceToDevice, st));
      *flag_list_out = keep;
    }
  }
  *n_flag_out = n_flag;
  return 0;
}

// raw scores of the E4M3 pass, s[q][n] = 4 * ham(q, n), for tests and bring-up
int hamming_tc_scores(int nq, int nb, int W, const unsigned long long *pb,
                      const unsigned long long *pq, float *scores, cudaStream_t st) {
  const int dfl = 16 * W;
  Tf32Plan plan = tf32_plan(nq, nb, dfl, 1);
  if (!plan.ok || plan.pair) return fail(3, "hamming tensor path does not support this shape");
  plan.kind = 1;
  const long padded = tf32_padded_rows(nb);
  ScratchScope ws(Carver::need(64ull * W * nb) + Carver::need(64ull * W * nq) +
                      Carver::need(4ull * padded) + Carver::need(plan.ws_bytes),
                  st);
  Carver c(ws.p);
  void *base8 = c.take<char>(64ull * W * nb);
  void *query8 = c.take<char>(64ull * W * nq);
  float *an = c.take<float>(padded);
  void *tfws = c.take<char>(plan.ws_bytes);
  int rc;
  if ((rc = expand_codes(pb, nb, W, 1, 0, nb, base8, st))) return rc;
  if ((rc = expand_codes(pq, nq, W, 1, 0, nq, query8, st))) return rc;
  if ((rc = fill_f32(an, nb, 128.0f * (float)W, st))) return rc;
  if ((rc = fill_f32(an + nb, padded - nb, __builtin_inff(), st))) return rc;
  return tf32_scores(plan, nq, nb, dfl, tf32_tiles(nb), 1, (const float *)base8,
                     (const float *)query8, an, scores, nb, tfws, st);
}

// raw accumulators of the PACKED E4M3 pass (bring-up / tests): out[q][c] = -2 * acc(q, combined
// row c), c < ceil(nb / slots)
int hamming_tc_packed_dump(int nq, int nb, int W, int slots, const unsigned long long *pb,
                           const unsigned long long *pq, float *out, cudaStream_t st) {
  if (slots < 1 || slots > 3 || slots * W > 8) return fail(3, "hamming packed dump: %d slots of %d words", slots, W);
  const int dfl = 16 * W * slots;
  const int nc = (nb + slots - 1) / slots;
  Tf32Plan plan = tf32_plan(nq, nc, dfl, 1);
  if (!plan.ok || plan.pair) return fail(3, "hamming tensor path does not support this shape");
  plan.kind = 1;
  const long padded = tf32_padded_rows(nc);
  const size_t rowb = 64ull * W * slots;
  ScratchScope ws(Carver::need(rowb * nc) + Carver::need(rowb * nq) + Carver::need(4ull * padded) +
                      Carver::need(plan.ws_bytes),
                  st);
  Carver c(ws.p);
  void *base8 = c.take<char>(rowb * nc);
  void *query8 = c.take<char>(rowb * nq);
  float *an = c.take<float>(padded);
  void *tfws = c.take<char>(plan.ws_bytes);
  int rc;
  if ((rc = expand_codes(pb, nb, W, slots, 1, nc, base8, st))) return rc;
  if ((rc = expand_codes(pq, nq, W, slots, 0, nq, query8, st))) return rc;
  if ((rc = fill_f32(an, padded, 0.0f, st))) return rc;
  return tf32_scores(plan, nq, nc, dfl, tf32_tiles(nc), 1, (const float *)base8,
                     (const float *)query8, an, out, nc, tfws, st);
}

}  // namespace yb
