// yb_select.cu -- k smallest of each row, ordered by (value, index).
//
// Replaces fvec_k_min / fvecs_k_min (yael/sorting.c:191-255) and, for the exact k-NN path,
// the fbinheap streaming selector (yael/binheap.c:139-211).  The reference picks between an
// argmin scan, a max-heap and a quickselect by regime; all three return the k smallest
// values ascending and differ only in how they order equal values (heap slot / arbitrary,
// yael/sorting.c:174-181).  Here the order is DEFINED as (value, index) for every regime.
//
// One CTA per row:  (1) 3-pass radix select (11+11+10 bits of the monotone float key) finds
// the k-th smallest key;  (2) compaction keeps everything below it plus the lowest-index
// ties;  (3) the survivors are bitonic-sorted as 64-bit (key, index) words in shared memory
// (k <= 4096) or in a global scratch slab (larger k).
#include "yb_common.cuh"
#include "yb_internal.cuh"

namespace yb {

constexpr int KT = 512;           // threads per CTA
constexpr int KSMEM_SORT = 4096;  // largest padded k sorted in shared memory

struct BlockScan {
  int warp_tot[KT / 32];
  int total;
};

// exclusive prefix sum over the CTA; returns this thread's offset, *total = CTA sum
__device__ __forceinline__ int block_scan_excl(int v, BlockScan &bs, int *total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) bs.warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < KT / 32 ? bs.warp_tot[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < KT / 32) bs.warp_tot[lane] = winc - w;
    if (lane == 31) bs.total = winc;
  }
  __syncthreads();
  int r = inc - v + bs.warp_tot[warp];
  *total = bs.total;
  __syncthreads();
  return r;
}

struct SelectShared {
  int hist[2048];
  BlockScan bs;
  uint32_t prefix;
  int rem;
  int bucket_cnt;
  int cnt;
  int nvalid;
  unsigned long long best;
};

__device__ __forceinline__ float signed_val(const float *row, long i, int sign) {
  float v = __ldg(row + i);
  return sign < 0 ? -v : v;
}

// flags: bit0 = k==1 follows nn_single_full (yael/nn.c:404-440): start from (-1, 1e30f),
// strict '<'
__global__ void __launch_bounds__(KT)
k_kmin_rows(const float *__restrict__ val, long n, long ld, int k, int sign,
            int *__restrict__ idx_out, float *__restrict__ val_out, int id_offset, int flags,
            unsigned long long *__restrict__ gsort, int k_pad) {
  __shared__ SelectShared sh;
  extern __shared__ unsigned long long ssort[];  // k_pad entries when k_pad <= KSMEM_SORT
  const int tid = threadIdx.x;
  const long rowi = blockIdx.x;
  const float *row = val + rowi * ld;
  int *io = idx_out + rowi * (long)k;
  float *vo = val_out ? val_out + rowi * (long)k : nullptr;

  // ---------------------------------------------------------------- k == 1: arg-min
  if (k == 1) {
    unsigned long long best = ~0ull;
    for (long i = tid; i < n; i += KT) {
      unsigned long long c = ((unsigned long long)float_key(signed_val(row, i, sign)) << 32) |
                             (unsigned int)i;
      best = c < best ? c : best;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
      best = t < best ? t : best;
    }
    if (tid == 0) sh.best = ~0ull;
    __syncthreads();
    if ((tid & 31) == 0) atomicMin(&sh.best, best);
    __syncthreads();
    if (tid == 0) {
      best = sh.best;
      uint32_t key = (uint32_t)(best >> 32);
      int id = (int)(uint32_t)best;
      bool ok = !is_nan_key(key);
      float v = ok ? __ldg(row + id) : 0.f;
      if ((flags & 1) && ok && !(v < 1e30f)) ok = false;
      if (ok) {
        io[0] = id + id_offset;
        if (vo) vo[0] = v;
      } else {
        io[0] = -1;
        if (vo) vo[0] = (flags & 1) ? 1e30f : __uint_as_float(0xffffffffu);
      }
    }
    return;
  }

  // ---------------------------------------------------------------- radix select
  if (tid == 0) {
    sh.prefix = 0;
    sh.cnt = 0;
    sh.nvalid = 0;
  }
  // count the non-NaN elements (NaN is never selected: yael/binheap.c:144,149)
  {
    int c = 0;
    for (long i = tid; i < n; i += KT) c += !is_nan_key(float_key(signed_val(row, i, sign)));
    c = warp_sum(c);
    __syncthreads();
    if ((tid & 31) == 0 && c) atomicAdd(&sh.nvalid, c);
    __syncthreads();
  }
  const int keff = min(k, sh.nvalid);
  if (tid == 0) sh.rem = keff;
  __syncthreads();

  if (keff > 0) {
    const int shifts[3] = {21, 10, 0};
    const int widths[3] = {11, 11, 10};
    uint32_t mask_prev = 0;
    for (int p = 0; p < 3; p++) {
      for (int h = tid; h < 2048; h += KT) sh.hist[h] = 0;
      __syncthreads();
      const uint32_t prefix = sh.prefix;
      const int sft = shifts[p];
      const uint32_t dm = (1u << widths[p]) - 1u;
      for (long i = tid; i < n; i += KT) {
        uint32_t key = float_key(signed_val(row, i, sign));
        if ((key & mask_prev) == prefix) atomicAdd(&sh.hist[(key >> sft) & dm], 1);
      }
      __syncthreads();
      // locate the bucket that holds rank `rem`
      int loc[4], s = 0;
#pragma unroll
      for (int u = 0; u < 4; u++) {
        loc[u] = sh.hist[tid * 4 + u];
        s += loc[u];
      }
      int tot;
      int off = block_scan_excl(s, sh.bs, &tot);
      const int rem = sh.rem;
      __syncthreads();
#pragma unroll
      for (int u = 0; u < 4; u++) {
        if (off < rem && rem <= off + loc[u]) {
          sh.prefix = prefix | ((uint32_t)(tid * 4 + u) << sft);
          sh.rem = rem - off;
          sh.bucket_cnt = loc[u];
        }
        off += loc[u];
      }
      mask_prev |= dm << sft;
      __syncthreads();
    }
  }
  const uint32_t pivot = sh.prefix;
  const int need_eq = sh.rem;        // how many elements equal to the pivot are kept
  const int have_eq = sh.bucket_cnt;  // how many exist
  unsigned long long *buf = (k_pad <= KSMEM_SORT) ? ssort : gsort + rowi * (long)k_pad;

  // ---------------------------------------------------------------- compaction
  if (keff > 0) {
    if (have_eq == need_eq) {
      // every tie is needed: order is irrelevant here, the sort fixes it
      for (long i = tid; i < n; i += KT) {
        uint32_t key = float_key(signed_val(row, i, sign));
        if (key <= pivot) {
          int pos = atomicAdd(&sh.cnt, 1);
          buf[pos] = ((unsigned long long)key << 32) | (unsigned int)i;
        }
      }
    } else {
      // only the lowest-index ties are kept: walk the row in index order
      int base_eq = 0;
      for (long i0 = 0; i0 < n; i0 += KT) {
        long i = i0 + tid;
        uint32_t key = i < n ? float_key(signed_val(row, i, sign)) : 0xffffffffu;
        int is_eq = (i < n && key == pivot) ? 1 : 0;
        int tot;
        int rank = block_scan_excl(is_eq, sh.bs, &tot);
        bool keep = (i < n) && (key < pivot || (is_eq && base_eq + rank < need_eq));
        if (keep) {
          int pos = atomicAdd(&sh.cnt, 1);
          buf[pos] = ((unsigned long long)key << 32) | (unsigned int)i;
        }
        base_eq += tot;
      }
    }
  }
  __syncthreads();
  for (int i = keff + tid; i < k_pad; i += KT) buf[i] = ~0ull;
  bitonic_sort_u64(buf, k_pad, tid, KT, [] { __syncthreads(); });

  for (int r = tid; r < k; r += KT) {
    if (r < keff) {
      int id = (int)(uint32_t)buf[r];
      io[r] = id + id_offset;
      if (vo) vo[r] = __ldg(row + id);
    } else {  // yael/nn.c:515-518
      io[r] = -1;
      if (vo) vo[r] = __uint_as_float(0xffffffffu);
    }
  }
}

size_t kmin_ws_bytes(long nrow, int k) {
  int k_pad = pow2_ceil(k < 2 ? 2 : k);
  if (k_pad <= KSMEM_SORT) return 256;
  return Carver::need(sizeof(unsigned long long) * (size_t)nrow * k_pad);
}

int kmin_rows(const float *val, long n, long ld, long nrow, int k, int sign, int *idx,
              float *vals, int id_offset, int flags, void *ws, cudaStream_t st) {
  if (nrow <= 0 || k <= 0) return 0;
  if (k > n) return fail(3, "k_min: k=%d exceeds n=%ld", k, n);
  int k_pad = pow2_ceil(k < 2 ? 2 : k);
  size_t smem = (k_pad <= KSMEM_SORT && k > 1) ? sizeof(unsigned long long) * (size_t)k_pad : 0;
  static bool attr_done[64] = {};
  once_per_device(attr_done, [] {
    cudaFuncSetAttribute(k_kmin_rows, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(sizeof(unsigned long long) * KSMEM_SORT));
  });
  k_kmin_rows<<<(unsigned)nrow, KT, smem, st>>>(val, n, ld, k, sign, idx, vals, id_offset, flags,
                                                (unsigned long long *)ws, k_pad);
  YB_LAUNCH_CHECK();
  return 0;
}

}  // namespace yb

using namespace yb;

extern "C" int yb_k_min_rows(const float *val, long n, long ld, long nrow, int k, int sign,
                              int *idx, float *vals, yb_stream_t s) {
  if (nrow <= 0 || k <= 0 || n <= 0) return 0;
  Guard g;
  cudaStream_t st = stream_of(s);
  ScratchScope ws(kmin_ws_bytes(nrow, k), st);
  return kmin_rows(val, n, ld, nrow, k, sign, idx, vals, 0, 0, ws.p, st);
}
