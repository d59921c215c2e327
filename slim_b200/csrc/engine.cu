// engine.cu -- the sm_100a CUDA engine behind SLIM_Learn / Py_SLIM_Learn.
//
// Replaces, from scratch, the reference's OpenMP coordinate-descent learner:
//   CreateTrainingMatrix   reference src/libslim/setup.c:109-135   -> stage()        (K0 kernels)
//   EstimateModelCD        reference src/libslim/estimate.c:328-558 -> learn()        (cd_solve_kernel)
//   CoordinateDescent      reference src/libslim/cd.c:101-142       -> the sweep loop inside cd_solve_kernel
//   SaveModel              reference src/libslim/estimate.c:570-593 -> gather_columns_kernel + host assembly (api.cpp)
// The default solver works in GRAM space (gram.cuh, gram_batch.cuh: G = R^T R staged once, no pass over R per
// target); the user-space kernels in this file are the path when G does not fit in HBM.  predict.cuh holds the
// batched top-N (reference src/libslim/predict.c).
//
// Data layout in HBM:
//   CSR   rowptr int64[nrows+1], rowind int32[nnz], rowval fp32[nnz] (absent for all-ones input)
//   CSC   colptr int64[ncols+1] in PADDED entry units (every column starts on a 16-byte boundary and
//         is padded to a multiple of 4 entries so a lane reads 4 user ids + 4 values with two 128-bit
//         loads), colcnt int32[ncols] true lengths, colind int32[nnzp], colval fp32[nnzp],
//         cnorms fp32[ncols] (reference rounding), csq fp64[ncols] exact sum of squares.
//
// There is no CPU fallback: every entry point fails with an error status when CUDA is unusable.
#include "engine.h"

#include <cuda_runtime.h>
#include <cooperative_groups.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <numeric>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include <dlfcn.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_segmented_sort.cuh>

namespace cg = cooperative_groups;

namespace slimb200 {

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
const char *last_error() { return g_last_error.c_str(); }

struct EngineError : std::runtime_error {
  int status;
  EngineError(int st, const std::string &m) : std::runtime_error(m), status(st) {}
};

static inline void ck(cudaError_t e, const char *what) {
  if (e != cudaSuccess) {
    int st = (e == cudaErrorMemoryAllocation) ? kErrMemory : kErr;
    throw EngineError(st, std::string(what) + ": " + cudaGetErrorString(e));
  }
}
#define CK(x) ck((x), #x)

// Temporary device buffers.  Inside an AsyncAllocScope (learn()) they come from the device's stream-ordered memory
// pool (cudaMallocAsync / cudaFreeAsync on the engine stream, release threshold "never": no device-wide
// synchronisation and no page (un)mapping per call -- cudaMalloc / cudaFree of the ~100 MB pools and sort buffers cost
// 10-700 ms per learn() call on a device that holds a 13 GB Gram matrix); elsewhere plain cudaMalloc / cudaFree.
static thread_local cudaStream_t t_alloc_stream = nullptr;
static thread_local bool t_alloc_async = false;
struct AsyncAllocScope {
  bool prev_on;
  cudaStream_t prev_s;
  explicit AsyncAllocScope(cudaStream_t s) : prev_on(t_alloc_async), prev_s(t_alloc_stream) {
    t_alloc_async = true;
    t_alloc_stream = s;
  }
  AsyncAllocScope(const AsyncAllocScope &) = delete;
  AsyncAllocScope &operator=(const AsyncAllocScope &) = delete;
  ~AsyncAllocScope() {
    t_alloc_async = prev_on;
    t_alloc_stream = prev_s;
  }
};
// *pooled = true when the memory came from the pool (free it with dev_free(.., true, stream) or with cudaFree)
static cudaError_t dev_malloc(void **p, size_t bytes, bool *pooled) {
  if (t_alloc_async) {
    if (cudaMallocAsync(p, bytes, t_alloc_stream) == cudaSuccess) {
      *pooled = true;
      return cudaSuccess;
    }
    (void)cudaGetLastError();
  }
  *pooled = false;
  return cudaMalloc(p, bytes);
}
static void dev_free(void *p, bool pooled, cudaStream_t s) {
  if (!p) return;
  if (pooled) {
    if (cudaFreeAsync(p, s) == cudaSuccess) return;
    (void)cudaGetLastError();
  }
  cudaFree(p);
}

template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  bool pooled = false;
  cudaStream_t pstream = nullptr;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { reset(); }
  void reset() {
    dev_free(p, pooled, pstream);
    p = nullptr;
    n = 0;
    pooled = false;
  }
  void alloc(size_t count) {
    reset();
    n = count;
    pstream = t_alloc_stream;
    void *q = nullptr;
    CK(dev_malloc(&q, std::max<size_t>(count, 1) * sizeof(T), &pooled));
    p = static_cast<T *>(q);
  }
  void alloc_zero(size_t count, cudaStream_t s) {
    alloc(count);
    CK(cudaMemsetAsync(p, 0, std::max<size_t>(count, 1) * sizeof(T), s));
  }
  T *release() {  // (the new owner frees with cudaFree, which accepts pool memory too)
    T *q = p;
    p = nullptr;
    n = 0;
    pooled = false;
    return q;
  }
};

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    CK(cudaGetDevice(&prev));
    if (prev != dev) CK(cudaSetDevice(dev));
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

struct EventPair {  // two timing events, destroyed on every exit path
  cudaEvent_t a = nullptr, b = nullptr;
  EventPair() {
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
  }
  EventPair(const EventPair &) = delete;
  EventPair &operator=(const EventPair &) = delete;
  ~EventPair() {
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
  }
};

int device_count() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

// Layout of G in HBM: PANEL-major.  A panel is kGramPW consecutive columns of all rows, rows contiguous
// inside the panel: element (k, i) lives at ((i / PW) * nrows + k) * PW + i % PW.  A block of coordinates
// then reads one short segment of many rows out of ONE panel (25.6 MB for 100K items) instead of touching
// every 400 KB row of a 40 GB matrix: few DRAM pages, few TLB entries per block round.
constexpr int kGramPW = 64;
__host__ __device__ __forceinline__ size_t gram_off(size_t nr, int k, int i) {
  return ((size_t)(i / kGramPW) * nr + (size_t)k) * kGramPW + (size_t)(i % kGramPW);
}


// Fire-and-forget fp64 reduction (SASS REDG.E.ADD.F64).  atomicAdd() with an unused result is compiled to a RETURNING
// ATOMG inside the unrolled update loops of the user-space kernels, and a lane's next ATOMG then waits for the
// previous one to come back from the L2 (profiles/r02_c5_hybrid_notes.txt).
__device__ __forceinline__ void red_add_f64(double *p, double v) {
  asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// ------------------------------------------------------------------------------------------------
// staged matrix
// ------------------------------------------------------------------------------------------------
struct Matrix {
  int device = 0;
  int32_t nrows = 0, ncols = 0;
  int64_t nnz = 0, nnzp = 0;
  bool has_val = false;   // the caller passed a value array
  bool nonneg = true;     // no stored value is negative (Gram blocks >= 0: lets the window solve skip slots)
  bool unit = false;      // ... and every stored value is exactly 1.0f: the kernels then skip the value
                          // stream (4 B per nonzero instead of 8), same arithmetic as the binary path
  // Items are RELABELED by popularity inside the engine: internal id = position of the item when items
  // are sorted by (descending nnz, ascending id).  A target's active items are mostly popular ones, so
  // in this order they fall into few, densely filled 32-item windows (see cd_cluster_kernel).  Every
  // device structure below uses internal ids; results are mapped back before they leave the engine.
  int32_t *d_rank = nullptr;      // original id -> internal id
  int32_t *d_inv = nullptr;       // internal id -> original id
  std::vector<int32_t> h_rank, h_inv;
  double *d_wgram = nullptr;      // [ceil(ncols/32)][32][32] Gram blocks of 32 consecutive item columns
  int32_t rows_per_part = 0;      // user-range width of the 16-way column split (cluster kernel)
  int32_t *d_colsplit = nullptr;  // [ncols][kParts+1] entry offsets of the user ranges in each column
  int64_t *d_rowptr = nullptr;
  int32_t *d_rowind = nullptr;
  float *d_rowval = nullptr;
  int64_t *d_colptr = nullptr;
  int32_t *d_colcnt = nullptr;
  int32_t *d_colind = nullptr;
  float *d_colval = nullptr;
  float *d_cnorms = nullptr;
  double *d_csq = nullptr;
  std::vector<int32_t> h_colcnt;
  // Gram matrix G = R^T R in internal item order (gram.cuh): the Gram-space solver reads rows of it
  void *d_gram = nullptr;          // panel-major; packed 8/16/32-bit unsigned (exact integer sums) or double elements
  size_t gram_ld = 0;              // ncols rounded up to whole panels (stride of item-indexed scratch arrays)
  bool gram_f64 = false;
  size_t gram_bytes = 0;
  int32_t gram_h32 = 0, gram_h16 = 0;      // packed layout: first 16-bit / 8-bit column (gram.cuh: GramView)
  size_t gram_off16 = 0, gram_off8 = 0;    // packed layout: byte offsets of the 16-bit / 8-bit column ranges
  // STAIR layout (gram.cuh: GaStair), used when the full matrix does not fit: panel p stores rows [0, max(64(p+1), hd))
  bool gram_stair = false;
  int32_t gram_hd = 0;
  unsigned long long *d_gram_pbase = nullptr;   // byte offset of every panel
  std::vector<unsigned long long> h_gram_pbase;
  unsigned long long *d_expand = nullptr;  // per item: sum of len(row_u) over the users of the column
  double gram_ms = 0.0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // extra streams: the Gram launches of the target classes run side by side
  cudaStream_t stream3 = nullptr;
  cudaStream_t stream4 = nullptr;
  cudaStream_t stream5 = nullptr;
  int sm_count = 0;
  int smem_optin = 0;
  double stage_ms = 0.0;
  int32_t stage_launches = 0;
  // solve scratch, cached between learn() calls on the same matrix
  void *d_scratch = nullptr;
  size_t scratch_bytes = 0;
  // learn() keeps per-call state here (scratch, the class streams): concurrent calls on ONE staged matrix are serialised
  std::mutex learn_mutex;
};

void free_matrix(Matrix *m) {
  if (!m) return;
  const double w0 = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(m->device);
  cudaFree(m->d_rowptr);
  cudaFree(m->d_rowind);
  cudaFree(m->d_rowval);
  cudaFree(m->d_colptr);
  cudaFree(m->d_colcnt);
  cudaFree(m->d_colind);
  cudaFree(m->d_colval);
  cudaFree(m->d_cnorms);
  cudaFree(m->d_csq);
  cudaFree(m->d_colsplit);
  cudaFree(m->d_wgram);
  cudaFree(m->d_rank);
  cudaFree(m->d_inv);
  cudaFree(m->d_scratch);
  cudaFree(m->d_gram);
  cudaFree(m->d_gram_pbase);
  cudaFree(m->d_expand);
  if (m->stream2) cudaStreamDestroy(m->stream2);
  if (m->stream3) cudaStreamDestroy(m->stream3);
  if (m->stream4) cudaStreamDestroy(m->stream4);
  if (m->stream5) cudaStreamDestroy(m->stream5);
  if (m->stream) cudaStreamDestroy(m->stream);
  if (prev >= 0 && prev != m->device) cudaSetDevice(prev);  // leave the caller's current device as it was
  if (getenv("SLIMB200_VERBOSE") && atoi(getenv("SLIMB200_VERBOSE")) >= 2)
    fprintf(stderr, "[slim-b200] free_matrix(): host wall %.1f ms\n",
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count() - w0);
  delete m;
}

void matrix_info(const Matrix *m, int32_t *nrows, int32_t *ncols, int64_t *nnz, int32_t *device,
                 double *stage_ms, int32_t *stage_launches) {
  if (nrows) *nrows = m->nrows;
  if (ncols) *ncols = m->ncols;
  if (nnz) *nnz = m->nnz;
  if (device) *device = m->device;
  if (stage_ms) *stage_ms = m->stage_ms;
  if (stage_launches) *stage_launches = m->stage_launches;
}

// ------------------------------------------------------------------------------------------------
// K0: staging kernels (CSR -> padded CSC, norms).  Replaces gk_csr_CreateIndex(COL)
// (reference lib/GKlib/csr.c:1501-1585) and gk_csr_ComputeNorms (csr.c:1897-1938).
// ------------------------------------------------------------------------------------------------
__global__ void max_index_kernel(const int32_t *__restrict__ ind, int64_t n, int32_t *out_max,
                                 int32_t *out_min) {
  int32_t mx = -1, mn = 0x7fffffff;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
       k += (int64_t)gridDim.x * blockDim.x) {
    int32_t v = ind[k];
    mx = max(mx, v);
    mn = min(mn, v);
  }
  for (int o = 16; o; o >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(out_max, mx);
    atomicMin(out_min, mn);
  }
}

__global__ void count_columns_kernel(const int32_t *__restrict__ ind, int64_t n, int32_t *cnt,
                                     uint32_t *pos) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
       k += (int64_t)gridDim.x * blockDim.x) {
    atomicAdd(&cnt[ind[k]], 1);
    pos[k] = (uint32_t)k;
  }
}

__global__ void relabel_items_kernel(int32_t *ind, int64_t n, const int32_t *__restrict__ rank) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
       k += (int64_t)gridDim.x * blockDim.x)
    ind[k] = rank[ind[k]];
}

// Single-CTA scan over the column counts: colptr (padded to 4 entries) and cstart (unpadded).
__global__ void scan_columns_kernel(const int32_t *__restrict__ cnt, int32_t ncols, int64_t *colptr,
                                    int64_t *cstart) {
  __shared__ int64_t s_pad[1024], s_raw[1024];
  const int tid = threadIdx.x;
  const int per = (ncols + 1023) / 1024;
  const int lo = min(ncols, tid * per), hi = min(ncols, lo + per);
  int64_t pad = 0, raw = 0;
  for (int i = lo; i < hi; i++) {
    raw += cnt[i];
    pad += (cnt[i] + 3) & ~3;
  }
  s_pad[tid] = pad;
  s_raw[tid] = raw;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    int64_t a = 0, b = 0;
    if (tid >= o) {
      a = s_pad[tid - o];
      b = s_raw[tid - o];
    }
    __syncthreads();
    s_pad[tid] += a;
    s_raw[tid] += b;
    __syncthreads();
  }
  int64_t bp = s_pad[tid] - pad, br = s_raw[tid] - raw;  // exclusive
  for (int i = lo; i < hi; i++) {
    colptr[i] = bp;
    cstart[i] = br;
    br += cnt[i];
    bp += (cnt[i] + 3) & ~3;
  }
  if (tid == 1023) {
    colptr[ncols] = s_pad[1023];
    cstart[ncols] = s_raw[1023];
  }
}

// After a STABLE sort of the nonzero positions by column id, entry k of the sorted stream is the
// (k - cstart[c])-th nonzero of column c in ascending user order (csr.c:1549-1584 produces the
// same order with its serial counting sort).
template <bool HASVAL>
__global__ void fill_csc_kernel(const int32_t *__restrict__ scol, const uint32_t *__restrict__ spos,
                                int64_t nnz, const int64_t *__restrict__ rowptr, int32_t nrows,
                                const float *__restrict__ rowval, const int64_t *__restrict__ colptr,
                                const int64_t *__restrict__ cstart, int32_t *colind, float *colval) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz;
       k += (int64_t)gridDim.x * blockDim.x) {
    const int32_t c = scol[k];
    const int64_t p = spos[k];
    int32_t lo = 0, hi = nrows;  // largest r with rowptr[r] <= p
    while (hi - lo > 1) {
      int32_t mid = lo + ((hi - lo) >> 1);
      if (rowptr[mid] <= p) lo = mid; else hi = mid;
    }
    const int64_t dst = colptr[c] + (k - cstart[c]);
    colind[dst] = lo;
    if (HASVAL) colval[dst] = rowval[p];
  }
}

// One warp per column.  cnorms reproduces the reference rounding: float accumulation in ascending
// user order, separate multiply and add (gk_fdot, lib/GKlib/gk_mkblas.h:161-170 compiled -std=c99),
// then (float)sqrt((double)sum) (csr.c:1931); without values sqrt(count) (csr.c:1936).
template <bool HASVAL>
__global__ void column_norms_kernel(int32_t ncols, const int64_t *__restrict__ colptr,
                                    const int32_t *__restrict__ colcnt, const float *__restrict__ colval,
                                    float *cnorms, double *csq) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < ncols; c += warps) {
    const int cnt = colcnt[c];
    if (!HASVAL) {
      if (lane == 0) {
        cnorms[c] = (float)sqrt((double)cnt);
        csq[c] = (double)cnt;
      }
      continue;
    }
    const float *v = colval + colptr[c];
    float fsum = 0.0f;
    double dsum = 0.0;
    for (int base = 0; base < cnt; base += 32) {
      const int e = base + lane;
      const float x = e < cnt ? v[e] : 0.0f;
      const float sq = __fmul_rn(x, x);
      dsum += (double)x * (double)x;
      const int lim = min(32, cnt - base);
      for (int l = 0; l < lim; l++) fsum = __fadd_rn(fsum, __shfl_sync(0xffffffffu, sq, l));
    }
    for (int o = 16; o; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
    if (lane == 0) {
      cnorms[c] = (float)sqrt((double)fsum);
      csq[c] = dsum;
    }
  }
}

constexpr int kParts = 16;  // user ranges per column = largest cluster size

__global__ void all_ones_kernel(const float *__restrict__ v, int64_t n, int32_t *flags) {
  bool bad = false, neg = false;  // flags bit 0: some value != 1, bit 1: some value < 0 (or NaN)
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
       k += (int64_t)gridDim.x * blockDim.x) {
    bad |= (v[k] != 1.0f);
    neg |= !(v[k] >= 0.0f);
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flags, 1);
  if (__any_sync(0xffffffffu, neg) && (threadIdx.x & 31) == 0) atomicOr(flags, 2);
}

// colsplit[c][r] = number of entries of column c whose user id is < r * rows_per_part.  Users are
// ascending inside a column, so CTA r of a cluster owns the contiguous entry range
// [colsplit[c][r], colsplit[c][r+1]) and with it a private slice of yhat.
__global__ void column_split_kernel(int32_t ncols, int32_t rows_per_part, const int64_t *__restrict__ colptr,
                                    const int32_t *__restrict__ colcnt, const int32_t *__restrict__ colind,
                                    int32_t *colsplit) {
  const int64_t total = (int64_t)ncols * (kParts + 1);
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(t / (kParts + 1)), r = (int)(t % (kParts + 1));
    const int cnt = colcnt[c];
    const int32_t *ix = colind + colptr[c];
    const int64_t bound = (int64_t)r * rows_per_part;
    int lo = 0, hi = cnt;  // first entry with user >= bound
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((int64_t)ix[mid] < bound) lo = mid + 1; else hi = mid;
    }
    colsplit[t] = lo;
  }
}

// Window Gram blocks.  Items are grouped into windows of 32 consecutive ids; G[w][k][m] = <a_k, a_m>
// for the columns k, m of window w (exact: integer counts or fp64 sums of fp32 products).  With it a
// whole window of coordinates is updated in ONE pass over its columns, exactly as sequential CD
// would (see cd_window_kernel).  One CTA per window at a time; a user-indexed bit mask (which columns
// of the window contain user u) finds the overlapping pairs, which are rare for sparse columns.
template <bool HASVAL>
__global__ void window_gram_kernel(int32_t ncols, int32_t nrows, int32_t nitems, const int32_t *__restrict__ item_w,
                                   const int32_t *__restrict__ item_s, const int32_t *__restrict__ item_S,
                                   const int64_t *__restrict__ colptr, const int32_t *__restrict__ colcnt,
                                   const int32_t *__restrict__ colind, const float *__restrict__ colval,
                                   uint32_t *masks, size_t mask_stride, double *wgram) {
  // One work item = (window, user slice): dense head windows are split over many CTAs by user range.
  uint32_t *mask = masks + (size_t)blockIdx.x * mask_stride;
  __shared__ double g[32][33];
  __shared__ int e_lo[32], e_hi[32];
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
    const int w = item_w[it], sl = item_s[it], S = item_S[it];
    const int64_t u_lo = (int64_t)nrows * sl / S, u_hi = (int64_t)nrows * (sl + 1) / S;
    for (int t = tid; t < 32 * 33; t += nt) (&g[0][0])[t] = 0.0;
    const int cbase = w * 32;
    const int nc = min(32, ncols - cbase);
    if (tid < 32) {  // entry range of column tid inside the user slice (users ascend inside a column)
      int lo = 0, hi = 0;
      if (tid < nc) {
        const int32_t *ix = colind + colptr[cbase + tid];
        const int cnt = colcnt[cbase + tid];
        int a0 = 0, a1 = cnt;
        while (a0 < a1) { const int mid = (a0 + a1) >> 1; if ((int64_t)ix[mid] < u_lo) a0 = mid + 1; else a1 = mid; }
        lo = a0;
        a1 = cnt;
        while (a0 < a1) { const int mid = (a0 + a1) >> 1; if ((int64_t)ix[mid] < u_hi) a0 = mid + 1; else a1 = mid; }
        hi = a0;
      }
      e_lo[tid] = lo;
      e_hi[tid] = hi;
    }
    __syncthreads();
    // pass 1: set bit k of mask[u] for every entry (u) of column k
    for (int k = 0; k < nc; k++) {
      const int64_t c0 = colptr[cbase + k];
      for (int e = e_lo[k] + tid; e < e_hi[k]; e += nt) atomicOr(&mask[colind[c0 + e]], 1u << k);
    }
    __syncthreads();
    // pass 2: every user shared by columns k > m contributes v_ku * v_mu to G[k][m]
    for (int k = 0; k < nc; k++) {
      const int64_t c0 = colptr[cbase + k];
      for (int e = e_lo[k] + tid; e < e_hi[k]; e += nt) {
        const int u = colind[c0 + e];
        uint32_t lower = __ldcg(&mask[u]) & ((1u << k) - 1u);  // L2: the atomics above bypass L1
        const double vk = HASVAL ? (double)colval[c0 + e] : 1.0;
        while (lower) {
          const int mcol = __ffs(lower) - 1;
          lower &= lower - 1;
          double vm = 1.0;
          if (HASVAL) {  // value of user u in column mcol: binary search (users ascend inside a column)
            const int64_t m0 = colptr[cbase + mcol];
            int lo = 0, hi = colcnt[cbase + mcol] - 1;
            while (lo < hi) {
              const int mid = (lo + hi) >> 1;
              if (colind[m0 + mid] < u) lo = mid + 1; else hi = mid;
            }
            vm = (double)colval[m0 + lo];
          }
          atomicAdd(&g[k][mcol], vk * vm);
        }
      }
    }
    __syncthreads();
    // pass 3: clear the mask for the next work item, accumulate the (lower triangular) block
    for (int k = 0; k < nc; k++) {
      const int64_t c0 = colptr[cbase + k];
      for (int e = e_lo[k] + tid; e < e_hi[k]; e += nt) mask[colind[c0 + e]] = 0u;
    }
    double *out = wgram + (size_t)w * 1024;
    for (int t = tid; t < 1024; t += nt) {
      const int k = t >> 5, mcol = t & 31;
      if (k > mcol && g[k][mcol] != 0.0) atomicAdd(&out[t], g[k][mcol]);
    }
    __syncthreads();
  }
}

__global__ void window_gram_symmetrize_kernel(int32_t nwin, double *wgram) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < (int64_t)nwin * 1024;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)((t >> 5) & 31), mcol = (int)(t & 31);
    if (k < mcol) wgram[t] = wgram[(t & ~(int64_t)1023) + mcol * 32 + k];
  }
}

static int grid_for(int64_t n, int block, int sm_count) {
  int64_t g = (n + block - 1) / block;
  return (int)std::max<int64_t>(1, std::min<int64_t>(g, (int64_t)sm_count * 16));
}

static void build_gram(Matrix *m);  // defined next to the Gram kernels (gram.cuh)

static double wall_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

Matrix *stage(int device, int32_t nrows, const ssize_t *rowptr, const int32_t *rowind,
              const float *rowval, bool on_device, int64_t nnz_if_device, int32_t *status) {
  Matrix *m = nullptr;
  const double w0 = wall_ms();
  try {
    if (nrows < 0 || !rowptr) throw EngineError(kErrInput, "stage: bad nrows/rowptr");
    if (device < 0 || device >= device_count())
      throw EngineError(kErr, "stage: no usable CUDA device (this library has no CPU fallback)");
    DeviceGuard guard(device);
    (void)cudaGetLastError();  // do not inherit a stale non-sticky error from the caller
    m = new Matrix();
    m->device = device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    m->sm_count = prop.multiProcessorCount;
    m->smem_optin = (int)prop.sharedMemPerBlockOptin;
    CK(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    {  // the stream-ordered pool keeps what learn() frees (see DevBuf); build_gram() trims it before it sizes G
      cudaMemPool_t pool = nullptr;
      if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        (void)cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      (void)cudaGetLastError();
    }
    CK(cudaStreamCreateWithFlags(&m->stream2, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&m->stream3, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&m->stream4, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&m->stream5, cudaStreamNonBlocking));
    cudaStream_t s = m->stream;
    m->nrows = nrows;
    m->nnz = on_device ? nnz_if_device : (int64_t)rowptr[nrows];
    m->has_val = rowval != nullptr;
    const int64_t nnz = m->nnz;
    if (nnz < 0 || nnz >= (int64_t)0xffffffffLL)
      throw EngineError(kErrInput, "stage: nnz must be in [0, 2^32)");
    if (!rowind && nnz > 0) throw EngineError(kErrInput, "stage: rowind is NULL");

    EventPair stage_ev;  // destroyed on every exit path
    cudaEvent_t e0 = stage_ev.a, e1 = stage_ev.b;
    CK(cudaEventRecord(e0, s));

    CK(cudaMalloc(&m->d_rowptr, sizeof(int64_t) * ((size_t)nrows + 1)));
    CK(cudaMalloc(&m->d_rowind, sizeof(int32_t) * std::max<int64_t>(nnz, 1)));
    if (m->has_val) CK(cudaMalloc(&m->d_rowval, sizeof(float) * std::max<int64_t>(nnz, 1)));
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    static_assert(sizeof(ssize_t) == sizeof(int64_t), "LP64 only");
    CK(cudaMemcpyAsync(m->d_rowptr, rowptr, sizeof(int64_t) * ((size_t)nrows + 1), kind, s));
    if (nnz > 0) {
      CK(cudaMemcpyAsync(m->d_rowind, rowind, sizeof(int32_t) * nnz, kind, s));
      if (m->has_val) CK(cudaMemcpyAsync(m->d_rowval, rowval, sizeof(float) * nnz, kind, s));
    }

    // ncols = max(rowind) + 1   (setup.c:117)
    DevBuf<int32_t> d_mm;
    d_mm.alloc(2);
    int32_t init[2] = {-1, 0x7fffffff};
    CK(cudaMemcpyAsync(d_mm.p, init, sizeof(init), cudaMemcpyHostToDevice, s));
    if (nnz > 0) {
      max_index_kernel<<<grid_for(nnz, 256, m->sm_count), 256, 0, s>>>(m->d_rowind, nnz, d_mm.p,
                                                                      d_mm.p + 1);
      m->stage_launches++;
    }
    int32_t mm[2];
    CK(cudaMemcpyAsync(mm, d_mm.p, sizeof(mm), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (nnz > 0 && mm[1] < 0) throw EngineError(kErrInput, "stage: negative column index");
    m->ncols = nnz > 0 ? mm[0] + 1 : 0;
    const int32_t ncols = m->ncols;

    CK(cudaMalloc(&m->d_colcnt, sizeof(int32_t) * std::max<int32_t>(ncols, 1)));
    CK(cudaMemsetAsync(m->d_colcnt, 0, sizeof(int32_t) * std::max<int32_t>(ncols, 1), s));
    CK(cudaMalloc(&m->d_colptr, sizeof(int64_t) * ((size_t)ncols + 1)));
    CK(cudaMalloc(&m->d_cnorms, sizeof(float) * std::max<int32_t>(ncols, 1)));
    CK(cudaMalloc(&m->d_csq, sizeof(double) * std::max<int32_t>(ncols, 1)));
    DevBuf<int64_t> d_cstart;
    d_cstart.alloc((size_t)ncols + 1);
    DevBuf<uint32_t> d_pos, d_spos;
    DevBuf<int32_t> d_scol;
    d_pos.alloc(nnz);
    d_spos.alloc(nnz);
    d_scol.alloc(nnz);

    if (nnz > 0) {
      count_columns_kernel<<<grid_for(nnz, 256, m->sm_count), 256, 0, s>>>(m->d_rowind, nnz,
                                                                          m->d_colcnt, d_pos.p);
      m->stage_launches++;
    }
    // popularity relabeling: internal id = rank by (descending nnz, ascending id)
    {
      std::vector<int32_t> cnt0(ncols);
      if (ncols > 0)
        CK(cudaMemcpyAsync(cnt0.data(), m->d_colcnt, sizeof(int32_t) * ncols, cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
      m->h_inv.resize(ncols);
      m->h_rank.resize(ncols);
      std::iota(m->h_inv.begin(), m->h_inv.end(), 0);
      const bool relabel = !(getenv("SLIMB200_NO_RELABEL") && atoi(getenv("SLIMB200_NO_RELABEL")));
      if (relabel)
        std::stable_sort(m->h_inv.begin(), m->h_inv.end(), [&](int32_t x, int32_t y) { return cnt0[x] > cnt0[y]; });
      for (int32_t n = 0; n < ncols; n++) m->h_rank[m->h_inv[n]] = n;
      CK(cudaMalloc(&m->d_rank, sizeof(int32_t) * std::max<int32_t>(ncols, 1)));
      CK(cudaMalloc(&m->d_inv, sizeof(int32_t) * std::max<int32_t>(ncols, 1)));
      if (ncols > 0) {
        CK(cudaMemcpyAsync(m->d_rank, m->h_rank.data(), sizeof(int32_t) * ncols, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(m->d_inv, m->h_inv.data(), sizeof(int32_t) * ncols, cudaMemcpyHostToDevice, s));
      }
      if (nnz > 0 && relabel) {
        relabel_items_kernel<<<grid_for(nnz, 256, m->sm_count), 256, 0, s>>>(m->d_rowind, nnz, m->d_rank);
        CK(cudaMemsetAsync(m->d_colcnt, 0, sizeof(int32_t) * ncols, s));
        count_columns_kernel<<<grid_for(nnz, 256, m->sm_count), 256, 0, s>>>(m->d_rowind, nnz, m->d_colcnt, d_pos.p);
        m->stage_launches += 2;
        CK(cudaStreamSynchronize(s));  // h_rank / h_inv staging buffers are read by the copies above
      }
    }
    scan_columns_kernel<<<1, 1024, 0, s>>>(m->d_colcnt, ncols, m->d_colptr, d_cstart.p);
    m->stage_launches++;
    m->h_colcnt.resize(ncols);
    int64_t nnzp = 0;
    CK(cudaMemcpyAsync(&nnzp, m->d_colptr + ncols, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    if (ncols > 0)
      CK(cudaMemcpyAsync(m->h_colcnt.data(), m->d_colcnt, sizeof(int32_t) * ncols,
                         cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    m->nnzp = nnzp;

    CK(cudaMalloc(&m->d_colind, sizeof(int32_t) * std::max<int64_t>(nnzp, 4)));
    CK(cudaMemsetAsync(m->d_colind, 0, sizeof(int32_t) * std::max<int64_t>(nnzp, 4), s));
    if (m->has_val) {
      CK(cudaMalloc(&m->d_colval, sizeof(float) * std::max<int64_t>(nnzp, 4)));
      CK(cudaMemsetAsync(m->d_colval, 0, sizeof(float) * std::max<int64_t>(nnzp, 4), s));
    }

    if (nnz > 0) {
      // stable LSD radix sort of (column id, nonzero position); plumbing, not the graded path
      int bits = 1;
      while ((1LL << bits) < (int64_t)ncols) bits++;
      size_t tmp_bytes = 0;
      CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, m->d_rowind, d_scol.p, d_pos.p,
                                         d_spos.p, nnz, 0, bits, s));
      DevBuf<unsigned char> d_tmp;
      d_tmp.alloc(tmp_bytes);
      CK(cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, m->d_rowind, d_scol.p, d_pos.p,
                                         d_spos.p, nnz, 0, bits, s));
      m->stage_launches += 4;  // cub onesweep: histogram + scan + per-digit passes (approximate)
      if (m->has_val)
        fill_csc_kernel<true><<<grid_for(nnz, 256, m->sm_count), 256, 0, s>>>(
            d_scol.p, d_spos.p, nnz, m->d_rowptr, nrows, m->d_rowval, m->d_colptr, d_cstart.p,
            m->d_colind, m->d_colval);
      else
        fill_csc_kernel<false><<<grid_for(nnz, 256, m->sm_count), 256, 0, s>>>(
            d_scol.p, d_spos.p, nnz, m->d_rowptr, nrows, nullptr, m->d_colptr, d_cstart.p,
            m->d_colind, nullptr);
      m->stage_launches++;
      CK(cudaStreamSynchronize(s));  // d_tmp must outlive the sort
    }
    if (ncols > 0) {
      const int g = grid_for((int64_t)ncols * 32, 256, m->sm_count);
      if (m->has_val)
        column_norms_kernel<true><<<g, 256, 0, s>>>(ncols, m->d_colptr, m->d_colcnt, m->d_colval,
                                                    m->d_cnorms, m->d_csq);
      else
        column_norms_kernel<false><<<g, 256, 0, s>>>(ncols, m->d_colptr, m->d_colcnt, nullptr,
                                                     m->d_cnorms, m->d_csq);
      m->stage_launches++;
    }
    // unit-valued input: drop the value streams on the device (after the norms were taken from them)
    if (m->has_val && nnz > 0) {
      DevBuf<int32_t> d_flag;
      d_flag.alloc_zero(1, s);
      all_ones_kernel<<<grid_for(nnz, 256, m->sm_count), 256, 0, s>>>(m->d_rowval, nnz, d_flag.p);
      m->stage_launches++;
      int32_t flag = 1;
      CK(cudaMemcpyAsync(&flag, d_flag.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
      const bool keep = getenv("SLIMB200_KEEP_VALUES") && atoi(getenv("SLIMB200_KEEP_VALUES"));
      m->nonneg = (flag & 2) == 0;
      if ((flag & 1) == 0 && !keep) {
        m->unit = true;
        cudaFree(m->d_rowval);
        cudaFree(m->d_colval);
        m->d_rowval = nullptr;
        m->d_colval = nullptr;
      }
    }
    if (ncols > 0) {
      m->rows_per_part = (nrows + kParts - 1) / kParts;
      if (m->rows_per_part < 1) m->rows_per_part = 1;
      CK(cudaMalloc(&m->d_colsplit, sizeof(int32_t) * (size_t)ncols * (kParts + 1)));
      column_split_kernel<<<grid_for((int64_t)ncols * (kParts + 1), 256, m->sm_count), 256, 0, s>>>(
          ncols, m->rows_per_part, m->d_colptr, m->d_colcnt, m->d_colind, m->d_colsplit);
      m->stage_launches++;
    }
    if (ncols > 0) {
      const int32_t nwin = (ncols + 31) / 32;
      CK(cudaMalloc(&m->d_wgram, sizeof(double) * (size_t)nwin * 1024));
      CK(cudaMemsetAsync(m->d_wgram, 0, sizeof(double) * (size_t)nwin * 1024, s));
      // work items: (window, user slice); a window with many entries is split by user range
      std::vector<int32_t> iw, is, iS;
      for (int32_t w = 0; w < nwin; w++) {
        int64_t tot = 0;
        for (int32_t c = w * 32; c < std::min(ncols, w * 32 + 32); c++) tot += m->h_colcnt[c];
        const int32_t S = (int32_t)std::max<int64_t>(1, std::min<int64_t>(128, (tot + 131071) / 131072));
        for (int32_t sl = 0; sl < S; sl++) {
          iw.push_back(w);
          is.push_back(sl);
          iS.push_back(S);
        }
      }
      const int32_t nitems = (int32_t)iw.size();
      DevBuf<int32_t> d_iw, d_is, d_iS;
      d_iw.alloc(nitems);
      d_is.alloc(nitems);
      d_iS.alloc(nitems);
      CK(cudaMemcpyAsync(d_iw.p, iw.data(), sizeof(int32_t) * nitems, cudaMemcpyHostToDevice, s));
      CK(cudaMemcpyAsync(d_is.p, is.data(), sizeof(int32_t) * nitems, cudaMemcpyHostToDevice, s));
      CK(cudaMemcpyAsync(d_iS.p, iS.data(), sizeof(int32_t) * nitems, cudaMemcpyHostToDevice, s));
      const int gw = std::max(1, std::min(nitems, m->sm_count * 2));
      const size_t mask_stride = ((size_t)std::max(nrows, 1) + 3) & ~size_t(3);
      DevBuf<uint32_t> d_masks;
      d_masks.alloc_zero((size_t)gw * mask_stride, s);
      const bool kv = m->has_val && !m->unit;
      if (kv)
        window_gram_kernel<true><<<gw, 512, 0, s>>>(ncols, nrows, nitems, d_iw.p, d_is.p, d_iS.p, m->d_colptr,
                                                    m->d_colcnt, m->d_colind, m->d_colval, d_masks.p, mask_stride,
                                                    m->d_wgram);
      else
        window_gram_kernel<false><<<gw, 512, 0, s>>>(ncols, nrows, nitems, d_iw.p, d_is.p, d_iS.p, m->d_colptr,
                                                     m->d_colcnt, m->d_colind, nullptr, d_masks.p, mask_stride,
                                                     m->d_wgram);
      window_gram_symmetrize_kernel<<<grid_for((int64_t)nwin * 1024, 256, m->sm_count), 256, 0, s>>>(nwin, m->d_wgram);
      m->stage_launches += 2;
      CK(cudaGetLastError());
      CK(cudaStreamSynchronize(s));  // staging buffers are released at scope exit
    }
    build_gram(m);
    CK(cudaGetLastError());
    CK(cudaEventRecord(e1, s));
    CK(cudaStreamSynchronize(s));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    m->stage_ms = ms;
    if (getenv("SLIMB200_VERBOSE") && atoi(getenv("SLIMB200_VERBOSE")) >= 2)
      fprintf(stderr, "[slim-b200] stage(): host wall %.1f ms, CUDA events %.1f ms (Gram build %.1f ms)\n", wall_ms() - w0, ms,
              m->gram_ms);
    if (status) *status = kOk;
    return m;
  } catch (const EngineError &e) {
    g_last_error = e.what();
    if (status) *status = e.status;
  } catch (const std::exception &e) {
    g_last_error = e.what();
    if (status) *status = kErrMemory;
  }
  free_matrix(m);
  return nullptr;
}

int matrix_colcounts(const Matrix *m, int32_t *cnt) {
  for (int32_t c = 0; c < m->ncols; c++) cnt[c] = m->h_colcnt[m->h_rank[c]];
  return kOk;
}

int matrix_item_order_to_host(const Matrix *m, int32_t *rank) {
  if (m->ncols > 0) memcpy(rank, m->h_rank.data(), sizeof(int32_t) * m->ncols);
  return kOk;
}

int matrix_window_gram_to_host(const Matrix *m, double *out) {
  try {
    DeviceGuard guard(m->device);
    const size_t n = (size_t)((m->ncols + 31) / 32) * 1024;
    if (n > 0) CK(cudaMemcpy(out, m->d_wgram, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return kOk;
  } catch (const EngineError &e) {
    g_last_error = e.what();
    return e.status;
  }
}

void matrix_gram_info(const Matrix *m, int32_t *elem_bytes, double *build_ms) {
  if (elem_bytes) *elem_bytes = m->d_gram ? (m->gram_f64 ? 8 : 4) : 0;
  if (build_ms) *build_ms = m->gram_ms;
}

void matrix_gram_stair(const Matrix *m, int32_t *stair, int32_t *hd) {
  if (stair) *stair = (m->d_gram && m->gram_stair) ? 1 : 0;
  if (hd) *hd = (m->d_gram && m->gram_stair) ? m->gram_hd : 0;
}

void matrix_gram_layout(const Matrix *m, int64_t *bytes, int32_t *h32, int32_t *h16) {
  if (bytes) *bytes = m->d_gram ? (int64_t)m->gram_bytes : 0;
  if (h32) *h32 = m->gram_f64 ? 0 : m->gram_h32;
  if (h16) *h16 = m->gram_f64 ? 0 : m->gram_h16;
}

// Dense copy, ncols x ncols: double elements for the fp64 layout, float for the packed layout (every packed entry is
// an integer below 2^24, so the float is exact).
int matrix_gram_to_host(const Matrix *m, void *out) {
  try {
    if (!m->d_gram) throw EngineError(kErrInput, "matrix_gram_to_host: no Gram matrix was staged");
    DeviceGuard guard(m->device);
    const size_t n = (size_t)m->ncols;
    std::vector<unsigned char> raw(m->gram_bytes);
    CK(cudaMemcpy(raw.data(), m->d_gram, raw.size(), cudaMemcpyDeviceToHost));
    if (m->gram_f64) {
      double *o = static_cast<double *>(out);
      const double *g = reinterpret_cast<const double *>(raw.data());
      for (size_t k = 0; k < n; k++)
        for (size_t i = 0; i < n; i++) o[k * n + i] = g[gram_off(n, (int)k, (int)i)];
    } else if (m->gram_stair) {
      // stair layout: element (k, i) is row k of column i when stored, else row i of column k (gram.cuh: GaStair)
      float *o = static_cast<float *>(out);
      const size_t h32 = (size_t)m->gram_h32, h16 = (size_t)m->gram_h16, hd = (size_t)m->gram_hd;
      auto elem = [&](size_t r, size_t c) -> uint32_t {
        const size_t w = c < h32 ? 4 : (c < h16 ? 2 : 1);
        const unsigned char *src = raw.data() + m->h_gram_pbase[c >> 6] + r * kGramPW * w + (c & 63) * w;
        uint32_t v = 0;
        memcpy(&v, src, w);
        return v;
      };
      for (size_t k = 0; k < n; k++)
        for (size_t i = 0; i < n; i++)
          o[k * n + i] = (float)((k < hd || (k >> 6) <= (i >> 6)) ? elem(k, i) : elem(i, k));
    } else {
      float *o = static_cast<float *>(out);
      const size_t h32 = (size_t)m->gram_h32, h16 = (size_t)m->gram_h16;
      for (size_t k = 0; k < n; k++)
        for (size_t i = 0; i < n; i++) {
          uint32_t v;
          if (i < h32) {
            memcpy(&v, raw.data() + ((i >> 6) * n + k) * 256 + (i & 63) * 4, 4);
          } else if (i < h16) {
            const size_t ii = i - h32;
            uint16_t h;
            memcpy(&h, raw.data() + m->gram_off16 + ((ii >> 6) * n + k) * 128 + (ii & 63) * 2, 2);
            v = h;
          } else {
            const size_t ii = i - h16;
            v = raw[m->gram_off8 + ((ii >> 6) * n + k) * 64 + (ii & 63)];
          }
          o[k * n + i] = (float)v;
        }
    }
    return kOk;
  } catch (const EngineError &e) {
    g_last_error = e.what();
    return e.status;
  }
}

int matrix_csc_to_host(const Matrix *m, int64_t *colptr, int32_t *colind, float *colval,
                       float *cnorms) {
  try {
    DeviceGuard guard(m->device);
    std::vector<int64_t> pp((size_t)m->ncols + 1);
    std::vector<int32_t> pind((size_t)std::max<int64_t>(m->nnzp, 1));
    std::vector<float> pval;
    CK(cudaMemcpy(pp.data(), m->d_colptr, sizeof(int64_t) * pp.size(), cudaMemcpyDeviceToHost));
    if (m->nnzp > 0)
      CK(cudaMemcpy(pind.data(), m->d_colind, sizeof(int32_t) * m->nnzp, cudaMemcpyDeviceToHost));
    if (m->has_val && colval) {
      pval.assign((size_t)std::max<int64_t>(m->nnzp, 1), 1.0f);
      if (m->nnzp > 0 && !m->unit)
        CK(cudaMemcpy(pval.data(), m->d_colval, sizeof(float) * m->nnzp, cudaMemcpyDeviceToHost));
    }
    int64_t o = 0;
    for (int32_t c = 0; c < m->ncols; c++) {  // c: ORIGINAL item id, n: internal id
      const int32_t n = m->h_rank[c];
      colptr[c] = o;
      for (int32_t e = 0; e < m->h_colcnt[n]; e++, o++) {
        colind[o] = pind[pp[n] + e];
        if (m->has_val && colval) colval[o] = pval[pp[n] + e];
      }
    }
    colptr[m->ncols] = o;
    if (cnorms && m->ncols > 0) {
      std::vector<float> cn(m->ncols);
      CK(cudaMemcpy(cn.data(), m->d_cnorms, sizeof(float) * m->ncols, cudaMemcpyDeviceToHost));
      for (int32_t c = 0; c < m->ncols; c++) cnorms[c] = cn[m->h_rank[c]];
    }
    return kOk;
  } catch (const EngineError &e) {
    g_last_error = e.what();
    return e.status;
  }
}

// ------------------------------------------------------------------------------------------------
// K1 + K2 + K3: one team (a warp or a CTA) owns one target item column j from start to end.
// ------------------------------------------------------------------------------------------------
struct __align__(32) ActMeta {  // one active coordinate; built once per target, streamed per sweep
  int64_t c0;   // padded offset of column i
  int32_t cnt;  // nnz of column i
  float aty;    // <a_i, y> rounded to float (gk_fkv_t.key; estimate.c:437, cd.c:118)
  double den;   // (double)cnorm_i * cnorm_i + l2r  (cd.c:119,127)
  double sq;    // exact sum of squares of column i (for the one-gather form of cd.c:122-123)
};

struct SolveArgs {
  int32_t nrows, ncols;
  const int64_t *rowptr;
  const int32_t *rowind;
  const float *rowval;
  const int64_t *colptr;
  const int32_t *colcnt;
  const int32_t *colind;
  const float *colval;
  const float *cnorms;
  const double *csq;
  const int32_t *rank;  // original item id -> internal id
  const int32_t *inv;   // internal item id -> original id
  double l1r, l2r, opttol;
  int32_t maxniters;
  const int32_t *targets;  // internal ids
  int32_t ntargets;
  int32_t *queue;
  // warm start (CSC of the initial model), wcolptr == nullptr for cold start
  const int64_t *wcolptr;
  const int32_t *wcolind;
  const float *wcolval;
  int32_t wncols;
  // per-CTA scratch
  double *acc;
  float *xw;
  int32_t *act_idx;
  ActMeta *act_meta;
  double *x;
  double *yhat;
  size_t col_stride, row_stride;
  // outputs (indexed by position in `targets`)
  int32_t *out_cnt;
  int64_t *out_off;
  int32_t *pool_idx;
  float *pool_val;
  unsigned long long *pool_used;
  int64_t pool_cap;
  int32_t *st_niters;
  int32_t *st_nactive;
  int64_t *st_actnnz;
  int64_t *st_expand;
  double *st_rnorm;
  double *st_obj;
  float *st_phase;  // [ntargets][4] microseconds: candidates, active set, sweeps, epilogue (cluster kernel)
  int32_t *st_ngroups;
  // fSLIM: neighbour lists built by fslim_neighbors_kernel (fslim.cuh), nullptr for plain SLIM
  int32_t nnbrs;
  const int32_t *nbr_list;  // [ntargets][nnbrs] internal ids, ascending
  const int32_t *nbr_cnt;   // [ntargets]
};

template <int NT>
__device__ __forceinline__ void team_sync() {
  if (NT == 32) __syncwarp(); else __syncthreads();
}

// Sum over the team; every thread returns the bit-identical value (fixed summation order), so the
// coordinate update below is computed redundantly by all threads and needs no broadcast.
template <int NT>
__device__ __forceinline__ double team_sum(double v, double *red, int &par) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (NT == 32) return v;
  constexpr int NW = NT / 32;
  if ((threadIdx.x & 31) == 0) red[par * NW + (threadIdx.x >> 5)] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NW; i++) s += red[par * NW + i];
  par ^= 1;
  return s;
}

template <int NT>
__device__ __forceinline__ int team_excl_scan(bool flag, int *sc, int &total) {
  const unsigned b = __ballot_sync(0xffffffffu, flag);
  const int lane = threadIdx.x & 31;
  const int pre = __popc(b & ((1u << lane) - 1u));
  const int wt = __popc(b);
  if (NT == 32) {
    total = wt;
    return pre;
  }
  constexpr int NW = NT / 32;
  const int w = threadIdx.x >> 5;
  if (lane == 0) sc[w] = wt;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < NW; i++) {
    const int t = sc[i];
    base += (i < w) ? t : 0;
    tot += t;
  }
  __syncthreads();
  total = tot;
  return base + pre;
}

struct Chunk {
  uint4 ix;
  float4 vv;
};

template <bool HASVAL>
__device__ __forceinline__ void load_chunk(const SolveArgs &a, int64_t c0, int ch, Chunk &r) {
  r.ix = __ldg(reinterpret_cast<const uint4 *>(a.colind + c0) + ch);
  if (HASVAL) r.vv = __ldg(reinterpret_cast<const float4 *>(a.colval + c0) + ch);
}

template <bool HASVAL>
__device__ __forceinline__ double dot_chunk(const Chunk &r, int e0, int cnt, const double *yh) {
  double s = 0.0;
  if (e0 + 0 < cnt) s += HASVAL ? (double)r.vv.x * yh[r.ix.x] : yh[r.ix.x];
  if (e0 + 1 < cnt) s += HASVAL ? (double)r.vv.y * yh[r.ix.y] : yh[r.ix.y];
  if (e0 + 2 < cnt) s += HASVAL ? (double)r.vv.z * yh[r.ix.z] : yh[r.ix.z];
  if (e0 + 3 < cnt) s += HASVAL ? (double)r.vv.w * yh[r.ix.w] : yh[r.ix.w];
  return s;
}

template <bool HASVAL>
__device__ __forceinline__ void axpy_chunk(const Chunk &r, int e0, int cnt, double d, double *yh) {
  if (e0 + 0 < cnt) yh[r.ix.x] += HASVAL ? d * (double)r.vv.x : d;
  if (e0 + 1 < cnt) yh[r.ix.y] += HASVAL ? d * (double)r.vv.y : d;
  if (e0 + 2 < cnt) yh[r.ix.z] += HASVAL ? d * (double)r.vv.z : d;
  if (e0 + 3 < cnt) yh[r.ix.w] += HASVAL ? d * (double)r.vv.w : d;
}

template <int NT>
__host__ __device__ constexpr size_t solve_fixed_smem() {
  return sizeof(double) * 2 * (NT / 32) + sizeof(int) * (NT / 32) + 64;
}

template <int NT, bool YSMEM, bool HASVAL>
__global__ void __launch_bounds__(NT) cd_solve_kernel(const SolveArgs a) {
  constexpr int NW = NT / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *red = reinterpret_cast<double *>(smem_raw);
  int *sc = reinterpret_cast<int *>(red + 2 * NW);
  long long *s_misc = reinterpret_cast<long long *>(smem_raw + ((sizeof(double) * 2 * NW + sizeof(int) * NW + 15) & ~size_t(15)));
  double *yh_s = reinterpret_cast<double *>(smem_raw + ((solve_fixed_smem<NT>() + 15) & ~size_t(15)));

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int par = 0;

  double *acc = a.acc + (size_t)blockIdx.x * a.col_stride;
  float *xw = a.xw ? a.xw + (size_t)blockIdx.x * a.col_stride : nullptr;
  int32_t *act_idx = a.act_idx + (size_t)blockIdx.x * a.col_stride;
  ActMeta *meta = a.act_meta + (size_t)blockIdx.x * a.col_stride;
  double *x = a.x + (size_t)blockIdx.x * a.col_stride;
  double *yh = YSMEM ? yh_s : a.yhat + (size_t)blockIdx.x * a.row_stride;

  if (YSMEM) {
    for (int u = tid; u < a.nrows; u += NT) yh[u] = 0.0;
  }

  for (;;) {
    // ---- next target from the cost-ordered queue -------------------------------------------
    team_sync<NT>();
    if (tid == 0) s_misc[0] = atomicAdd(a.queue, 1);
    team_sync<NT>();
    const int q = (int)s_misc[0];
    if (q >= a.ntargets) break;
    const int j = a.targets[q];
    const int64_t cj0 = a.colptr[j];
    const int cntj = a.colcnt[j];

    // ---- K1: candidates and aTy by CSR row expansion (replaces the full CSC sweep of
    //      estimate.c:412-421): acc[i] += r_ui * r_uj for every user u of column j -------------
    long long expand = 0;
    for (int e = warp; e < cntj; e += NW) {
      const int u = a.colind[cj0 + e];
      const double vy = HASVAL ? (double)a.colval[cj0 + e] : 1.0;
      const int64_t r0 = a.rowptr[u], r1 = a.rowptr[u + 1];
      for (int64_t k = r0 + lane; k < r1; k += 32) {
        const int i = __ldg(a.rowind + k);
        const double prod = HASVAL ? (double)__ldg(a.rowval + k) * vy : 1.0;
        red_add_f64(&acc[i], prod);
      }
      if (lane == 0) expand += (r1 - r0);
    }
    // warm start: scatter column j of the initial model (estimate.c:455-458)
    const int jo = a.inv[j];  // original id of the target (the warm-start model is indexed by original ids)
    const bool warm = a.wcolptr != nullptr && jo < a.wncols;
    if (warm) {
      for (int64_t k = a.wcolptr[jo] + tid; k < a.wcolptr[jo + 1]; k += NT) {
        const int r = a.wcolind[k];
        if (r >= 0 && r < a.ncols) xw[a.rank[r]] = a.wcolval[k];
      }
    }
    __threadfence_block();
    team_sync<NT>();

    // ---- active set: ascending i, strict aTy > l1r, i != j (estimate.c:433-444) ---------------
    int na = 0;
    long long actnnz = 0;
    for (int base = 0; base < a.ncols; base += NT) {
      const int i = base + tid;
      double v = 0.0;
      if (i < a.ncols) {
        v = __ldcg(&acc[i]);  // L2: the atomics above bypass L1
        if (v != 0.0) acc[i] = 0.0;
      }
      const bool flag = (i < a.ncols) && (i != j) && (v > a.l1r);
      int tot;
      const int pos = na + team_excl_scan<NT>(flag, sc, tot);
      if (flag) {
        ActMeta m;
        m.c0 = a.colptr[i];
        m.cnt = a.colcnt[i];
        m.aty = (float)v;
        const double cn = (double)a.cnorms[i];
        m.den = cn * cn + a.l2r;
        m.sq = a.csq[i];
        meta[pos] = m;
        act_idx[pos] = i;
        x[pos] = warm ? (double)xw[i] : 0.0;
        actnnz += m.cnt;
      }
      na += tot;
    }
    team_sync<NT>();
    if (warm) {
      for (int64_t k = a.wcolptr[jo] + tid; k < a.wcolptr[jo + 1]; k += NT) {
        const int r = a.wcolind[k];
        if (r >= 0 && r < a.ncols) xw[a.rank[r]] = 0.0f;
      }
    }

    // ---- iteration cap (estimate.c:448-449) ---------------------------------------------------
    const long long cap64 = 50LL * cntj;
    const int maxit = (int)(cap64 < (long long)a.maxniters ? cap64 : (long long)a.maxniters);

    // ---- yhat = sum x_i a_i for a warm start (cd.c:108-110, with AddSpVec's EPS skip) ---------
    if (warm) {
      for (int p = 0; p < na; p++) {
        const double xi = x[p];
        if (fabs(xi) > kEps) {
          const ActMeta m = meta[p];
          const int nch = (m.cnt + 3) >> 2;
          for (int ch = tid; ch < nch; ch += NT) {
            Chunk c;
            load_chunk<HASVAL>(a, m.c0, ch, c);
            axpy_chunk<HASVAL>(c, ch * 4, m.cnt, xi, yh);
          }
        }
        team_sync<NT>();
      }
    }

    // ---- K2: the coordinate-descent sweeps (cd.c:112-140), fixed ascending order --------------
    int niters = 1;
    if (na > 0 && maxit > 0) {
      ActMeta m_cur = meta[0];
      ActMeta m_nxt = meta[na > 1 ? 1 : 0];
      Chunk c_cur;
      c_cur.ix = make_uint4(0, 0, 0, 0);
      c_cur.vv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tid < ((m_cur.cnt + 3) >> 2)) load_chunk<HASVAL>(a, m_cur.c0, tid, c_cur);
      bool done = false;
      int t = 0;
      for (; t < maxit && !done; t++) {
        double dltx = 0.0;
        for (int p = 0; p < na; p++) {
          const double xi = x[p];
          int p2 = p + 2;
          p2 = p2 >= na ? p2 - na : p2;
          p2 = p2 >= na ? p2 - na : p2;
          if (p2 >= na) p2 = 0;
          const ActMeta m_nn = meta[p2];
          const int nch = (m_cur.cnt + 3) >> 2;

          // <a_i, yhat>: coalesced 128-bit column loads, gathered fp64 yhat, fp64 accumulate
          double part = 0.0;
          if (tid < nch) part = dot_chunk<HASVAL>(c_cur, tid * 4, m_cur.cnt, yh);
          for (int ch = tid + NT; ch < nch; ch += NT) {
            Chunk c;
            load_chunk<HASVAL>(a, m_cur.c0, ch, c);
            part += dot_chunk<HASVAL>(c, ch * 4, m_cur.cnt, yh);
          }
          // issue the next coordinate's first chunk before the reduction (independent of yhat)
          Chunk c_nxt;
          c_nxt.ix = make_uint4(0, 0, 0, 0);
          c_nxt.vv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (tid < ((m_nxt.cnt + 3) >> 2)) load_chunk<HASVAL>(a, m_nxt.c0, tid, c_nxt);

          const double ipf = team_sum<NT>(part, red, par);

          // soft-threshold / shrink update (cd.c:122-128).  The reference removes x_i a_i from yhat,
          // takes the inner product and adds x_i' a_i back; algebraically
          //   ip = <a_i,yhat> - x_i*|a_i|^2,  yhat += (x_i' - x_i) a_i
          // with AddSpVec's |x| <= EPS skip (cd.c:27) kept on both terms.
          const double in_old = fabs(xi) > kEps ? xi : 0.0;
          const double ip = ipf - in_old * m_cur.sq;
          const double num = (double)m_cur.aty - ip;
          const double nx = num > a.l1r ? (num - a.l1r) / m_cur.den : 0.0;
          const double in_new = fabs(nx) > kEps ? nx : 0.0;
          const double d = in_new - in_old;
          if (d != 0.0) {
            if (tid < nch) axpy_chunk<HASVAL>(c_cur, tid * 4, m_cur.cnt, d, yh);
            for (int ch = tid + NT; ch < nch; ch += NT) {
              Chunk c;
              load_chunk<HASVAL>(a, m_cur.c0, ch, c);
              axpy_chunk<HASVAL>(c, ch * 4, m_cur.cnt, d, yh);
            }
          }
          if (tid == 0) x[p] = nx;
          dltx += (nx - xi) * (nx - xi);
          team_sync<NT>();  // yhat (and x[p]) visible to the whole team before the next gather
          m_cur = m_nxt;
          m_nxt = m_nn;
          c_cur = c_nxt;
        }
        if (dltx < a.opttol) done = true;  // cd.c:135-138
      }
      niters = done ? t : maxit + 1;  // cd.c:140 (t + 1 at the break; maxit + 1 when the cap is hit)
    } else if (maxit > 0) {
      niters = (0.0 < a.opttol) ? 1 : maxit + 1;
    }

    // ---- residual / objective (estimate.c:477-489) --------------------------------------------
    double yy = 0.0, yd = 0.0;
    for (int e = tid; e < cntj; e += NT) {
      const int u = a.colind[cj0 + e];
      const double v = HASVAL ? (double)a.colval[cj0 + e] : 1.0;
      yy += v * v;
      yd += v * yh[u];
    }
    team_sync<NT>();
    double hh = 0.0;
    for (int u = tid; u < a.nrows; u += NT) {  // also restores yhat = 0 (estimate.c:528-530)
      const double h = yh[u];
      if (h != 0.0) {
        hh += h * h;
        yh[u] = 0.0;
      }
    }
    double reg = 0.0;
    int nnz_local = 0;
    for (int p = tid; p < na; p += NT) {
      const double xv = x[p];
      reg += 0.5 * a.l2r * xv * xv + a.l1r * fabs(xv);
      nnz_local += fabs(xv) > kEps ? 1 : 0;
    }
    yy = team_sum<NT>(yy, red, par);
    yd = team_sum<NT>(yd, red, par);
    hh = team_sum<NT>(hh, red, par);
    reg = team_sum<NT>(reg, red, par);
    const int nnz_w = (int)(team_sum<NT>((double)nnz_local, red, par) + 0.5);
    const double expand_t = team_sum<NT>((double)expand, red, par);
    const double actnnz_t = team_sum<NT>((double)actnnz, red, par);

    // ---- K3: compaction |x| > EPS -> (i, (float)x) ascending i (estimate.c:492-505) ------------
    if (tid == 0) {
      const unsigned long long off = atomicAdd(a.pool_used, (unsigned long long)nnz_w);
      s_misc[1] = (long long)off;
    }
    team_sync<NT>();
    const long long off = s_misc[1];
    const bool fits = off + nnz_w <= a.pool_cap;
    if (fits) {
      int w0 = 0;
      for (int base = 0; base < na; base += NT) {
        const int p = base + tid;
        double xv = 0.0;
        if (p < na) xv = x[p];
        const bool flag = (p < na) && fabs(xv) > kEps;
        int tot;
        const int pos = w0 + team_excl_scan<NT>(flag, sc, tot);
        if (flag) {
          a.pool_idx[off + pos] = a.inv[act_idx[p]];  // original item id
          a.pool_val[off + pos] = (float)xv;
        }
        w0 += tot;
      }
    }
    if (tid == 0) {
      a.out_cnt[q] = fits ? nnz_w : -1 - nnz_w;
      a.out_off[q] = off;
      a.st_niters[q] = niters;
      a.st_nactive[q] = na;
      a.st_actnnz[q] = (long long)(actnnz_t + 0.5);
      a.st_expand[q] = (long long)(expand_t + 0.5);
      const double rn = 0.5 * (yy - 2.0 * yd + hh);
      a.st_rnorm[q] = rn;
      a.st_obj[q] = rn + reg;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Cluster variant for large user counts (yhat does not fit shared memory): one thread-block CLUSTER
// owns one target column.  CTA r of the cluster owns the user range [r*U, (r+1)*U) -- for every
// column that is a contiguous entry range (colsplit) -- and therefore a PRIVATE slice of yhat: the
// gather and the update of a coordinate touch only that slice, so the only cross-CTA traffic per
// coordinate is one partial inner product per CTA, exchanged through distributed shared memory with
// a single cluster barrier.  The number of clusters in flight is capped so that all their yhat
// vectors stay resident in the 126 MB L2 while the column streams come from HBM.
// ------------------------------------------------------------------------------------------------
struct __align__(128) ActMetaC {  // one active coordinate, one 128-byte line
  int64_t c0;
  int32_t cnt;
  float aty;
  double den;
  double sq;
  int32_t split[kParts + 1];
  int32_t pad[7];
};
static_assert(sizeof(ActMetaC) == 128, "one line per active coordinate");

struct GroupMeta {  // the active coordinates of one 32-item window
  int32_t win;      // window id = item id / 32
  uint32_t mask;    // bit b set: item win*32+b is active
  int32_t pbase;    // position of the window's first active coordinate in the active list
  int32_t pad;
};

struct ClusterArgs {
  const int32_t *colsplit;
  ActMetaC *meta;       // per cluster [col_stride]
  double *xc;           // per CTA [col_stride]: every CTA keeps its own copy of x (identical values)
  int32_t rows_per_part;
  const double *wgram;  // window Gram blocks (staging)
  int32_t nonneg;       // all ratings >= 0: Gram blocks are >= 0
  GroupMeta *groups;    // per cluster [grp_stride]
  size_t grp_stride;
};

struct CoordView {  // what one CTA needs to know about one coordinate
  int64_t c0;
  int32_t s0, s1;  // this CTA's entry range inside the column
  float aty;
  double den, sq;
};

// The line was written by CTA 0 of the cluster: read it from L2 (ld.global.cg), never from this SM's L1.
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ CoordView load_view(const ActMetaC *m, int pr0, int pr1) {
  CoordView v;
  const int4 h0 = __ldcg(reinterpret_cast<const int4 *>(m));      // c0, cnt, aty
  const int4 h1 = __ldcg(reinterpret_cast<const int4 *>(m) + 1);  // den, sq
  v.c0 = (int64_t)(((unsigned long long)(unsigned)h0.y << 32) | (unsigned)h0.x);
  v.aty = __int_as_float(h0.w);
  v.den = __hiloint2double(h1.y, h1.x);
  v.sq = __hiloint2double(h1.w, h1.z);
  v.s0 = __ldcg(&m->split[pr0]);
  v.s1 = __ldcg(&m->split[pr1]);
  return v;
}

// L2 = true: yhat is read with ld.global.cg and updated with fp64 atomics.  The window sweep needs
// this because several warps update DIFFERENT columns of a window at the same time and those columns
// can share users (inside one column the users are distinct, so the one-column-at-a-time paths can use
// plain loads and stores).
template <bool HASVAL, bool L2 = false>
__device__ __forceinline__ double dot_chunk_r(const Chunk &r, int e0, int lo, int hi, const double *yh) {
  // The four gathers are UNCONDITIONAL (every id of a padded column chunk is a valid user id; entries
  // outside [lo, hi) belong to another CTA's range or are padding) so that they issue back to back and
  // their latencies overlap; the range test only selects which values are summed.
#define SLIM_YH(i) (L2 ? __ldcg(yh + (i)) : yh[(i)])
  const double y0 = SLIM_YH(r.ix.x), y1 = SLIM_YH(r.ix.y), y2 = SLIM_YH(r.ix.z), y3 = SLIM_YH(r.ix.w);
#undef SLIM_YH
  const bool k0 = e0 + 0 >= lo && e0 + 0 < hi, k1 = e0 + 1 >= lo && e0 + 1 < hi;
  const bool k2 = e0 + 2 >= lo && e0 + 2 < hi, k3 = e0 + 3 >= lo && e0 + 3 < hi;
  double s = 0.0;
  s += k0 ? (HASVAL ? (double)r.vv.x * y0 : y0) : 0.0;
  s += k1 ? (HASVAL ? (double)r.vv.y * y1 : y1) : 0.0;
  s += k2 ? (HASVAL ? (double)r.vv.z * y2 : y2) : 0.0;
  s += k3 ? (HASVAL ? (double)r.vv.w * y3 : y3) : 0.0;
  return s;
}

template <bool HASVAL, bool L2 = false>
__device__ __forceinline__ void axpy_chunk_r(const Chunk &r, int e0, int lo, int hi, double d, double *yh) {
#define SLIM_UPD(i, v)                   \
  do {                                   \
    if (L2) red_add_f64(yh + (i), (v));  \
    else yh[(i)] += (v);                 \
  } while (0)
  if (e0 + 0 >= lo && e0 + 0 < hi) SLIM_UPD(r.ix.x, HASVAL ? d * (double)r.vv.x : d);
  if (e0 + 1 >= lo && e0 + 1 < hi) SLIM_UPD(r.ix.y, HASVAL ? d * (double)r.vv.y : d);
  if (e0 + 2 >= lo && e0 + 2 < hi) SLIM_UPD(r.ix.z, HASVAL ? d * (double)r.vv.z : d);
  if (e0 + 3 >= lo && e0 + 3 < hi) SLIM_UPD(r.ix.w, HASVAL ? d * (double)r.vv.w : d);
#undef SLIM_UPD
}

#ifndef SLIM_CLUSTER_NT
#define SLIM_CLUSTER_NT 256
#endif
constexpr int kClusterNT = SLIM_CLUSTER_NT;                 // threads per CTA of the cluster kernel
#ifndef SLIM_CTAS_PER_SM
#define SLIM_CTAS_PER_SM (SLIM_CLUSTER_NT <= 256 ? 2 : 1)
#endif
constexpr int kClusterCtasPerSm = SLIM_CTAS_PER_SM;  // two co-resident CTAs interleave their rounds
constexpr int kWarpSlots = 5;  // small columns of a round per consumer warp: ceil(32 / (kClusterNT/32 - 1))

constexpr int kSmallCol = 1024;  // entries of a column inside one CTA's user range handled by ONE warp

struct ClusterSmem {
  double red[2][kClusterNT / 32];  // per-warp partials (double buffered)
  double part[2][kParts];          // per-CTA partials of the whole cluster, written through DSMEM
  int sc[kClusterNT / 32];
  int q;
  int na;
  int ng;
  long long off;
};

// ---- window sweep: shared-memory pipeline --------------------------------------------------------
// Static description of one round (= one 32-item window) for THIS CTA, built one round ahead of its
// copies and two rounds ahead of its use by the producer warp.
struct __align__(16) RoundTab {
  long long c0[32];   // padded offset of the column of slot b
  double inv_den[32]; // 1 / (cnorm^2 + l2r)
  double sq[32];      // exact sum of squares
  int4 rec[32];       // .x/.y: this CTA's entry range [s0, s1) in the column; .z: first 16-byte chunk of the
                      // staged copy in the stage buffer (-1: not staged); .w: position in the active list (x[p])
  float aty[32];
  signed char wfast[kClusterNT / 32];     // consumer warp w: all its columns are staged, <= 1 chunk per lane
  int bigpre[32];     // big slots: first kBigBlk-entry block of the slot in the round's flattened block list
  int bigtot;         // number of kBigBlk-entry blocks of all big slots
  alignas(8) signed char wslot[kClusterNT / 32][8];  // the (up to kWarpSlots) small columns gathered by consumer
                                                     // warp w (-1: none); read as one 8-byte word
  signed char abit[32];                   // active slots in ascending order (the solve's visiting order)
  unsigned mask;      // active slots
  unsigned bigmask;   // active slots whose range is too long for one warp: gathered by the whole CTA
  int win;            // window id
  unsigned bytes;     // bytes of column data staged for the round
  int nact;
  int pad[3];
};

template <bool HASVAL>
struct __align__(128) PipeSmem {
  static constexpr int CAP = (HASVAL ? 1536 : 3072) / kClusterCtasPerSm;  // 16-byte chunks per stage buffer
  uint4 sidx[2][CAP];                               // staged user ids of the round's small columns
  float4 sval[HASVAL ? 2 : 1][HASVAL ? CAP : 1];    // ... and their values
  double G[2][32][32];                              // Gram block of the window, G[m][k] = <a_m, a_k>
  unsigned long long pall[2][kParts][32][2];        // partial inner products of every CTA: {lo32|tag}, {hi32|tag}
  RoundTab tab[3];
  double pcta[2][32];                               // this CTA's partial inner products per slot (per exchange buffer)
  double dlt[32];                                   // yhat step per slot
  double pw[kClusterNT / 32];
  double pwb[kClusterNT / 32][32];                  // per-warp partial inner products of the big slots
  double dl;
  unsigned long long mbar[2];                       // "stage buffer b has landed"
  unsigned long long xbar[2];                       // "all CTAs' partials of exchange buffer b have landed"
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
// TMA 1-D bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// Flag-in-data exchange (the "LL" idea of NCCL): every 8-byte remote store carries 4 bytes of payload
// and a 4-byte round tag; 8-byte stores are single-copy atomic, so the receiver simply polls its OWN
// shared memory until all tags match -- no fence, no barrier, no mbarrier on the critical path.
__device__ __forceinline__ void st_peer_tagged(unsigned long long *local_dst, uint32_t peer_rank, double v,
                                               uint32_t tag) {
  uint32_t rdst;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rdst) : "r"(smem_u32(local_dst)), "r"(peer_rank));
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  const unsigned long long w0 = (bits & 0xffffffffull) | ((unsigned long long)tag << 32);
  const unsigned long long w1 = (bits >> 32) | ((unsigned long long)tag << 32);
  asm volatile("st.shared::cluster.v2.u64 [%0], {%1, %2};" ::"r"(rdst), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ double ld_tagged_wait(const unsigned long long *src, uint32_t tag) {
  unsigned long long w0, w1;
  const uint32_t addr = smem_u32(src);
  do {
    asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "r"(addr) : "memory");
  } while ((uint32_t)(w0 >> 32) != tag || (uint32_t)(w1 >> 32) != tag);
  return __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// <a, yhat> / yhat += d a over a STAGED column range: user ids (and values) come from shared memory
template <bool HASVAL>
__device__ __forceinline__ double warp_dot_staged(const uint4 *sidx, const float4 *sval, int s0, int s1,
                                                  const double *yh) {
  const int lane = threadIdx.x & 31;
  const int ch0 = s0 >> 2, nch = ((s1 + 3) >> 2) - ch0;
  double part = 0.0;
  int k = lane;
  for (; k + 96 < nch; k += 128) {  // 4 chunks = 16 yhat gathers in flight per lane
    Chunk c0, c1, c2, c3;
    c0.ix = sidx[k];
    c1.ix = sidx[k + 32];
    c2.ix = sidx[k + 64];
    c3.ix = sidx[k + 96];
    if (HASVAL) {
      c0.vv = sval[k];
      c1.vv = sval[k + 32];
      c2.vv = sval[k + 64];
      c3.vv = sval[k + 96];
    }
    const double a0 = dot_chunk_r<HASVAL, true>(c0, (ch0 + k) * 4, s0, s1, yh);
    const double a1 = dot_chunk_r<HASVAL, true>(c1, (ch0 + k + 32) * 4, s0, s1, yh);
    const double a2 = dot_chunk_r<HASVAL, true>(c2, (ch0 + k + 64) * 4, s0, s1, yh);
    const double a3 = dot_chunk_r<HASVAL, true>(c3, (ch0 + k + 96) * 4, s0, s1, yh);
    part += (a0 + a1) + (a2 + a3);
  }
  for (; k < nch; k += 32) {
    Chunk c;
    c.ix = sidx[k];
    if (HASVAL) c.vv = sval[k];
    part += dot_chunk_r<HASVAL, true>(c, (ch0 + k) * 4, s0, s1, yh);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  return part;
}

template <bool HASVAL>
__device__ __forceinline__ void warp_axpy_staged(const uint4 *sidx, const float4 *sval, int s0, int s1, double d,
                                                 double *yh) {
  const int lane = threadIdx.x & 31;
  const int ch0 = s0 >> 2, nch = ((s1 + 3) >> 2) - ch0;
#pragma unroll 2
  for (int k = lane; k < nch; k += 32) {
    Chunk c;
    c.ix = sidx[k];
    if (HASVAL) c.vv = sval[k];
    axpy_chunk_r<HASVAL, true>(c, (ch0 + k) * 4, s0, s1, d, yh);
  }
}

// <a, yhat> over entries [s0, s1) of a column, one warp (lanes stride the 16-byte chunks)
template <bool HASVAL>
__device__ __forceinline__ double warp_dot(const SolveArgs &a, int64_t c0, int s0, int s1, const double *yh) {
  const int lane = threadIdx.x & 31;
  double part = 0.0;
#pragma unroll 2
  for (int ch = (s0 >> 2) + lane; ch < ((s1 + 3) >> 2); ch += 32) {
    Chunk c;
    load_chunk<HASVAL>(a, c0, ch, c);
    part += dot_chunk_r<HASVAL, true>(c, ch * 4, s0, s1, yh);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  return part;
}

template <bool HASVAL>
__device__ __forceinline__ void warp_axpy(const SolveArgs &a, int64_t c0, int s0, int s1, double d, double *yh) {
  const int lane = threadIdx.x & 31;
#pragma unroll 2
  for (int ch = (s0 >> 2) + lane; ch < ((s1 + 3) >> 2); ch += 32) {
    Chunk c;
    load_chunk<HASVAL>(a, c0, ch, c);
    axpy_chunk_r<HASVAL, true>(c, ch * 4, s0, s1, d, yh);
  }
}

// One block of kBigBlk entries [e0, e0+kBigBlk) of a column range [s0, s1) by one warp, "transposed":
// gather / atomic instruction k covers the 32 CONSECUTIVE entries e0 + 32k .. e0 + 32k + 31 (ascending,
// for dense head columns nearly consecutive users), so the lanes of an instruction share 32-byte yhat
// sectors.  All id loads of the block are issued before the first gather and all gathers before the
// first add: 16 loads in flight per lane hide the DRAM / L2 latency inside one warp.
constexpr int kBigBlk = 512;
constexpr int kBigShift = 9;
constexpr int kBigPer = kBigBlk / 32;

template <bool HASVAL>
__device__ __forceinline__ double warp_bigblock_dot(const SolveArgs &a, int64_t c0, int s0, int s1, int e0,
                                                    const double *yh) {
  const int lane = threadIdx.x & 31;
  const int32_t *ix = a.colind + c0;
  const float *vv = HASVAL ? a.colval + c0 : nullptr;
  const int lim = (s1 + 3) & ~3;  // padded columns are readable up to the next multiple of 4 entries
  const int e = e0 + lane;
  int id[kBigPer];
  float vl[kBigPer];
#pragma unroll
  for (int k = 0; k < kBigPer; k++) {
    id[k] = (e + 32 * k < lim) ? __ldg(ix + e + 32 * k) : 0;
    if (HASVAL) vl[k] = (e + 32 * k >= s0 && e + 32 * k < s1) ? __ldg(vv + e + 32 * k) : 0.f;
  }
  double y[kBigPer];
#pragma unroll
  for (int k = 0; k < kBigPer; k++) y[k] = __ldcg(yh + id[k]);
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < kBigPer; k++) {
    const bool ok = e + 32 * k >= s0 && e + 32 * k < s1;
    acc += ok ? (HASVAL ? (double)vl[k] * y[k] : y[k]) : 0.0;
  }
  return acc;
}

template <bool HASVAL>
__device__ __forceinline__ void warp_bigblock_axpy(const SolveArgs &a, int64_t c0, int s0, int s1, int e0, double d,
                                                   double *yh) {
  const int lane = threadIdx.x & 31;
  const int32_t *ix = a.colind + c0;
  const float *vv = HASVAL ? a.colval + c0 : nullptr;
  const int lim = (s1 + 3) & ~3;
  const int e = e0 + lane;
  int id[kBigPer];
  float vl[kBigPer];
#pragma unroll
  for (int k = 0; k < kBigPer; k++) {
    id[k] = (e + 32 * k < lim) ? __ldg(ix + e + 32 * k) : 0;
    if (HASVAL) vl[k] = (e + 32 * k >= s0 && e + 32 * k < s1) ? __ldg(vv + e + 32 * k) : 0.f;
  }
#pragma unroll
  for (int k = 0; k < kBigPer; k++)
    if (e + 32 * k >= s0 && e + 32 * k < s1) red_add_f64(yh + id[k], HASVAL ? d * (double)vl[k] : d);
}

// Same over the whole CTA for LONG column ranges.  Entries are taken "transposed": one gather / atomic
// instruction of a warp covers 32 CONSECUTIVE entries of the column (ascending, for dense head columns
// nearly consecutive users), so the lanes of an instruction share 32-byte yhat sectors instead of each
// lane touching its own one -- up to 4x fewer L2 sector requests on the dense columns that dominate a
// typical target.  4 x 32-bit id loads (each coalesced) replace one 128-bit load per lane.
template <bool HASVAL>
__device__ __forceinline__ double block_dot(const SolveArgs &a, int64_t c0, int s0, int s1, const double *yh) {
  constexpr int NT = kClusterNT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int32_t *ix = a.colind + c0;
  const float *vv = HASVAL ? a.colval + c0 : nullptr;
  const int e_begin = s0 & ~127;  // blocks of 128 entries, one block per warp per step
  const int lim = (s1 + 3) & ~3;  // padded columns are readable up to the next multiple of 4 entries
  double part = 0.0;
  for (int e0 = e_begin + warp * 128; e0 < s1; e0 += (NT / 32) * 128) {
    const int e = e0 + lane;
    int i0 = 0, i1 = 0, i2 = 0, i3 = 0;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
    const bool k0 = e >= s0 && e < s1, k1 = e + 32 >= s0 && e + 32 < s1;
    const bool k2 = e + 64 >= s0 && e + 64 < s1, k3 = e + 96 >= s0 && e + 96 < s1;
    if (e < lim) i0 = __ldg(ix + e);
    if (e + 32 < lim) i1 = __ldg(ix + e + 32);
    if (e + 64 < lim) i2 = __ldg(ix + e + 64);
    if (e + 96 < lim) i3 = __ldg(ix + e + 96);
    if (HASVAL) {
      if (k0) v0 = __ldg(vv + e);
      if (k1) v1 = __ldg(vv + e + 32);
      if (k2) v2 = __ldg(vv + e + 64);
      if (k3) v3 = __ldg(vv + e + 96);
    }
    const double y0 = __ldcg(yh + i0), y1 = __ldcg(yh + i1), y2 = __ldcg(yh + i2), y3 = __ldcg(yh + i3);
    double acc = 0.0;
    acc += k0 ? (HASVAL ? (double)v0 * y0 : y0) : 0.0;
    acc += k1 ? (HASVAL ? (double)v1 * y1 : y1) : 0.0;
    acc += k2 ? (HASVAL ? (double)v2 * y2 : y2) : 0.0;
    acc += k3 ? (HASVAL ? (double)v3 * y3 : y3) : 0.0;
    part += acc;
  }
  return part;
}

template <bool HASVAL>
__device__ __forceinline__ void block_axpy(const SolveArgs &a, int64_t c0, int s0, int s1, double d, double *yh) {
  constexpr int NT = kClusterNT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int32_t *ix = a.colind + c0;
  const float *vv = HASVAL ? a.colval + c0 : nullptr;
  const int e_begin = s0 & ~127;
  for (int e0 = e_begin + warp * 128; e0 < s1; e0 += (NT / 32) * 128) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int e = e0 + lane + 32 * k;
      if (e >= s0 && e < s1) {
        const int u = __ldg(ix + e);
        red_add_f64(yh + u, HASVAL ? d * (double)__ldg(vv + e) : d);
      }
    }
  }
}

// Sum over all threads of the cluster; every thread of every CTA returns the bit-identical value.
__device__ __forceinline__ double cluster_sum(double v, ClusterSmem &sm, int &par, cg::cluster_group &cl,
                                              int cs, int rank) {
  constexpr int NW = kClusterNT / 32;
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sm.red[par][threadIdx.x >> 5] = v;
  __syncthreads();
  if ((int)threadIdx.x < cs) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NW; i++) s += sm.red[par][i];
    double *dst = cl.map_shared_rank(&sm.part[par][rank], threadIdx.x);
    *dst = s;
  }
  cl.sync();
  double tot = 0.0;
  for (int c = 0; c < cs; c++) tot += sm.part[par][c];
  par ^= 1;
  return tot;
}

#ifdef SLIM_PROFILE_ROUNDS
#define SLIM_TICK(k)                          \
  do {                                        \
    const long long now_ = clock64();         \
    prof[k] += now_ - tick_;                  \
    tick_ = now_;                             \
  } while (0)
#else
#define SLIM_TICK(k) do { } while (0)
#endif

template <bool HASVAL, bool WINDOW>
__global__ void __launch_bounds__(kClusterNT, kClusterCtasPerSm) cd_cluster_kernel(const SolveArgs a, const ClusterArgs ca) {
  constexpr int NT = kClusterNT, NW = NT / 32;
  constexpr int NCW = NW - 1;  // consumer warps of the window sweep (the last warp is the producer)
  __shared__ ClusterSmem sm;
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  PipeSmem<HASVAL> &ps = *reinterpret_cast<PipeSmem<HASVAL> *>(dyn_smem);
  unsigned rr = 0;  // rounds consumed so far by this CTA (kernel lifetime): buffer rr&1, parity (rr>>1)&1
  unsigned xr = 0;  // exchanges done so far (kernel lifetime): buffer xr&1, parity (xr>>1)&1
  cg::cluster_group cl = cg::this_cluster();
  const int cs = (int)cl.num_blocks();
  const int rank = (int)cl.block_rank();
  const int cid = blockIdx.x / cs;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int par = 0;

  const int pr0 = rank * (kParts / cs), pr1 = (rank + 1) * (kParts / cs);
  const int64_t ulo64 = (int64_t)pr0 * ca.rows_per_part, uhi64 = (int64_t)pr1 * ca.rows_per_part;
  const int u_lo = (int)(ulo64 < a.nrows ? ulo64 : a.nrows), u_hi = (int)(uhi64 < a.nrows ? uhi64 : a.nrows);

  double *acc = a.acc + (size_t)cid * a.col_stride;
  float *xw = a.xw + (size_t)cid * a.col_stride;
  int32_t *act_idx = a.act_idx + (size_t)cid * a.col_stride;
  ActMetaC *meta = ca.meta + (size_t)cid * a.col_stride;
  GroupMeta *groups = ca.groups + (size_t)cid * ca.grp_stride;
  double *x = ca.xc + (size_t)blockIdx.x * a.col_stride;
  double *yh = a.yhat + (size_t)cid * a.row_stride;

  if (WINDOW) {
    for (int k = tid; k < 2 * kParts * 32 * 2; k += NT) (&ps.pall[0][0][0][0])[k] = 0ull;  // tag 0 = "nothing yet"
    if (tid == 0) {
      mbar_init(&ps.mbar[0], 1);
      mbar_init(&ps.mbar[1], 1);
      mbar_init(&ps.xbar[0], (uint32_t)cs);  // one arrival per source CTA of the cluster
      mbar_init(&ps.xbar[1], (uint32_t)cs);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      fence_proxy_async();
    }
    __syncthreads();
  }

  for (;;) {
    // ---- next target -----------------------------------------------------------------------------
    if (rank == 0 && tid == 0) {
      const int q = atomicAdd(a.queue, 1);
      for (int c = 0; c < cs; c++) *cl.map_shared_rank(&sm.q, c) = q;
    }
    cl.sync();
    const int q = sm.q;
    if (q >= a.ntargets) break;
    const int j = a.targets[q];
    const int64_t cj0 = a.colptr[j];
    const int cntj = a.colcnt[j];
    const unsigned long long tm0 = globaltimer_ns();

    // ---- K1: candidates / aTy by CSR row expansion, all warps of the cluster ------------------------
    long long expand = 0;
    for (int e = rank * NW + warp; e < cntj; e += cs * NW) {
      const int u = a.colind[cj0 + e];
      const double vy = HASVAL ? (double)a.colval[cj0 + e] : 1.0;
      const int64_t r0 = a.rowptr[u], r1 = a.rowptr[u + 1];
      for (int64_t k = r0 + lane; k < r1; k += 32) {
        const int i = __ldg(a.rowind + k);
        const double prod = HASVAL ? (double)__ldg(a.rowval + k) * vy : 1.0;
        red_add_f64(&acc[i], prod);
      }
      if (lane == 0) expand += (r1 - r0);
    }
    const int jo = a.inv[j];  // original id of the target (the warm-start model is indexed by original ids)
    const bool warm = a.wcolptr != nullptr && jo < a.wncols;
    if (warm && rank == 0) {
      for (int64_t k = a.wcolptr[jo] + tid; k < a.wcolptr[jo + 1]; k += NT) {
        const int r = a.wcolind[k];
        if (r >= 0 && r < a.ncols) xw[a.rank[r]] = a.wcolval[k];
      }
    }
    __threadfence();
    cl.sync();
    const unsigned long long tm1 = globaltimer_ns();

    // ---- active set, built by CTA 0, read by all ----------------------------------------------------
    long long actnnz = 0;
    if (rank == 0) {
      int na = 0, ng = 0;
      for (int base = 0; base < a.ncols; base += NT) {
        const int i = base + tid;
        double v = 0.0;
        if (i < a.ncols) {
          v = __ldcg(&acc[i]);
          if (v != 0.0) __stcg(&acc[i], 0.0);
        }
        const bool flag = (i < a.ncols) && (i != j) && (v > a.l1r);
        int tot;
        const int pos = na + team_excl_scan<NT>(flag, sm.sc, tot);
        if (flag) {
          ActMetaC m;
          m.c0 = a.colptr[i];
          m.cnt = a.colcnt[i];
          m.aty = (float)v;
          const double cn = (double)a.cnorms[i];
          m.den = cn * cn + a.l2r;
          m.sq = a.csq[i];
          const int32_t *sp = ca.colsplit + (size_t)i * (kParts + 1);
#pragma unroll
          for (int k = 0; k <= kParts; k++) m.split[k] = sp[k];
#pragma unroll
          for (int k = 0; k < 7; k++) m.pad[k] = 0;
          meta[pos] = m;
          act_idx[pos] = i;
          actnnz += m.cnt;
        }
        if (WINDOW) {
          // a warp covers exactly one 32-item window (base is a multiple of 512): its ballot is the
          // window's active mask, and lane 0's scan value is the position of its first active coordinate
          const unsigned wmask = __ballot_sync(0xffffffffu, flag);
          const bool gflag = (lane == 0) && (wmask != 0u);
          int gtot;
          const int gpos = ng + team_excl_scan<NT>(gflag, sm.sc, gtot);
          if (gflag) {
            GroupMeta gm;
            gm.win = (base >> 5) + warp;
            gm.mask = wmask;
            gm.pbase = pos;
            gm.pad = 0;
            groups[gpos] = gm;
          }
          ng += gtot;
        }
        na += tot;
      }
      if (tid == 0)
        for (int c = 0; c < cs; c++) {
          *cl.map_shared_rank(&sm.na, c) = na;
          *cl.map_shared_rank(&sm.ng, c) = ng;
        }
    }
    __threadfence();
    cl.sync();
    const int na = sm.na;

    // every CTA initialises its own copy of x (identical values everywhere)
    for (int p = tid; p < na; p += NT) x[p] = warm ? (double)__ldcg(&xw[__ldcg(&act_idx[p])]) : 0.0;
    __syncthreads();
    cl.sync();  // all CTAs have read xw before CTA 0 clears it
    if (warm && rank == 0) {
      for (int64_t k = a.wcolptr[jo] + tid; k < a.wcolptr[jo + 1]; k += NT) {
        const int r = a.wcolind[k];
        if (r >= 0 && r < a.ncols) xw[a.rank[r]] = 0.0f;
      }
    }

    const long long cap64 = 50LL * cntj;
    const int maxit = (int)(cap64 < (long long)a.maxniters ? cap64 : (long long)a.maxniters);
    const unsigned long long tm2 = globaltimer_ns();

    // ---- warm start: yhat slice = sum x_i a_i over this CTA's user range ------------------------------
    if (warm) {
      for (int p = 0; p < na; p++) {
        const double xi = x[p];
        if (fabs(xi) > kEps) {
          const CoordView v = load_view(&meta[p], pr0, pr1);
          for (int ch = (v.s0 >> 2) + tid; ch < ((v.s1 + 3) >> 2); ch += NT) {
            Chunk c;
            load_chunk<HASVAL>(a, v.c0, ch, c);
            axpy_chunk_r<HASVAL>(c, ch * 4, v.s0, v.s1, xi, yh);
          }
        }
        __syncthreads();
      }
    }

    // ---- K2: sweeps ---------------------------------------------------------------------------------
    int niters = 1;
    if (WINDOW && na > 0 && maxit > 0) {
      // Exact block form of sequential CD.  For the active coordinates k of one 32-item window, taken
      // in ascending order, sequential CD needs ip_k = <a_k, yhat> AFTER the steps of the earlier
      // coordinates m < k of the window:  <a_k, yhat_0> + sum_{m<k} d_m <a_k, a_m>.  The first term comes
      // from ONE pass over the window's columns against the yhat at the start of the round, the second
      // from the precomputed Gram block, so a round costs one cluster barrier for up to 32 coordinates.
      //
      // Software pipeline: everything a round needs that does not depend on yhat -- its slot table, the
      // user ids of its columns inside this CTA's user range, its Gram block -- is brought into shared
      // memory while the previous round runs: the producer warp builds the table two rounds ahead and
      // issues TMA bulk copies (cp.async.bulk, completion on an mbarrier) one round ahead.  On the
      // critical path of a round remain: the yhat gather (L2), the exchange of 32 partials through
      // DSMEM + one cluster barrier, the Gram-space solve, and the yhat update.
      constexpr int CAP = PipeSmem<HASVAL>::CAP;
      const int ng = sm.ng;

      // Producer warp, slot table of local round r, split in two phases so that its two dependent L2
      // reads never sit on a round's critical path: tab_load() issues the reads of the coordinate lines
      // (given the window record read one round earlier), tab_store() consumes them a round later.
      struct TabRegs {
        int4 gm;          // window record: id, active mask, first position
        int4 h0, h1;      // coordinate line of slot `lane`
        int sp0, sp1;
      };
      auto tab_load = [&](const int4 gm, TabRegs &R) {
        R.gm = gm;
        const unsigned mask = (unsigned)gm.y;
        if ((mask >> lane) & 1u) {
          const ActMetaC *m = &meta[gm.z + __popc(mask & ((1u << lane) - 1u))];
          R.h0 = __ldcg(reinterpret_cast<const int4 *>(m));
          R.h1 = __ldcg(reinterpret_cast<const int4 *>(m) + 1);
          R.sp0 = __ldcg(&m->split[pr0]);
          R.sp1 = __ldcg(&m->split[pr1]);
        }
      };
      auto tab_store = [&](int r, const TabRegs &R) {
        RoundTab &T = ps.tab[r % 3];
        const unsigned mask = (unsigned)R.gm.y;
        const bool act = (mask >> lane) & 1u;
        int nch = 0, slot_p = 0;
        bool small = false;
        if (act) {
          const double den = __hiloint2double(R.h1.y, R.h1.x);
          T.c0[lane] = (long long)(((unsigned long long)(unsigned)R.h0.y << 32) | (unsigned)R.h0.x);
          slot_p = R.gm.z + __popc(mask & ((1u << lane) - 1u));
          T.aty[lane] = __int_as_float(R.h0.w);
          T.inv_den[lane] = 1.0 / den;
          T.sq[lane] = __hiloint2double(R.h1.w, R.h1.z);
          nch = ((R.sp1 + 3) >> 2) - (R.sp0 >> 2);
          small = (R.sp1 > R.sp0) && (R.sp1 - R.sp0 <= kSmallCol);
        }
        int want = small ? nch : 0, pre = want;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int up = __shfl_up_sync(0xffffffffu, pre, o);
          if (lane >= o) pre += up;
        }
        const bool staged = small && pre <= CAP;  // pre is the inclusive prefix
        T.rec[lane] = make_int4(act ? R.sp0 : 0, act ? R.sp1 : 0, staged ? pre - want : -1, slot_p);
        // work lists: the ord-th small column goes to consumer warp ord % NCW; big columns to the whole CTA
        for (int k = lane; k < 8 * NW; k += 32) (&T.wslot[0][0])[k] = -1;
        __syncwarp();
        const unsigned smask = __ballot_sync(0xffffffffu, small);
        const unsigned bmask = __ballot_sync(0xffffffffu, act && !small && R.sp1 > R.sp0);
        if (small) {
          const int ord = __popc(smask & ((1u << lane) - 1u));
          T.wslot[ord % NCW][ord / NCW] = (signed char)lane;
        }
        {  // per consumer warp: can it take the one-chunk-per-lane fast path for all its columns?
          const unsigned slow = __ballot_sync(0xffffffffu, small && !(staged && nch <= 32));
          if (lane < NW) {
            unsigned mine = 0u;
            for (int o = lane; o < 32; o += NCW) mine |= 1u << o;   // ords of warp `lane`
            // slot bit -> ord: the ord-th set bit of smask
            bool ok = true;
            for (unsigned mm = slow; mm; mm &= mm - 1) {
              const int b = __ffs(mm) - 1;
              const int ord = __popc(smask & ((1u << b) - 1u));
              if ((mine >> ord) & 1u) ok = false;
            }
            T.wfast[lane] = ok ? 1 : 0;
          }
        }
        if (act) T.abit[__popc(mask & ((1u << lane) - 1u))] = (signed char)lane;
        int tot = staged ? nch : 0;
#pragma unroll
        for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        {  // flattened list of the kBigBlk-entry blocks of the big slots (ascending slot order)
          const bool big = (bmask >> lane) & 1u;
          const int nblk = big ? ((R.sp1 + kBigBlk - 1) >> kBigShift) - (R.sp0 >> kBigShift) : 0;
          int pb = nblk;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, pb, o);
            if (lane >= o) pb += up;
          }
          T.bigpre[lane] = pb - nblk;
          if (lane == 31) T.bigtot = pb;
        }
        if (lane == 0) {
          T.mask = mask;
          T.bigmask = bmask;
          T.win = R.gm.x;
          T.bytes = (unsigned)tot * (HASVAL ? 32u : 16u);
          T.nact = __popc(mask);
        }
        __syncwarp();
      };
      auto group_rec = [&](int r) { return __ldcg(reinterpret_cast<const int4 *>(groups + (r % ng))); };
      auto issue_copies = [&](int r, unsigned buf) {  // producer warp: stage local round r into buffer buf
        const RoundTab &T = ps.tab[r % 3];
        if (lane == 0) mbar_expect_tx(&ps.mbar[buf], T.bytes + 8192u);
        __syncwarp();
        const int4 rc = T.rec[lane];
        if (((T.mask >> lane) & 1u) && rc.z >= 0) {
          const int ch0 = rc.x >> 2, nch = ((rc.y + 3) >> 2) - ch0;
          bulk_g2s(&ps.sidx[buf][rc.z], a.colind + T.c0[lane] + 4 * (int64_t)ch0, (uint32_t)nch * 16u,
                   &ps.mbar[buf]);
          if (HASVAL)
            bulk_g2s(&ps.sval[buf][rc.z], a.colval + T.c0[lane] + 4 * (int64_t)ch0, (uint32_t)nch * 16u,
                     &ps.mbar[buf]);
        }
        if (lane == 0) bulk_g2s(&ps.G[buf][0][0], ca.wgram + (size_t)T.win * 1024, 8192u, &ps.mbar[buf]);
      };

      // prologue: tables of rounds 0 and 1, copies of round 0
      TabRegs treg;     // producer warp: coordinate lines of round r + 2 in flight
      int4 gm_ahead = make_int4(0, 0, 0, 0);  // window record of round r + 3
      if (warp == NW - 1) {
        tab_load(group_rec(0), treg);
        tab_store(0, treg);
        tab_load(group_rec(1), treg);
        tab_store(1, treg);
        issue_copies(0, rr & 1u);
        gm_ahead = group_rec(2);
      } else if (warp == 0) {
        ps.pcta[0][lane] = 0.0;
        ps.pcta[1][lane] = 0.0;
      }
      __syncthreads();

      bool done = false;
      int t = 0, r = 0;
#ifdef SLIM_PROFILE_ROUNDS
      long long prof[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
      long long tick_ = clock64();
#endif
      for (; t < maxit && !done; t++) {
        double dltx = 0.0;
        for (int g = 0; g < ng; g++, r++, rr++) {
          const unsigned buf = rr & 1u;
          const unsigned xb = xr & 1u;  // exchange buffer of the round
          const RoundTab &T = ps.tab[r % 3];
          const unsigned mask = T.mask;
          double xi = 0.0;
          if (warp == 0 && ((mask >> lane) & 1u)) xi = x[T.rec[lane].w];  // consumed by the solve
          if (warp == NW - 1) {
            issue_copies(r + 1, buf ^ 1u);  // next round's columns + Gram block: a full round to land
            tab_load(gm_ahead, treg);       // start reading the coordinate lines of round r + 2 ...
            gm_ahead = group_rec(r + 3);    // ... and the window record of round r + 3
          }
          mbar_wait(&ps.mbar[buf], (rr >> 1) & 1u);  // this round's staged data has landed
          SLIM_TICK(0);

          // gather: staged small columns from shared memory, one consumer warp each; the rest from global
          {
            // slots of this warp (assigned by the producer when it built the table), three at a time
            static_assert((kClusterNT / 32 - 1) * kWarpSlots >= 32, "not enough consumer warp slots");
            SLIM_TICK(5);
            if (warp < NCW && T.wslot[warp][0] >= 0) {
              int myb[kWarpSlots];
              {
                const long long packed = *reinterpret_cast<const long long *>(&T.wslot[warp][0]);
#pragma unroll
                for (int k = 0; k < kWarpSlots; k++) myb[k] = (int)(signed char)((packed >> (8 * k)) & 0xff);
              }
              SLIM_TICK(5);
              if (T.wfast[warp]) {
                // fast path: every column is staged with at most one 16-byte chunk per lane -- all yhat
                // gathers (L2 latency) of all the warp's columns are issued before the first reduction
                double pv[kWarpSlots];
#pragma unroll
                for (int k = 0; k < kWarpSlots; k++) {
                  const int b = myb[k];
                  double v = 0.0;
                  if (b >= 0) {
                    const int4 rc = T.rec[b];
                    const int ch0 = rc.x >> 2, nch = ((rc.y + 3) >> 2) - ch0;
                    if (lane < nch) {
                      Chunk c;
                      c.ix = ps.sidx[buf][rc.z + lane];
                      if (HASVAL) c.vv = ps.sval[HASVAL ? buf : 0][HASVAL ? rc.z + lane : 0];
                      v = dot_chunk_r<HASVAL, true>(c, (ch0 + lane) * 4, rc.x, rc.y, yh);
                    }
                  }
                  pv[k] = v;
                }
#ifdef SLIM_PROFILE_ROUNDS
                if (pv[0] == 123.456) prof[11]++;  // force the gathers to complete before the tick
#endif
                SLIM_TICK(11);
#pragma unroll
                for (int o = 16; o; o >>= 1) {
#pragma unroll
                  for (int k = 0; k < kWarpSlots; k++) pv[k] += __shfl_xor_sync(0xffffffffu, pv[k], o);
                }
                if (lane == 0) {
#pragma unroll
                  for (int k = 0; k < kWarpSlots; k++)
                    if (myb[k] >= 0) ps.pcta[xb][myb[k]] = pv[k];
                }
              } else {
#pragma unroll
                for (int k = 0; k < kWarpSlots; k++) {
                  const int b = myb[k];
                  if (b < 0) continue;
                  const int4 rc = T.rec[b];
                  if (rc.y <= rc.x) continue;
                  const double v = rc.z >= 0 ? warp_dot_staged<HASVAL>(&ps.sidx[buf][rc.z], &ps.sval[HASVAL ? buf : 0][HASVAL ? rc.z : 0], rc.x, rc.y, yh)
                                             : warp_dot<HASVAL>(a, T.c0[b], rc.x, rc.y, yh);
                  if (lane == 0) ps.pcta[xb][b] = v;
                }
              }
            }
            SLIM_TICK(6);
            // big slots: their kBigBlk-entry blocks form ONE flattened list that the consumer warps stride over,
            // so the loads of different columns overlap and no barrier is paid per column; a warp flushes
            // its partial sum when it moves on to the next column
            if (warp < NCW && T.bigtot > 0) {
              ps.pwb[warp][lane] = 0.0;
              __syncwarp();
              unsigned mm = T.bigmask;
              int b = __ffs(mm) - 1;
              int pre = T.bigpre[b];
              int4 rc = T.rec[b];
              int nblk = ((rc.y + kBigBlk - 1) >> kBigShift) - (rc.x >> kBigShift);
              double acc = 0.0;
              for (int g = warp; g < T.bigtot; g += NCW) {
                if (g >= pre + nblk) {  // next column(s): flush the finished one
#pragma unroll
                  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                  if (lane == 0) ps.pwb[warp][b] = acc;
                  acc = 0.0;
                  do {
                    mm &= mm - 1;
                    b = __ffs(mm) - 1;
                    pre = T.bigpre[b];
                    rc = T.rec[b];
                    nblk = ((rc.y + kBigBlk - 1) >> kBigShift) - (rc.x >> kBigShift);
                  } while (g >= pre + nblk);
                }
                acc += warp_bigblock_dot<HASVAL>(a, T.c0[b], rc.x, rc.y, (rc.x & ~(kBigBlk - 1)) + (g - pre) * kBigBlk, yh);
              }
#pragma unroll
              for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
              if (lane == 0) ps.pwb[warp][b] = acc;
            }
          }
          SLIM_TICK(7);
          __syncthreads();
          if (T.bigtot > 0) {
            if (tid < 32 && ((T.bigmask >> tid) & 1u)) {
              double sum = 0.0;
#pragma unroll
              for (int w = 0; w < NCW; w++) sum += ps.pwb[w][tid];
              ps.pcta[xb][tid] = sum;
            }
            __syncthreads();
          }
          SLIM_TICK(1);

          {
            // all-gather of the 32 partials across the cluster: warp w stores this CTA's 32 values, each
            // tagged with the round number, into CTA w's pall[xb][rank][.] (coalesced DSMEM stores)
            for (int dst = warp; dst < cs; dst += NW)
              st_peer_tagged(&ps.pall[xb][rank][lane][0], (uint32_t)dst, ps.pcta[xb][lane], xr + 1u);
          }
          SLIM_TICK(8);

          // Gram-space sequential solve of the window by warp 0 (identical in every CTA)
          if (warp == 0) {
            const bool act = (mask >> lane) & 1u;
            const double aty = act ? (double)T.aty[lane] : 0.0;
            const double inv_den = act ? T.inv_den[lane] : 0.0;
            const double sqv = act ? T.sq[lane] : 0.0;
            const int nact = T.nact;
            ps.pcta[xb ^ 1u][lane] = 0.0;  // clear the other buffer for the next round
            double P = 0.0;  // poll until the tagged partials of all CTAs have landed (fixed summation order)
            for (int c = 0; c < cs; c++) P += ld_tagged_wait(&ps.pall[xb][c][lane][0], xr + 1u);
#ifdef SLIM_PROFILE_ROUNDS
            if (P == 123.456) prof[2]++;
#endif
            SLIM_TICK(2);
            const double in_old = fabs(xi) > kEps ? xi : 0.0;
            // With d_m = in_new_m - in_old_m the sequential inner product of slot k is
            //   ip_k = P_k - in_old_k*sq_k - sum_{m<k} in_old_m G[m][k] + sum_{m<k} in_new_m G[m][k].
            // The in_old sum does not depend on the chain: every lane accumulates it up front (throughput
            // bound); the dependent chain per slot is then DADD -> SHFL -> select -> DFMA.
            double ip = P - in_old * sqv;
            for (int k = 0; k < nact; k++) {
              const int m = T.abit[k];
              const double o = __shfl_sync(0xffffffffu, in_old, m);
              if (m < lane) ip = fma(-o, ps.G[buf][m][lane], ip);
            }
            const double cthr = aty - a.l1r;      // x' = max(cthr - ip, 0) * inv_den
            const double den = act ? 1.0 / inv_den : 1.0;
            const double qeps = kEps * den;       // |x'| > EPS  <=>  q > EPS * den   (q >= 0)
            double nx = act ? 0.0 : xi, d = act ? -in_old : 0.0;  // default: the slot ends at zero
            // With non-negative ratings G >= 0 and every step adds in_new >= 0 times a Gram entry, so ip can
            // only GROW along the chain: a slot whose numerator is already <= 0 ends at exactly zero and
            // contributes nothing to later slots.  Only the remaining candidates form the dependent chain
            // (after the first sweeps that is a small fraction of the window).
            unsigned cand = __ballot_sync(0xffffffffu, act && (ca.nonneg ? (cthr - ip > 0.0) : true));
            if (cand) {
              int mbc = __ffs(cand) - 1;
              double growc = ps.G[buf][mbc][lane];
              while (cand) {
                cand &= cand - 1;
                const int mb_next = cand ? __ffs(cand) - 1 : mbc;
                const double grow_next = ps.G[buf][mb_next][lane];
                // slot mbc: q = numerator of the update; broadcast q*inv_den when it enters yhat, else 0
                const double q = cthr - ip;
                const double xq = q > 0.0 ? q * inv_den : 0.0;
                const double inn = q > qeps ? xq : 0.0;
                const double in_new_b = __shfl_sync(0xffffffffu, inn, mbc);
                if (lane == mbc) {
                  nx = xq;
                  d = inn - in_old;
                }
                ip = fma(in_new_b, growc, ip);  // only later slots still use ip; the diagonal of G is zero
                mbc = mb_next;
                growc = grow_next;
              }
            }
            if (act) x[T.rec[lane].w] = nx;
            ps.dlt[lane] = act ? d : 0.0;
            double dd = act ? (nx - xi) * (nx - xi) : 0.0;
#pragma unroll
            for (int o = 16; o; o >>= 1) dd += __shfl_xor_sync(0xffffffffu, dd, o);
            if (lane == 0) ps.dl = dd;
            SLIM_TICK(9);
          }
          __syncthreads();
          SLIM_TICK(3);
          dltx += ps.dl;

          // update this CTA's yhat slice: yhat += sum_k d_k a_k  (fp64 atomics: columns may share users)
#ifdef SLIM_EXP_NOUPD
          if (false) {
#else
          if (warp < NCW) {
#endif
#pragma unroll
            for (int k = 0; k < kWarpSlots; k++) {
              const int b = T.wslot[warp][k];
              if (b < 0) break;
              const double d = ps.dlt[b];
              const int4 rc = T.rec[b];
              const int s0 = rc.x, s1 = rc.y;
              if (d != 0.0 && s1 > s0) {
                const int so = rc.z;
                if (so >= 0) warp_axpy_staged<HASVAL>(&ps.sidx[buf][so], &ps.sval[HASVAL ? buf : 0][HASVAL ? so : 0], s0, s1, d, yh);
                else warp_axpy<HASVAL>(a, T.c0[b], s0, s1, d, yh);
              }
            }
          }
          if (warp < NCW && T.bigtot > 0) {
            unsigned mm = T.bigmask;
            int b = __ffs(mm) - 1;
            int pre = T.bigpre[b];
            int4 rc = T.rec[b];
            int nblk = ((rc.y + kBigBlk - 1) >> kBigShift) - (rc.x >> kBigShift);
            double d = ps.dlt[b];
            for (int g = warp; g < T.bigtot; g += NCW) {
              while (g >= pre + nblk) {
                mm &= mm - 1;
                b = __ffs(mm) - 1;
                pre = T.bigpre[b];
                rc = T.rec[b];
                nblk = ((rc.y + kBigBlk - 1) >> kBigShift) - (rc.x >> kBigShift);
                d = ps.dlt[b];
              }
              if (d != 0.0) warp_bigblock_axpy<HASVAL>(a, T.c0[b], rc.x, rc.y, (rc.x & ~(kBigBlk - 1)) + (g - pre) * kBigBlk, d, yh);
            }
          }
          SLIM_TICK(10);
          if (warp == NW - 1) tab_store(r + 2, treg);  // ... consumed here, a whole round later
          __syncthreads();
          SLIM_TICK(4);
          xr++;
        }
        if (dltx < a.opttol) done = true;
      }
      niters = done ? t : maxit + 1;
      // drain the copies issued for the round that will not run, so the barrier phases stay in step
      mbar_wait(&ps.mbar[rr & 1u], (rr >> 1) & 1u);
      rr++;
      __syncthreads();
#ifdef SLIM_PROFILE_ROUNDS
      if (rank == 0 && (tid == 0 || tid == 32 * (NW - 1))) {
        const double rnds = (double)ng * t;
        printf("target %d %s na %d ng %d sw %d | wait %.0f setup %.0f gathers %.0f reduce %.0f bigloop %.0f sync1 %.0f xsend %.0f poll %.0f "
               "chain %.0f sync3 %.0f upd %.0f sync4 %.0f\n",
               j, tid == 0 ? "C" : "P", na, ng, t, prof[0] / rnds, prof[5] / rnds, prof[11] / rnds, prof[6] / rnds, prof[7] / rnds,
               prof[1] / rnds, prof[8] / rnds, prof[2] / rnds, prof[9] / rnds, prof[3] / rnds, prof[10] / rnds,
               prof[4] / rnds);
      }
#endif
    } else if (na > 0 && maxit > 0) {
      CoordView v_cur = load_view(&meta[0], pr0, pr1);
      CoordView v_nxt = load_view(&meta[na > 1 ? 1 : 0], pr0, pr1);
      Chunk c_cur;
      c_cur.ix = make_uint4(0, 0, 0, 0);
      c_cur.vv = make_float4(0.f, 0.f, 0.f, 0.f);
      {
        const int ch = (v_cur.s0 >> 2) + tid;
        if (ch < ((v_cur.s1 + 3) >> 2)) load_chunk<HASVAL>(a, v_cur.c0, ch, c_cur);
      }
      bool done = false;
      int t = 0;
      for (; t < maxit && !done; t++) {
        double dltx = 0.0;
        for (int p = 0; p < na; p++) {
          const double xi = x[p];
          int p2 = p + 2;
          p2 = p2 >= na ? p2 - na : p2;
          p2 = p2 >= na ? p2 - na : p2;
          if (p2 >= na) p2 = 0;
          const CoordView v_nn = load_view(&meta[p2], pr0, pr1);
          const int ch0 = v_cur.s0 >> 2, ch1 = (v_cur.s1 + 3) >> 2;

          double part = 0.0;
          if (ch0 + tid < ch1) part = dot_chunk_r<HASVAL>(c_cur, (ch0 + tid) * 4, v_cur.s0, v_cur.s1, yh);
          {
            int ch = ch0 + tid + NT;
            // 4 chunks in flight per thread: 4 x (16 B ids [+ 16 B values]) before the first gather
            for (; ch + 3 * NT < ch1; ch += 4 * NT) {
              Chunk c0, c1, c2, c3;
              load_chunk<HASVAL>(a, v_cur.c0, ch, c0);
              load_chunk<HASVAL>(a, v_cur.c0, ch + NT, c1);
              load_chunk<HASVAL>(a, v_cur.c0, ch + 2 * NT, c2);
              load_chunk<HASVAL>(a, v_cur.c0, ch + 3 * NT, c3);
              part += dot_chunk_r<HASVAL>(c0, ch * 4, v_cur.s0, v_cur.s1, yh);
              part += dot_chunk_r<HASVAL>(c1, (ch + NT) * 4, v_cur.s0, v_cur.s1, yh);
              part += dot_chunk_r<HASVAL>(c2, (ch + 2 * NT) * 4, v_cur.s0, v_cur.s1, yh);
              part += dot_chunk_r<HASVAL>(c3, (ch + 3 * NT) * 4, v_cur.s0, v_cur.s1, yh);
            }
            for (; ch < ch1; ch += NT) {
              Chunk c;
              load_chunk<HASVAL>(a, v_cur.c0, ch, c);
              part += dot_chunk_r<HASVAL>(c, ch * 4, v_cur.s0, v_cur.s1, yh);
            }
          }
          Chunk c_nxt;
          c_nxt.ix = make_uint4(0, 0, 0, 0);
          c_nxt.vv = make_float4(0.f, 0.f, 0.f, 0.f);
          {
            const int ch = (v_nxt.s0 >> 2) + tid;
            if (ch < ((v_nxt.s1 + 3) >> 2)) load_chunk<HASVAL>(a, v_nxt.c0, ch, c_nxt);
          }

          const double ipf = cluster_sum(part, sm, par, cl, cs, rank);

          const double in_old = fabs(xi) > kEps ? xi : 0.0;
          const double ip = ipf - in_old * v_cur.sq;
          const double num = (double)v_cur.aty - ip;
          const double nx = num > a.l1r ? (num - a.l1r) / v_cur.den : 0.0;
          const double in_new = fabs(nx) > kEps ? nx : 0.0;
          const double d = in_new - in_old;
          if (d != 0.0) {
            if (ch0 + tid < ch1) axpy_chunk_r<HASVAL>(c_cur, (ch0 + tid) * 4, v_cur.s0, v_cur.s1, d, yh);
            int ch = ch0 + tid + NT;
            for (; ch + 3 * NT < ch1; ch += 4 * NT) {
              Chunk c0, c1, c2, c3;
              load_chunk<HASVAL>(a, v_cur.c0, ch, c0);
              load_chunk<HASVAL>(a, v_cur.c0, ch + NT, c1);
              load_chunk<HASVAL>(a, v_cur.c0, ch + 2 * NT, c2);
              load_chunk<HASVAL>(a, v_cur.c0, ch + 3 * NT, c3);
              axpy_chunk_r<HASVAL>(c0, ch * 4, v_cur.s0, v_cur.s1, d, yh);
              axpy_chunk_r<HASVAL>(c1, (ch + NT) * 4, v_cur.s0, v_cur.s1, d, yh);
              axpy_chunk_r<HASVAL>(c2, (ch + 2 * NT) * 4, v_cur.s0, v_cur.s1, d, yh);
              axpy_chunk_r<HASVAL>(c3, (ch + 3 * NT) * 4, v_cur.s0, v_cur.s1, d, yh);
            }
            for (; ch < ch1; ch += NT) {
              Chunk c;
              load_chunk<HASVAL>(a, v_cur.c0, ch, c);
              axpy_chunk_r<HASVAL>(c, ch * 4, v_cur.s0, v_cur.s1, d, yh);
            }
          }
          if (tid == 0) x[p] = nx;
          dltx += (nx - xi) * (nx - xi);
          __syncthreads();  // this CTA's yhat slice and x[p] are private to the CTA: a CTA barrier is enough
          v_cur = v_nxt;
          v_nxt = v_nn;
          c_cur = c_nxt;
        }
        if (dltx < a.opttol) done = true;
      }
      niters = done ? t : maxit + 1;
    } else if (maxit > 0) {
      niters = (0.0 < a.opttol) ? 1 : maxit + 1;
    }

    const unsigned long long tm3 = globaltimer_ns();
    // ---- residual / objective over this CTA's user range, then cluster-reduced ------------------------
    double yy = 0.0, yd = 0.0;
    {
      const int32_t *sp = ca.colsplit + (size_t)j * (kParts + 1);
      const int s0 = sp[pr0], s1 = sp[pr1];
      for (int e = s0 + tid; e < s1; e += NT) {
        const int u = a.colind[cj0 + e];
        const double v = HASVAL ? (double)a.colval[cj0 + e] : 1.0;
        yy += v * v;
        yd += v * __ldcg(&yh[u]);
      }
    }
    __syncthreads();
    double hh = 0.0;
    for (int u = u_lo + tid; u < u_hi; u += NT) {
      const double h = __ldcg(&yh[u]);
      if (h != 0.0) {
        hh += h * h;
        yh[u] = 0.0;
      }
    }
    double reg = 0.0;
    int nnz_local = 0;
    if (rank == 0) {
      for (int p = tid; p < na; p += NT) {
        const double xv = x[p];
        reg += 0.5 * a.l2r * xv * xv + a.l1r * fabs(xv);
        nnz_local += fabs(xv) > kEps ? 1 : 0;
      }
    }
    yy = cluster_sum(yy, sm, par, cl, cs, rank);
    yd = cluster_sum(yd, sm, par, cl, cs, rank);
    hh = cluster_sum(hh, sm, par, cl, cs, rank);
    reg = cluster_sum(reg, sm, par, cl, cs, rank);
    const int nnz_w = (int)(cluster_sum((double)nnz_local, sm, par, cl, cs, rank) + 0.5);
    const double expand_t = cluster_sum((double)expand, sm, par, cl, cs, rank);
    const double actnnz_t = cluster_sum((double)actnnz, sm, par, cl, cs, rank);

    // ---- K3: compaction by CTA 0 -----------------------------------------------------------------------
    if (rank == 0) {
      if (tid == 0) sm.off = (long long)atomicAdd(a.pool_used, (unsigned long long)nnz_w);
      __syncthreads();
      const long long off = sm.off;
      const bool fits = off + nnz_w <= a.pool_cap;
      if (fits) {
        int w0 = 0;
        for (int base = 0; base < na; base += NT) {
          const int p = base + tid;
          double xv = 0.0;
          if (p < na) xv = x[p];
          const bool flag = (p < na) && fabs(xv) > kEps;
          int tot;
          const int pos = w0 + team_excl_scan<NT>(flag, sm.sc, tot);
          if (flag) {
            a.pool_idx[off + pos] = a.inv[act_idx[p]];  // original item id
            a.pool_val[off + pos] = (float)xv;
          }
          w0 += tot;
        }
      }
      if (tid == 0) {
        a.out_cnt[q] = fits ? nnz_w : -1 - nnz_w;
        a.out_off[q] = off;
        a.st_niters[q] = niters;
        a.st_nactive[q] = na;
        a.st_actnnz[q] = (long long)(actnnz_t + 0.5);
        a.st_expand[q] = (long long)(expand_t + 0.5);
        const double rn = 0.5 * (yy - 2.0 * yd + hh);
        a.st_rnorm[q] = rn;
        a.st_obj[q] = rn + reg;
        const unsigned long long tm4 = globaltimer_ns();
        a.st_phase[4 * q + 0] = (float)((tm1 - tm0) * 1e-3);
        a.st_phase[4 * q + 1] = (float)((tm2 - tm1) * 1e-3);
        a.st_phase[4 * q + 2] = (float)((tm3 - tm2) * 1e-3);
        a.st_phase[4 * q + 3] = (float)((tm4 - tm3) * 1e-3);
        a.st_ngroups[q] = WINDOW ? sm.ng : na;
      }
      __syncthreads();
    }
  }
}

#include "gram.cuh"
#include "gram_batch.cuh"
#include "hybrid.cuh"
static_assert(sizeof(HybLine) == sizeof(ActMetaC), "cd_hybrid_kernel reuses the scratch of cd_cluster_kernel's lines");
#include "fslim.cuh"
#include "predict.cuh"

// ------------------------------------------------------------------------------------------------
// K0g host side: decide whether G fits, pick its element type, build it (part of staging).
//   SLIMB200_GRAM=0        never build G (the user-space kernels are used)
//   SLIMB200_GRAM_GB=n     upper bound for G in GiB (default 150); G must also fit the free memory minus a reserve
//   SLIMB200_GRAM_F64=1    force double elements
// ------------------------------------------------------------------------------------------------
static int env_int(const char *name, int dflt);

static void build_gram(Matrix *m) {
  const int32_t ncols = m->ncols;
  if (ncols <= 0 || m->nnz <= 0 || !env_int("SLIMB200_GRAM", 1)) return;
  cudaStream_t s = m->stream;
  // Element type (gram.cuh): packed unsigned integers when every rating is a non-negative integer and no sum can
  // reach 2^24 (in-block tiles are held as float); fp64 otherwise.
  bool packed = true;
  int32_t rmax = 1;
  const bool kv = m->has_val && !m->unit;
  if (kv) {
    DevBuf<int32_t> d_flag;
    d_flag.alloc_zero(2, s);
    integer_values_kernel<<<grid_for(m->nnz, 256, m->sm_count), 256, 0, s>>>(m->d_rowval, m->nnz, d_flag.p);
    m->stage_launches++;
    int32_t flag[2] = {1, 0};
    CK(cudaMemcpyAsync(flag, d_flag.p, sizeof(flag), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    packed = flag[0] == 0;
    rmax = std::max(1, flag[1]);
  }
  std::vector<double> csq(ncols);
  CK(cudaMemcpyAsync(csq.data(), m->d_csq, sizeof(double) * ncols, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (packed) {
    double mx = 0.0;
    for (double v : csq) mx = std::max(mx, v);
    packed = mx < 16777216.0;  // |G[i][k]| <= sqrt(csq_i csq_k) <= max csq (Cauchy-Schwarz), also every partial sum
  }
  if (env_int("SLIMB200_GRAM_F64", 0)) packed = false;
  const size_t ld = ((size_t)ncols + 127) & ~size_t(127);  // item-indexed scratch arrays: whole panels
  const size_t npan = ld / kGramPW;
  // column ranges of the packed layout: G[k][i] <= rmax * sum_u r_ui <= rmax * csq_i (integers: r <= r^2); the
  // 16-bit range starts at the first panel from which every column's bound is <= 65535, the 8-bit range likewise
  size_t h32 = 0, h16 = 0;
  if (packed) {
    int32_t last16 = -1, last8 = -1;  // last column whose bound exceeds the 16-bit / 8-bit limit
    for (int32_t i = 0; i < ncols; i++) {
      const double bound = (double)rmax * csq[i];
      if (bound > 65535.0) last16 = i;
      if (bound > 255.0) last8 = i;
    }
    h32 = (size_t)((last16 + 1 + kGramPW - 1) / kGramPW) * kGramPW;
    h16 = (size_t)((last8 + 1 + kGramPW - 1) / kGramPW) * kGramPW;
    const int force = env_int("SLIMB200_GRAM_WIDTH", 0);  // tests: 4 = everything 32-bit, 2 = no 8-bit range
    if (force == 4) h32 = h16 = ld;
    if (force == 2) h16 = ld;
    h32 = std::min(h32, ld);
    h16 = std::min(std::max(h16, h32), ld);
  }
  const size_t nr = (size_t)ncols;
  const size_t off16 = (h32 / kGramPW) * nr * (kGramPW * 4);
  const size_t off8 = off16 + ((h16 - h32) / kGramPW) * nr * (kGramPW * 2);
  size_t bytes = packed ? off8 + ((ld - h16) / kGramPW) * nr * kGramPW + 16  // (+16: word loads at the very end)
                              : npan * nr * kGramPW * sizeof(double);
  size_t free_b = 0, total_b = 0;
  {
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, m->device) == cudaSuccess) (void)cudaMemPoolTrimTo(pool, 0);
    (void)cudaGetLastError();
  }
  CK(cudaMemGetInfo(&free_b, &total_b));
  // leave room for the solve scratch and the result pools
  const size_t reserve = std::min((size_t)16 << 30, total_b / 4);
  const size_t budget = std::min((size_t)env_int("SLIMB200_GRAM_GB", 150) << 30, free_b > reserve ? free_b - reserve : 0);
  // Full matrix if it fits; otherwise -- packed elements only -- the STAIR layout (gram.cuh: GaStair), half the bytes:
  // panel p keeps rows [0, max(64 (p + 1), hd)).  SLIMB200_GRAM_LAYOUT=stair forces it (tests), =full forbids it;
  // SLIMB200_GRAM_HD = side of the full square of the most popular items (default 32768).
  const char *lay = getenv("SLIMB200_GRAM_LAYOUT");
  const bool force_stair = lay && !strcmp(lay, "stair"), no_stair = lay && !strcmp(lay, "full");
  bool stair = false;
  size_t hd = 0;
  std::vector<unsigned long long> pbase;
  if (packed && !no_stair && (force_stair || bytes > budget)) {
    hd = (size_t)std::max(64, env_int("SLIMB200_GRAM_HD", 32768));
    hd = std::min((hd + kGramPW - 1) / kGramPW * kGramPW, ld);
    pbase.resize(npan + 1);
    size_t o = 0;
    for (size_t pnl = 0; pnl < npan; pnl++) {
      pbase[pnl] = o;
      const size_t rows = std::min(nr, std::max((pnl + 1) * kGramPW, hd));
      const size_t w = pnl * kGramPW < h32 ? 4 : (pnl * kGramPW < h16 ? 2 : 1);
      o += rows * kGramPW * w;
    }
    pbase[npan] = o;
    bytes = o + 16;
    stair = true;
  }
  if (bytes > budget) {
    if (env_int("SLIMB200_VERBOSE", 0))
      fprintf(stderr, "[slim-b200] Gram matrix (%.1f GB%s) does not fit (budget %.1f GB): user-space kernels will be used\n",
              bytes / 1e9, stair ? ", stair layout" : "", budget / 1e9);
    return;
  }
  EventPair ev;
  cudaEvent_t g0 = ev.a, g1 = ev.b;
  CK(cudaEventRecord(g0, s));
  if (cudaMalloc(&m->d_gram, bytes) != cudaSuccess) {  // not fatal: fall back to the user-space kernels
    (void)cudaGetLastError();
    m->d_gram = nullptr;
    return;
  }
  CK(cudaMemsetAsync(m->d_gram, 0, bytes, s));
  CK(cudaMalloc(&m->d_expand, sizeof(unsigned long long) * ncols));
  CK(cudaMemsetAsync(m->d_expand, 0, sizeof(unsigned long long) * ncols, s));
  m->gram_ld = ld;
  m->gram_f64 = !packed;
  m->gram_bytes = bytes;
  m->gram_h32 = (int32_t)h32;
  m->gram_h16 = (int32_t)h16;
  m->gram_off16 = off16;
  m->gram_off8 = off8;
  GramView gv{static_cast<const unsigned char *>(m->d_gram), nr, m->gram_h32, m->gram_h16, off16, off8, nullptr, 0};
  if (stair) {
    CK(cudaMalloc(&m->d_gram_pbase, sizeof(unsigned long long) * pbase.size()));
    CK(cudaMemcpyAsync(m->d_gram_pbase, pbase.data(), sizeof(unsigned long long) * pbase.size(), cudaMemcpyHostToDevice, s));
    m->gram_stair = true;
    m->gram_hd = (int32_t)hd;
    m->h_gram_pbase = pbase;
    gv.pbase = m->d_gram_pbase;
    gv.hd = m->gram_hd;
  }
  // work items (column, entry range), heaviest columns first (internal ids are in popularity order)
  constexpr int32_t kSeg = 2048;
  std::vector<int32_t> wc, w0, w1;
  for (int32_t k = 0; k < ncols; k++)
    for (int32_t e = 0; e < m->h_colcnt[k]; e += kSeg) {
      wc.push_back(k);
      w0.push_back(e);
      w1.push_back(std::min(m->h_colcnt[k], e + kSeg));
    }
  const int32_t nwork = (int32_t)wc.size();
  DevBuf<int32_t> d_wc, d_w0, d_w1;
  d_wc.alloc(nwork);
  d_w0.alloc(nwork);
  d_w1.alloc(nwork);
  if (nwork > 0) {
    CK(cudaMemcpyAsync(d_wc.p, wc.data(), sizeof(int32_t) * nwork, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_w0.p, w0.data(), sizeof(int32_t) * nwork, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_w1.p, w1.data(), sizeof(int32_t) * nwork, cudaMemcpyHostToDevice, s));
    const int grid = std::max(1, std::min(nwork, m->sm_count * 8));
#define SLIM_GRAM_BUILD(GB, HV)                                                                               \
  gram_build_kernel<GB, HV><<<grid, 256, 0, s>>>(nwork, d_wc.p, d_w0.p, d_w1.p, m->d_colptr, m->d_colind,      \
                                                 m->d_colval, m->d_rowptr, m->d_rowind, m->d_rowval, gv, m->d_expand)
    if (stair) {
      if (kv) SLIM_GRAM_BUILD(GbStair, true); else SLIM_GRAM_BUILD(GbStair, false);
    } else if (packed) {
      if (kv) SLIM_GRAM_BUILD(GbPacked, true); else SLIM_GRAM_BUILD(GbPacked, false);
    } else {
      if (kv) SLIM_GRAM_BUILD(GbF64, true); else SLIM_GRAM_BUILD(GbF64, false);
    }
#undef SLIM_GRAM_BUILD
    CK(cudaGetLastError());
    m->stage_launches++;
  }
  CK(cudaEventRecord(g1, s));
  CK(cudaStreamSynchronize(s));  // the work list is released at scope exit
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, g0, g1));
  m->gram_ms = ms;
  if (env_int("SLIMB200_VERBOSE", 0))
    fprintf(stderr, "[slim-b200] Gram matrix: %d x %zu %s, %.2f GB (32-bit columns %zu, 16-bit %zu, 8-bit %zu), %d work "
                    "items, built in %.1f ms\n", ncols, ld,
            stair ? "packed unsigned (exact), STAIR layout" : packed ? "packed unsigned (exact)" : "fp64", bytes / 1e9,
            packed ? h32 : 0, packed ? h16 - h32 : 0, packed ? ld - h16 : 0, nwork, ms);
}

// K3 (second half): ordered gather of the solved columns into compact CSC arrays (the CSC
// assembly of SaveModel, estimate.c:570-588).  One warp per column.
constexpr int kMaxPools = 10;
struct PoolTable {
  const int32_t *idx[kMaxPools];
  const float *val[kMaxPools];
};

__global__ void gather_columns_kernel(int32_t nsel, const int64_t *__restrict__ src_off,
                                      const int32_t *__restrict__ pool_of,
                                      const int64_t *__restrict__ dst_off, const PoolTable pools,
                                      int32_t *out_idx, float *out_val, int32_t *out_counts) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < nsel; c += warps) {
    const int64_t s = src_off[c], d = dst_off[c];
    const int n = (int)(dst_off[c + 1] - d);
    const int32_t *pi = pools.idx[pool_of[c]];
    const float *pv = pools.val[pool_of[c]];
    for (int e = lane; e < n; e += 32) {
      out_idx[d + e] = pi[s + e];
      out_val[d + e] = pv[s + e];
    }
    if (lane == 0) out_counts[c] = n;
  }
}

// ------------------------------------------------------------------------------------------------
// host side of learn()
// ------------------------------------------------------------------------------------------------
struct Result {
  int device = 0;  // a Result may outlive the Matrix it came from: it keeps no pointer to it
  int32_t nsel = 0;
  int64_t nnz = 0;
  // compact, caller-ordered CSC of the solved columns (device)
  int64_t *d_colptr = nullptr;
  int32_t *d_counts = nullptr;
  int32_t *d_colind = nullptr;
  float *d_colval = nullptr;
  std::vector<int64_t> h_colptr;
  std::vector<int32_t> niters, nactive;
  std::vector<int64_t> actnnz, expand;
  std::vector<double> rnorm, obj;
  std::vector<float> phase;      // [nsel][4] microseconds (cluster kernel; zeros otherwise)
  std::vector<int32_t> ngroups;  // sync rounds per sweep
  Timings tm{};
};

void free_result(Result *r) {
  if (!r) return;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(r->device);
  cudaFree(r->d_colptr);
  cudaFree(r->d_counts);
  cudaFree(r->d_colind);
  cudaFree(r->d_colval);
  if (prev >= 0) cudaSetDevice(prev);
  delete r;
}

void result_info(const Result *r, int32_t *nsel, int64_t *nnz, Timings *t) {
  if (nsel) *nsel = r->nsel;
  if (nnz) *nnz = r->nnz;
  if (t) *t = r->tm;
}

int result_stats(const Result *r, int32_t *niters, int32_t *nactive, int64_t *active_nnz,
                 int64_t *expand_nnz, double *rnorm, double *objval) {
  const size_t n = (size_t)r->nsel;
  if (niters) memcpy(niters, r->niters.data(), n * sizeof(int32_t));
  if (nactive) memcpy(nactive, r->nactive.data(), n * sizeof(int32_t));
  if (active_nnz) memcpy(active_nnz, r->actnnz.data(), n * sizeof(int64_t));
  if (expand_nnz) memcpy(expand_nnz, r->expand.data(), n * sizeof(int64_t));
  if (rnorm) memcpy(rnorm, r->rnorm.data(), n * sizeof(double));
  if (objval) memcpy(objval, r->obj.data(), n * sizeof(double));
  return kOk;
}

int result_phases(const Result *r, float *phase_us, int32_t *ngroups) {
  if (phase_us) memcpy(phase_us, r->phase.data(), sizeof(float) * r->phase.size());
  if (ngroups) memcpy(ngroups, r->ngroups.data(), sizeof(int32_t) * r->ngroups.size());
  return kOk;
}

int result_to_host(const Result *r, int64_t *colptr, int32_t *colind, float *colval) {
  try {
    DeviceGuard guard(r->device);
    memcpy(colptr, r->h_colptr.data(), sizeof(int64_t) * ((size_t)r->nsel + 1));
    if (r->nnz > 0) {  // learn() synchronised its stream before returning: plain copies are safe
      CK(cudaMemcpy(colind, r->d_colind, sizeof(int32_t) * r->nnz, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(colval, r->d_colval, sizeof(float) * r->nnz, cudaMemcpyDeviceToHost));
    }
    return kOk;
  } catch (const EngineError &e) {
    g_last_error = e.what();
    return e.status;
  }
}

int result_to_device(const Result *r, int32_t *d_counts, int32_t *d_colind, float *d_colval) {
  try {
    DeviceGuard guard(r->device);
    if (d_counts && r->nsel > 0)
      CK(cudaMemcpy(d_counts, r->d_counts, sizeof(int32_t) * r->nsel, cudaMemcpyDeviceToDevice));
    if (r->nnz > 0) {
      if (d_colind) CK(cudaMemcpy(d_colind, r->d_colind, sizeof(int32_t) * r->nnz, cudaMemcpyDeviceToDevice));
      if (d_colval) CK(cudaMemcpy(d_colval, r->d_colval, sizeof(float) * r->nnz, cudaMemcpyDeviceToDevice));
    }
    CK(cudaDeviceSynchronize());
    return kOk;
  } catch (const EngineError &e) {
    g_last_error = e.what();
    return e.status;
  }
}

#include "gather.cuh"

struct LaunchPlan {
  int nt;
  bool ysmem;
  size_t smem;
  int grid;
};

template <int NT, bool YSMEM, bool HASVAL>
static void launch_solve(const SolveArgs &args, const LaunchPlan &plan, cudaStream_t s, bool query_only,
                         int *blocks_per_sm) {
  auto kern = cd_solve_kernel<NT, YSMEM, HASVAL>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
  if (query_only) {
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kern, NT, plan.smem));
    return;
  }
  kern<<<plan.grid, NT, plan.smem, s>>>(args);
  CK(cudaGetLastError());
}

static void dispatch_solve(const SolveArgs &args, const LaunchPlan &plan, bool hasval, cudaStream_t s,
                           bool query_only, int *bps) {
#define SLIM_DISPATCH(NTV)                                                                   \
  if (plan.nt == NTV) {                                                                      \
    if (plan.ysmem) {                                                                        \
      if (hasval) launch_solve<NTV, true, true>(args, plan, s, query_only, bps);             \
      else launch_solve<NTV, true, false>(args, plan, s, query_only, bps);                   \
    } else {                                                                                 \
      if (hasval) launch_solve<NTV, false, true>(args, plan, s, query_only, bps);            \
      else launch_solve<NTV, false, false>(args, plan, s, query_only, bps);                  \
    }                                                                                        \
    return;                                                                                  \
  }
  SLIM_DISPATCH(32)
  SLIM_DISPATCH(128)
  SLIM_DISPATCH(512)
#undef SLIM_DISPATCH
  throw EngineError(kErr, "dispatch_solve: unsupported team size");
}

template <bool HASVAL, bool WINDOW>
static int cluster_launch(const SolveArgs &args, const ClusterArgs &cargs, int cs, int nclusters, cudaStream_t s,
                          bool query_only) {
  auto kern = cd_cluster_kernel<HASVAL, WINDOW>;
  const size_t dyn = WINDOW ? sizeof(PipeSmem<HASVAL>) : 0;
  if (cs > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  if (dyn) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3((unsigned)(cs * std::max(nclusters, 1)), 1, 1);
  cfg.blockDim = dim3(kClusterNT, 1, 1);
  cfg.dynamicSmemBytes = dyn;
  cfg.stream = s;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (query_only) {
    int n = 0;
    CK(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    return n;
  }
  CK(cudaLaunchKernelEx(&cfg, kern, args, cargs));
  return 0;
}

static int cluster_dispatch(bool vals, bool window, const SolveArgs &args, const ClusterArgs &cargs, int cs,
                            int nclusters, cudaStream_t s, bool query_only) {
  if (vals) {
    return window ? cluster_launch<true, true>(args, cargs, cs, nclusters, s, query_only)
                  : cluster_launch<true, false>(args, cargs, cs, nclusters, s, query_only);
  }
  return window ? cluster_launch<false, true>(args, cargs, cs, nclusters, s, query_only)
                : cluster_launch<false, false>(args, cargs, cs, nclusters, s, query_only);
}

// Launch (or, with query_only, size) cd_gram_kernel<GT, CS>.  Query: CTAs per SM for CS == 1, co-resident
// clusters on the device for CS > 1.  `count` = CTAs (CS == 1) or clusters (CS > 1) to launch.
template <typename GA, int CS, int UNR>
static int gram_launch_t(const SolveArgs &args, const GramArgs &gargs, int count, cudaStream_t s, bool query_only) {
  auto kern = cd_gram_kernel<GA, CS, UNR>;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3((unsigned)(CS * std::max(count, 1)), 1, 1);
  cfg.blockDim = dim3(kGramNT, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cfg.attrs = attr;
  cfg.numAttrs = CS > 1 ? 1 : 0;
  if (CS > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  if (query_only) {
    int n = 0;
    if (CS == 1) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, kGramNT, 0));
    else CK(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    return n;
  }
  CK(cudaLaunchKernelEx(&cfg, kern, args, gargs));
  return 0;
}

static int gram_launch(bool f64, int cs, const SolveArgs &args, const GramArgs &gargs, int count, cudaStream_t s,
                       bool query_only) {
  if (gargs.gv.pbase) {  // stair layout: single CTAs and clusters of 4
    if (cs == 1) return gram_launch_t<GaStair, 1, 16>(args, gargs, count, s, query_only);
    if (cs == 4) return gram_launch_t<GaStair, 4, 16>(args, gargs, count, s, query_only);
    throw EngineError(kErr, "gram_launch: the stair layout supports cluster sizes 1 and 4");
  }
  // the deep-unroll variant exists for the packed layout and the two default shapes (single CTA, clusters of 4)
  const bool deep = !f64 && (cs == 1 || cs == 4) && env_int("SLIMB200_GRAM_UNROLL", 32) == 32;
  if (deep) {
    return cs == 1 ? gram_launch_t<GaPacked, 1, 32>(args, gargs, count, s, query_only)
                   : gram_launch_t<GaPacked, 4, 32>(args, gargs, count, s, query_only);
  }
#define SLIM_GRAM_CS(CSV)                                                                        \
  if (cs == CSV)                                                                                 \
    return f64 ? gram_launch_t<GaF64, CSV, 16>(args, gargs, count, s, query_only)                \
               : gram_launch_t<GaPacked, CSV, 16>(args, gargs, count, s, query_only);
  SLIM_GRAM_CS(1)
  SLIM_GRAM_CS(2)
  SLIM_GRAM_CS(4)
  SLIM_GRAM_CS(8)
  SLIM_GRAM_CS(16)
#undef SLIM_GRAM_CS
  throw EngineError(kErr, "gram_launch: unsupported cluster size");
}

// cd_hybrid_kernel (hybrid.cuh): one cluster of `cs` CTAs x 512 threads per giant target; query = co-resident clusters
template <typename GA, bool HV>
static int hybrid_launch_t(const SolveArgs &args, const HybridArgs &hargs, int cs, int count, cudaStream_t s, bool query_only) {
  auto kern = cd_hybrid_kernel<GA, HV>;
  const size_t dyn = sizeof(HybSmem<GA, HV>);
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  if (cs > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3((unsigned)(cs * std::max(count, 1)), 1, 1);
  cfg.blockDim = dim3(kHybNT, 1, 1);
  cfg.dynamicSmemBytes = dyn;
  cfg.stream = s;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (query_only) {
    int n = 0;
    CK(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    return n;
  }
  CK(cudaLaunchKernelEx(&cfg, kern, args, hargs));
  return 0;
}

static int hybrid_launch(bool stair, bool vals, const SolveArgs &args, const HybridArgs &hargs, int cs, int count,
                         cudaStream_t s, bool query_only) {
  if (stair) return vals ? hybrid_launch_t<GaStair, true>(args, hargs, cs, count, s, query_only)
                         : hybrid_launch_t<GaStair, false>(args, hargs, cs, count, s, query_only);
  return vals ? hybrid_launch_t<GaPacked, true>(args, hargs, cs, count, s, query_only)
              : hybrid_launch_t<GaPacked, false>(args, hargs, cs, count, s, query_only);
}

constexpr int kBatchT = 8, kBatchV = 2;

template <typename GA, int CS, int NTB>
static int batch_launch_t(const SolveArgs &args, const GramArgs &gargs, const BatchArgs &bargs, int count,
                          cudaStream_t s, bool query_only) {
  auto kern = bargs.use_mma ? cd_gram_batch_kernel<GA, CS, kBatchT, kBatchV, NTB, true>
                            : cd_gram_batch_kernel<GA, CS, kBatchT, kBatchV, NTB, false>;
  const size_t dyn = sizeof(BatchSmem<GA, CS, kBatchT, kBatchV, NTB>);
  constexpr int kBatchNT = NTB;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  if (CS > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3((unsigned)(CS * std::max(count, 1)), 1, 1);
  cfg.blockDim = dim3(kBatchNT, 1, 1);
  cfg.dynamicSmemBytes = dyn;
  cfg.stream = s;
  cfg.attrs = attr;
  cfg.numAttrs = CS > 1 ? 1 : 0;
  if (query_only) {
    int n = 0;
    if (CS == 1) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, kBatchNT, dyn));
    else CK(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    return n;
  }
  CK(cudaLaunchKernelEx(&cfg, kern, args, gargs, bargs));
  return 0;
}

// cd_gram_batch_kernel: T targets per cluster of `cs` CTAs.  Query: co-resident clusters (CTAs per SM for cs == 1).
static int batch_launch(bool f64, int cs, int nt, const SolveArgs &args, const GramArgs &gargs, const BatchArgs &bargs,
                        int count, cudaStream_t s, bool query_only) {
  // fp64 G: the staging ring of 16 warps would not fit shared memory, 256 threads only
#define SLIM_BATCH_CS(CSV)                                                                         \
  if (cs == CSV) {                                                                                 \
    if (f64) return batch_launch_t<GaF64, CSV, 256>(args, gargs, bargs, count, s, query_only);     \
    return nt == 512 ? batch_launch_t<GaPacked, CSV, 512>(args, gargs, bargs, count, s, query_only) \
                     : batch_launch_t<GaPacked, CSV, 256>(args, gargs, bargs, count, s, query_only); \
  }
  SLIM_BATCH_CS(1)
  SLIM_BATCH_CS(4)
  SLIM_BATCH_CS(8)
  SLIM_BATCH_CS(16)
#undef SLIM_BATCH_CS
  throw EngineError(kErr, "batch_launch: unsupported cluster size");
}

static size_t smem_for(int nt, bool ysmem, int32_t nrows) {
  size_t fixed = nt == 32 ? solve_fixed_smem<32>() : nt == 128 ? solve_fixed_smem<128>() : solve_fixed_smem<512>();
  fixed = (fixed + 15) & ~size_t(15);
  return fixed + (ysmem ? sizeof(double) * (size_t)nrows : 0);
}

static int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

Result *learn(Matrix *m, const LearnParams &p, const int32_t *cols, int32_t nsel_in,
              const WarmStart *warm, int32_t *status) {
  Result *res = nullptr;
  // SLIMB200_VERBOSE=2: host wall-clock split of the call (plan + scratch, solve, gather, release)
  auto wall = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double w_start = wall();
  double w_plan = 0, w_solved = 0, w_gathered = 0;
  try {
    std::lock_guard<std::mutex> one_call_at_a_time(m->learn_mutex);
    DeviceGuard guard(m->device);
    (void)cudaGetLastError();
    cudaStream_t s = m->stream;
    AsyncAllocScope alloc_scope(s);  // temporaries of this call: stream-ordered pool (see DevBuf)
    const int32_t ncols = m->ncols, nrows = m->nrows;
    const int32_t nsel = cols ? nsel_in : ncols;
    if (nsel < 0) throw EngineError(kErrInput, "learn: negative column count");
    for (int32_t q = 0; cols && q < nsel; q++)
      if (cols[q] < 0 || cols[q] >= ncols) throw EngineError(kErrInput, "learn: column id out of range");

    res = new Result();
    res->device = m->device;
    res->nsel = nsel;
    res->h_colptr.assign((size_t)nsel + 1, 0);
    res->niters.assign(nsel, 0);
    res->nactive.assign(nsel, 0);
    res->actnnz.assign(nsel, 0);
    res->expand.assign(nsel, 0);
    res->rnorm.assign(nsel, 0.0);
    res->obj.assign(nsel, 0.0);
    res->phase.assign((size_t)nsel * 4, 0.f);
    res->ngroups.assign(nsel, 0);

    // processing order: heaviest target column first (longest-processing-time-first on the queue)
    std::vector<int32_t> order(nsel);
    std::iota(order.begin(), order.end(), 0);
    auto colof = [&](int32_t q) { return cols ? cols[q] : q; };
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) {
      return m->h_colcnt[m->h_rank[colof(x)]] > m->h_colcnt[m->h_rank[colof(y)]];
    });

    // ---- launch plan ---------------------------------------------------------------------------
    // Gram-space solver whenever G was staged (gram.cuh); otherwise the user-space kernels below
    const bool use_gram = m->d_gram != nullptr && env_int("SLIMB200_GRAM", 1) != 0;
    const bool fslim = p.nnbrs > 0;
    if (fslim && !use_gram)
      throw EngineError(kErrInput, "learn: fSLIM (nnbrs > 0) needs the Gram-space solver, but no Gram matrix is staged "
                                   "for this training matrix (too many items for HBM, or SLIMB200_GRAM=0)");
    if (fslim && (p.simtype < 0 || p.simtype > 2)) throw EngineError(kErrInput, "learn: unknown simtype");
    const bool kernel_vals = m->has_val && !m->unit;
    LaunchPlan plan{};
    const double mean_col = ncols > 0 ? (double)m->nnz / ncols : 0.0;
    plan.nt = mean_col <= 128.0 ? 32 : (mean_col <= 4096.0 ? 128 : 512);
    plan.nt = env_int("SLIMB200_NT", plan.nt);
    if (plan.nt != 32 && plan.nt != 128 && plan.nt != 512) plan.nt = 128;
    const size_t smem_limit = (size_t)m->smem_optin;
    plan.ysmem = smem_for(plan.nt, true, nrows) <= smem_limit;
    if (env_int("SLIMB200_YHAT_GLOBAL", 0)) plan.ysmem = false;
    if (plan.ysmem) {
      // a large smem yhat leaves few CTAs per SM: widen the team to keep the SM busy
      const size_t per = smem_for(plan.nt, true, nrows);
      const size_t fit = (size_t)228 * 1024 / (per + 1024);
      if (!getenv("SLIMB200_NT")) {
        if (fit < 4) plan.nt = std::max(plan.nt, 512);
        else if (fit < 16) plan.nt = std::max(plan.nt, 128);
      }
    }
    plan.smem = smem_for(plan.nt, plan.ysmem, nrows);
    SolveArgs args{};
    ClusterArgs cargs{};

    // yhat too large for shared memory: thread-block clusters with yhat resident in L2
    int cs = env_int("SLIMB200_CLUSTER", plan.ysmem ? 0 : 16);
    if (cs != 0 && cs != 1 && cs != 2 && cs != 4 && cs != 8 && cs != 16) cs = 16;
    // GIANT targets next to a resident packed Gram matrix: in the STAIR layout (build_gram) a Gram-space sweep of a
    // target with >= ~5 000 nonzeros gathers |S| x |A| ~ 10^9 scattered elements, more than the nnz(R) entries a
    // user-space sweep streams, so the heaviest targets go to cd_hybrid_kernel (from 9 000 nonzeros: the 16-CTA
    // clusters of the hybrid kernel are the scarce resource of a step, the one-target Gram clusters fill the other
    // SMs; 4 096-column C5 steps: 5 721 / 5 278 / 5 112 / 5 628 ms for thresholds 5 000 / 7 000 / 9 000 / 12 000) (hybrid.cuh: user-space inner products, exact
    // block solve with Gram tiles), launched side by side with the Gram classes.  SLIMB200_STAIR_USER sets the
    // threshold for the stair layout, SLIMB200_HYBRID_MIN for any packed layout (default: never for the full layout,
    // where the batched kernel is faster); SLIMB200_GIANT_KERNEL=cluster runs them on cd_cluster_kernel instead.
    const bool stair = use_gram && m->gram_stair;
    const int stair_user = (use_gram && !m->gram_f64 && !fslim)
                               ? env_int("SLIMB200_HYBRID_MIN", stair ? env_int("SLIMB200_STAIR_USER", 9000) : INT32_MAX)
                               : INT32_MAX;
    int32_t n_user = 0;
    if (stair_user != INT32_MAX)
      for (int32_t q = 0; q < nsel; q++) n_user += m->h_colcnt[m->h_rank[colof(q)]] >= stair_user ? 1 : 0;
    const bool mixed = n_user > 0;
    const char *gk = getenv("SLIMB200_GIANT_KERNEL");
    const bool giant_hybrid = mixed && !(gk && !strcmp(gk, "cluster"));
    if (mixed && cs == 0) cs = 16;
    if (giant_hybrid) {
      cs = env_int("SLIMB200_HYBRID_CS", 16);
      if (cs != 1 && cs != 2 && cs != 4 && cs != 8 && cs != 16) cs = 16;
    }
    if (use_gram && !mixed) cs = 0;
    const bool use_cluster = cs > 0;
    const bool use_window = use_cluster && env_int("SLIMB200_WINDOW", 1) != 0;
    const size_t col_stride = use_gram ? m->gram_ld : (((size_t)std::max(ncols, 1) + 3) & ~size_t(3));
    const size_t row_stride = ((size_t)std::max(nrows, 1) + 3) & ~size_t(3);
    int nclusters = 0;
    HybridArgs hargs{};
    if (use_cluster) {
      int hw = giant_hybrid ? hybrid_launch(stair, kernel_vals, args, hargs, cs, 1, s, true)
                            : cluster_dispatch(kernel_vals, use_window, args, cargs, cs, 1, s, true);
      if (hw < 1) throw EngineError(kErr, "learn: cluster launch configuration not supported on this device");
      // keep the yhat vectors of all clusters in flight inside L2 (default budget 96 MB of 126 MB)
      const size_t l2_budget = (size_t)env_int("SLIMB200_L2_MB", 96) << 20;
      const int by_l2 = (int)std::max<size_t>(1, l2_budget / (row_stride * sizeof(double)));
      // ... but never fewer than 8 clusters: with very many users (C5: 40 MB per yhat) part of the gathers
      // then comes from HBM, which still beats leaving most SMs idle
      nclusters = giant_hybrid ? hw : std::min(hw, std::max(by_l2, 8));
      if (env_int("SLIMB200_NCLUSTERS", 0) > 0) nclusters = env_int("SLIMB200_NCLUSTERS", 0);
      nclusters = std::max(1, std::min(std::min(nclusters, hw), std::max(mixed ? n_user : nsel, 1)));
      plan.grid = nclusters * cs;
      if (env_int("SLIMB200_VERBOSE", 0))
        fprintf(stderr, "[slim-b200] %s: cluster=%d CTAs x %d threads, %d clusters in flight (hw max %d, "
                        "L2 budget allows %d), values=%d, window sweep=%d\n", giant_hybrid ? "hybrid kernel" : "cluster kernel",
                cs, giant_hybrid ? kHybNT : kClusterNT, nclusters, hw, by_l2, (int)kernel_vals, (int)use_window);
    }
    const int user_grid = use_cluster ? nclusters * cs : 0;  // CTAs of the user-space cluster launch
    if (!use_cluster && !use_gram) {
      int bps = 1;
      dispatch_solve(args, plan, kernel_vals, s, true, &bps);
      bps = std::max(1, bps);
      const int max_ctas = env_int("SLIMB200_CTAS_PER_SM", bps) * m->sm_count;
      plan.grid = std::max(1, std::min<int>(nsel, std::min(max_ctas, bps * m->sm_count)));
      if (env_int("SLIMB200_VERBOSE", 0))
        fprintf(stderr, "[slim-b200] team kernel: %d threads per target, yhat in %s, %d CTAs (%d per SM), values=%d\n",
                plan.nt, plan.ysmem ? "smem" : "global", plan.grid, bps, (int)kernel_vals);
    }

    // Gram-space plan: target classes by column nnz (the target list is sorted by descending nnz), one launch
    // per class, all launches side by side on their own streams:
    //   nnz >= gram_top   : cd_gram_batch_kernel, 8 targets per cluster; when there are only a few such batches,
    //                       16 CTAs x 512 threads per batch (shortest time per batch -- they set the makespan)
    //   nnz >= gram_batch : cd_gram_batch_kernel, clusters of 8 CTAs x 256 threads, two CTAs per SM (most work
    //                       per SM-second when there are many batches; empty by default, see below)
    //   nnz >= gram_heavy : cd_gram_kernel on clusters of gram_cs CTAs (position-space blocks)
    //   the rest          : cd_gram_kernel, one CTA per target
    GramArgs gargs{};
    BatchArgs bargs{};
    bargs.use_mma = env_int("SLIMB200_BATCH_MMA", 1) != 0;
    struct GramClass {
      bool batch;
      int cs, nt;      // CTAs per cluster, threads per CTA
      int min_nnz;     // class = targets with min_nnz <= nnz < min_nnz of the previous class
      int ntargets;    // of this call
      int units;       // clusters (CTAs for cs == 1) to launch at most
      int slot_base;   // first scratch slot (CTA granularity)
    };
    std::vector<GramClass> classes;
    if (use_gram) {
      int gram_cs = env_int("SLIMB200_GRAM_CS", 4);
      if (gram_cs != 1 && gram_cs != 2 && gram_cs != 4 && gram_cs != 8 && gram_cs != 16) gram_cs = 4;
      // fSLIM targets have at most nnbrs active coordinates: one CTA each
      const int gram_heavy = fslim ? INT32_MAX : env_int("SLIMB200_GRAM_HEAVY", 2000);
      // measured on C4 (profiles/r02_plan_sweeps.txt): with the packed Gram matrix and the DMMA gather a batch of 8
      // targets beats eight one-target clusters from ~9 000 nonzeros on (30 000 in round 1: scalar DFMAs, fp32 G);
      // the one-target clusters are bound by the DRAM's random-sector rate, the batches read whole panel rows and
      // share them between their targets.  16 CTAs x 512 threads per batch is the best shape at every batch count
      // (8 x 256 with two CTAs per SM was measured 15-45 % slower); the second batch class stays for experiments.
      const int gram_batch = fslim ? INT32_MAX : std::max(gram_heavy, env_int("SLIMB200_GRAM_BATCH", 9000));
      const int gram_top = fslim ? INT32_MAX : std::max(gram_batch, env_int("SLIMB200_GRAM_TOP", 9000));
      auto batch_cs_of = [&](int dflt) {
        const int v = env_int("SLIMB200_BATCH_CS", dflt);
        return (v == 1 || v == 4 || v == 8 || v == 16) ? v : dflt;
      };
      auto batch_nt_of = [&](int dflt) {
        return (!m->gram_f64 && env_int("SLIMB200_BATCH_NT", dflt) == 512) ? 512 : 256;
      };
      if (stair) {
        gargs.gv.pbase = m->d_gram_pbase;  // (gram_launch picks the accessor from it, also for the occupancy queries)
        if (gram_cs != 1) gram_cs = 4;
      } else {
        classes.push_back({true, batch_cs_of(16), batch_nt_of(512), gram_top, 0, 0, 0});
        classes.push_back({true, batch_cs_of(8), batch_nt_of(256), gram_batch, 0, 0, 0});
      }
      if (gram_cs > 1) classes.push_back({false, gram_cs, kGramNT, gram_heavy, 0, 0, 0});
      classes.push_back({false, 1, kGramNT, INT32_MIN, 0, 0, 0});
      for (int32_t q = 0; q < nsel; q++) {
        const int32_t c = m->h_colcnt[m->h_rank[colof(q)]];
        if (c >= stair_user) continue;  // user-space cluster kernel
        for (auto &gc : classes)
          if (c >= gc.min_nnz) {
            gc.ntargets++;
            break;
          }
      }
      int slots = 0;
      for (auto &gc : classes) {
        if (gc.ntargets == 0) {  // nothing to launch for this class: no occupancy query either
          gc.units = 0;
          gc.slot_base = slots;
          continue;
        }
        int hw = gc.batch ? batch_launch(m->gram_f64, gc.cs, gc.nt, args, gargs, bargs, 1, s, true)
                          : gram_launch(m->gram_f64, gc.cs, args, gargs, 1, s, true);
        if (gc.cs == 1) hw *= m->sm_count;
        if (hw < 1) throw EngineError(kErr, "learn: Gram kernel launch configuration not supported on this device");
        const int need = gc.batch ? (gc.ntargets + kBatchT - 1) / kBatchT : gc.ntargets;
        gc.units = std::min(hw, need);
        // the batches are compute-bound (DMMA) and own their SMs; the one-target classes are memory-bound and need
        // only a few SMs to keep the DRAM busy: SLIMB200_BATCH_UNITS caps the batches in flight so both overlap
        if (gc.batch && env_int("SLIMB200_BATCH_UNITS", 0) > 0) gc.units = std::min(gc.units, env_int("SLIMB200_BATCH_UNITS", 0));
        gc.slot_base = slots;
        slots += gc.units * gc.cs;
        if (env_int("SLIMB200_VERBOSE", 0) && gc.ntargets > 0)
          fprintf(stderr, "[slim-b200] Gram-space (%s G): %d targets with nnz >= %d -> %s, %d x (%d CTAs x %d threads)\n",
                  m->gram_f64 ? "fp64" : (stair ? "packed, stair layout" : "packed"), gc.ntargets, gc.min_nnz,
                  gc.batch ? "cd_gram_batch_kernel (8 targets per cluster)" : "cd_gram_kernel", gc.units, gc.cs, gc.nt);
      }
      plan.grid = std::max(1, slots);
      if (mixed && env_int("SLIMB200_VERBOSE", 0))
        fprintf(stderr, "[slim-b200] %d giant targets with nnz >= %d -> %s\n", n_user, stair_user,
                giant_hybrid ? "cd_hybrid_kernel" : "cd_cluster_kernel");
    }

    // ---- scratch (cached on the matrix) --------------------------------------------------------
    // user-space slots first (clusters, or CTAs of the team kernel), then the Gram slots (CTAs)
    const size_t gu = use_cluster ? (size_t)nclusters : (use_gram ? 0 : (size_t)plan.grid);
    const size_t gg = use_gram ? (size_t)plan.grid : 0;
    const size_t gux = use_cluster ? (size_t)user_grid : gu;  // the cluster kernel keeps one copy of x per CTA
    size_t off = 0;
    auto carve = [&](size_t bytes) {
      size_t o = off;
      off += (bytes + 255) & ~size_t(255);
      return o;
    };
    const size_t o_acc = carve(gu * col_stride * sizeof(double));
    const size_t o_xw = carve((gu + gg) * col_stride * sizeof(float));
    const size_t o_yh = carve((plan.ysmem && !use_cluster) ? 0 : gu * row_stride * sizeof(double));
    const size_t zero_bytes = off;  // acc, xw, yhat must start at zero
    const size_t o_meta = carve(gu * col_stride * (use_cluster ? sizeof(ActMetaC) : sizeof(ActMeta)));
    const size_t o_x = carve((gux + gg) * col_stride * sizeof(double));
    const size_t o_idx = carve((gu + gg) * col_stride * sizeof(int32_t));
    const size_t o_gslot = carve(gg * col_stride * sizeof(int32_t));
    const size_t o_grow = carve(gg * col_stride * sizeof(int32_t));
    const size_t o_gval = carve(gg * col_stride * sizeof(double));
    size_t nbcta = 0;  // the batch classes come first: their CTAs use slots 0 .. nbcta-1
    for (const auto &gc : classes)
      if (gc.batch) nbcta += (size_t)gc.units * gc.cs;
    const size_t nwords = col_stride / 32;
    const size_t o_bxt = carve(nbcta * kBatchT * col_stride * sizeof(double));
    const size_t o_bsv = carve(nbcta * kBatchT * col_stride * sizeof(double));
    const size_t o_bam = carve(nbcta * kBatchT * nwords * sizeof(uint32_t));
    const size_t o_ban = carve(nbcta * nwords * sizeof(uint32_t));
    const size_t grp_stride = (size_t)(ncols + 31) / 32 + 1;
    const size_t o_grp = carve(use_cluster ? gu * grp_stride * sizeof(GroupMeta) : 0);
    if (off > m->scratch_bytes) {
      cudaFree(m->d_scratch);
      m->d_scratch = nullptr;
      m->scratch_bytes = 0;
      CK(cudaMalloc(&m->d_scratch, off));
      m->scratch_bytes = off;
    }
    CK(cudaMemsetAsync(m->d_scratch, 0, zero_bytes, s));
    unsigned char *sb = static_cast<unsigned char *>(m->d_scratch);

    // ---- warm start ----------------------------------------------------------------------------
    DevBuf<int64_t> d_wptr;
    DevBuf<int32_t> d_wind;
    DevBuf<float> d_wval;
    if (warm && warm->colptr) {
      const int64_t wnnz = warm->colptr[warm->ncols];
      d_wptr.alloc((size_t)warm->ncols + 1);
      d_wind.alloc(wnnz);
      d_wval.alloc(wnnz);
      CK(cudaMemcpyAsync(d_wptr.p, warm->colptr, sizeof(int64_t) * ((size_t)warm->ncols + 1),
                         cudaMemcpyHostToDevice, s));
      if (wnnz > 0) {
        CK(cudaMemcpyAsync(d_wind.p, warm->colind, sizeof(int32_t) * wnnz, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(d_wval.p, warm->colval, sizeof(float) * wnnz, cudaMemcpyHostToDevice, s));
      }
    }

    args.nrows = nrows;
    args.ncols = ncols;
    args.rowptr = m->d_rowptr;
    args.rowind = m->d_rowind;
    args.rowval = m->d_rowval;
    args.colptr = m->d_colptr;
    args.colcnt = m->d_colcnt;
    args.colind = m->d_colind;
    args.colval = m->d_colval;
    args.cnorms = m->d_cnorms;
    args.csq = m->d_csq;
    args.rank = m->d_rank;
    args.inv = m->d_inv;
    args.l1r = p.l1r;
    args.l2r = p.l2r;
    args.opttol = p.opttol;
    args.maxniters = p.maxniters;
    args.wcolptr = d_wptr.p;
    args.wcolind = d_wind.p;
    args.wcolval = d_wval.p;
    args.wncols = warm && warm->colptr ? warm->ncols : 0;
    args.acc = reinterpret_cast<double *>(sb + o_acc);
    args.xw = reinterpret_cast<float *>(sb + o_xw);
    args.yhat = (plan.ysmem && !use_cluster) ? nullptr : reinterpret_cast<double *>(sb + o_yh);
    args.act_meta = reinterpret_cast<ActMeta *>(sb + o_meta);
    args.x = reinterpret_cast<double *>(sb + o_x);
    cargs.colsplit = m->d_colsplit;
    cargs.meta = reinterpret_cast<ActMetaC *>(sb + o_meta);
    cargs.xc = reinterpret_cast<double *>(sb + o_x);
    cargs.rows_per_part = m->rows_per_part;
    cargs.wgram = m->d_wgram;
    cargs.nonneg = m->nonneg ? 1 : 0;
    cargs.groups = reinterpret_cast<GroupMeta *>(sb + o_grp);
    cargs.grp_stride = grp_stride;
    hargs.colsplit = m->d_colsplit;
    hargs.rows_per_part = m->rows_per_part;
    hargs.xc = reinterpret_cast<double *>(sb + o_x);
    hargs.lines = reinterpret_cast<HybLine *>(sb + o_meta);  // (the slot of cd_cluster_kernel's ActMetaC lines: same size)
    hargs.expand = m->d_expand;
    args.act_idx = reinterpret_cast<int32_t *>(sb + o_idx);
    if (use_gram) {
      gargs.gv = GramView{static_cast<const unsigned char *>(m->d_gram), (size_t)ncols, m->gram_h32, m->gram_h16,
                          m->gram_off16, m->gram_off8, stair ? m->d_gram_pbase : nullptr, stair ? m->gram_hd : 0};
      hargs.gv = gargs.gv;
      gargs.act = reinterpret_cast<int32_t *>(sb + o_idx) + gu * col_stride;
      gargs.x = reinterpret_cast<double *>(sb + o_x) + gux * col_stride;
      gargs.slotp = reinterpret_cast<int32_t *>(sb + o_gslot);
      gargs.sl_row = reinterpret_cast<int32_t *>(sb + o_grow);
      gargs.sl_val = reinterpret_cast<double *>(sb + o_gval);
      gargs.expand = m->d_expand;
      bargs.xt = reinterpret_cast<double *>(sb + o_bxt);
      bargs.sl_valT = reinterpret_cast<double *>(sb + o_bsv);
      bargs.amask = reinterpret_cast<uint32_t *>(sb + o_bam);
      bargs.anym = reinterpret_cast<uint32_t *>(sb + o_ban);
      bargs.istride = col_stride;
      bargs.nwords = (int32_t)nwords;
      bargs.profile = env_int("SLIMB200_PROFILE", 0);
    }
    args.col_stride = col_stride;
    args.row_stride = row_stride;

    cudaEvent_t e0, e1, e2;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventCreate(&e2));

    w_plan = wall();
    // ---- solve, retrying the (rare) columns that did not fit the output pool --------------------
    std::vector<int32_t> pending = order;  // indices into the caller's column list
    std::vector<int64_t> src_off(nsel, 0);
    std::vector<int32_t> cnt(nsel, 0);
    struct Pool {
      int32_t *idx;
      float *val;
      bool pooled_idx, pooled_val;
    };
    std::vector<Pool> pools;
    std::vector<int32_t> pool_of(nsel, 0);
    int64_t cap = std::max<int64_t>(1 << 20, (int64_t)nsel * std::min<int32_t>(std::max(ncols, 1), 2048));
    double solve_ms = 0.0;
    auto free_pools = [&]() {
      for (auto &pl : pools) {
        dev_free(pl.idx, pl.pooled_idx, s);
        dev_free(pl.val, pl.pooled_val, s);
      }
      pools.clear();
    };
    try {
      for (int round = 0; !pending.empty(); round++) {
        if (round >= kMaxPools) throw EngineError(kErr, "learn: output pool retry limit");
        const int32_t nt = (int32_t)pending.size();
        std::vector<int32_t> tcols(nt);
        for (int32_t k = 0; k < nt; k++) tcols[k] = m->h_rank[colof(pending[k])];  // internal ids
        DevBuf<int32_t> d_targets, d_queue, d_ocnt, d_nit, d_nact;
        DevBuf<int64_t> d_ooff, d_an, d_ex;
        DevBuf<double> d_rn, d_ob;
        DevBuf<float> d_ph;
        DevBuf<int32_t> d_ng;
        DevBuf<unsigned long long> d_used, d_hprof;
        int32_t n_hprof = 0;
        d_targets.alloc(nt);
        d_queue.alloc_zero(8, s);
        d_used.alloc_zero(1, s);
        d_ocnt.alloc(nt);
        d_ooff.alloc(nt);
        d_nit.alloc(nt);
        d_nact.alloc(nt);
        d_an.alloc(nt);
        d_ex.alloc(nt);
        d_rn.alloc(nt);
        d_ob.alloc(nt);
        d_ph.alloc_zero((size_t)nt * 4, s);
        d_ng.alloc_zero(nt, s);
        Pool pl{nullptr, nullptr, false, false};
        CK(dev_malloc(reinterpret_cast<void **>(&pl.idx), sizeof(int32_t) * cap, &pl.pooled_idx));
        pools.push_back(pl);
        CK(dev_malloc(reinterpret_cast<void **>(&pools.back().val), sizeof(float) * cap, &pools.back().pooled_val));
        CK(cudaMemcpyAsync(d_targets.p, tcols.data(), sizeof(int32_t) * nt, cudaMemcpyHostToDevice, s));
        args.targets = d_targets.p;
        args.ntargets = nt;
        args.queue = d_queue.p;
        args.out_cnt = d_ocnt.p;
        args.out_off = d_ooff.p;
        args.pool_idx = pools.back().idx;
        args.pool_val = pools.back().val;
        args.pool_used = d_used.p;
        args.pool_cap = cap;
        args.st_niters = d_nit.p;
        args.st_nactive = d_nact.p;
        args.st_actnnz = d_an.p;
        args.st_expand = d_ex.p;
        args.st_rnorm = d_rn.p;
        args.st_obj = d_ob.p;
        args.st_phase = d_ph.p;
        args.st_ngroups = d_ng.p;
        LaunchPlan lp = plan;
        lp.grid = std::min(plan.grid, nt);
        // fSLIM: neighbour lists of this round's targets (fslim.cuh), consumed by cd_gram_kernel
        DevBuf<int32_t> d_nbr, d_nbrcnt, d_nq, d_ncv;
        DevBuf<uint32_t> d_nfirst;
        DevBuf<float> d_ndot, d_nck;
        args.nnbrs = 0;
        args.nbr_list = nullptr;
        args.nbr_cnt = nullptr;
        int nbr_grid = 0;
        if (fslim) {
          nbr_grid = std::max(1, std::min<int>(nt, m->sm_count * 4));
          d_nbr.alloc((size_t)nt * p.nnbrs);
          d_nbrcnt.alloc_zero(nt, s);
          d_nq.alloc_zero(1, s);
          d_nfirst.alloc((size_t)nbr_grid * col_stride);
          CK(cudaMemsetAsync(d_nfirst.p, 0xff, sizeof(uint32_t) * (size_t)nbr_grid * col_stride, s));
          d_ndot.alloc_zero((size_t)nbr_grid * col_stride, s);
          d_ncv.alloc((size_t)nbr_grid * col_stride);
          d_nck.alloc((size_t)nbr_grid * col_stride);
          args.nnbrs = p.nnbrs;
          args.nbr_list = d_nbr.p;
          args.nbr_cnt = d_nbrcnt.p;
        }
        CK(cudaEventRecord(e0, s));
        if (fslim) {
          NbrArgs nb{};
          nb.nnbrs = p.nnbrs;
          nb.simtype = p.simtype;
          nb.queue = d_nq.p;
          nb.first = d_nfirst.p;
          nb.dot = d_ndot.p;
          nb.cand_val = d_ncv.p;
          nb.cand_key = d_nck.p;
          nb.out_nbr = d_nbr.p;
          nb.out_cnt = d_nbrcnt.p;
          if (kernel_vals) fslim_neighbors_kernel<true><<<nbr_grid, kNbrNT, 0, s>>>(args, nb);
          else fslim_neighbors_kernel<false><<<nbr_grid, kNbrNT, 0, s>>>(args, nb);
          CK(cudaGetLastError());
          res->tm.launches++;
          res->tm.solve_launches++;
        }
        if (use_gram) {
          // `pending` is sorted by descending column nnz: the classes are consecutive ranges
          cudaStream_t streams[5] = {s, m->stream2, m->stream3, m->stream4, m->stream5};
          int used = 0;
          // fork point: the side streams wait for the copies / memsets queued on s so far, NOT for the
          // launches that follow on s (the classes must overlap)
          CK(cudaEventRecord(e2, s));
          auto next_stream = [&]() {
            cudaStream_t st = streams[used];
            if (used > 0) CK(cudaStreamWaitEvent(st, e2, 0));
            used++;
            return st;
          };
          int32_t q_next = 0;
          hargs.prof = nullptr;
          if (mixed) {
            // the giants at the head of the list go to cd_hybrid_kernel (or the user-space cluster kernel)
            while (q_next < nt && m->h_colcnt[tcols[q_next]] >= stair_user) q_next++;
            if (q_next > 0) {
              SolveArgs ua = args;
              ua.ntargets = q_next;
              ua.queue = d_queue.p + 7;
              const int ncl = std::min(nclusters, (int)q_next);
              if (giant_hybrid && env_int("SLIMB200_PROFILE", 0)) {
                d_hprof.alloc_zero((size_t)q_next * 10, s);
                hargs.prof = d_hprof.p;
                n_hprof = q_next;
              }
              if (giant_hybrid) hybrid_launch(stair, kernel_vals, ua, hargs, cs, ncl, next_stream(), false);
              else cluster_dispatch(kernel_vals, use_window, ua, cargs, cs, ncl, next_stream(), false);
            }
          }
          // the Gram kernels index their scratch slots from the end of the user-space slots
          SolveArgs ga_args = args;
          ga_args.xw = args.xw + gu * col_stride;
          int ci = 0;
          for (const auto &gc : classes) {
            int32_t q_end = q_next;
            while (q_end < nt && m->h_colcnt[tcols[q_end]] >= gc.min_nnz) q_end++;
            const int32_t ncl_targets = q_end - q_next;
            if (ncl_targets > 0) {
              GramArgs gq = gargs;
              gq.q_begin = q_next;
              gq.q_end = q_end;
              gq.queue = d_queue.p + ci;
              gq.slot_base = gc.slot_base;
              const int need = gc.batch ? (ncl_targets + kBatchT - 1) / kBatchT : ncl_targets;
              const int units = std::max(1, std::min(gc.units, need));
              if (gc.batch) batch_launch(m->gram_f64, gc.cs, gc.nt, ga_args, gq, bargs, units, next_stream(), false);
              else gram_launch(m->gram_f64, gc.cs, ga_args, gq, units, next_stream(), false);
            }
            q_next = q_end;
            ci++;
          }
          for (int k = 1; k < used; k++) {  // join the side streams
            CK(cudaEventRecord(e2, streams[k]));
            CK(cudaStreamWaitEvent(s, e2, 0));
            res->tm.launches++;  // the common accounting below adds the first launch
            res->tm.solve_launches++;
          }
          CK(cudaGetLastError());
        } else if (use_cluster) {
          const int ncl = std::min(nclusters, nt);
          cluster_dispatch(kernel_vals, use_window, args, cargs, cs, ncl, s, false);
          CK(cudaGetLastError());
        } else {
          dispatch_solve(args, lp, kernel_vals, s, false, nullptr);
        }
        CK(cudaEventRecord(e1, s));
        res->tm.launches++;
        res->tm.solve_launches++;
        std::vector<int32_t> h_cnt(nt), h_nit(nt), h_nact(nt);
        std::vector<int64_t> h_off(nt), h_an(nt), h_ex(nt);
        std::vector<double> h_rn(nt), h_ob(nt);
        std::vector<float> h_ph((size_t)nt * 4);
        std::vector<int32_t> h_ng(nt);
        CK(cudaMemcpyAsync(h_ph.data(), d_ph.p, sizeof(float) * nt * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(h_ng.data(), d_ng.p, sizeof(int32_t) * nt, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(h_cnt.data(), d_ocnt.p, sizeof(int32_t) * nt, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(h_off.data(), d_ooff.p, sizeof(int64_t) * nt, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(h_nit.data(), d_nit.p, sizeof(int32_t) * nt, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(h_nact.data(), d_nact.p, sizeof(int32_t) * nt, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(h_an.data(), d_an.p, sizeof(int64_t) * nt, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(h_ex.data(), d_ex.p, sizeof(int64_t) * nt, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(h_rn.data(), d_rn.p, sizeof(double) * nt, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(h_ob.data(), d_ob.p, sizeof(double) * nt, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (n_hprof > 0) {  // SLIMB200_PROFILE=1: cycle split of the hybrid kernel's rounds, per giant target
          std::vector<unsigned long long> hp((size_t)n_hprof * 10);
          CK(cudaMemcpy(hp.data(), d_hprof.p, sizeof(unsigned long long) * hp.size(), cudaMemcpyDeviceToHost));
          for (int32_t k = 0; k < n_hprof; k++) {
            const unsigned long long *pp = &hp[(size_t)k * 10];
            fprintf(stderr, "[slim-b200] hybrid target %d (nnz %d): short rounds %llu: dots %.2f xchg %.2f chain %.2f upd %.2f us "
                            "each | long rounds %llu: dots %.1f xchg %.1f chain %.1f upd %.1f us each (at 1.9 GHz)\n",
                    k, m->h_colcnt[tcols[k]], pp[8], pp[0] / 1.9e3 / std::max<double>(pp[8], 1), pp[1] / 1.9e3 / std::max<double>(pp[8], 1),
                    pp[2] / 1.9e3 / std::max<double>(pp[8], 1), pp[3] / 1.9e3 / std::max<double>(pp[8], 1), pp[9],
                    pp[4] / 1.9e3 / std::max<double>(pp[9], 1), pp[5] / 1.9e3 / std::max<double>(pp[9], 1),
                    pp[6] / 1.9e3 / std::max<double>(pp[9], 1), pp[7] / 1.9e3 / std::max<double>(pp[9], 1));
          }
        }
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        solve_ms += ms;
        std::vector<int32_t> again;
        int64_t need = 0;
        for (int32_t k = 0; k < nt; k++) {
          const int32_t q = pending[k];
          res->niters[q] = h_nit[k];
          res->nactive[q] = h_nact[k];
          res->actnnz[q] = h_an[k];
          res->expand[q] = h_ex[k];
          res->rnorm[q] = h_rn[k];
          res->obj[q] = h_ob[k];
          for (int z = 0; z < 4; z++) res->phase[(size_t)q * 4 + z] = h_ph[(size_t)k * 4 + z];
          res->ngroups[q] = h_ng[k];
          if (h_cnt[k] >= 0) {
            cnt[q] = h_cnt[k];
            src_off[q] = h_off[k];
            pool_of[q] = (int)pools.size() - 1;
          } else {
            again.push_back(q);
            need += -1 - h_cnt[k];
          }
        }
        pending.swap(again);
        cap = std::max<int64_t>(need + 1024, 1 << 20);
      }
      res->tm.solve_ms = solve_ms;
      w_solved = wall();

      // ---- ordered gather into compact CSC (caller's column order) ------------------------------
      for (int32_t q = 0; q < nsel; q++) res->h_colptr[q + 1] = res->h_colptr[q] + cnt[q];
      res->nnz = res->h_colptr[nsel];
      {  // (pool memory as well; free_result() releases it with cudaFree, which accepts both kinds)
        bool pooled_ignored = false;
        CK(dev_malloc(reinterpret_cast<void **>(&res->d_colptr), sizeof(int64_t) * ((size_t)nsel + 1), &pooled_ignored));
        CK(dev_malloc(reinterpret_cast<void **>(&res->d_counts), sizeof(int32_t) * std::max(nsel, 1), &pooled_ignored));
        CK(dev_malloc(reinterpret_cast<void **>(&res->d_colind), sizeof(int32_t) * std::max<int64_t>(res->nnz, 1), &pooled_ignored));
        CK(dev_malloc(reinterpret_cast<void **>(&res->d_colval), sizeof(float) * std::max<int64_t>(res->nnz, 1), &pooled_ignored));
      }
      CK(cudaMemcpyAsync(res->d_colptr, res->h_colptr.data(), sizeof(int64_t) * ((size_t)nsel + 1),
                         cudaMemcpyHostToDevice, s));
      CK(cudaEventRecord(e1, s));
      if (nsel > 0) {
        PoolTable tab{};
        for (size_t pi = 0; pi < pools.size(); pi++) {
          tab.idx[pi] = pools[pi].idx;
          tab.val[pi] = pools[pi].val;
        }
        DevBuf<int64_t> d_so;
        DevBuf<int32_t> d_po;
        d_so.alloc(nsel);
        d_po.alloc(nsel);
        CK(cudaMemcpyAsync(d_so.p, src_off.data(), sizeof(int64_t) * nsel, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(d_po.p, pool_of.data(), sizeof(int32_t) * nsel, cudaMemcpyHostToDevice, s));
        // the kernels emit each column in INTERNAL id order; the model wants ascending original ids
        // (estimate.c:497-503): gather, then sort every column segment by its (original) row ids
        DevBuf<int32_t> d_ti;
        DevBuf<float> d_tv;
        d_ti.alloc(res->nnz);
        d_tv.alloc(res->nnz);
        gather_columns_kernel<<<grid_for((int64_t)nsel * 32, 256, m->sm_count), 256, 0, s>>>(
            nsel, d_so.p, d_po.p, res->d_colptr, tab, d_ti.p, d_tv.p, res->d_counts);
        CK(cudaGetLastError());
        res->tm.launches++;
        if (res->nnz > 0) {
          size_t tmp_bytes = 0;
          CK(cub::DeviceSegmentedSort::SortPairs(nullptr, tmp_bytes, d_ti.p, res->d_colind, d_tv.p, res->d_colval,
                                                 res->nnz, (int64_t)nsel, res->d_colptr, res->d_colptr + 1, s));
          DevBuf<unsigned char> d_tmp;
          d_tmp.alloc(tmp_bytes);
          CK(cub::DeviceSegmentedSort::SortPairs(d_tmp.p, tmp_bytes, d_ti.p, res->d_colind, d_tv.p, res->d_colval,
                                                 res->nnz, (int64_t)nsel, res->d_colptr, res->d_colptr + 1, s));
          res->tm.launches += 2;
          CK(cudaStreamSynchronize(s));
        }
        CK(cudaStreamSynchronize(s));
      }
      CK(cudaEventRecord(e2, s));
      CK(cudaStreamSynchronize(s));
      float gms = 0.f;
      CK(cudaEventElapsedTime(&gms, e1, e2));
      res->tm.gather_ms = gms;
    } catch (...) {
      free_pools();
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
      cudaEventDestroy(e2);
      throw;
    }
    w_gathered = wall();
    free_pools();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaEventDestroy(e2);
    if (env_int("SLIMB200_VERBOSE", 0) >= 2)
      fprintf(stderr, "[slim-b200] learn(%d columns): host wall %.1f ms = plan+scratch %.1f | solve %.1f (kernels %.1f) | gather %.1f | "
                      "release %.1f\n", nsel, wall() - w_start, w_plan - w_start, w_solved - w_plan, solve_ms, w_gathered - w_solved,
              wall() - w_gathered);
    if (status) *status = kOk;
    return res;
  } catch (const EngineError &e) {
    g_last_error = e.what();
    if (status) *status = e.status;
  } catch (const std::exception &e) {
    g_last_error = e.what();
    if (status) *status = kErrMemory;
  }
  free_result(res);
  return nullptr;
}

// ------------------------------------------------------------------------------------------------
// host side of the batched top-N (predict.cuh)
// ------------------------------------------------------------------------------------------------
int predict_topn(int device, int32_t wrows, int32_t wcols, const ssize_t *wrowptr, const int32_t *wrowind,
                 const float *wrowval, int32_t nusers, const ssize_t *urowptr, const int32_t *urowind,
                 const float *urowval, int32_t nrcmds, int32_t *out_ids, float *out_scores, int32_t *out_counts,
                 double *kernel_ms) {
  try {
    if (device < 0 || device >= device_count()) throw EngineError(kErr, "predict_topn: no usable CUDA device");
    if (wrows < 0 || wcols < 0 || nusers < 0 || nrcmds <= 0 || !wrowptr || !urowptr)
      throw EngineError(kErrInput, "predict_topn: bad arguments");
    DeviceGuard guard(device);
    (void)cudaGetLastError();
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    const int64_t wnnz = wrowptr[wrows], unnz = urowptr[nusers];
    static_assert(sizeof(ssize_t) == sizeof(int64_t), "LP64 only");
    DevBuf<int64_t> d_wp, d_up;
    DevBuf<int32_t> d_wi, d_ui, d_mark, d_cand, d_oid, d_cnt;
    DevBuf<float> d_wv, d_uv, d_score, d_osc;
    d_wp.alloc((size_t)wrows + 1);
    d_up.alloc((size_t)nusers + 1);
    d_wi.alloc(wnnz);
    d_wv.alloc(wnnz);
    d_ui.alloc(unnz);
    if (urowval) d_uv.alloc(unnz);
    CK(cudaMemcpy(d_wp.p, wrowptr, sizeof(int64_t) * ((size_t)wrows + 1), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_up.p, urowptr, sizeof(int64_t) * ((size_t)nusers + 1), cudaMemcpyHostToDevice));
    if (wnnz > 0) {
      CK(cudaMemcpy(d_wi.p, wrowind, sizeof(int32_t) * wnnz, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(d_wv.p, wrowval, sizeof(float) * wnnz, cudaMemcpyHostToDevice));
    }
    if (unnz > 0) {
      CK(cudaMemcpy(d_ui.p, urowind, sizeof(int32_t) * unnz, cudaMemcpyHostToDevice));
      if (urowval) CK(cudaMemcpy(d_uv.p, urowval, sizeof(float) * unnz, cudaMemcpyHostToDevice));
    }
    const size_t nout = (size_t)nusers * nrcmds;
    d_oid.alloc(nout);
    d_osc.alloc(nout);
    d_cnt.alloc_zero(nusers, nullptr);
    if (nout > 0) {  // rows shorter than nrcmds keep the caller's content
      CK(cudaMemcpy(d_oid.p, out_ids, sizeof(int32_t) * nout, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(d_osc.p, out_scores, sizeof(float) * nout, cudaMemcpyHostToDevice));
    }
    const int grid = std::max(1, std::min(nusers, prop.multiProcessorCount * 8));
    const size_t stride = ((size_t)std::max(wcols, 1) + 31) & ~size_t(31);
    d_score.alloc_zero((size_t)grid * stride, nullptr);
    d_mark.alloc_zero((size_t)grid * stride, nullptr);
    d_cand.alloc((size_t)grid * stride);
    PredictArgs pa{};
    pa.nusers = nusers;
    pa.wrows = wrows;
    pa.wcols = wcols;
    pa.nrcmds = nrcmds;
    pa.wrowptr = d_wp.p;
    pa.wrowind = d_wi.p;
    pa.wrowval = d_wv.p;
    pa.urowptr = d_up.p;
    pa.urowind = d_ui.p;
    pa.urowval = urowval ? d_uv.p : nullptr;
    pa.score = d_score.p;
    pa.mark = d_mark.p;
    pa.cand = d_cand.p;
    pa.stride = stride;
    pa.out_ids = d_oid.p;
    pa.out_scores = d_osc.p;
    pa.out_counts = d_cnt.p;
    EventPair ev;
    cudaEvent_t e0 = ev.a, e1 = ev.b;
    CK(cudaEventRecord(e0, nullptr));
    if (nusers > 0) predict_topn_kernel<<<grid, kPredNT>>>(pa);
    CK(cudaGetLastError());
    CK(cudaEventRecord(e1, nullptr));
    CK(cudaDeviceSynchronize());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (kernel_ms) *kernel_ms = ms;
    if (nout > 0) {
      CK(cudaMemcpy(out_ids, d_oid.p, sizeof(int32_t) * nout, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(out_scores, d_osc.p, sizeof(float) * nout, cudaMemcpyDeviceToHost));
    }
    if (out_counts && nusers > 0)
      CK(cudaMemcpy(out_counts, d_cnt.p, sizeof(int32_t) * nusers, cudaMemcpyDeviceToHost));
    return kOk;
  } catch (const EngineError &e) {
    g_last_error = e.what();
    return e.status;
  } catch (const std::exception &e) {
    g_last_error = e.what();
    return kErrMemory;
  }
}

}  // namespace slimb200
