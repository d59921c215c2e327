// gram.cuh -- Gram-space coordinate descent ("covariance updates"); included by engine.cu.
//
// The reference sweeps in USER space: for every active coordinate i it streams column a_i three times
// against the dense vector yhat = sum_k x_k a_k (AddSpVec / SpVecInnerProduct / AddSpVec, reference
// src/libslim/cd.c:122-129).  The only quantity the update rule needs from that is the inner product
//     <a_i, yhat> = sum_{k : x_k != 0} x_k <a_i, a_k> = sum_k x_k G[k][i],        G = R^T R,
// and aTy_i = <a_i, a_j> = G[j][i] (estimate.c:412-421) is one row of the same matrix.  With G staged in
// HBM once per training matrix (gram_build_kernel; N^2 fp32, 40 GB for 100K items) a target column is
// solved without touching R again:
//   * the active set {i != j : G[j][i] > l1r} is a filtered copy of row j (estimate.c:433-444);
//   * the coordinates are visited in the same fixed cyclic order as the other kernels, in blocks of 32
//     consecutive active coordinates.  For a block the 32 inner products against the CURRENT iterate are
//     sum_{k in S} x_k G[k][block] over the nonzero set S (one gathered 32-wide load per nonzero, split
//     over all warps of the CTA or cluster); inside the block the sequential dependence is resolved
//     exactly with the 32x32 Gram block (ip_m += d_k G[k][m] for m after k) by one warp, which only
//     visits the coordinates whose value actually changes.
// The iterates are those of sequential CD (cd.c:117-133) up to fp64 rounding: same update rule, same
// EPSILON rule of AddSpVec (cd.c:27), same stop rule (cd.c:135-138), same iteration cap
// (estimate.c:448-449), same compaction (estimate.c:492-505).
//
// One CTA (CS == 1) or one thread-block cluster (CS > 1, heavy targets) owns a target.  In a cluster
// every CTA keeps a private copy of the small per-target state and runs the (deterministic) chain
// redundantly; only the sum over S is divided among the CTAs, and the 32 partial sums are all-gathered
// with tagged DSMEM stores (no cluster barrier on the per-block path).
#pragma once

constexpr int kGramNT = 256;
constexpr int kGramNW = kGramNT / 32;

// ------------------------------------------------------------------------------------------------
// Layout of G in HBM.  PANEL-major: a panel is kGramPW = 64 consecutive columns of all rows, rows contiguous
// inside the panel, so a block of coordinates reads short segments of many rows out of ONE panel.
//
// Element width.  For integer, non-negative ratings every entry is a non-negative integer bounded by
//     G[k][i] = sum_u r_uk r_ui  <=  rmax * sum_u r_ui  <=  rmax * csq_i        (r <= r^2 for integers),
// i.e. by a per-COLUMN bound that falls with the item's popularity.  Items are stored in popularity order, so
// the columns split into three contiguous ranges (boundaries rounded to whole panels):
//     [0, h32)  : 32-bit unsigned     [h32, h16) : bound <= 65535, 16-bit     [h16, ld) : bound <= 255, 8-bit
// For C4 (100 K items) that is 192 / 29 120 / 70 688 columns: 130 KB per row instead of 400 KB, 13 GB instead of
// 40 GB, and -- what matters -- 2-3x fewer DRAM sectors per gathered block (profiles/r02_typed_gram_model.txt).
// All sums are exact.  Ratings that are not non-negative integers (or sums >= 2^32) use fp64 elements instead.
// ------------------------------------------------------------------------------------------------
struct GramView {
  const unsigned char *base;
  size_t nr;           // rows of G (= ncols)
  int32_t h32, h16;    // packed layout: first column of the 16-bit / 8-bit range (multiples of kGramPW)
  size_t off16, off8;  // packed layout: byte offsets of the 16-bit / 8-bit ranges
  // STAIR layout (GaStair below; nullptr / 0 for the full layouts): byte offset of every panel, and the number of
  // leading rows that every panel stores
  const unsigned long long *pbase;
  int32_t hd;
};

// fp64 elements, one range: element (k, i) at gram_off(nr, k, i)
struct GatherStage {  // (row, value) pairs of a 32-entry chunk of the nonzero list, one slice per warp
  int row[32];
  double val[32];
};

struct GaF64 {
  static constexpr bool kStair = false;
  using Stage = GatherStage;
  using Tile = double;                          // element type of the in-block Gram tiles in shared memory
  static constexpr int kMaxRowBytes = kGramPW * 8;  // bytes of one panel row
  struct Col {
    const double *p;
  };
  static __device__ __forceinline__ Col col(const GramView &g, int i) {
    return Col{reinterpret_cast<const double *>(g.base) + gram_off(g.nr, 0, i)};
  }
  using Raw = double;  // what a gathered load leaves in a register until it is consumed
  static __device__ __forceinline__ Raw raw(const Col &c, int r) { return __ldg(c.p + (size_t)(uint32_t)r * kGramPW); }
  static __device__ __forceinline__ double cvt(const Col &, Raw w) { return w; }
  static __device__ __forceinline__ double at(const Col &c, int r) { return raw(c, r); }
  static __device__ __forceinline__ double at(const GramView &, const Col &c, int r) { return raw(c, r); }
  static __device__ __forceinline__ double at(const GramView &g, int k, int i) {
    return __ldg(reinterpret_cast<const double *>(g.base) + gram_off(g.nr, k, i));
  }
  // two adjacent items (item0 even) of row k
  static __device__ __forceinline__ void at2(const GramView &g, int k, int item0, double (&o)[2]) {
    const double2 v = __ldg(reinterpret_cast<const double2 *>(reinterpret_cast<const double *>(g.base) + gram_off(g.nr, k, item0)));
    o[0] = v.x;
    o[1] = v.y;
  }
  static __device__ __forceinline__ int row_bytes(const GramView &, int) { return kGramPW * 8; }
  static __device__ __forceinline__ const unsigned char *panel(const GramView &g, int item0) {
    return g.base + gram_off(g.nr, 0, item0) * 8;
  }
};

// packed unsigned elements, three column ranges
struct GaPacked {
  static constexpr bool kStair = false;
  using Stage = GatherStage;
  using Tile = float;  // exact: every entry is an integer below 2^24 ... or the tile holds it as float anyway
  static constexpr int kMaxRowBytes = kGramPW * 4;
  struct Col {
    const unsigned char *p;  // 4-byte aligned address of the word that holds (row 0, column i)
    uint32_t stride;         // bytes per panel row: 64 x element width
    uint32_t sel;            // PRMT selector that moves the element's bytes to the bottom of a register, zero above
  };
  // byte offset of (row 0, column i), bytes per panel row, log2 element width
  static __device__ __forceinline__ size_t col_byte(const GramView &g, int i, uint32_t &stride, uint32_t &wlog) {
    if (i < g.h32) {
      stride = kGramPW * 4;
      wlog = 2;
      return (size_t)(i >> 6) * g.nr * (kGramPW * 4) + (size_t)(i & 63) * 4;
    }
    if (i < g.h16) {
      const int ii = i - g.h32;
      stride = kGramPW * 2;
      wlog = 1;
      return g.off16 + (size_t)(ii >> 6) * g.nr * (kGramPW * 2) + (size_t)(ii & 63) * 2;
    }
    const int ii = i - g.h16;
    stride = kGramPW;
    wlog = 0;
    return g.off8 + (size_t)(ii >> 6) * g.nr * kGramPW + (size_t)(ii & 63);
  }
  static __device__ __forceinline__ Col col(const GramView &g, int i) {
    uint32_t stride, wlog;
    const size_t byte = col_byte(g, i, stride, wlog);
    Col c;
    c.p = g.base + (byte & ~size_t(3));
    c.stride = stride;
    const uint32_t b = (uint32_t)(byte & 3);  // first byte of the element inside its word
    // selector nibbles 0-3 pick a byte of the loaded word, 4 picks a byte of the second PRMT operand (zero)
    c.sel = wlog == 2 ? 0x3210u : (wlog == 1 ? (0x4400u | ((b + 1) << 4) | b) : (0x4440u | b));
    return c;
  }
  using Raw = uint32_t;
  // one mad.wide (row * row bytes + column base) and one 32-bit load per gathered element
  static __device__ __forceinline__ Raw raw(const Col &c, int r) {
    unsigned long long a;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(a) : "r"((uint32_t)r), "r"(c.stride), "l"(c.p));
    return __ldg(reinterpret_cast<const uint32_t *>(a));
  }
  static __device__ __forceinline__ double cvt(const Col &c, Raw w) { return (double)__byte_perm(w, 0u, c.sel); }
  static __device__ __forceinline__ double at(const Col &c, int r) { return cvt(c, raw(c, r)); }
  static __device__ __forceinline__ double at(const GramView &, const Col &c, int r) { return cvt(c, raw(c, r)); }
  static __device__ __forceinline__ double at(const GramView &g, int k, int i) {
    uint32_t stride, wlog;
    const size_t byte = col_byte(g, i, stride, wlog) + (size_t)k * stride;
    const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(g.base + (byte & ~size_t(3))));
    const uint32_t m = wlog == 2 ? 0xffffffffu : (wlog == 1 ? 0xffffu : 0xffu);
    return (double)((w >> ((uint32_t)(byte & 3) * 8u)) & m);
  }
  static __device__ __forceinline__ void at2(const GramView &g, int k, int item0, double (&o)[2]) {
    uint32_t stride, wlog;
    const size_t byte = col_byte(g, item0, stride, wlog) + (size_t)k * stride;  // item0 even: both in one panel row
    if (wlog == 2) {
      const uint2 v = __ldg(reinterpret_cast<const uint2 *>(g.base + byte));
      o[0] = (double)v.x;
      o[1] = (double)v.y;
    } else if (wlog == 1) {
      const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(g.base + byte));
      o[0] = (double)(w & 0xffffu);
      o[1] = (double)(w >> 16);
    } else {
      const uint32_t w = (uint32_t)__ldg(reinterpret_cast<const unsigned short *>(g.base + byte));
      o[0] = (double)(w & 0xffu);
      o[1] = (double)(w >> 8);
    }
  }
  static __device__ __forceinline__ int row_bytes(const GramView &g, int item0) {
    return item0 < g.h32 ? kGramPW * 4 : (item0 < g.h16 ? kGramPW * 2 : kGramPW);
  }
  static __device__ __forceinline__ const unsigned char *panel(const GramView &g, int item0) {  // item0 % 64 == 0
    uint32_t stride, wlog;
    return g.base + col_byte(g, item0, stride, wlog);
  }
};

// ------------------------------------------------------------------------------------------------
// STAIR layout: the packed layout for item counts whose full N x N matrix does not fit in HBM (C5: 500 K items would
// need 291 GB).  G is symmetric, so panel p (columns 64p .. 64p+63) stores only the rows
//     [0, max(64 (p + 1), hd))            (clamped to ncols),
// i.e. the upper triangle by panels plus a full square of the hd most popular items.  Element (k, i) is read as
// row k of column i when that is stored (k < hd or panel(k) <= panel(i)) and as row i of column k otherwise.
// Active sets consist mostly of popular items, so almost every gathered pair falls into the square (same access
// pattern as the full layout); the triangle costs one DRAM sector per gathered element.  Half the bytes of the full
// layout: 128 GB for C5.  Panel offsets come from a table (pbase, 8 bytes per panel).
// ------------------------------------------------------------------------------------------------
struct GatherStageStair : GatherStage {
  const unsigned char *kp[32];  // the entry's own column: word address of (row 0, column k) ...
  uint2 ks[32];                 // ... its row stride and PRMT selector
};

struct GaStair {
  static constexpr bool kStair = true;
  using Stage = GatherStageStair;
  using Tile = float;
  static constexpr int kMaxRowBytes = kGramPW * 4;
  struct Col {
    const unsigned char *p;  // 4-byte aligned address of the word that holds (row 0, column i)
    uint32_t stride;         // bytes per panel row
    uint32_t sel;            // PRMT selector of the element inside its word
    int32_t pan;             // panel of the column
    int32_t item;            // the column itself
  };
  static __device__ __forceinline__ Col col(const GramView &g, int i) {
    const uint32_t wlog = i < g.h32 ? 2u : (i < g.h16 ? 1u : 0u);
    const size_t byte = (size_t)__ldg(g.pbase + (i >> 6)) + ((size_t)(i & 63) << wlog);
    Col c;
    c.p = g.base + (byte & ~size_t(3));
    c.stride = (uint32_t)kGramPW << wlog;
    const uint32_t b = (uint32_t)(byte & 3);
    c.sel = wlog == 2 ? 0x3210u : (wlog == 1 ? (0x4400u | ((b + 1) << 4) | b) : (0x4440u | b));
    c.pan = i >> 6;
    c.item = i;
    return c;
  }
  static __device__ __forceinline__ bool stored(const GramView &g, int k, int pan_i) { return k < g.hd || (k >> 6) <= pan_i; }
  static __device__ __forceinline__ uint32_t word(const unsigned char *p, uint32_t stride, int r) {
    unsigned long long a;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(a) : "r"((uint32_t)r), "r"(stride), "l"(p));
    return __ldg(reinterpret_cast<const uint32_t *>(a));
  }
  // G[ck.item][ci.item] from two column descriptors
  static __device__ __forceinline__ double at(const GramView &g, const Col &ck, const Col &ci) {
    const bool d = stored(g, ck.item, ci.pan);
    const uint32_t w = d ? word(ci.p, ci.stride, ck.item) : word(ck.p, ck.stride, ci.item);
    return (double)__byte_perm(w, 0u, d ? ci.sel : ck.sel);
  }
  static __device__ __forceinline__ double at(const GramView &g, const Col &ci, int k) {
    if (stored(g, k, ci.pan)) return (double)__byte_perm(word(ci.p, ci.stride, k), 0u, ci.sel);
    const Col ck = col(g, k);
    return (double)__byte_perm(word(ck.p, ck.stride, ci.item), 0u, ck.sel);
  }
  static __device__ __forceinline__ double at(const GramView &g, int k, int i) { return at(g, col(g, i), k); }
};

struct GramArgs {
  GramView gv;    // the Gram matrix
  int32_t q_begin, q_end;  // positions in SolveArgs::targets served by this launch
  int32_t *queue;          // work counter of this launch (starts at 0)
  int32_t slot_base;       // first scratch slot of this launch (slot = slot_base + blockIdx.x)
  // per-CTA scratch, stride SolveArgs::col_stride
  int32_t *act;     // active coordinates (internal item ids), ascending
  double *x;        // current iterate per active coordinate
  int32_t *slotp;   // position of the coordinate in the nonzero list, -1 when it never entered it
  int32_t *sl_row;  // nonzero list: item id ...
  double *sl_val;   // ... and effective value (0 for a coordinate that went back to zero)
  const unsigned long long *expand;  // per item: sum of row lengths over the users of the column
};

// ------------------------------------------------------------------------------------------------
// K0g: G = R^T R by CSR row expansion, one Gram ROW per work item: for every user u of column k
// and every item i of row u, G[k][i] += r_uk * r_ui.  Work items are (column, entry range) in
// ascending internal id, so the CTAs in flight write a narrow band of rows that stays L2-resident;
// every row reaches HBM once.  fp32 sums are exact when the ratings are integers and the largest
// column sum of squares is below 2^24 (checked at staging); otherwise GT = double.
// ------------------------------------------------------------------------------------------------
// element update of the build: fp64 atomic add, or -- packed layout -- an integer add into the 32-bit word that
// holds the 8 / 16 / 32-bit field (no carry can leave a field: every partial sum is <= the final, bounded, sum)
struct GbF64 {
  static __device__ __forceinline__ void add(const GramView &g, int k, int i, float a, float b) {
    atomicAdd(const_cast<double *>(reinterpret_cast<const double *>(g.base)) + gram_off(g.nr, k, i), (double)a * (double)b);
  }
};
struct GbPacked {
  static __device__ __forceinline__ void add(const GramView &g, int k, int i, float a, float b) {
    uint32_t stride, wlog;
    const size_t byte = GaPacked::col_byte(g, i, stride, wlog) + (size_t)k * stride;
    const uint32_t inc = (uint32_t)(a * b);  // integer ratings: exact
    atomicAdd(reinterpret_cast<uint32_t *>(const_cast<unsigned char *>(g.base) + (byte & ~size_t(3))),
              inc << ((uint32_t)(byte & 3) * 8u));
  }
};

// stair layout: pairs whose element is not stored are skipped (the mirrored pair is stored)
struct GbStair {
  static __device__ __forceinline__ void add(const GramView &g, int k, int i, float a, float b) {
    if (!GaStair::stored(g, k, i >> 6)) return;
    const uint32_t wlog = i < g.h32 ? 2u : (i < g.h16 ? 1u : 0u);
    const size_t byte = (size_t)__ldg(g.pbase + (i >> 6)) + (size_t)k * ((size_t)kGramPW << wlog) + ((size_t)(i & 63) << wlog);
    const uint32_t inc = (uint32_t)(a * b);
    atomicAdd(reinterpret_cast<uint32_t *>(const_cast<unsigned char *>(g.base) + (byte & ~size_t(3))),
              inc << ((uint32_t)(byte & 3) * 8u));
  }
};

template <typename GB, bool HASVAL>
__global__ void __launch_bounds__(256) gram_build_kernel(int32_t nwork, const int32_t *__restrict__ wk_col,
                                                         const int32_t *__restrict__ wk_e0,
                                                         const int32_t *__restrict__ wk_e1,
                                                         const int64_t *__restrict__ colptr,
                                                         const int32_t *__restrict__ colind,
                                                         const float *__restrict__ colval,
                                                         const int64_t *__restrict__ rowptr,
                                                         const int32_t *__restrict__ rowind,
                                                         const float *__restrict__ rowval, const GramView gv,
                                                         unsigned long long *expand) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
    const int k = wk_col[w];
    const int e0 = wk_e0[w], e1 = wk_e1[w];
    const int64_t c0 = colptr[k];
    unsigned long long ex = 0;
    // two users per warp iteration keep more loads in flight
    for (int e = e0 + warp * 2; e < e1; e += 16) {
      const int ua = __ldg(colind + c0 + e);
      const bool two = e + 1 < e1;
      const int ub = two ? __ldg(colind + c0 + e + 1) : ua;
      const float va = HASVAL ? __ldg(colval + c0 + e) : 1.0f;
      const float vb = HASVAL ? (two ? __ldg(colval + c0 + e + 1) : 0.0f) : 1.0f;
      const int64_t a0 = __ldg(rowptr + ua), a1 = __ldg(rowptr + ua + 1);
      const int64_t b0 = __ldg(rowptr + ub), b1 = two ? __ldg(rowptr + ub + 1) : b0;
      ex += (unsigned long long)((a1 - a0) + (b1 - b0));
      int64_t ta = a0 + lane, tb = b0 + lane;
      while (ta < a1 || tb < b1) {
        const bool pa = ta < a1, pb = tb < b1;
        int ia = 0, ib = 0;
        float ra = 1.0f, rb = 1.0f;
        if (pa) ia = __ldg(rowind + ta);
        if (pb) ib = __ldg(rowind + tb);
        if (HASVAL) {
          if (pa) ra = __ldg(rowval + ta);
          if (pb) rb = __ldg(rowval + tb);
        }
        if (pa) GB::add(gv, k, ia, ra, va);
        if (pb) GB::add(gv, k, ib, rb, vb);
        ta += 32;
        tb += 32;
      }
    }
    if (expand) {
#pragma unroll
      for (int o = 16; o; o >>= 1) ex += __shfl_xor_sync(0xffffffffu, ex, o);
      if (lane == 0 && ex) atomicAdd(expand + k, ex / 32ull);  // every lane counted the same rows
    }
  }
}

// flags[0] bit 0: some rating is not an integer (or is not finite), bit 1: some rating is negative;
// flags[1] = largest rating (as an integer, when they all are)
__global__ void integer_values_kernel(const float *__restrict__ v, int64_t n, int32_t *flags) {
  bool bad = false, neg = false;
  int mx = 0;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const float x = v[k];
    const bool isint = x == rintf(x) && fabsf(x) < 16777216.0f;
    bad |= !isint;
    neg |= x < 0.0f;
    if (isint) mx = max(mx, (int)x);
  }
  const unsigned f = (__any_sync(0xffffffffu, bad) ? 1u : 0u) | (__any_sync(0xffffffffu, neg) ? 2u : 0u);
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) {
    if (f) atomicOr(flags, (int)f);
    atomicMax(flags + 1, mx);
  }
}

// ------------------------------------------------------------------------------------------------
// the solve kernel
// ------------------------------------------------------------------------------------------------
template <int CS>
struct __align__(16) GramSmem {
  double part[kGramNW][32];
  unsigned long long xchg[2][CS > 1 ? CS : 1][32][2];  // tagged all-gather slots (flag in data)
  double red[2 * kGramNW];
  int sc[kGramNW];
  long long misc[4];
  int len;    // entries in the nonzero list
  int nzero;  // ... of which currently zero
  int done;
};

__device__ __forceinline__ uint32_t gram_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

// Sum of one double per lane over the CTAs of the cluster; called by warp 0 of every CTA with the same
// tag sequence.  Every CTA adds the partials in rank order, so all of them get the bit-identical sum.
template <int CS>
__device__ __forceinline__ double gram_allsum(GramSmem<CS> &sm, double v, uint32_t rank, uint32_t &tag) {
  if (CS == 1) return v;
  const int lane = threadIdx.x & 31;
  tag++;
  const int buf = tag & 1;
#pragma unroll
  for (int r = 0; r < CS; r++) st_peer_tagged(&sm.xchg[buf][rank][lane][0], (uint32_t)r, v, tag);
  double s = 0.0;
#pragma unroll
  for (int r = 0; r < CS; r++) s += ld_tagged_wait(&sm.xchg[buf][r][lane][0], tag);
  return s;
}

// partial (this warp's share) of  sum_{e < len} val[e] * G[row[e]][col]  for the lane's column `col`.
// The (row, value) pairs of a 32-entry chunk are staged in this warp's slice of shared memory and read back as
// broadcasts: per gathered element one LDS (row), one mad.wide, one LDG, one PRMT, one I2F, one LDS.64 (value) and
// the DFMA.
template <typename GA, int UNR>
__device__ __forceinline__ double gram_gather_sum(const GramView &gv, const typename GA::Col &col, const int32_t *sl_row,
                                                  const double *sl_val, int len, int first_chunk, int chunk_stride,
                                                  typename GA::Stage &st) {
  const int lane = threadIdx.x & 31;
  double acc = 0.0;
  for (int c = first_chunk; c * 32 < len; c += chunk_stride) {
    const int e = c * 32 + lane;
    __syncwarp();
    const int k_mine = e < len ? sl_row[e] : 0;
    st.row[lane] = k_mine;
    st.val[lane] = e < len ? sl_val[e] : 0.0;  // entries past the end contribute 0 * G[0][col]
    if constexpr (GA::kStair) {
      const typename GA::Col ck = GA::col(gv, k_mine);
      st.kp[lane] = ck.p;
      st.ks[lane] = make_uint2(ck.stride, ck.sel);
    }
    __syncwarp();
    const int cnt = min(32, len - c * 32);
    if constexpr (GA::kStair) {
      // per element: row k of the lane's column when that is stored, row `item` of column k otherwise
      for (int i0 = 0; i0 < cnt; i0 += UNR) {
        uint32_t g[UNR];
#pragma unroll
        for (int u = 0; u < UNR; u++) {
          const int k = st.row[i0 + u];
          const bool d = GA::stored(gv, k, col.pan);
          g[u] = GA::word(d ? col.p : st.kp[i0 + u], d ? col.stride : st.ks[i0 + u].x, d ? k : col.item);
        }
#pragma unroll
        for (int u = 0; u < UNR; u++) {
          const bool d = GA::stored(gv, st.row[i0 + u], col.pan);
          acc = fma(st.val[i0 + u], (double)__byte_perm(g[u], 0u, d ? col.sel : st.ks[i0 + u].y), acc);
        }
      }
    } else {
      for (int i0 = 0; i0 < cnt; i0 += UNR) {
        typename GA::Raw g[UNR];  // UNR independent gathered loads in flight per lane
#pragma unroll
        for (int u = 0; u < UNR; u++) g[u] = GA::raw(col, st.row[i0 + u]);
#pragma unroll
        for (int u = 0; u < UNR; u++) acc = fma(st.val[i0 + u], GA::cvt(col, g[u]), acc);
      }
    }
  }
  return acc;
}

// UNR = gathered loads in flight per lane: 16 with 64 registers (4 CTAs per SM) or 32 with 80 registers (3 CTAs per SM);
// the stair accessor needs 80 registers at UNR = 16
template <typename GA, int CS, int UNR>
__global__ void __launch_bounds__(kGramNT, (UNR > 16 || GA::kStair) ? 3 : 4) cd_gram_kernel(const SolveArgs a, const GramArgs ga) {
  constexpr int NT = kGramNT, NW = kGramNW;
  using Tile = typename GA::Tile;
  __shared__ GramSmem<CS> sm;
  __shared__ Tile s_gbb[32][33];
  __shared__ typename GA::Stage s_stage[NW];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = CS > 1 ? gram_cluster_rank() : 0u;
  uint32_t tag = 0;
  int par = 0;

  const GramView &gv = ga.gv;
  const size_t slot = (size_t)(ga.slot_base + blockIdx.x) * a.col_stride;
  int32_t *act = ga.act + slot;
  double *x = ga.x + slot;
  int32_t *slotp = ga.slotp + slot;
  int32_t *sl_row = ga.sl_row + slot;
  double *sl_val = ga.sl_val + slot;
  float *xw = a.xw ? a.xw + slot : nullptr;

  if (CS > 1) {
    for (int i = tid; i < 2 * CS * 32 * 2; i += NT) (&sm.xchg[0][0][0][0])[i] = 0ull;
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }

  for (;;) {
    // ---- next target of this launch's queue (rank 0 fetches, the cluster agrees through the all-gather)
    __syncthreads();
    if (warp == 0) {
      double qv = 0.0;
      if (rank == 0) {
        int v = 0;
        if (lane == 0) v = atomicAdd(ga.queue, 1);
        qv = (double)__shfl_sync(0xffffffffu, v, 0);
      }
      qv = gram_allsum<CS>(sm, qv, rank, tag);
      if (lane == 0) sm.misc[0] = (long long)qv;
    }
    __syncthreads();
    const int q = ga.q_begin + (int)sm.misc[0];
    if (q >= ga.q_end) break;
    const int j = a.targets[q];
    const int cntj = a.colcnt[j];
    const typename GA::Col colj = GA::col(gv, j);
    auto gj_at = [&](int i) {  // aTy_i = G[j][i]
      if constexpr (GA::kStair) return GA::at(gv, colj, GA::col(gv, i));
      else return GA::at(gv, j, i);
    };
    const bool timer = rank == 0 && tid == 0;
    unsigned long long t_start = 0, t_act = 0, t_sweep = 0;
    if (timer) t_start = globaltimer_ns();

    // ---- warm start: scatter column j of the initial model (estimate.c:455-458)
    const int jo = a.inv[j];
    const bool warm = a.wcolptr != nullptr && jo < a.wncols;
    if (warm) {
      for (int64_t k = a.wcolptr[jo] + tid; k < a.wcolptr[jo + 1]; k += NT) {
        const int r = a.wcolind[k];
        if (r >= 0 && r < a.ncols) xw[a.rank[r]] = a.wcolval[k];
      }
      __syncthreads();
    }

    // ---- active set: ascending i, strict aTy > l1r, i != j (estimate.c:433-444); aTy_i = G[j][i]
    int na = 0;
    long long actnnz = 0;
    if (a.nnbrs > 0) {
      // fSLIM (estimate.c:424-431): the active set is the neighbour list of fslim_neighbors_kernel, no > l1r filter;
      // the reference sets no x = -0.1 flags on this branch, so a warm-start model has no effect (estimate.c:453-464)
      na = a.nbr_cnt[q];
      for (int p = tid; p < na; p += NT) {
        const int i = a.nbr_list[(size_t)q * a.nnbrs + p];
        act[p] = i;
        x[p] = 0.0;
        actnnz += a.colcnt[i];
      }
    } else {
      for (int base = 0; base < a.ncols; base += NT) {
        const int i = base + tid;
        double v = 0.0;
        if (i < a.ncols) v = gj_at(i);
        const bool flag = (i < a.ncols) && (i != j) && (v > a.l1r);
        int tot;
        const int pos = na + team_excl_scan<NT>(flag, sm.sc, tot);
        if (flag) {
          act[pos] = i;
          x[pos] = warm ? (double)xw[i] : 0.0;
          actnnz += a.colcnt[i];
        }
        na += tot;
      }
    }
    __syncthreads();
    if (warm) {
      for (int64_t k = a.wcolptr[jo] + tid; k < a.wcolptr[jo + 1]; k += NT) {
        const int r = a.wcolind[k];
        if (r >= 0 && r < a.ncols) xw[a.rank[r]] = 0.0f;
      }
    }

    // ---- nonzero list S from the starting iterate (empty for a cold start)
    auto rebuild_list = [&]() {
      int len = 0;
      for (int base = 0; base < na; base += NT) {
        const int p = base + tid;
        double xv = 0.0;
        if (p < na) xv = x[p];
        const bool flag = (p < na) && fabs(xv) > kEps;
        int tot;
        const int pos = len + team_excl_scan<NT>(flag, sm.sc, tot);
        if (p < na) slotp[p] = flag ? pos : -1;
        if (flag) {
          sl_row[pos] = act[p];
          sl_val[pos] = xv;
        }
        len += tot;
      }
      __syncthreads();
      if (tid == 0) {
        sm.len = len;
        sm.nzero = 0;
      }
      __syncthreads();
    };
    rebuild_list();

    // ---- iteration cap (estimate.c:448-449)
    const long long cap64 = 50LL * cntj;
    const int maxit = (int)(cap64 < (long long)a.maxniters ? cap64 : (long long)a.maxniters);

    // ---- the sweeps (cd.c:112-140)
    if (timer) t_act = globaltimer_ns();
    int niters = 1;
    const int nblk = (na + 31) >> 5;
    if (na > 0 && maxit > 0) {
      bool done = false;
      int t = 0;
      for (; t < maxit && !done; t++) {
        double dl = 0.0;  // warp 0: this lane's share of sum (x' - x)^2
        for (int b = 0; b < nblk; b++) {
          const int p0 = b * 32, pm = p0 + lane;
          const bool valid = pm < na;
          const int ab = act[valid ? pm : p0];
          const typename GA::Col cab = GA::col(gv, ab);  // column ab: row k at + k * row stride
          // operands of the chain, requested early (warp 0 only uses them)
          double xv = 0.0, sq = 0.0, den = 1.0, aty = 0.0;
          int myslot = -1;
          if (warp == 0) {
            if (valid) {
              xv = x[pm];
              myslot = slotp[pm];
            }
            const double cn = (double)__ldg(a.cnorms + ab);
            den = 1.0 / (cn * cn + a.l2r);  // reciprocal of cnorm^2 + l2r (cd.c:127), taken off the chain's critical path
            sq = __ldg(a.csq + ab);
            aty = (double)(float)GA::at(gv, cab, j);  // G[j][ab]; gk_fkv_t.key is a float (estimate.c:437)
          }
          // in-block Gram rows: warp w stages rows 4w .. 4w+3
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int r = warp * 4 + u;
            if (p0 + r < na) s_gbb[r][lane] = (Tile)GA::at(gv, cab, act[p0 + r]);
          }
          // <a_m, yhat> for the 32 coordinates of the block: this warp's share of the sum over S
          const int len = sm.len;
          sm.part[warp][lane] = gram_gather_sum<GA, UNR>(gv, cab, sl_row, sl_val, len, (int)rank * NW + warp, CS * NW, s_stage[warp]);
          __syncthreads();
          if (warp == 0) {
            double ipf = 0.0;
#pragma unroll
            for (int w = 0; w < NW; w++) ipf += sm.part[w][lane];
            ipf = gram_allsum<CS>(sm, ipf, rank, tag);

            // exact sequential CD inside the block; only coordinates whose value changes are visited
            double xn = xv;
            int k = 0;
            for (;;) {
              const double in_old = fabs(xn) > kEps ? xn : 0.0;
              const double ip = ipf - in_old * sq;  // cd.c:122-123 in one step
              const double num = aty - ip;
              const double nx = num > a.l1r ? (num - a.l1r) * den : 0.0;
              const unsigned want = __ballot_sync(0xffffffffu, valid && lane >= k && nx != xn);
              if (!want) break;
              const int kk = __ffs(want) - 1;
              const double in_new = fabs(nx) > kEps ? nx : 0.0;
              const double d = __shfl_sync(0xffffffffu, in_new - in_old, kk);
              if (lane == kk) {
                dl += (nx - xn) * (nx - xn);
                xn = nx;
              }
              if (d != 0.0) ipf = fma(d, (double)s_gbb[kk][lane], ipf);
              k = kk + 1;
            }

            // write back x and keep the nonzero list in step
            const double was = fabs(xv) > kEps ? xv : 0.0;
            const double now = fabs(xn) > kEps ? xn : 0.0;
            if (valid && xn != xv) x[pm] = xn;
            const bool chg = valid && now != was;
            const bool app = chg && myslot < 0;  // enters the list for the first time (now != 0)
            const unsigned appm = __ballot_sync(0xffffffffu, app);
            const unsigned zerm = __ballot_sync(0xffffffffu, chg && myslot >= 0 && now == 0.0);
            const unsigned revm = __ballot_sync(0xffffffffu, chg && myslot >= 0 && was == 0.0);
            if (chg && myslot >= 0) sl_val[myslot] = now;
            if (app) {
              const int pos = len + __popc(appm & ((1u << lane) - 1u));
              sl_row[pos] = ab;
              sl_val[pos] = now;
              slotp[pm] = pos;
            }
            if (lane == 0) {
              sm.len = len + __popc(appm);
              sm.nzero += __popc(zerm) - __popc(revm);
            }
          }
          __syncthreads();
        }
        // ---- end of sweep: stop rule (cd.c:135-138)
        if (warp == 0) {
#pragma unroll
          for (int o = 16; o; o >>= 1) dl += __shfl_xor_sync(0xffffffffu, dl, o);
          if (lane == 0) sm.done = dl < a.opttol ? 1 : 0;
        }
        __syncthreads();
        done = sm.done != 0;
        const bool compact = sm.nzero * 8 > sm.len;
        __syncthreads();
        if (!done && compact) rebuild_list();
      }
      niters = done ? t : maxit + 1;  // cd.c:140
    } else if (maxit > 0) {
      niters = (0.0 < a.opttol) ? 1 : maxit + 1;
    }

    // ---- residual / objective (estimate.c:477-489) in Gram space:
    //      |y - yhat|^2 = |y|^2 - 2 sum_k x_k aTy_k + sum_k x_k <a_k, yhat>
    if (timer) t_sweep = globaltimer_ns();
    const int len = sm.len;
    double hh = 0.0;
    {
      const int nch = (len + 31) >> 5;
      for (int cb = (int)rank; cb < nch; cb += CS) {
        const int e = cb * 32 + lane;
        const int col = sl_row[e < len ? e : 0];
        const double vk = e < len ? sl_val[e] : 0.0;
        const double s = gram_gather_sum<GA, UNR>(gv, GA::col(gv, col), sl_row, sl_val, len, warp, NW, s_stage[warp]);
        hh = fma(vk, s, hh);
      }
    }
    hh = team_sum<NT>(hh, sm.red, par);
    if (CS > 1) {
      if (warp == 0) {
        const double tot = gram_allsum<CS>(sm, lane == 0 ? hh : 0.0, rank, tag);
        if (lane == 0) sm.misc[2] = __double_as_longlong(tot);
      }
      __syncthreads();
      hh = __longlong_as_double(sm.misc[2]);
    }

    if (rank == 0) {
      double yd = 0.0, reg = 0.0;
      int nnz_local = 0;
      for (int p = tid; p < na; p += NT) {
        const double xv = x[p];
        const double in = fabs(xv) > kEps ? xv : 0.0;
        yd = fma(in, gj_at(act[p]), yd);
        reg += 0.5 * a.l2r * xv * xv + a.l1r * fabs(xv);
        nnz_local += in != 0.0 ? 1 : 0;
      }
      yd = team_sum<NT>(yd, sm.red, par);
      reg = team_sum<NT>(reg, sm.red, par);
      const int nnz_w = (int)(team_sum<NT>((double)nnz_local, sm.red, par) + 0.5);
      const double actnnz_t = team_sum<NT>((double)actnnz, sm.red, par);

      // ---- compaction |x| > EPS -> (i, (float)x) in visiting order (estimate.c:492-505)
      if (tid == 0) sm.misc[1] = (long long)atomicAdd(a.pool_used, (unsigned long long)nnz_w);
      __syncthreads();
      const long long off = sm.misc[1];
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
            a.pool_idx[off + pos] = a.inv[act[p]];
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
        a.st_expand[q] = ga.expand ? (long long)ga.expand[j] : 0;
        const double yy = a.csq[j];
        const double rn = 0.5 * (yy - 2.0 * yd + hh);
        a.st_rnorm[q] = rn;
        a.st_obj[q] = rn + reg;
        a.st_ngroups[q] = nblk;
        const unsigned long long t_end = globaltimer_ns();
        a.st_phase[(size_t)q * 4 + 0] = 0.f;
        a.st_phase[(size_t)q * 4 + 1] = (float)(t_act - t_start) * 1e-3f;
        a.st_phase[(size_t)q * 4 + 2] = (float)(t_sweep - t_act) * 1e-3f;
        a.st_phase[(size_t)q * 4 + 3] = (float)(t_end - t_sweep) * 1e-3f;
      }
    }
  }
  if (CS > 1) {
    // no CTA may exit while a peer can still write into its shared memory
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}
