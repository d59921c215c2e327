// hybrid.cuh -- the GIANT targets of a matrix whose Gram matrix is resident; included by engine.cu.
//
// A giant target (tens of thousands of users, an active set of 10^5 coordinates and a nonzero set S of a third of
// that) is expensive in both formulations: a Gram-space sweep gathers |S| x |A| ~ 10^9-10^10 elements of G, a
// user-space sweep streams nnz(active columns) ~ nnz(R) entries -- fewer, but cd_cluster_kernel spends a fixed
// ~12 us per 32-ITEM window on its exchange / solve chain, 15 625 windows per sweep for 500 K items.
// This kernel takes what is cheap from each side:
//   * <a_i, yhat> comes from user space (reference cd.c:122-123): one pass over the columns of a block of
//     kHybBK = 128 consecutive ACTIVE coordinates against the yhat at the start of the round; CTA r of the cluster
//     owns user range r (colsplit) and with it a private slice of yhat;
//   * the sequential dependence inside the block (cd.c:117-133 visits the coordinates one after the other) is
//     resolved exactly with the 128 x 128 Gram tile G[block][block] -- 8 K gathered elements of the resident G above
//     the diagonal, read ahead while the previous block's sums are exchanged and its chain runs -- by one warp that
//     only visits the coordinates whose value changes AND whose tile row is not zero ("emitters");
//   * yhat += d_i a_i (cd.c:129) for the coordinates that changed, fp64 reductions (REDG) into the CTA's own slice.
// A round costs one cluster barrier per 128 active coordinates (~10 us for short column ranges, measured with
// SLIMB200_PROFILE=1: profiles/r02_c5_hybrid_notes.txt); inactive items cost nothing.  Iterates, stop rule, cap and
// compaction are those of the other kernels (estimate.c:433-505).
// Timestamps of the profile: BAR.SYNC is compiled DEFER_BLOCKING, so a clock read right after __syncthreads() can
// precede the barrier's completion -- the wait then shows up in the NEXT interval.
#pragma once

constexpr int kHybNT = 512;
constexpr int kHybNW = kHybNT / 32;
constexpr int kHybBK = 128;
constexpr int kHybLane = 16;    // entries of a column inside the CTA's user range: one LANE takes the column
constexpr int kHybGroup = 128;  // ... 8 lanes
constexpr int kHybWarp = 8192;  // ... one warp; longer ranges are split over all warps of the CTA

// One active coordinate of the current target, built once per target by the whole cluster (one 128-byte line):
// a round's read-ahead is then ONE load latency instead of a chain act -> colptr / colsplit / pbase -> G.
struct __align__(128) HybLine {
  long long c0;               // padded offset of the column
  double inv_den;             // 1 / (cnorm^2 + l2r)
  double sq;                  // exact sum of squares
  const unsigned char *gp;    // Gram column descriptor (GA::Col): word address of (row 0, column item) ...
  uint32_t gstride, gsel;     // ... row stride and PRMT selector
  float aty;                  // (float)G[j][item]: gk_fkv_t.key is a float (estimate.c:437)
  int32_t item;
  int32_t split[kParts + 1];  // entry offsets of the 16 user ranges inside the column
  int32_t pad[3];
};
static_assert(sizeof(HybLine) == 128, "one line per active coordinate");

struct HybridArgs {
  GramView gv;
  const int32_t *colsplit;  // [ncols][kParts + 1]
  int32_t rows_per_part;
  double *xc;               // per CTA [col_stride]: every CTA keeps its own copy of x (identical values)
  HybLine *lines;           // per cluster [col_stride]
  const unsigned long long *expand;
  // SLIMB200_PROFILE=1: per target 10 counters (clock64 cycles of CTA 0's thread 0): inner products / exchange /
  // chain + read-ahead / yhat update, for rounds without and with long column ranges, and the two round counts
  unsigned long long *prof;
};

template <typename GA>
struct __align__(16) HybMeta {  // one block of active coordinates, as seen by this CTA
  typename GA::Col dcol[kHybBK];  // where column act[p] of G lives (gathers of the tile need no dependent load)
  int item[kHybBK];
  long long c0[kHybBK];
  double inv_den[kHybBK];
  double sq[kHybBK];
  int s0[kHybBK], s1[kHybBK];  // this CTA's entry range inside the column
  float aty[kHybBK];
  unsigned char cls[kHybBK];  // 0: empty range, 1: lane, 2: group of 8 lanes, 3: warp, 4: whole CTA
  unsigned bmask[4];          // class-4 columns
  unsigned wmask[4];          // class-3 columns
  unsigned gmask[4];          // class-2 columns
};

template <typename GA, bool HASVAL>
struct __align__(16) HybSmem {
  float tile[2][kHybBK][kHybBK];  // G[block][block] (upper triangle used), double buffered
  HybMeta<GA> meta[3];            // ring: the block of this round, of the next round and of the round after
  int sid[2][kHybLane][kHybBK];                                      // user ids of the class-1 columns, read ahead
  float sval[HASVAL ? 2 : 1][HASVAL ? kHybLane : 1][HASVAL ? kHybBK : 1];  // ... and their values
  double mine[2][kHybBK];         // this CTA's partial inner products; peers read them through DSMEM
  double pwb[kHybNW][kHybBK];     // per-warp partials of the class-4 columns
  double tot[kHybBK];             // cluster-wide sums
  double dlt[kHybBK];             // yhat step per coordinate
  double red[kHybNW];
  unsigned dep[2][4];             // coordinates of the block whose tile row has a nonzero above the diagonal ("emitters")
  int sc[kHybNW];
  int q, na, done;
  double dl;
  long long off;
};

__device__ __forceinline__ void hyb_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void hyb_cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void hyb_cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void hyb_bar_warps03() {  // named barrier 1: the four warps that own the block's sums
  asm volatile("bar.sync 1, 128;" ::: "memory");
}
__device__ __forceinline__ double hyb_ld_peer(const double *local, uint32_t peer) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local)), "r"(peer));
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
  return v;
}
__device__ __forceinline__ void hyb_st_peer_int(int *local, uint32_t peer, int v) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local)), "r"(peer));
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(ra), "r"(v) : "memory");
}

// entries e_k = e0 + sub + G k, k < 16, of column range [.., s1): 16 id loads, then 16 yhat gathers in flight per lane
template <bool HASVAL, int G>
__device__ __forceinline__ double hyb_dot16(const SolveArgs &a, long long c0, int e0, int s1, int sub, const double *yh) {
  const int32_t *ix = a.colind + c0;
  int id[16];
  float vl[16];
#pragma unroll
  for (int k = 0; k < 16; k++) {
    const int e = e0 + sub + G * k;
    id[k] = e < s1 ? __ldg(ix + e) : -1;
    if (HASVAL) vl[k] = e < s1 ? __ldg(a.colval + c0 + e) : 0.f;
  }
  double y[16];
#pragma unroll
  for (int k = 0; k < 16; k++) y[k] = __ldcg(yh + (id[k] < 0 ? 0 : id[k]));
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < 16; k++) acc += id[k] < 0 ? 0.0 : (HASVAL ? (double)vl[k] * y[k] : y[k]);
  return acc;
}

template <bool HASVAL, int G>
__device__ __forceinline__ void hyb_axpy16(const SolveArgs &a, long long c0, int e0, int s1, int sub, double d, double *yh) {
  const int32_t *ix = a.colind + c0;
  int id[16];
  float vl[16];
#pragma unroll
  for (int k = 0; k < 16; k++) {
    const int e = e0 + sub + G * k;
    id[k] = e < s1 ? __ldg(ix + e) : -1;
    if (HASVAL) vl[k] = e < s1 ? __ldg(a.colval + c0 + e) : 0.f;
  }
#pragma unroll
  for (int k = 0; k < 16; k++)
    if (id[k] >= 0) red_add_f64(yh + id[k], HASVAL ? d * (double)vl[k] : d);
}

// One warp over the entries e_begin + lane + 32 k + i * step of a LONG column range, software-pipelined: the user ids
// of step i+1 are requested together with the yhat gathers of step i, so a step costs one load latency, not two.
template <bool HASVAL>
__device__ __forceinline__ double hyb_dot_range(const SolveArgs &a, long long c0, int e_begin, int s1, int step, int lane,
                                                const double *yh) {
  const int32_t *ix = a.colind + c0;
  int id[16];
  float vl[16];
#pragma unroll
  for (int k = 0; k < 16; k++) {
    const int e = e_begin + lane + 32 * k;
    id[k] = e < s1 ? __ldg(ix + e) : -1;
    if (HASVAL) vl[k] = e < s1 ? __ldg(a.colval + c0 + e) : 0.f;
  }
  double acc = 0.0;
  for (int e0 = e_begin; e0 < s1; e0 += step) {
    double y[16];
#pragma unroll
    for (int k = 0; k < 16; k++) y[k] = __ldcg(yh + (id[k] < 0 ? 0 : id[k]));
    int idn[16];
    float vln[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const int e = e0 + step + lane + 32 * k;
      idn[k] = e < s1 ? __ldg(ix + e) : -1;
      if (HASVAL) vln[k] = e < s1 ? __ldg(a.colval + c0 + e) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 16; k++) {
      acc += id[k] < 0 ? 0.0 : (HASVAL ? (double)vl[k] * y[k] : y[k]);
      id[k] = idn[k];
      if (HASVAL) vl[k] = vln[k];
    }
  }
  return acc;
}

template <bool HASVAL>
__device__ __forceinline__ void hyb_axpy_range(const SolveArgs &a, long long c0, int e_begin, int s1, int step, int lane,
                                               double d, double *yh) {
  const int32_t *ix = a.colind + c0;
  int id[16];
  float vl[16];
#pragma unroll
  for (int k = 0; k < 16; k++) {
    const int e = e_begin + lane + 32 * k;
    id[k] = e < s1 ? __ldg(ix + e) : -1;
    if (HASVAL) vl[k] = e < s1 ? __ldg(a.colval + c0 + e) : 0.f;
  }
  for (int e0 = e_begin; e0 < s1; e0 += step) {
    int idn[16];
    float vln[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const int e = e0 + step + lane + 32 * k;
      idn[k] = e < s1 ? __ldg(ix + e) : -1;
      if (HASVAL) vln[k] = e < s1 ? __ldg(a.colval + c0 + e) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 16; k++) {
      if (id[k] >= 0) red_add_f64(yh + id[k], HASVAL ? d * (double)vl[k] : d);
      id[k] = idn[k];
      if (HASVAL) vl[k] = vln[k];
    }
  }
}

template <typename GA, bool HASVAL>
__global__ void __launch_bounds__(kHybNT, 1) cd_hybrid_kernel(const SolveArgs a, const HybridArgs ha) {
  constexpr int NT = kHybNT, NW = kHybNW, BK = kHybBK;
  extern __shared__ __align__(128) unsigned char hyb_dyn[];
  HybSmem<GA, HASVAL> &sm = *reinterpret_cast<HybSmem<GA, HASVAL> *>(hyb_dyn);
  cg::cluster_group cl = cg::this_cluster();
  const int cs = (int)cl.num_blocks();
  const uint32_t rank = cl.block_rank();
  const int cid = blockIdx.x / cs;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const GramView &gv = ha.gv;

  const int pr0 = (int)rank * (kParts / cs), pr1 = ((int)rank + 1) * (kParts / cs);
  const int64_t ulo64 = (int64_t)pr0 * ha.rows_per_part, uhi64 = (int64_t)pr1 * ha.rows_per_part;
  const int u_lo = (int)(ulo64 < a.nrows ? ulo64 : a.nrows), u_hi = (int)(uhi64 < a.nrows ? uhi64 : a.nrows);

  float *xw = a.xw + (size_t)cid * a.col_stride;
  int32_t *act = a.act_idx + (size_t)cid * a.col_stride;
  double *x = ha.xc + (size_t)blockIdx.x * a.col_stride;
  double *yh = a.yhat + (size_t)cid * a.row_stride;
  HybLine *lines = ha.lines + (size_t)cid * a.col_stride;
  unsigned xb = 0;  // exchange buffer of the next round (kernel lifetime)
  hyb_cluster_sync();  // every CTA of the cluster has started: its shared memory may be written from now on

  for (;;) {
    // ---- next target -------------------------------------------------------------------------------
    if (rank == 0 && tid == 0) {
      const int q = atomicAdd(a.queue, 1);
      for (int c = 0; c < cs; c++) hyb_st_peer_int(&sm.q, (uint32_t)c, q);
    }
    hyb_cluster_sync();
    const int q = sm.q;
    if (q >= a.ntargets) break;
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

    // ---- warm start (estimate.c:455-458) and active set (estimate.c:433-444), built by CTA 0 ---------
    const int jo = a.inv[j];
    const bool warm = a.wcolptr != nullptr && jo < a.wncols;
    long long actnnz = 0;
    if (rank == 0) {
      if (warm) {
        for (int64_t k = a.wcolptr[jo] + tid; k < a.wcolptr[jo + 1]; k += NT) {
          const int r = a.wcolind[k];
          if (r >= 0 && r < a.ncols) xw[a.rank[r]] = a.wcolval[k];
        }
      }
      int na = 0;
      for (int base = 0; base < a.ncols; base += NT) {
        const int i = base + tid;
        double v = 0.0;
        if (i < a.ncols) v = gj_at(i);
        const bool flag = (i < a.ncols) && (i != j) && (v > a.l1r);
        int tot;
        const int pos = na + team_excl_scan<NT>(flag, sm.sc, tot);
        if (flag) {
          act[pos] = i;
          actnnz += a.colcnt[i];
        }
        na += tot;
      }
      if (tid == 0)
        for (int c = 0; c < cs; c++) hyb_st_peer_int(&sm.na, (uint32_t)c, na);
    }
    __threadfence();
    hyb_cluster_sync();
    const int na = sm.na;
    for (int p = tid; p < na; p += NT) x[p] = warm ? (double)__ldcg(&xw[__ldcg(&act[p])]) : 0.0;
    // one line per active coordinate, filled by all CTAs of the cluster
    for (int p = (int)rank * NT + tid; p < na; p += cs * NT) {
      const int i = __ldcg(&act[p]);
      const typename GA::Col ci = GA::col(gv, i);
      const long long c0 = __ldg(a.colptr + i);
      const double cn = (double)__ldg(a.cnorms + i);
      const double inv_den = 1.0 / (cn * cn + a.l2r);
      const double sq = __ldg(a.csq + i);
      double aty;
      if constexpr (GA::kStair) aty = GA::at(gv, colj, ci);
      else aty = GA::at(gv, ci, j);
      const int32_t *sp = ha.colsplit + (size_t)i * (kParts + 1);
      int sv[kParts + 1];
#pragma unroll
      for (int k = 0; k <= kParts; k++) sv[k] = __ldg(sp + k);
      const unsigned long long gp = (unsigned long long)ci.p;
      int4 *dst = reinterpret_cast<int4 *>(&lines[p]);  // field order of HybLine
      __stcg(dst + 0, make_int4((int)(unsigned)c0, (int)(c0 >> 32), __double2loint(inv_den), __double2hiint(inv_den)));
      __stcg(dst + 1, make_int4(__double2loint(sq), __double2hiint(sq), (int)(unsigned)gp, (int)(gp >> 32)));
      __stcg(dst + 2, make_int4((int)ci.stride, (int)ci.sel, __float_as_int((float)aty), i));
      __stcg(dst + 3, make_int4(sv[0], sv[1], sv[2], sv[3]));
      __stcg(dst + 4, make_int4(sv[4], sv[5], sv[6], sv[7]));
      __stcg(dst + 5, make_int4(sv[8], sv[9], sv[10], sv[11]));
      __stcg(dst + 6, make_int4(sv[12], sv[13], sv[14], sv[15]));
      __stcg(dst + 7, make_int4(sv[16], 0, 0, 0));
    }
    __threadfence();
    __syncthreads();
    hyb_cluster_sync();  // every CTA has read xw before CTA 0 clears it; the lines are complete
    if (warm && rank == 0) {
      for (int64_t k = a.wcolptr[jo] + tid; k < a.wcolptr[jo + 1]; k += NT) {
        const int r = a.wcolind[k];
        if (r >= 0 && r < a.ncols) xw[a.rank[r]] = 0.0f;
      }
    }
    const long long cap64 = 50LL * cntj;
    const int maxit = (int)(cap64 < (long long)a.maxniters ? cap64 : (long long)a.maxniters);
    const int nblk = (na + BK - 1) / BK;

    // Read-ahead, every part ONE load latency deep, issued right after a round's partial sums are on their way and
    // overlapped with the exchange and the chain of that round:
    //   prefetch_meta (warps 4-7): the coordinate lines of the block TWO rounds ahead -> meta ring slot `ms`;
    //   prefetch_ids (warps 4-7): user ids of the short (class-1) column ranges of the NEXT block (its meta slot was
    //     filled a round earlier) -> id buffer `ts`;
    //   prefetch_tile (warps 8-15): G[block][block] of the next block through the descriptors of its meta slot -> tile `ts`.
    auto prefetch_meta = [&](int b, int ms) {
      if (warp < 4 || warp > 7) return;
      HybMeta<GA> &M = sm.meta[ms];
      const int p0 = b * BK;
      const int n = min(BK, na - p0);
      const int m = tid - 128;
      int cls = 0;
      if (m < n) {
        const HybLine *L = &lines[p0 + m];
        const int4 h0 = __ldcg(reinterpret_cast<const int4 *>(L));      // c0, inv_den
        const int4 h1 = __ldcg(reinterpret_cast<const int4 *>(L) + 1);  // sq, gp
        const int4 h2 = __ldcg(reinterpret_cast<const int4 *>(L) + 2);  // gstride, gsel, aty, item
        const int s0 = __ldcg(&L->split[pr0]), s1 = __ldcg(&L->split[pr1]);
        typename GA::Col dc;
        dc.p = reinterpret_cast<const unsigned char *>(((unsigned long long)(unsigned)h1.w << 32) | (unsigned)h1.z);
        dc.stride = (uint32_t)h2.x;
        dc.sel = (uint32_t)h2.y;
        if constexpr (GA::kStair) {
          dc.pan = h2.w >> 6;
          dc.item = h2.w;
        }
        M.dcol[m] = dc;
        M.item[m] = h2.w;
        M.c0[m] = (long long)(((unsigned long long)(unsigned)h0.y << 32) | (unsigned)h0.x);
        M.inv_den[m] = __hiloint2double(h0.w, h0.z);
        M.sq[m] = __hiloint2double(h1.y, h1.x);
        M.aty[m] = __int_as_float(h2.z);
        M.s0[m] = s0;
        M.s1[m] = s1;
        const int len = s1 - s0;
        cls = len <= 0 ? 0 : (len <= kHybLane ? 1 : (len <= kHybGroup ? 2 : (len <= kHybWarp ? 3 : 4)));
      }
      M.cls[m] = (unsigned char)cls;
      const unsigned b4 = __ballot_sync(0xffffffffu, cls == 4), b3 = __ballot_sync(0xffffffffu, cls == 3);
      const unsigned b2 = __ballot_sync(0xffffffffu, cls == 2);
      if (lane == 0) {
        M.bmask[warp - 4] = b4;
        M.wmask[warp - 4] = b3;
        M.gmask[warp - 4] = b2;
      }
    };
    auto prefetch_ids = [&](int ms, int nb) {
      if (warp < 4 || warp > 7) return;
      const HybMeta<GA> &M = sm.meta[ms];
      const int m = tid - 128;
      if (M.cls[m] == 1) {
        const long long c0 = M.c0[m];
        const int s0 = M.s0[m], len = M.s1[m] - s0;
        int id[kHybLane];
        float vl[kHybLane];
#pragma unroll
        for (int k = 0; k < kHybLane; k++) {
          id[k] = k < len ? __ldg(a.colind + c0 + s0 + k) : 0;
          if (HASVAL) vl[k] = k < len ? __ldg(a.colval + c0 + s0 + k) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < kHybLane; k++) {
          sm.sid[nb][k][m] = id[k];
          if (HASVAL) sm.sval[nb][k][m] = vl[k];
        }
      }
    };
    auto prefetch_tile = [&](int b, int ms, int nb) {
      if (warp < 8) return;
      const HybMeta<GA> &M = sm.meta[ms];
      const int n = min(BK, na - b * BK);
      // the 8128 elements above the diagonal, flattened: rows P and 127 - P together hold exactly 127 of them
      constexpr int NTH = (NW - 8) * 32, PER = (BK * (BK - 1) / 2 + NTH - 1) / NTH;
      const int t = (warp - 8) * 32 + lane;
      constexpr int HALF = PER / 2;
      static_assert(PER % 2 == 0, "two batches");
#pragma unroll 1
      for (int k0 = 0; k0 < PER; k0 += HALF) {  // two batches of HALF independent gathers per lane
        // Above the diagonal the row item is the more popular one (actives ascend), so the element is always stored
        // as row r of column c, in both layouts.  Addresses first, then all loads, then the conversions: no branch
        // between two loads, HALF of them in flight per lane.
        uint32_t w[HALF];
#pragma unroll
        for (int k = 0; k < HALF; k++) {
          const int e = t + (k0 + k) * NTH;
          const int P = e / (BK - 1), o = e - P * (BK - 1);
          const bool top = o < BK - 1 - P;
          const int r = top ? P : BK - 1 - P;
          const int c = top ? P + 1 + o : BK - P + (o - (BK - 1 - P));
          const bool ok = e < BK * (BK - 1) / 2 && c < n;
          const int cs_ = ok ? c : 0, rs_ = ok ? r : 0;
          unsigned long long adr;
          asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(adr) : "r"((uint32_t)M.item[rs_]), "r"((uint32_t)M.dcol[cs_].stride), "l"(M.dcol[cs_].p));
          w[k] = ok ? __ldg(reinterpret_cast<const uint32_t *>(adr)) : 0u;
        }
#pragma unroll
        for (int k = 0; k < HALF; k++) {
          const int e = t + (k0 + k) * NTH;
          const int P = e / (BK - 1), o = e - P * (BK - 1);
          const bool top = o < BK - 1 - P;
          const int r = top ? P : BK - 1 - P;
          const int c = top ? P + 1 + o : BK - P + (o - (BK - 1 - P));
          if (e < BK * (BK - 1) / 2 && c < n) {
            const uint32_t v = __byte_perm(w[k], 0u, M.dcol[c].sel);
            sm.tile[nb][r][c] = (float)v;
            if (v != 0u) atomicOr(&sm.dep[nb][r >> 5], 1u << (r & 31));
          }
        }
      }
    };

    // yhat slice += d * (this CTA's range of the block's columns), for the coordinates with d != 0
    auto update_yhat = [&](int ms, int buf, int n) {
      const HybMeta<GA> &M = sm.meta[ms];
      {  // class 1: user ids staged in shared memory, one entry per thread and pass (4 entries of every column per pass)
        const int m = tid & (BK - 1);
        if (m < n && M.cls[m] == 1) {
          const double d = sm.dlt[m];
          const int len = M.s1[m] - M.s0[m];
          if (d != 0.0)
            for (int k = tid >> 7; k < len; k += NT / BK)
              red_add_f64(yh + sm.sid[buf][k][m], HASVAL ? d * (double)sm.sval[HASVAL ? buf : 0][HASVAL ? k : 0][HASVAL ? m : 0] : d);
        }
      }
      if ((M.gmask[0] | M.gmask[1] | M.gmask[2] | M.gmask[3]) != 0u) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int m = (tid >> 3) + 64 * h;
          if (m < n && M.cls[m] == 2) {
            const double d = sm.dlt[m];
            if (d != 0.0) hyb_axpy16<HASVAL, 8>(a, M.c0[m], M.s0[m], M.s1[m], lane & 7, d, yh);
          }
        }
      }
      if ((M.wmask[0] | M.wmask[1] | M.wmask[2] | M.wmask[3]) != 0u) {
        for (int m = warp; m < n; m += NW)
          if (M.cls[m] == 3) {
            const double d = sm.dlt[m];
            if (d != 0.0)
              hyb_axpy_range<HASVAL>(a, M.c0[m], M.s0[m], M.s1[m], 512, lane, d, yh);
          }
      }
#pragma unroll
      for (int w4 = 0; w4 < 4; w4++) {
        unsigned mm = M.bmask[w4];
        while (mm) {
          const int m = 32 * w4 + __ffs(mm) - 1;
          mm &= mm - 1;
          const double d = sm.dlt[m];
          if (d != 0.0) hyb_axpy_range<HASVAL>(a, M.c0[m], M.s0[m] + warp * 512, M.s1[m], NW * 512, lane, d, yh);
        }
      }
    };

    int mc = 0, tc = 0;  // meta ring slot and tile / id buffer of the current block
    if (tid < 8) sm.dep[tid >> 2][tid & 3] = 0u;
    // ---- warm start: yhat slice = sum_k x_k a_k over this CTA's user range -------------------------------
    if (warm) {
      for (int b = 0; b < nblk; b++) {
        const int p0 = b * BK, n = min(BK, na - p0);
        prefetch_meta(b, mc);
        if (tid < BK) {
          const double xi = tid < n ? x[p0 + tid] : 0.0;
          sm.dlt[tid] = fabs(xi) > kEps ? xi : 0.0;
        }
        __syncthreads();
        prefetch_ids(mc, tc);
        __syncthreads();
        update_yhat(mc, tc, n);
        __syncthreads();
      }
    }
    if (nblk > 0) {
      prefetch_meta(0, 0);
      __syncthreads();  // (the same warps fill both slots)
      prefetch_meta(nblk > 1 ? 1 : 0, 1);
    }
    __syncthreads();
    if (nblk > 0) {
      prefetch_ids(0, 0);
      prefetch_tile(0, 0, 0);
    }
    __syncthreads();

    // ---- the sweeps (cd.c:112-140) ---------------------------------------------------------------------
    if (timer) t_act = globaltimer_ns();
    int niters = 1;
    if (na > 0 && maxit > 0) {
      bool done = false;
      int t = 0;
      for (; t < maxit && !done; t++) {
        double dl = 0.0;  // warp 0: this lane's share of sum (x' - x)^2
        for (int b = 0; b < nblk; b++) {
          const HybMeta<GA> &M = sm.meta[mc];
          const int p0 = b * BK, n = min(BK, na - p0);
          const int bn1 = b + 1 == nblk ? 0 : b + 1, bn2 = bn1 + 1 == nblk ? 0 : bn1 + 1;
          const int mn1 = mc == 2 ? 0 : mc + 1, mn2 = mn1 == 2 ? 0 : mn1 + 1, tn = tc ^ 1;
          const bool prof_on = ha.prof != nullptr && rank == 0 && tid == 0;
          const int pheavy = ((M.bmask[0] | M.bmask[1] | M.bmask[2] | M.bmask[3] | M.wmask[0] | M.wmask[1] | M.wmask[2] |
                               M.wmask[3]) != 0u) ? 4 : 0;
          long long pt0 = 0, pt1 = 0, pt2 = 0, pt3 = 0;
          if (prof_on) pt0 = clock64();
          double xv[4] = {0.0, 0.0, 0.0, 0.0};
          if (warp == 0) {
#pragma unroll
            for (int s = 0; s < 4; s++)
              if (32 * s + lane < n) xv[s] = x[p0 + 32 * s + lane];
          }
          // ---- partial inner products <a_m, yhat> over this CTA's user range
          if (tid < n && M.cls[tid] <= 1) {
            double v = 0.0;
            if (M.cls[tid] == 1) {  // user ids staged in shared memory: one load latency
              const int len = M.s1[tid] - M.s0[tid];
              double y[kHybLane];
#pragma unroll
              for (int k = 0; k < kHybLane; k++) y[k] = k < len ? __ldcg(yh + sm.sid[tc][k][tid]) : 0.0;
#pragma unroll
              for (int k = 0; k < kHybLane; k++)
                v += HASVAL ? (double)sm.sval[HASVAL ? tc : 0][HASVAL ? k : 0][HASVAL ? tid : 0] * y[k] : y[k];
            }
            sm.mine[xb][tid] = v;
          }
          if ((M.gmask[0] | M.gmask[1] | M.gmask[2] | M.gmask[3]) != 0u) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
              const int m = (tid >> 3) + 64 * h;
              const bool on = m < n && M.cls[m] == 2;
              double v = on ? hyb_dot16<HASVAL, 8>(a, M.c0[m], M.s0[m], M.s1[m], lane & 7, yh) : 0.0;
              v += __shfl_xor_sync(0xffffffffu, v, 1);
              v += __shfl_xor_sync(0xffffffffu, v, 2);
              v += __shfl_xor_sync(0xffffffffu, v, 4);
              if (on && (lane & 7) == 0) sm.mine[xb][m] = v;
            }
          }
          if ((M.wmask[0] | M.wmask[1] | M.wmask[2] | M.wmask[3]) != 0u) {
            for (int m = warp; m < n; m += NW)
              if (M.cls[m] == 3) {
                double v = hyb_dot_range<HASVAL>(a, M.c0[m], M.s0[m], M.s1[m], 512, lane, yh);
#pragma unroll
                for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) sm.mine[xb][m] = v;
              }
          }
          const bool anybig = (M.bmask[0] | M.bmask[1] | M.bmask[2] | M.bmask[3]) != 0u;
          if (anybig) {
#pragma unroll
            for (int w4 = 0; w4 < 4; w4++) {
              unsigned mm = M.bmask[w4];
              while (mm) {
                const int m = 32 * w4 + __ffs(mm) - 1;
                mm &= mm - 1;
                double v = hyb_dot_range<HASVAL>(a, M.c0[m], M.s0[m] + warp * 512, M.s1[m], NW * 512, lane, yh);
#pragma unroll
                for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) sm.pwb[warp][m] = v;
              }
            }
            __syncthreads();
            if (tid < n && M.cls[tid] == 4) {
              double s = 0.0;
#pragma unroll
              for (int w = 0; w < NW; w++) s += sm.pwb[w][tid];
              sm.mine[xb][tid] = s;
            }
          }
          // ---- sum over the CTAs of the cluster (rank order: bit-identical everywhere)
          if (prof_on) pt1 = clock64();
          hyb_cluster_arrive();
          prefetch_meta(bn2, mn2);
          prefetch_ids(mn1, tn);
          prefetch_tile(bn1, mn1, tn);
          hyb_cluster_wait();
          if (tid < n) {
            double s = 0.0;
            for (int c = 0; c < cs; c++) s += hyb_ld_peer(&sm.mine[xb][tid], (uint32_t)c);
            sm.tot[tid] = s;
          }
          xb ^= 1u;
          if (warp < 4) hyb_bar_warps03();
          if (prof_on) pt2 = clock64();

          if (warp == 0) {
            // ---- exact sequential CD inside the block (cd.c:117-133); only coordinates whose value changes are visited
            const float(*T)[BK] = sm.tile[tc];
            double ipf[4], xn[4], sq[4], den[4], aty[4];
            bool valid[4];
#pragma unroll
            for (int s = 0; s < 4; s++) {
              const int m = 32 * s + lane;
              valid[s] = m < n;
              ipf[s] = valid[s] ? sm.tot[m] : 0.0;
              sq[s] = valid[s] ? M.sq[m] : 0.0;
              den[s] = valid[s] ? M.inv_den[m] : 1.0;
              aty[s] = valid[s] ? (double)M.aty[m] : 0.0;
              xn[s] = xv[s];
            }
            // Only a coordinate whose tile row has a nonzero after the diagonal ("emitter") can change a later inner
            // product: the emitters that change are visited one after the other, every other coordinate of the
            // sub-block takes its step in parallel afterwards (its ip then holds the steps of all earlier emitters,
            // T is applied above the diagonal only).  Unpopular items rarely co-occur: few emitters per block.
#pragma unroll
            for (int s = 0; s < 4; s++) {
              if (32 * s >= n) break;
              const unsigned em = sm.dep[tc][s];
              int k = 0;
              for (;;) {
                const double in_old = fabs(xn[s]) > kEps ? xn[s] : 0.0;
                const double ip = ipf[s] - in_old * sq[s];  // cd.c:122-123 in one step
                const double num = aty[s] - ip;
                const double nx = num > a.l1r ? (num - a.l1r) * den[s] : 0.0;
                const unsigned want = __ballot_sync(0xffffffffu, valid[s] && lane >= k && nx != xn[s]) & em;
                if (!want) {
                  if (valid[s] && !((em >> lane) & 1u) && nx != xn[s]) {
                    dl += (nx - xn[s]) * (nx - xn[s]);
                    xn[s] = nx;
                  }
                  break;
                }
                const int kk = __ffs(want) - 1;
                const double in_new = fabs(nx) > kEps ? nx : 0.0;
                const double d = __shfl_sync(0xffffffffu, in_new - in_old, kk);
                if (lane == kk) {
                  dl += (nx - xn[s]) * (nx - xn[s]);
                  xn[s] = nx;
                }
                if (d != 0.0) {
                  const float *row = T[32 * s + kk];
#pragma unroll
                  for (int s2 = 0; s2 < 4; s2++)
                    if (s2 > s || (s2 == s && lane > kk)) ipf[s2] = fma(d, (double)row[32 * s2 + lane], ipf[s2]);
                }
                k = kk + 1;
              }
            }
#pragma unroll
            for (int s = 0; s < 4; s++) {
              const int m = 32 * s + lane;
              const double was = fabs(xv[s]) > kEps ? xv[s] : 0.0;
              const double now = fabs(xn[s]) > kEps ? xn[s] : 0.0;
              if (valid[s] && xn[s] != xv[s]) x[p0 + m] = xn[s];
              sm.dlt[m] = valid[s] ? now - was : 0.0;
            }
          }
          __syncthreads();
          if (prof_on) pt3 = clock64();
          if (tid < 4) sm.dep[tc][tid] = 0u;  // (free again: the gather of the round after the next one sets its bits)
          update_yhat(mc, tc, n);
          __syncthreads();  // (yhat slices are private to a CTA: block-level visibility is all the next round needs)
          if (prof_on) {
            unsigned long long *pp = ha.prof + (size_t)q * 10;
            pp[pheavy + 0] += (unsigned long long)(pt1 - pt0);
            pp[pheavy + 1] += (unsigned long long)(pt2 - pt1);
            pp[pheavy + 2] += (unsigned long long)(pt3 - pt2);
            pp[pheavy + 3] += (unsigned long long)(clock64() - pt3);
            pp[8 + (pheavy >> 2)] += 1ull;
          }
          mc = mn1;
          tc = tn;
        }
        // ---- end of sweep: stop rule (cd.c:135-138)
        if (warp == 0) {
#pragma unroll
          for (int o = 16; o; o >>= 1) dl += __shfl_xor_sync(0xffffffffu, dl, o);
          if (lane == 0) sm.done = dl < a.opttol ? 1 : 0;
        }
        __syncthreads();
        done = sm.done != 0;
        __syncthreads();
      }
      niters = done ? t : maxit + 1;  // cd.c:140
    } else if (maxit > 0) {
      niters = (0.0 < a.opttol) ? 1 : maxit + 1;
    }
    if (timer) t_sweep = globaltimer_ns();

    // ---- residual (estimate.c:477-489): |y - yhat|^2 = sum_u yhat_u^2 + sum_{u in col j} (r_uj^2 - 2 r_uj yhat_u);
    //      the pass over the slice also clears it for the next target
    double jdot = 0.0, ssq = 0.0;
    {
      const int32_t *sp = ha.colsplit + (size_t)j * (kParts + 1);
      const int s0 = sp[pr0], s1 = sp[pr1];
      const int64_t cj0 = a.colptr[j];
      for (int e = s0 + tid; e < s1; e += NT) {
        const double r = HASVAL ? (double)a.colval[cj0 + e] : 1.0;
        jdot = fma(r, __ldcg(yh + a.colind[cj0 + e]), jdot);
      }
      __syncthreads();
      for (int u = u_lo + tid; u < u_hi; u += NT) {
        const double v = __ldcg(yh + u);
        if (v != 0.0) {
          ssq = fma(v, v, ssq);
          __stcg(yh + u, 0.0);
        }
      }
    }
    auto cta_sum = [&](double v) {
#pragma unroll
      for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      __syncthreads();
      if (lane == 0) sm.red[warp] = v;
      __syncthreads();
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < NW; w++) s += sm.red[w];
      return s;
    };
    jdot = cta_sum(jdot);
    ssq = cta_sum(ssq);
    if (tid == 0) {
      sm.mine[xb][0] = jdot;
      sm.mine[xb][1] = ssq;
    }
    __threadfence();
    hyb_cluster_sync();
    double jd_all = 0.0, ss_all = 0.0;
    if (rank == 0)
      for (int c = 0; c < cs; c++) {
        jd_all += hyb_ld_peer(&sm.mine[xb][0], (uint32_t)c);
        ss_all += hyb_ld_peer(&sm.mine[xb][1], (uint32_t)c);
      }
    xb ^= 1u;

    if (rank == 0) {
      double reg = 0.0;
      int nnz_local = 0;
      for (int p = tid; p < na; p += NT) {
        const double xv = x[p];
        reg += 0.5 * a.l2r * xv * xv + a.l1r * fabs(xv);
        nnz_local += fabs(xv) > kEps ? 1 : 0;
      }
      reg = cta_sum(reg);
      const int nnz_w = (int)(cta_sum((double)nnz_local) + 0.5);
      const double actnnz_t = cta_sum((double)actnnz);
      // ---- compaction |x| > EPS -> (i, (float)x) in visiting order (estimate.c:492-505)
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
        a.st_expand[q] = ha.expand ? (long long)ha.expand[j] : 0;
        const double rn = 0.5 * (ss_all - 2.0 * jd_all + a.csq[j]);
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
    __threadfence();
  }
  // no CTA may exit while a peer can still read its shared memory
  hyb_cluster_sync();
}
