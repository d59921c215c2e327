// gram_batch.cuh -- Gram-space coordinate descent for HEAVY target columns, T targets per cluster.
// Included by engine.cu after gram.cuh (same formulation; read that header first).
//
// A heavy target (tens of thousands of active coordinates, thousands of nonzero weights) spends its
// time in   <a_m, yhat> = sum_{k in S} x_k G[k][m]   : every block of coordinates re-reads one segment
// of every Gram row in S.  Heavy targets share most of S (the popular items), so this kernel solves T
// targets at once and works in ITEM space: a block is 32*V consecutive (internal) item ids, the same
// for every target, so one 32*V-wide segment of row k -- one coalesced vector load per lane -- feeds
// the T inner products (T fp64 FMAs per loaded element).  HBM traffic and load latency per target drop
// by up to T; the sweeps stay exact sequential CD per target:
//   * warp t owns target t's chain over the block (same chain as cd_gram_kernel, V coordinates per
//     lane, visiting order = ascending item id); coordinates that are not active for target t
//     (G[j_t][i] <= l1r, estimate.c:433-444) are masked out;
//   * the nonzero list S is the UNION over the T targets, one entry per item with T values.
// The CTAs of the cluster each take a share of the entries of S; the T*V*32 partial sums of a block
// are reduce-scattered and all-gathered through tagged DSMEM stores (no cluster barrier), after which
// every CTA runs the T chains redundantly on private copies of the state (deterministic: all CTAs
// stay bit-identical).
#pragma once

constexpr int kBatchGroup = 8;   // entries per cp.async commit group
constexpr int kBatchStages = 3;  // groups in a warp's ring (two in flight while one is consumed)

struct BatchArgs {
  double *xt;        // per CTA: [T][istride] current iterate by item
  double *sl_valT;   // per CTA: [istride][T] effective values of the nonzero-list entries
  uint32_t *amask;   // per CTA: [T][nwords] active-coordinate bits, word w = items 32w .. 32w+31
  uint32_t *anym;    // per CTA: [nwords] OR of amask over the targets of the batch
  size_t istride;    // >= ncols, multiple of 128 (whole panels of G)
  int32_t nwords;    // istride / 32
  int32_t profile;   // SLIMB200_PROFILE=1: thread 0 of rank 0 prints the cycle split of every batch
};

template <typename GT, int V>
struct GVecLoad;
template <>
struct GVecLoad<float, 1> {
  static __device__ __forceinline__ void ld(const float *p, float (&o)[1]) { o[0] = __ldg(p); }
};
template <>
struct GVecLoad<float, 2> {
  static __device__ __forceinline__ void ld(const float *p, float (&o)[2]) {
    const float2 v = __ldg(reinterpret_cast<const float2 *>(p));
    o[0] = v.x;
    o[1] = v.y;
  }
};
template <>
struct GVecLoad<float, 4> {
  static __device__ __forceinline__ void ld(const float *p, float (&o)[4]) {
    const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
    o[0] = v.x;
    o[1] = v.y;
    o[2] = v.z;
    o[3] = v.w;
  }
};
template <>
struct GVecLoad<double, 1> {
  static __device__ __forceinline__ void ld(const double *p, double (&o)[1]) { o[0] = __ldg(p); }
};
template <>
struct GVecLoad<double, 2> {
  static __device__ __forceinline__ void ld(const double *p, double (&o)[2]) {
    const double2 v = __ldg(reinterpret_cast<const double2 *>(p));
    o[0] = v.x;
    o[1] = v.y;
  }
};
template <>
struct GVecLoad<double, 4> {
  static __device__ __forceinline__ void ld(const double *p, double (&o)[4]) {
    const double2 v = __ldg(reinterpret_cast<const double2 *>(p));
    const double2 w = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    o[0] = v.x;
    o[1] = v.y;
    o[2] = w.x;
    o[3] = w.y;
  }
};

template <typename GT, int CS, int T, int V, int NTB>
struct __align__(16) BatchSmem {
  static constexpr int kBatchNW = NTB / 32;
  static constexpr int TV = T * V;
  static constexpr int OWN = CS > 1 ? TV / CS : 1;  // (target, v) pairs this CTA reduces
  struct Ring {  // per-warp cp.async staging: row segments + entry values
    GT g[kBatchStages * kBatchGroup][32][V];
    double v[kBatchStages * kBatchGroup][1][T];
  };
  union W {  // a warp's partial sums overwrite its OWN (drained) ring
    Ring ring;
    double part[T][V][32];
  } w[kBatchNW];
  GT gbb[32 * V][32 * V];
  unsigned long long rs[2][CS > 1 ? CS : 1][OWN][32][2];  // reduce-scatter slots  (tagged)
  unsigned long long ag[2][CS > 1 ? TV : 1][32][2];       // all-gather slots      (tagged)
  double red[2 * kBatchNW];
  int sc[kBatchNW];
  long long misc[4];
  uint32_t newmask[T][V];
  int target[T];   // internal id of target t, -1 when the batch is short
  int maxit[T];
  int done[T];     // 1 once target t stopped sweeping
  int niters[T];
  int len;         // entries of the union nonzero list
};

// add tagged peer store / wait from engine.cu: st_peer_tagged(), ld_tagged_wait()

template <typename GT, int CS, int T, int V, int NTB>
__global__ void __launch_bounds__(NTB, NTB <= 256 ? 2 : 1) cd_gram_batch_kernel(const SolveArgs a, const GramArgs ga,
                                                                               const BatchArgs ba) {
  constexpr int NT = NTB, NW = NTB / 32, TV = T * V, BW = 32 * V;  // BW = items per block
  static_assert(T <= NW, "one chain warp per target");
  static_assert(CS == 1 || (TV % CS == 0 && TV / CS <= NW && CS <= TV), "exchange layout");
  using Smem = BatchSmem<GT, CS, T, V, NTB>;
  extern __shared__ __align__(16) unsigned char batch_smem_raw[];
  Smem &sm = *reinterpret_cast<Smem *>(batch_smem_raw);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = CS > 1 ? gram_cluster_rank() : 0u;
  uint32_t tag = 0;
  int par = 0;

  const GT *__restrict__ G = static_cast<const GT *>(ga.G);
  const size_t nr = ga.nr;
  static_assert(32 * V == kGramPW, "a block of coordinates is one panel of G");
  const size_t istride = ba.istride;
  const int nwords = ba.nwords;
  const int nblk = (a.ncols + BW - 1) / BW;
  const size_t cta = (size_t)(ga.slot_base + blockIdx.x);
  double *xt = ba.xt + cta * T * istride;
  double *slv = ba.sl_valT + cta * istride * T;
  uint32_t *amask = ba.amask + cta * (size_t)T * nwords;
  uint32_t *anym = ba.anym + cta * (size_t)nwords;
  int32_t *slotp = ga.slotp + cta * istride;
  int32_t *sl_row = ga.sl_row + cta * istride;

  if (CS > 1) {
    for (int i = tid; i < (int)(sizeof(sm.rs) / 8); i += NT) (&sm.rs[0][0][0][0][0])[i] = 0ull;
    for (int i = tid; i < (int)(sizeof(sm.ag) / 8); i += NT) (&sm.ag[0][0][0][0])[i] = 0ull;
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }

  // cluster-wide sum of one double per lane contributed by warp 0 of every CTA (all-gather through ag[])
  // (every thread of the CTA advances `tag` before the call, warp 0 makes the call)
  auto allsum_w0 = [&](double v) -> double {
    if (CS == 1) return v;
    const int buf = tag & 1;
#pragma unroll
    for (int r = 0; r < CS; r++) st_peer_tagged(&sm.ag[buf][rank][lane][0], (uint32_t)r, v, tag);
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < CS; r++) s += ld_tagged_wait(&sm.ag[buf][r][lane][0], tag);
    return s;
  };

  for (;;) {
    // ---- next batch: T consecutive positions of the (descending nnz) target list ----------------
    __syncthreads();
    tag++;
    if (warp == 0) {
      double qv = 0.0;
      if (rank == 0) {
        int v = 0;
        if (lane == 0) v = atomicAdd(ga.queue, 1);
        qv = (double)__shfl_sync(0xffffffffu, v, 0);
      }
      qv = allsum_w0(qv);
      if (lane == 0) sm.misc[0] = (long long)qv;
    }
    __syncthreads();
    const int q0 = ga.q_begin + (int)sm.misc[0] * T;
    if (q0 >= ga.q_end) break;
    const int teff = min(T, ga.q_end - q0);
    const bool timer = rank == 0 && tid == 0;
    unsigned long long t_start = 0, t_act = 0, t_sweep = 0;
    if (timer) t_start = globaltimer_ns();
    if (tid < T) {
      const int j = tid < teff ? a.targets[q0 + tid] : -1;
      sm.target[tid] = j;
      const long long cap64 = j >= 0 ? 50LL * a.colcnt[j] : 0;  // estimate.c:448-449
      sm.maxit[tid] = (int)(cap64 < (long long)a.maxniters ? cap64 : (long long)a.maxniters);
      sm.done[tid] = j < 0 ? 1 : 0;
      sm.niters[tid] = 1;
    }
    if (tid == 0) sm.len = 0;
    // ---- reset the per-batch state
    for (size_t i = tid; i < (size_t)T * istride; i += NT) xt[i] = 0.0;
    for (size_t i = tid; i < istride; i += NT) slotp[i] = -1;
    __syncthreads();

    // ---- active sets: bit i of target t <=> G[j_t][i] > l1r, i != j_t (estimate.c:433-444) -------
    long long actnnz[T];
    int nact[T];
    bool anywarm = false;
#pragma unroll
    for (int t = 0; t < T; t++) {
      actnnz[t] = 0;
      nact[t] = 0;
      const int j = sm.target[t];
      if (j < 0) {
        for (int w = tid; w < nwords; w += NT) amask[(size_t)t * nwords + w] = 0u;
        continue;
      }
      for (int base = 0; base < (int)istride; base += NT) {
        const int i = base + tid;  // istride is a multiple of 128: all warps stay in range
        const bool f = i < a.ncols && i != j && (double)__ldg(G + gram_off(nr, j, i)) > a.l1r;
        const uint32_t m = __ballot_sync(0xffffffffu, f);
        if (lane == 0 && i < (int)istride) amask[(size_t)t * nwords + (i >> 5)] = m;
        if (f) {
          actnnz[t] += a.colcnt[i];
          nact[t]++;
        }
      }
      // warm start: column j of the initial model, active coordinates only (estimate.c:453-464)
      const int jo = a.inv[j];
      if (a.wcolptr != nullptr && jo < a.wncols) {
        anywarm = true;
        for (int64_t k = a.wcolptr[jo] + tid; k < a.wcolptr[jo + 1]; k += NT) {
          const int r = a.wcolind[k];
          if (r >= 0 && r < a.ncols) {
            const int i = a.rank[r];
            if (i != j && (double)__ldg(G + gram_off(nr, j, i)) > a.l1r) xt[(size_t)t * istride + i] = (double)a.wcolval[k];
          }
        }
      }
    }
    __syncthreads();
    for (int w = tid; w < nwords; w += NT) {
      uint32_t m = 0;
#pragma unroll
      for (int t = 0; t < T; t++) m |= amask[(size_t)t * nwords + w];
      anym[w] = m;
    }
    __syncthreads();
    if (anywarm) {  // union nonzero list of the starting iterate (ascending item id)
      int len = 0;
      for (int base = 0; base < a.ncols; base += NT) {
        const int i = base + tid;
        bool f = false;
        if (i < a.ncols) {
#pragma unroll
          for (int t = 0; t < T; t++) f |= fabs(xt[(size_t)t * istride + i]) > kEps;
        }
        int tot;
        const int pos = len + team_excl_scan<NT>(f, sm.sc, tot);
        if (f) {
          slotp[i] = pos;
          sl_row[pos] = i;
#pragma unroll
          for (int t = 0; t < T; t++) {
            const double xv = xt[(size_t)t * istride + i];
            slv[(size_t)pos * T + t] = fabs(xv) > kEps ? xv : 0.0;
          }
        }
        len += tot;
      }
      __syncthreads();
      if (tid == 0) sm.len = len;
      __syncthreads();
    }
    if (timer) t_act = globaltimer_ns();

    // ---- one pass over the item blocks: SWEEP = the CD sweep, otherwise only hh_t = <yhat_t, yhat_t>
    double hh[T];
#pragma unroll
    for (int t = 0; t < T; t++) hh[t] = 0.0;
    int nvisited = 0;
    long long pc[6] = {0, 0, 0, 0, 0, 0};  // cycles: gather, wait, exchange, chain, append/publish, rounds
    long long pt = 0;
#define SLIM_PT(k)                          \
  if (ba.profile && tid == 0) {             \
    const long long now_ = clock64();       \
    pc[k] += now_ - pt;                     \
    pt = now_;                              \
  }

    auto gather = [&](int b, double (&acc)[T][V]) {
      // this warp's share of  sum_{e < len} val[e][t] * G[row[e]][block b]  for the lane's V items.
      // The row segments (32*V elements, one per entry) and the T values of each entry are streamed
      // through a per-warp shared-memory ring with cp.async (LDGSTS): kBatchStages groups of
      // kBatchGroup entries, all but one in flight while one is consumed.  A lane reads back exactly
      // the bytes it copied, so only the values need a warp-level hand-over.
      constexpr int GS = kBatchGroup, NG = kBatchStages, GPC = 32 / GS;  // GPC groups per 32-entry chunk
      constexpr int LB = V * (int)sizeof(GT);                           // bytes per lane and entry
      static_assert(LB == 8 || LB == 16, "cp.async element size");
      const GT *__restrict__ Gblk = G + gram_off(nr, 0, b * BW) + (size_t)lane * V;  // panel b, row r at + r * PW
      const int len = sm.len;
      const int first = (int)rank * NW + warp, stride = CS * NW;
#pragma unroll
      for (int t = 0; t < T; t++)
#pragma unroll
        for (int v = 0; v < V; v++) acc[t][v] = 0.0;
      const int nchunk_all = (len + 31) >> 5;
      const int nch = nchunk_all > first ? (nchunk_all - first + stride - 1) / stride : 0;  // my chunks
      const int ngr = nch * GPC;
      typename Smem::Ring &ring = sm.w[warp].ring;
      const uint32_t g_base = smem_u32(&ring.g[0][0][0]) + (uint32_t)lane * LB;
      const uint32_t v_base = smem_u32(&ring.v[0][0][0]) + (uint32_t)lane * 8u;
      int row_cur = 0, row_nxt = 0;
      if (nch > 0) {
        const int e = first * 32 + lane;
        row_nxt = e < len ? sl_row[e] : 0;
      }
      // loop-invariant pieces of the addresses, kept in registers (the compiler otherwise rebuilds them per entry)
      const char *const g_src0 = reinterpret_cast<const char *>(Gblk);
      constexpr uint32_t kRowBytes = (uint32_t)(kGramPW * sizeof(GT));  // one row of the panel
      constexpr uint32_t kSlotG = GS * 32 * LB, kSlotV = GS * T * 8;     // ring bytes per group
      // `slot` is a compile-time constant at every call site (the main loop is unrolled over the ring),
      // so all shared-memory addresses are base + immediate.  Per entry: one shuffle (row id), one
      // mad.wide (row * row bytes + panel base) and one LDGSTS; the 8 x T values of a whole group are
      // contiguous in the list (512 B) and travel with ONE 16-byte cp.async per lane.
      unsigned long long g_src_u64 = reinterpret_cast<unsigned long long>(g_src0);
      asm volatile("" : "+l"(g_src_u64));  // keep the panel base in registers (do not rebuild it per entry)
      static_assert(GS * T * 8 == 32 * 16, "one 16-byte cp.async per lane moves the values of a group");
      const char *const v_grp0 = reinterpret_cast<const char *>(slv) + (size_t)lane * 16;
      const uint32_t v_base16 = smem_u32(&ring.v[0][0][0]) + (uint32_t)lane * 16u;
      auto issue = [&](int g, const int slot) {
        if (g < ngr) {
          const int k = g / GPC, i0 = (g % GPC) * GS;
          const int c = first + k * stride;
          if (i0 == 0) {
            row_cur = row_nxt;
            const int e = (c + stride) * 32 + lane;  // rows of my next chunk, one chunk ahead
            row_nxt = e < len ? sl_row[e] : 0;
          }
          const uint32_t gd = g_base + (uint32_t)slot * kSlotG;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(v_base16 + (uint32_t)slot * kSlotV),
                       "l"(v_grp0 + (size_t)(c * 32 + i0) * (T * 8))
                       : "memory");
#pragma unroll
          for (int u = 0; u < GS; u++) {
            const uint32_t r = (uint32_t)__shfl_sync(0xffffffffu, row_cur, i0 + u);
            unsigned long long src;
            asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(src) : "r"(r), "r"(kRowBytes), "l"(g_src_u64));
            if (LB == 8)
              asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(gd + (uint32_t)(u * 32 * LB)), "l"(src) : "memory");
            else
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(gd + (uint32_t)(u * 32 * LB)), "l"(src) : "memory");
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      const GT *const ring_g = &ring.g[0][lane][0];  // this lane's V elements of ring entry 0
      const double *const ring_v = &ring.v[0][0][0];
      auto consume_entry = [&](const int se) {  // se: compile-time ring entry
        GT gv[V];
#pragma unroll
        for (int v = 0; v < V; v++) gv[v] = ring_g[se * 32 * V + v];
#pragma unroll
        for (int t = 0; t < T; t++) {
          const double val = ring_v[se * T + t];
#pragma unroll
          for (int v = 0; v < V; v++) acc[t][v] = fma(val, (double)gv[v], acc[t][v]);
        }
      };
#pragma unroll
      for (int p = 0; p < NG - 1; p++) issue(p, p);
      int e0 = first * 32;  // first entry of group g
      for (int gb = 0; gb < ngr; gb += NG) {
#pragma unroll
        for (int sidx = 0; sidx < NG; sidx++) {
          const int g = gb + sidx;
          if (g < ngr) {  // warp-uniform
            issue(g + NG - 1, (sidx + NG - 1) % NG);
            asm volatile("cp.async.wait_group %0;" ::"n"(NG - 1) : "memory");
            __syncwarp();
            if (e0 + GS <= len) {  // whole group valid (warp-uniform): no per-entry tests
#pragma unroll
              for (int u = 0; u < GS; u++) consume_entry(sidx * GS + u);
            } else {
#pragma unroll
              for (int u = 0; u < GS; u++)
                if (e0 + u < len) consume_entry(sidx * GS + u);
            }
            __syncwarp();
            e0 += GS;
            if ((g + 1) % GPC == 0) e0 += (stride - 1) * 32;  // next chunk of this warp
          }
        }
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
    };

    // ---- the sweeps (cd.c:112-140) -------------------------------------------------------------
    for (int sweep = 0;; sweep++) {
      // which targets still sweep?  (uniform over the CTA and over the cluster)
      bool any = false;
#pragma unroll
      for (int t = 0; t < T; t++) any |= (sm.done[t] == 0 && sweep < sm.maxit[t]);
      if (!any) break;
      double dl = 0.0;  // warp t: this lane's share of sum (x' - x)^2 of target t
      const bool mine_live = warp < T && sm.done[warp < T ? warp : 0] == 0 && sweep < sm.maxit[warp < T ? warp : 0];
      __syncthreads();
      for (int b = 0; b < nblk; b++) {
        uint32_t anyw = 0;
#pragma unroll
        for (int v = 0; v < V; v++) anyw |= anym[b * V + v];
        if (anyw == 0) continue;  // no target has an active coordinate in this block
        if (sweep == 0) nvisited++;
        const int item0 = b * BW + lane * V;  // the lane's first item

        // chain operands of warp t, requested before the gather
        double xv[V], sq[V], den[V], aty[V];
        bool act[V];
        const int myj = warp < T ? sm.target[warp] : -1;
#pragma unroll
        for (int v = 0; v < V; v++) {
          xv[v] = 0.0;
          sq[v] = 0.0;
          den[v] = 1.0;
          aty[v] = 0.0;
          act[v] = false;
        }
        if (mine_live) {
          GT gj[V];
          GVecLoad<GT, V>::ld(G + gram_off(nr, myj, item0), gj);
#pragma unroll
          for (int v = 0; v < V; v++) {
            const int i = item0 + v;
            act[v] = (amask[(size_t)warp * nwords + (i >> 5)] >> (i & 31)) & 1u;
            if (i < a.ncols) {
              const double cn = (double)__ldg(a.cnorms + i);
              den[v] = cn * cn + a.l2r;
              sq[v] = __ldg(a.csq + i);
            }
            aty[v] = (double)(float)(double)gj[v];  // gk_fkv_t.key is a float (estimate.c:437)
            xv[v] = xt[(size_t)warp * istride + i];
          }
        }
        // in-block Gram rows (shared by the T chains): warp w stages rows w, w+NW, ...
        for (int r = warp; r < BW; r += NW) {
          GT row[V];
#pragma unroll
          for (int v = 0; v < V; v++) row[v] = (GT)0;
          if (b * BW + r < a.ncols) GVecLoad<GT, V>::ld(G + gram_off(nr, b * BW + r, item0), row);
#pragma unroll
          for (int v = 0; v < V; v++) sm.gbb[r][lane * V + v] = row[v];
        }
        double acc[T][V];
        if (ba.profile && tid == 0) pt = clock64();
        gather(b, acc);
        SLIM_PT(0)
#pragma unroll
        for (int t = 0; t < T; t++)
#pragma unroll
          for (int v = 0; v < V; v++) sm.w[warp].part[t][v][lane] = acc[t][v];
        __syncthreads();
        SLIM_PT(1)

        // reduce over the warps (and the CTAs of the cluster): warp t ends with target t's V sums
        double ipf[V];
        if (warp < T) {
#pragma unroll
          for (int v = 0; v < V; v++) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < NW; w++) s += sm.w[w].part[warp][v][lane];
            ipf[v] = s;
          }
        }
        if (CS > 1) {
          tag++;
          const int buf = tag & 1;
          constexpr int OWN = Smem::OWN;
          if (warp < T) {
#pragma unroll
            for (int v = 0; v < V; v++) {
              const int p = warp * V + v;  // pair id; owner CTA = p % CS, slot p / CS
              st_peer_tagged(&sm.rs[buf][rank][p / CS][lane][0], (uint32_t)(p % CS), ipf[v], tag);
            }
          }
          if (warp < OWN) {  // this CTA reduces pair p = warp * CS + rank and broadcasts the sum
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < CS; r++) s += ld_tagged_wait(&sm.rs[buf][r][warp][lane][0], tag);
            const int p = warp * CS + (int)rank;
#pragma unroll
            for (int r = 0; r < CS; r++) st_peer_tagged(&sm.ag[buf][p][lane][0], (uint32_t)r, s, tag);
          }
          if (warp < T) {
#pragma unroll
            for (int v = 0; v < V; v++) ipf[v] = ld_tagged_wait(&sm.ag[buf][warp * V + v][lane][0], tag);
          }
        }

        SLIM_PT(2)
        // ---- chains: warp t, exact sequential CD over the block in ascending item order ---------
        double xn[V];
        uint32_t appm[V];
#pragma unroll
        for (int v = 0; v < V; v++) {
          xn[v] = xv[v];
          appm[v] = 0u;
        }
        if (mine_live) {
          int kpos = 0;  // coordinates with order index lane * V + v >= kpos are still to be visited
          for (;;) {
            double nx[V], in_old[V];
            uint32_t wm[V];
#pragma unroll
            for (int v = 0; v < V; v++) {
              in_old[v] = fabs(xn[v]) > kEps ? xn[v] : 0.0;
              const double ip = ipf[v] - in_old[v] * sq[v];
              const double num = aty[v] - ip;
              nx[v] = num > a.l1r ? (num - a.l1r) / den[v] : 0.0;
              wm[v] = __ballot_sync(0xffffffffu, act[v] && (lane * V + v >= kpos) && nx[v] != xn[v]);
            }
            int ostar = 1 << 30;
#pragma unroll
            for (int v = 0; v < V; v++)
              if (wm[v]) ostar = min(ostar, (__ffs(wm[v]) - 1) * V + v);
            if (ostar == (1 << 30)) break;
            const int ls = ostar / V, vs = ostar % V;
            double dloc = 0.0;
#pragma unroll
            for (int v = 0; v < V; v++) {
              if (v == vs) {
                const double in_new = fabs(nx[v]) > kEps ? nx[v] : 0.0;
                dloc = in_new - in_old[v];
                if (lane == ls) {
                  dl += (nx[v] - xn[v]) * (nx[v] - xn[v]);
                  xn[v] = nx[v];
                }
              }
            }
            const double d = __shfl_sync(0xffffffffu, dloc, ls);
            if (d != 0.0) {
#pragma unroll
              for (int v = 0; v < V; v++) ipf[v] = fma(d, (double)sm.gbb[ostar][lane * V + v], ipf[v]);
            }
            kpos = ostar + 1;
          }
          // write back x; coordinates that become nonzero for the first time need a list entry
#pragma unroll
          for (int v = 0; v < V; v++) {
            const double was = fabs(xv[v]) > kEps ? xv[v] : 0.0;
            const double now = fabs(xn[v]) > kEps ? xn[v] : 0.0;
            if (xn[v] != xv[v]) xt[(size_t)warp * istride + item0 + v] = xn[v];
            const bool chg = act[v] && now != was;
            appm[v] = __ballot_sync(0xffffffffu, chg && slotp[item0 + v] < 0);
          }
        }
        if (warp < T && lane == 0) {
#pragma unroll
          for (int v = 0; v < V; v++) sm.newmask[warp][v] = appm[v];
        }
        SLIM_PT(3)
        __syncthreads();
        // ---- warp 0 appends the new entries (union over targets) in ascending item order ----------
        if (warp == 0) {
          uint32_t um[V];
          int before = 0;
#pragma unroll
          for (int v = 0; v < V; v++) {
            uint32_t m = 0;
#pragma unroll
            for (int t = 0; t < T; t++) m |= sm.newmask[t][v];
            um[v] = m;
          }
          int total = 0;
#pragma unroll
          for (int v = 0; v < V; v++) {
            total += __popc(um[v]);
            before += __popc(um[v] & ((1u << lane) - 1u));
          }
          if (total) {
            const int len = sm.len;
            int mine = 0;
#pragma unroll
            for (int v = 0; v < V; v++) {
              if ((um[v] >> lane) & 1u) {
                const int pos = len + before + mine;
                slotp[item0 + v] = pos;
                sl_row[pos] = item0 + v;
#pragma unroll
                for (int t = 0; t < T; t++) slv[(size_t)pos * T + t] = 0.0;
                mine++;
              }
            }
            if (lane == 0) sm.len = len + total;
          }
        }
        __syncthreads();
        // ---- chain warps publish the new effective values ---------------------------------------
        if (mine_live) {
#pragma unroll
          for (int v = 0; v < V; v++) {
            const double was = fabs(xv[v]) > kEps ? xv[v] : 0.0;
            const double now = fabs(xn[v]) > kEps ? xn[v] : 0.0;
            if (act[v] && now != was) slv[(size_t)slotp[item0 + v] * T + warp] = now;
          }
        }
        __syncthreads();
        SLIM_PT(4)
        pc[5]++;
      }
      // ---- end of sweep: stop rule per target (cd.c:135-138)
      if (warp < T) {
#pragma unroll
        for (int o = 16; o; o >>= 1) dl += __shfl_xor_sync(0xffffffffu, dl, o);
        if (lane == 0 && mine_live) {
          if (dl < a.opttol) {
            sm.done[warp] = 1;
            sm.niters[warp] = sweep + 1;
          } else if (sweep + 1 >= sm.maxit[warp]) {
            sm.done[warp] = 1;
            sm.niters[warp] = sm.maxit[warp] + 1;  // cd.c:140 when the cap is hit
          }
        }
      }
      __syncthreads();
    }
    if (tid < T && sm.target[tid] >= 0 && sm.maxit[tid] <= 0) sm.niters[tid] = 1;
    if (timer) t_sweep = globaltimer_ns();
    if (ba.profile && rank == 0 && tid == 0)
      printf("[batch q0=%d teff=%d len=%d] rounds %lld  cycles/round: gather %lld  wait %lld  exchange %lld  chain %lld  "
             "append+publish %lld\n", q0, teff, sm.len, pc[5], pc[0] / max(pc[5], 1LL), pc[1] / max(pc[5], 1LL),
             pc[2] / max(pc[5], 1LL), pc[3] / max(pc[5], 1LL), pc[4] / max(pc[5], 1LL));
#undef SLIM_PT

    // ---- residual / objective (estimate.c:477-489): hh_t = sum_i x_t[i] <a_i, yhat_t> ------------
    for (int b = 0; b < nblk; b++) {
      uint32_t anyw = 0;
#pragma unroll
      for (int v = 0; v < V; v++) anyw |= anym[b * V + v];
      if (anyw == 0) continue;
      double acc[T][V];
      gather(b, acc);
      const int item0 = b * BW + lane * V;
#pragma unroll
      for (int t = 0; t < T; t++)
#pragma unroll
        for (int v = 0; v < V; v++) {
          const double xv = xt[(size_t)t * istride + item0 + v];
          const double in = fabs(xv) > kEps ? xv : 0.0;
          hh[t] = fma(in, acc[t][v], hh[t]);
        }
    }
    double hh_tot[T];
#pragma unroll
    for (int t = 0; t < T; t++) hh_tot[t] = team_sum<NT>(hh[t], sm.red, par);
    if (CS > 1) {
      tag++;
      if (warp == 0) {
        double mine = 0.0;
#pragma unroll
        for (int t = 0; t < T; t++)
          if (lane == t) mine = hh_tot[t];
        const double tot = allsum_w0(mine);
        if (lane < T) sm.w[0].part[0][0][lane] = tot;
      }
      __syncthreads();
#pragma unroll
      for (int t = 0; t < T; t++) hh_tot[t] = sm.w[0].part[0][0][t];
      __syncthreads();
    }

    if (rank == 0) {
#pragma unroll 1
      for (int t = 0; t < teff; t++) {
        const int j = sm.target[t];
        const int q = q0 + t;
        const double *x = xt + (size_t)t * istride;
        double yd = 0.0, reg = 0.0;
        int nnz_local = 0;
        for (int i = tid; i < a.ncols; i += NT) {
          const double xv = x[i];
          if (xv != 0.0) {
            const double in = fabs(xv) > kEps ? xv : 0.0;
            yd = fma(in, (double)__ldg(G + gram_off(nr, j, i)), yd);
            reg += 0.5 * a.l2r * xv * xv + a.l1r * fabs(xv);
            nnz_local += in != 0.0 ? 1 : 0;
          }
        }
        yd = team_sum<NT>(yd, sm.red, par);
        reg = team_sum<NT>(reg, sm.red, par);
        const int nnz_w = (int)(team_sum<NT>((double)nnz_local, sm.red, par) + 0.5);
        double an = 0.0, na = 0.0;
#pragma unroll
        for (int tt = 0; tt < T; tt++)
          if (tt == t) {
            an = (double)actnnz[tt];
            na = (double)nact[tt];
          }
        const double actnnz_t = team_sum<NT>(an, sm.red, par);
        const int na_t = (int)(team_sum<NT>(na, sm.red, par) + 0.5);
        double hh_t = 0.0;
#pragma unroll
        for (int tt = 0; tt < T; tt++)
          if (tt == t) hh_t = hh_tot[tt];

        // compaction |x| > EPS -> (i, (float)x) in ascending internal id (estimate.c:492-505)
        if (tid == 0) sm.misc[1] = (long long)atomicAdd(a.pool_used, (unsigned long long)nnz_w);
        __syncthreads();
        const long long off = sm.misc[1];
        const bool fits = off + nnz_w <= a.pool_cap;
        if (fits) {
          int w0 = 0;
          for (int base = 0; base < a.ncols; base += NT) {
            const int i = base + tid;
            double xv = 0.0;
            if (i < a.ncols) xv = x[i];
            const bool flag = (i < a.ncols) && fabs(xv) > kEps;
            int tot;
            const int pos = w0 + team_excl_scan<NT>(flag, sm.sc, tot);
            if (flag) {
              a.pool_idx[off + pos] = a.inv[i];
              a.pool_val[off + pos] = (float)xv;
            }
            w0 += tot;
          }
        }
        if (tid == 0) {
          a.out_cnt[q] = fits ? nnz_w : -1 - nnz_w;
          a.out_off[q] = off;
          a.st_niters[q] = sm.niters[t];
          a.st_nactive[q] = na_t;
          a.st_actnnz[q] = (long long)(actnnz_t + 0.5);
          a.st_expand[q] = ga.expand ? (long long)ga.expand[j] : 0;
          const double yy = a.csq[j];
          const double rn = 0.5 * (yy - 2.0 * yd + hh_t);
          a.st_rnorm[q] = rn;
          a.st_obj[q] = rn + reg;
          a.st_ngroups[q] = nvisited;
          const unsigned long long t_end = globaltimer_ns();
          // the batch shares its time: every target reports the batch's phases divided by the batch size
          a.st_phase[(size_t)q * 4 + 0] = 0.f;
          a.st_phase[(size_t)q * 4 + 1] = (float)(t_act - t_start) * 1e-3f / teff;
          a.st_phase[(size_t)q * 4 + 2] = (float)(t_sweep - t_act) * 1e-3f / teff;
          a.st_phase[(size_t)q * 4 + 3] = (float)(t_end - t_sweep) * 1e-3f / teff;
        }
        __syncthreads();
      }
    }
  }
  if (CS > 1) {
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}
