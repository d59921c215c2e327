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
  int32_t use_mma;   // the gather runs on the fp64 tensor-core path (DMMA m8n8k4) instead of scalar DFMAs
};

// Row segments in the staging ring are kRingPad bytes apart from their natural stride so that the four entries a
// DMMA B-fragment load touches (four consecutive ring rows, same 8 items) fall into different banks.
constexpr int kRingPad = 32;

template <typename GA, int CS, int T, int V, int NTB>
struct __align__(16) BatchSmem {
  using GT = typename GA::Tile;
  static constexpr int kBatchNW = NTB / 32;
  static constexpr int TV = T * V;
  static constexpr int OWN = CS > 1 ? TV / CS : 1;  // (target, v) pairs this CTA reduces
  struct Ring {  // per-warp cp.async staging: row segments (one panel row per entry, at most kMaxRowBytes) + entry values
    unsigned char g[kBatchStages * kBatchGroup][GA::kMaxRowBytes + kRingPad];
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

template <typename GA, int CS, int T, int V, int NTB, bool MMA>
__global__ void __launch_bounds__(NTB, NTB <= 256 ? 2 : 1) cd_gram_batch_kernel(const SolveArgs a, const GramArgs ga,
                                                                               const BatchArgs ba) {
  constexpr int NT = NTB, NW = NTB / 32, TV = T * V, BW = 32 * V;  // BW = items per block
  static_assert(T <= NW, "one chain warp per target");
  static_assert(CS == 1 || (TV % CS == 0 && TV / CS <= NW && CS <= TV), "exchange layout");
  static_assert(V == 2, "a lane owns two adjacent items of the 64-item panel");
  using Smem = BatchSmem<GA, CS, T, V, NTB>;
  using GT = typename GA::Tile;
  extern __shared__ __align__(16) unsigned char batch_smem_raw[];
  Smem &sm = *reinterpret_cast<Smem *>(batch_smem_raw);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = CS > 1 ? gram_cluster_rank() : 0u;
  uint32_t tag = 0;
  int par = 0;

  const GramView &gv = ga.gv;
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
        const bool f = i < a.ncols && i != j && GA::at(gv, j, i) > a.l1r;
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
            if (i != j && GA::at(gv, j, i) > a.l1r) xt[(size_t)t * istride + i] = (double)a.wcolval[k];
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

    // this warp's share of  sum_{e < len} val[e][t] * G[row[e]][block b]  for the lane's V items.
    // The row segments (one 64-item panel row per entry: 64 x W bytes, W = element width of the panel) and the T
    // values of each entry are streamed through a per-warp shared-memory ring with cp.async (LDGSTS): kBatchStages
    // groups of kBatchGroup entries, all but one in flight while one is consumed.
    auto gather_w = [&](int b, double (&acc)[T][V], auto wtag) {
      constexpr int W = decltype(wtag)::value;                           // bytes per element: 1, 2, 4 (packed) or 8 (fp64)
      constexpr int GS = kBatchGroup, NG = kBatchStages, GPC = 32 / GS;  // GPC groups per 32-entry chunk
      constexpr int RB = kGramPW * W;                                    // bytes of one panel row
      constexpr int RS = RB + (W == 4 || W == 8 ? 32 : 16);              // ring stride of an entry (bank spreading)
      static_assert(RS <= GA::kMaxRowBytes + kRingPad, "ring row");
      constexpr int LB = RB / 32;                                        // bytes a lane consumes per entry
      const unsigned char *const panel = GA::panel(gv, b * BW);           // row r of the panel at + r * RB
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
      // copy layout: W >= 2: every lane copies the LB bytes it will consume; W == 1 (64-byte rows): lanes 0-15 copy
      // one entry and lanes 16-31 the next one, 4 bytes each (cp.async moves at least 4 bytes)
      constexpr int CB = LB < 4 ? 4 : LB;                    // bytes per lane and cp.async
      constexpr int EPI = (32 * CB) / RB;                    // entries per cp.async instruction (1, or 2 for W == 1)
      const int sub = EPI == 2 ? (lane >> 4) : 0;            // which of the EPI entries this lane copies
      const uint32_t lane_off = EPI == 2 ? (uint32_t)(lane & 15) * CB : (uint32_t)lane * CB;
      // dst: entries are RS apart; with two entries per instruction the upper half-warp writes the next entry
      const uint32_t g_base = smem_u32(&ring.g[0][0]) + (EPI == 2 ? (uint32_t)sub * RS + (uint32_t)(lane & 15) * CB
                                                                  : (uint32_t)lane * CB);
      int row_cur = 0, row_nxt = 0;
      if (nch > 0) {
        const int e = first * 32 + lane;
        row_nxt = e < len ? sl_row[e] : 0;
      }
      constexpr uint32_t kSlotG = GS * RS, kSlotV = GS * T * 8;  // ring bytes per group
      // `slot` is a compile-time constant at every call site (the main loop is unrolled over the ring), so all
      // shared-memory addresses are base + immediate.  Per copy: one shuffle (row id), one mad.wide (row * row
      // bytes + panel base) and one LDGSTS; the 8 x T values of a whole group are contiguous in the list (512 B) and
      // travel with ONE 16-byte cp.async per lane.
      unsigned long long g_src_u64 = reinterpret_cast<unsigned long long>(panel) + lane_off;
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
          for (int u = 0; u < GS; u += EPI) {
            const uint32_t r = (uint32_t)__shfl_sync(0xffffffffu, row_cur, i0 + u + sub);
            unsigned long long src;
            asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(src) : "r"(r), "r"((uint32_t)RB), "l"(g_src_u64));
            const uint32_t dst = gd + (uint32_t)(u * RS);
            if (CB == 4)
              asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
            else if (CB == 8)
              asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
            else
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      const unsigned char *const ring_g = &ring.g[0][0] + lane * LB;  // this lane's V elements of ring entry 0
      const double *const ring_v = &ring.v[0][0][0];
      auto consume_entry = [&](const int se) {  // se: compile-time ring entry
        double gd2[V];
        if (W == 8) {
          const double2 q = *reinterpret_cast<const double2 *>(ring_g + se * RS);
          gd2[0] = q.x;
          gd2[1] = q.y;
        } else if (W == 4) {
          const uint2 q = *reinterpret_cast<const uint2 *>(ring_g + se * RS);
          gd2[0] = (double)q.x;
          gd2[1] = (double)q.y;
        } else if (W == 2) {
          const uint32_t q = *reinterpret_cast<const uint32_t *>(ring_g + se * RS);
          gd2[0] = (double)(q & 0xffffu);
          gd2[1] = (double)(q >> 16);
        } else {
          const uint32_t q = (uint32_t)*reinterpret_cast<const unsigned short *>(ring_g + se * RS);
          gd2[0] = (double)(q & 0xffu);
          gd2[1] = (double)(q >> 8);
        }
#pragma unroll
        for (int t = 0; t < T; t++) {
          const double val = ring_v[se * T + t];
#pragma unroll
          for (int v = 0; v < V; v++) acc[t][v] = fma(val, gd2[v], acc[t][v]);
        }
      };
      // Tensor-core form of the same sums: one DMMA m8n8k4 multiplies the 8 x 4 block val[target][entry] (A, one
      // double per lane: target = lane / 4, entry = lane % 4) with the 4 x 8 block G[entry][item] (B, one element per
      // lane: entry = lane % 4, item = 8 n + lane / 4) and accumulates the 8 x 8 block of partial sums (two doubles
      // per lane: target = lane / 4, items 8 n + 2 (lane % 4) + {0, 1}).  A block of 64 items is eight such column
      // tiles; four ring entries are one k-step.  Same products and the same fp64 accumulation as the scalar form,
      // 1/8 of the issue slots.
      static_assert(!MMA || (T == 8 && V == 2), "DMMA tiling: 8 targets x 64 items");
      double dm[8][2];
      if (MMA) {
#pragma unroll
        for (int n = 0; n < 8; n++) dm[n][0] = dm[n][1] = 0.0;
      }
      const int kq = lane & 3, cq = lane >> 2;
      const unsigned char *const ring_b = &ring.g[0][0] + kq * RS + cq * W;  // B fragment of ring entry 0, tile 0
      const double *const ring_a = &ring.v[0][0][0] + kq * T + cq;           // A fragment of ring entry 0
      auto consume_kstep = [&](const int se0, const bool valid) {  // se0: compile-time first ring entry of the k-step
        const double av = valid ? ring_a[se0 * T] : 0.0;
#pragma unroll
        for (int n = 0; n < 8; n++) {
          double bv;
          const unsigned char *src = ring_b + se0 * RS + n * 8 * W;
          if (W == 8) bv = *reinterpret_cast<const double *>(src);
          else if (W == 4) bv = (double)*reinterpret_cast<const uint32_t *>(src);
          else if (W == 2) bv = (double)*reinterpret_cast<const unsigned short *>(src);
          else bv = (double)*src;
          if (!valid) bv = 0.0;
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                       : "+d"(dm[n][0]), "+d"(dm[n][1])
                       : "d"(av), "d"(bv));
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
            if (MMA) {
              if (e0 + GS <= len) {  // whole group valid (warp-uniform)
#pragma unroll
                for (int u = 0; u < GS; u += 4) consume_kstep(sidx * GS + u, true);
              } else {
#pragma unroll
                for (int u = 0; u < GS; u += 4)
                  if (e0 + u < len) consume_kstep(sidx * GS + u, e0 + u + kq < len);
              }
            } else if (e0 + GS <= len) {  // whole group valid (warp-uniform): no per-entry tests
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
      if (MMA) {
        // hand the tile fragments over in the layout the rest of the kernel expects: acc[t][v] of lane L is the sum
        // for target t and item 2 L + v.  This lane holds target cq, items 8 n + 2 kq + {0, 1} = 2 (4 n + kq) + v.
        // The exchange goes through the warp's own (drained) ring.
        double *xb = reinterpret_cast<double *>(&ring.g[0][0]);  // [T][32][V] doubles = 4 KB
        __syncwarp();
#pragma unroll
        for (int n = 0; n < 8; n++) {
          xb[(cq * 32 + 4 * n + kq) * V + 0] = dm[n][0];
          xb[(cq * 32 + 4 * n + kq) * V + 1] = dm[n][1];
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < T; t++) {
          const double2 q = *reinterpret_cast<const double2 *>(xb + (t * 32 + lane) * V);
          acc[t][0] = q.x;
          acc[t][1] = q.y;
        }
        __syncwarp();
      }
    };
    auto gather = [&](int b, double (&acc)[T][V]) {
      if constexpr (GA::kMaxRowBytes == kGramPW * 8) {
        gather_w(b, acc, std::integral_constant<int, 8>());
      } else {
        const int rb = GA::row_bytes(gv, b * BW);  // uniform over the block: the ranges are whole panels
        if (rb == kGramPW * 4) gather_w(b, acc, std::integral_constant<int, 4>());
        else if (rb == kGramPW * 2) gather_w(b, acc, std::integral_constant<int, 2>());
        else gather_w(b, acc, std::integral_constant<int, 1>());
      }
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
          double gj[V];
          GA::at2(gv, myj, item0, gj);
#pragma unroll
          for (int v = 0; v < V; v++) {
            const int i = item0 + v;
            act[v] = (amask[(size_t)warp * nwords + (i >> 5)] >> (i & 31)) & 1u;
            if (i < a.ncols) {
              const double cn = (double)__ldg(a.cnorms + i);
              den[v] = 1.0 / (cn * cn + a.l2r);  // reciprocal (cd.c:127), off the chain's critical path
              sq[v] = __ldg(a.csq + i);
            }
            aty[v] = (double)(float)gj[v];  // gk_fkv_t.key is a float (estimate.c:437)
            xv[v] = xt[(size_t)warp * istride + i];
          }
        }
        // in-block Gram rows (shared by the T chains): warp w stages rows w, w+NW, ...
        for (int r = warp; r < BW; r += NW) {
          double row[V];
#pragma unroll
          for (int v = 0; v < V; v++) row[v] = 0.0;
          if (b * BW + r < a.ncols) GA::at2(gv, b * BW + r, item0, row);
#pragma unroll
          for (int v = 0; v < V; v++) sm.gbb[r][lane * V + v] = (GT)row[v];
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
              nx[v] = num > a.l1r ? (num - a.l1r) * den[v] : 0.0;
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
            yd = fma(in, GA::at(gv, j, i), yd);
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
