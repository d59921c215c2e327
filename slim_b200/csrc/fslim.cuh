// fslim.cuh -- fSLIM neighbour search on the GPU; included by engine.cu after gram.cuh.
//
// With nnbrs > 0 (and ordered == 0) the reference restricts the active set of target column j to its nnbrs most
// similar columns and drops the aTy > l1r filter (src/libslim/estimate.c:424-431); the search is FindColumnNeighbors
// (src/libslim/neighbors.c:16-125): candidates = every item that shares a user with j, similarity accumulated in
// FLOAT, cosine divided by the candidate's norm only, "jaccard" built from norms (not squared norms), then
// gk_dfkvkselect + gk_fkvsortd.  Similarity ties at the nnbrs boundary are common on sparse data (209 of the 1683
// ml100k columns for cosine, nnbrs = 10), and which tied candidate survives the reference's quickselect depends on
// the ORDER of the candidate array -- first encounter while walking the users of j in ascending order and each
// user's row in row order.  To return the reference's model and not merely an equally good one, this kernel
// rebuilds that order and runs the same selection:
//   pass 1  every (user of j, item of the user's row) pair gets its flat position in the walk; atomicMin keeps an
//           item's FIRST position, a float atomicAdd accumulates the dot product (exact for integer ratings);
//   pass 2  the pairs that are first encounters are compacted in walk order -> the candidate array of neighbors.c;
//   select  one thread runs the selection of lib/GKlib/fkvkselect.c on it (Lomuto partition around a three-way
//           pivot); the winners are handed to cd_gram_kernel sorted by internal item id.
// One CTA per target, persistent over a queue.  The rows of the training matrix keep the caller's order inside the
// engine (items are relabelled in place, rows are not re-sorted), which is what the walk order needs.
#pragma once

struct NbrArgs {
  int32_t nnbrs, simtype;
  int32_t *queue;
  // per-CTA scratch, stride SolveArgs::col_stride
  uint32_t *first;    // flat position of the first encounter, 0xffffffff = not a candidate
  float *dot;         // similarity numerator
  int32_t *cand_val;  // candidate array in encounter order: item ...
  float *cand_key;    // ... and similarity
  // output, indexed by position in SolveArgs::targets
  int32_t *out_nbr;   // [ntargets][nnbrs] internal item ids, ascending
  int32_t *out_cnt;   // [ntargets]
};

constexpr int kNbrNT = 256;

template <int NT>
__device__ __forceinline__ int team_excl_scan_int(int v, int *sc, int &total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  constexpr int NW = NT / 32;
  if (lane == 31) sc[w] = incl;
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
  return base + incl - v;
}

template <bool HASVAL>
__global__ void __launch_bounds__(kNbrNT) fslim_neighbors_kernel(const SolveArgs a, const NbrArgs na) {
  constexpr int NT = kNbrNT, NW = NT / 32;
  __shared__ int s_sc[NW];
  __shared__ int s_off[NT];   // flat offset of a user's row inside the chunk
  __shared__ int s_pos[NT];   // candidate-array offset of a user's first encounters inside the chunk
  __shared__ int s_q, s_n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t slot = (size_t)blockIdx.x * a.col_stride;
  uint32_t *first = na.first + slot;
  float *dot = na.dot + slot;
  int32_t *cval = na.cand_val + slot;
  float *ckey = na.cand_key + slot;

  for (;;) {
    __syncthreads();
    if (tid == 0) s_q = atomicAdd(na.queue, 1);
    __syncthreads();
    const int q = s_q;
    if (q >= a.ntargets) break;
    const int j = a.targets[q];
    const int cntj = a.colcnt[j];
    const int64_t c0 = a.colptr[j];
    int ncand = 0;
    for (int pass = 0; pass < 2; pass++) {
      uint32_t base = 0;  // flat position of the chunk's first pair
      for (int e0 = 0; e0 < cntj; e0 += NT) {
        const int e = e0 + tid;
        int64_t r0 = 0;
        int len = 0;
        if (e < cntj) {
          const int u = a.colind[c0 + e];
          r0 = a.rowptr[u];
          len = (int)(a.rowptr[u + 1] - r0);
        }
        int tot;
        s_off[tid] = team_excl_scan_int<NT>(len, s_sc, tot);
        __syncthreads();
        const int nu = min(NT, cntj - e0);
        if (pass == 0) {
          for (int t = warp; t < nu; t += NW) {  // one warp per user: coalesced walk of the row
            const int u = a.colind[c0 + e0 + t];
            const int64_t rr = a.rowptr[u];
            const int ln = (int)(a.rowptr[u + 1] - rr);
            const float cv = HASVAL ? a.colval[c0 + e0 + t] : 1.0f;
            const uint32_t f0 = base + (uint32_t)s_off[t];
            for (int p = lane; p < ln; p += 32) {
              const int k = a.rowind[rr + p];
              if (k == j) continue;
              atomicMin(first + k, f0 + (uint32_t)p);
              atomicAdd(dot + k, HASVAL ? __fmul_rn(a.rowval[rr + p], cv) : cv);  // neighbors.c:52-55
            }
          }
        } else {
          // first encounters per user, then their offsets in the candidate array, then the ordered write
          int mine = 0;
          for (int t = warp; t < nu; t += NW) {
            const int u = a.colind[c0 + e0 + t];
            const int64_t rr = a.rowptr[u];
            const int ln = (int)(a.rowptr[u + 1] - rr);
            const uint32_t f0 = base + (uint32_t)s_off[t];
            int c = 0;
            for (int p0 = 0; p0 < ln; p0 += 32) {
              const int p = p0 + lane;
              bool fl = false;
              if (p < ln) {
                const int k = a.rowind[rr + p];
                fl = k != j && first[k] == f0 + (uint32_t)p;
              }
              c += __popc(__ballot_sync(0xffffffffu, fl));
            }
            if (lane == 0) s_pos[t] = c;
            (void)mine;
          }
          __syncthreads();
          int tot2;
          const int my = tid < nu ? s_pos[tid] : 0;
          const int ex = team_excl_scan_int<NT>(my, s_sc, tot2);
          __syncthreads();
          s_pos[tid] = ex;
          __syncthreads();
          for (int t = warp; t < nu; t += NW) {
            const int u = a.colind[c0 + e0 + t];
            const int64_t rr = a.rowptr[u];
            const int ln = (int)(a.rowptr[u + 1] - rr);
            const uint32_t f0 = base + (uint32_t)s_off[t];
            int w0 = ncand + s_pos[t];
            for (int p0 = 0; p0 < ln; p0 += 32) {
              const int p = p0 + lane;
              bool fl = false;
              int k = 0;
              if (p < ln) {
                k = a.rowind[rr + p];
                fl = k != j && first[k] == f0 + (uint32_t)p;
              }
              const unsigned m = __ballot_sync(0xffffffffu, fl);
              if (fl) {
                const int o = w0 + __popc(m & ((1u << lane) - 1u));
                cval[o] = k;
                ckey[o] = dot[k];
              }
              w0 += __popc(m);
            }
          }
          ncand += tot2;
        }
        base += (uint32_t)tot;
        __syncthreads();
      }
      __syncthreads();
    }
    // similarity (neighbors.c:82-83, 108-110), float arithmetic; reset the per-item scratch on the way
    const float cnj = a.cnorms[j];
    for (int i = tid; i < ncand; i += NT) {
      const int k = cval[i];
      float key = ckey[i];
      if (na.simtype == 0) key = __fdiv_rn(key, a.cnorms[k]);
      else if (na.simtype == 1) key = __fdiv_rn(key, __fsub_rn(__fadd_rn(a.cnorms[k], cnj), key));
      ckey[i] = key;
      first[k] = 0xffffffffu;
      dot[k] = 0.0f;
    }
    __syncthreads();
    // selection + hand-over (one thread: the outcome on ties depends on the exact sequence of swaps)
    if (tid == 0) {
      const int want = min(na.nnbrs, ncand);
      if (ncand > want) {
        int lo = 0, hi = ncand - 1;
        while (lo < hi) {
          int mid = lo + ((hi - lo) >> 1);
          if (ckey[lo] < ckey[mid]) mid = lo;
          if (ckey[hi] > ckey[mid]) {
            mid = hi;
            if (ckey[lo] < ckey[mid]) mid = lo;
          }
          float tk = ckey[mid];
          int tv = cval[mid];
          ckey[mid] = ckey[hi];
          cval[mid] = cval[hi];
          ckey[hi] = tk;
          cval[hi] = tv;
          const float pivot = tk;
          int store = lo - 1;
          for (int scan = lo; scan < hi; scan++) {
            const float ks = ckey[scan];
            if (ks >= pivot) {
              store++;
              const float k2 = ckey[store];
              const int v2 = cval[store];
              ckey[store] = ks;
              cval[store] = cval[scan];
              ckey[scan] = k2;
              cval[scan] = v2;
            }
          }
          store++;
          tk = ckey[store];
          tv = cval[store];
          ckey[store] = ckey[hi];
          cval[store] = cval[hi];
          ckey[hi] = tk;
          cval[hi] = tv;
          if (store > want) hi = store - 1;
          else if (store < want) lo = store + 1;
          else break;
        }
      }
      // ascending internal id (the visiting order of the solver)
      int32_t *out = na.out_nbr + (size_t)q * na.nnbrs;
      for (int i = 0; i < want; i++) {
        const int v = cval[i];
        int b = i - 1;
        while (b >= 0 && out[b] > v) {
          out[b + 1] = out[b];
          b--;
        }
        out[b + 1] = v;
      }
      na.out_cnt[q] = want;
    }
  }
}
