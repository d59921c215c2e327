// predict.cuh -- batched top-N recommendation on the GPU; included by engine.cu.
//
// Replaces the per-user loop of Py_SLIM_Predict (reference src/libslim/pyapi.c:530-563) around
// GetRecommendations (src/libslim/predict.c:15-71): for every user row of the history matrix, the scores
//     s[c] = sum_{i in history} r_ui * W[i][c]        (float, accumulated in history order, predict.c:41-52)
// of all items c outside the history, then the nrcmds best by (descending score, ascending item id).
// One CTA per user.  The history items are processed one after the other with a CTA barrier in between
// and the entries of one model row (distinct columns) are spread over the threads, so every score sees
// exactly the reference's sequence of separate float multiplies and adds: the lists AND the scores are
// bit-identical to the host restatement (api.cpp: recommend()).
#pragma once

constexpr int kPredNT = 256;

struct PredictArgs {
  int32_t nusers, wrows, wcols, nrcmds;
  const int64_t *wrowptr;
  const int32_t *wrowind;
  const float *wrowval;
  const int64_t *urowptr;
  const int32_t *urowind;
  const float *urowval;  // nullptr: all ratings 1.0
  // per-CTA scratch, stride `stride` (>= wcols), zero between users
  float *score;
  int32_t *mark;  // 0 untouched, 1 candidate, 2 history, 3 already emitted
  int32_t *cand;
  size_t stride;
  int32_t *out_ids;    // [nusers][nrcmds]
  float *out_scores;   // [nusers][nrcmds]
  int32_t *out_counts; // [nusers]
};

__global__ void __launch_bounds__(kPredNT) predict_topn_kernel(const PredictArgs a) {
  __shared__ int s_ncand;
  __shared__ float s_bs[kPredNT / 32];
  __shared__ int s_bc[kPredNT / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float *score = a.score + (size_t)blockIdx.x * a.stride;
  int32_t *mark = a.mark + (size_t)blockIdx.x * a.stride;
  int32_t *cand = a.cand + (size_t)blockIdx.x * a.stride;
  if (tid == 0) s_ncand = 0;
  __syncthreads();
  for (int u = blockIdx.x; u < a.nusers; u += gridDim.x) {
    const int64_t h0 = a.urowptr[u], h1 = a.urowptr[u + 1];
    // history items never enter the list (predict.c:34-37)
    for (int64_t r = h0 + tid; r < h1; r += kPredNT) {
      const int it = a.urowind[r];
      if (it >= 0 && it < a.wcols) mark[it] = 2;
    }
    __syncthreads();
    // scores, one history item at a time (float adds in the reference's order)
    for (int64_t r = h0; r < h1; r++) {
      const int i = a.urowind[r];
      if (i < 0 || i >= a.wrows) continue;  // block-uniform
      const float rating = a.urowval ? a.urowval[r] : 1.0f;
      const int64_t k0 = a.wrowptr[i], k1 = a.wrowptr[i + 1];
      for (int64_t k = k0 + tid; k < k1; k += kPredNT) {
        const int c = a.wrowind[k];
        const int mk = mark[c];
        if (mk == 2) continue;
        if (mk == 0) {
          mark[c] = 1;
          cand[atomicAdd(&s_ncand, 1)] = c;
        }
        score[c] = __fadd_rn(score[c], __fmul_rn(rating, a.wrowval[k]));  // no fused multiply-add (c99 build)
      }
      __syncthreads();
    }
    const int ncand = s_ncand;
    const int nout = min(ncand, a.nrcmds);
    // nout rounds of a CTA-wide arg-max over the candidates: (score desc, item id asc)
    for (int t = 0; t < nout; t++) {
      float bs = 0.f;
      int bc = -1;
      for (int j = tid; j < ncand; j += kPredNT) {
        const int c = cand[j];
        if (mark[c] != 1) continue;
        const float sc = score[c];
        if (bc < 0 || sc > bs || (sc == bs && c < bc)) {
          bs = sc;
          bc = c;
        }
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        const float os = __shfl_xor_sync(0xffffffffu, bs, o);
        const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
        if (oc >= 0 && (bc < 0 || os > bs || (os == bs && oc < bc))) {
          bs = os;
          bc = oc;
        }
      }
      if (lane == 0) {
        s_bs[warp] = bs;
        s_bc[warp] = bc;
      }
      __syncthreads();
      if (tid == 0) {
        for (int w = 1; w < kPredNT / 32; w++) {
          const float os = s_bs[w];
          const int oc = s_bc[w];
          if (oc >= 0 && (bc < 0 || os > bs || (os == bs && oc < bc))) {
            bs = os;
            bc = oc;
          }
        }
        a.out_ids[(size_t)u * a.nrcmds + t] = bc;
        a.out_scores[(size_t)u * a.nrcmds + t] = bs;
        mark[bc] = 3;
      }
      __syncthreads();
    }
    if (tid == 0) a.out_counts[u] = nout;
    // restore the scratch for the next user
    for (int j = tid; j < ncand; j += kPredNT) {
      const int c = cand[j];
      score[c] = 0.f;
      mark[c] = 0;
    }
    for (int64_t r = h0 + tid; r < h1; r += kPredNT) {
      const int it = a.urowind[r];
      if (it >= 0 && it < a.wcols) mark[it] = 0;
    }
    __syncthreads();
    if (tid == 0) s_ncand = 0;
    __syncthreads();
  }
}
