// On-device per-step metrics: what the reference computes on the host every training step with
// `evals.compute_metrics(indiv_prob.cpu().data.numpy(), input_label.cpu().data.numpy(), 0.5, all_metrics=False)`
// (train.py:131, fairsoft_train.py:149; evals.py:178-239) -- a device->host sync, a deepcopy and a numpy pass per step.
// Here: integer tp / fp / fn per label and per-row statistics in two small kernels, a third one folds them into
// ACC, HA, ebF1, miF1, maF1 and p@1/3/5 as device scalars (no host sync).  Counts are integers, hence exact.
#include "common.cuh"
#include "rows.h"

namespace mpv {
namespace {

// thread per label: walk the batch (coalesced across threads for a fixed row)
__global__ void __launch_bounds__(256)
label_counts_kernel(const float* __restrict__ prob, const float* __restrict__ y, int B, int L, float thr,
                    int* __restrict__ counts /* [3][L]: tp, fp, fn */) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    int tp = 0, fp = 0, fn = 0;
    const int b0 = blockIdx.y * 64, b1 = min(B, b0 + 64);     // 64-row slabs; integer atomics keep the sum exact
    for (int b = b0; b < b1; ++b) {
        const bool p = prob[(size_t)b * L + l] >= thr;          // evals.py:201-202
        const bool t = y[(size_t)b * L + l] != 0.0f;            // targets are {0,1}: t * p, !t * p, t * !p (evals.py:61-67)
        tp += (t && p); fp += (!t && p); fn += (t && !p);
    }
    if (tp) atomicAdd(&counts[l], tp);
    if (fp) atomicAdd(&counts[L + l], fp);
    if (fn) atomicAdd(&counts[2 * L + l], fn);
}

// CTA (4 warps) per row: exact-match flag, xor count, tp, |pred|, |target|, and the hits among the top-1/3/5 scores.
// ONE pass over the row: every thread keeps the five best of its own elements in registers (ordered by score, ties
// towards the HIGHER index: np.argsort ascending, then reversed, evals.py:37), the warp merges the 32 lists with five
// rounds of an arg-max over their heads, warp 0 merges the four warps' lists the same way.
constexpr int kRowStatThreads = 128;

__device__ __forceinline__ bool ahead(float v, int i, float w, int j) { return v > w || (v == w && i > j); }

__device__ __forceinline__ void warp_top5(float (&tv)[5], int (&ti)[5], float (&ov)[5], int (&oi)[5], int lane) {
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        float bv = tv[0];
        int bi = ti[0], bl = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float v = __shfl_xor_sync(0xffffffffu, bv, o);
            const int i = __shfl_xor_sync(0xffffffffu, bi, o);
            const int l = __shfl_xor_sync(0xffffffffu, bl, o);
            if (ahead(v, i, bv, bi) || (v == bv && i == bi && l < bl)) { bv = v; bi = i; bl = l; }
        }
        ov[k] = bv; oi[k] = bi;
        if (lane == bl) {                      // the winner's list moves up by one
#pragma unroll
            for (int q = 0; q < 4; ++q) { tv[q] = tv[q + 1]; ti[q] = ti[q + 1]; }
            tv[4] = -INFINITY; ti[4] = -1;
        }
    }
}

__global__ void __launch_bounds__(kRowStatThreads)
row_stats_kernel(const float* __restrict__ prob, const float* __restrict__ y, int B, int L, float thr,
                 int* __restrict__ rows /* [B][8]: xor, tp, npred, ntarg, hit1, hit3, hit5, - */) {
    constexpr int kWarps = kRowStatThreads / 32;
    __shared__ float s_v[kWarps][5];
    __shared__ int s_i[kWarps][5];
    __shared__ int s_cnt[kWarps][4];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* __restrict__ pr = prob + (size_t)b * L;
    const float* __restrict__ yr = y + (size_t)b * L;
    int nx = 0, tp = 0, np = 0, nt = 0;
    float tv[5];
    int ti[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) { tv[k] = -INFINITY; ti[k] = -1; }
#pragma unroll 4
    for (int l = tid; l < L; l += kRowStatThreads) {
        const float v = pr[l];
        const bool p = v >= thr, t = yr[l] != 0.0f;
        nx += (p != t); tp += (p && t); np += p; nt += t;
        if (v >= tv[4]) {                       // a thread walks its labels upwards: an equal score goes AHEAD of the kept one
            tv[4] = v; ti[4] = l;
            bool moving = true;                 // only the new element moves; kept equal scores stay in their order
#pragma unroll
            for (int k = 4; k > 0; --k) {
                moving = moving && tv[k] >= tv[k - 1];
                if (moving) {
                    const float fv = tv[k]; tv[k] = tv[k - 1]; tv[k - 1] = fv;
                    const int fi = ti[k]; ti[k] = ti[k - 1]; ti[k - 1] = fi;
                }
            }
        }
    }
    nx = warp_sum(nx); tp = warp_sum(tp); np = warp_sum(np); nt = warp_sum(nt);
    float wv[5];
    int wi[5];
    warp_top5(tv, ti, wv, wi, lane);
    if (lane == 0) {
        s_cnt[warp][0] = nx; s_cnt[warp][1] = tp; s_cnt[warp][2] = np; s_cnt[warp][3] = nt;
#pragma unroll
        for (int k = 0; k < 5; ++k) { s_v[warp][k] = wv[k]; s_i[warp][k] = wi[k]; }
    }
    __syncthreads();
    if (warp != 0) return;
    // lane w < kWarps holds warp w's (already ordered) list
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        tv[k] = lane < kWarps ? s_v[lane][k] : -INFINITY;
        ti[k] = lane < kWarps ? s_i[lane][k] : -1;
    }
    warp_top5(tv, ti, wv, wi, lane);
    if (lane == 0) {
        int hits[3] = {0, 0, 0};
        const int kmax = L < 5 ? L : 5;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            if (k < kmax) {
                const int rel = (wi[k] >= 0 && yr[wi[k]] != 0.0f) ? 1 : 0;
                if (k < 1) hits[0] += rel;
                if (k < 3) hits[1] += rel;
                hits[2] += rel;
            }
        }
        int c[4] = {0, 0, 0, 0};
        for (int w = 0; w < kWarps; ++w)
            for (int q = 0; q < 4; ++q) c[q] += s_cnt[w][q];
        int* r = rows + (size_t)b * 8;
        r[0] = c[0]; r[1] = c[1]; r[2] = c[2]; r[3] = c[3]; r[4] = hits[0]; r[5] = hits[1]; r[6] = hits[2]; r[7] = 0;
    }
}

// one CTA: fold the integer statistics into the eight scalars (fixed order, fp64 accumulation)
__global__ void __launch_bounds__(256)
metrics_finalize_kernel(const int* __restrict__ counts, const int* __restrict__ rows, int B, int L, double* __restrict__ out) {
    __shared__ double s_red[8][9];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double acc = 0, ham = 0, ebf = 0, ebn = 0, p1 = 0, p3 = 0, p5 = 0;
    for (int b = tid; b < B; b += 256) {
        const int* r = rows + (size_t)b * 8;
        acc += (r[0] == 0);
        ham += (double)r[0] / (double)L;
        const int den = r[2] + r[3];
        if (den != 0) { ebf += (double)(2.0f * (float)r[1] / (float)den); ebn += 1.0; }     // fp32 ratio as evals.py:84
        p1 += (double)r[4]; p3 += (double)r[5] / 3.0; p5 += (double)r[6] / 5.0;
    }
    double stp = 0, sfp = 0, sfn = 0, maf = 0, man = 0;
    for (int l = tid; l < L; l += 256) {
        const float tp = (float)counts[l], fp = (float)counts[L + l], fn = (float)counts[2 * L + l];
        stp += tp; sfp += fp; sfn += fn;
        const float c = (2.0f * tp) / (2.0f * tp + fp + fn + 1e-6f);                           // evals.py:108, fp32
        if (isfinite(c)) { maf += (double)c; man += 1.0; }
    }
    double v[12] = {acc, ham, ebf, ebn, p1, p3, p5, stp, sfp, sfn, maf, man};
#pragma unroll
    for (int i = 0; i < 12; ++i) v[i] = warp_sum(v[i]);
    __shared__ double s_all[8][12];
    if (lane == 0)
        for (int i = 0; i < 12; ++i) s_all[warp][i] = v[i];
    __syncthreads();
    if (tid == 0) {
        double t[12];
        for (int i = 0; i < 12; ++i) { t[i] = 0; for (int w = 0; w < 8; ++w) t[i] += s_all[w][i]; }
        out[0] = t[0] / B;                                   // ACC   subset accuracy
        out[1] = 1.0 - t[1] / B;                             // HA    1 - hamming loss
        out[2] = t[3] > 0 ? t[2] / t[3] : NAN;               // ebF1
        out[3] = (double)((float)(2.0 * t[7]) / (float)(2.0 * t[7] + t[8] + t[9]));   // miF1 (evals.py:97-99)
        out[4] = t[11] > 0 ? t[10] / t[11] : NAN;            // maF1
        out[5] = t[4] / B; out[6] = t[5] / B; out[7] = t[6] / B;   // p@1, p@3, p@5
    }
    (void)s_red;
}

// the per-label curves of compute_metrics(..., all_metrics=True) (evals.py:129-175, which calls
// scikit-learn 1.9: roc_auc_score, precision_recall_curve + auc, and the "FDR" recall).  Input: every label's scores
// sorted in DECREASING order with the targets carried along ((N, L) row-major, column l = label l).  One thread per
// label walks its column (coalesced across labels) and emits one curve point per DISTINCT score:
//   ROC  (fps / n_neg, tps / n_pos) from (0, 0), trapezoid area; NaN when a class is absent
//   PR   (recall, precision) = (tps / n_pos, tps / (tps + fps)) from the closing point (0, 1), trapezoid area;
//        recall = 1 throughout for a label without positives (-> 0.5)
//   FDR  recall of the lowest-threshold point with 1 - precision <= cutoff (the closing point always qualifies)
// Counts are integers, the arithmetic fp64: the results equal the numpy restatement to rounding (1e-15).
__global__ void __launch_bounds__(128)
label_curves_kernel(const float* __restrict__ s, const float* __restrict__ y, int N, int L, double cutoff,
                    double* __restrict__ out /* [3][L] */) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    int n_pos = 0;
    for (int i = 0; i < N; ++i) n_pos += (y[(size_t)i * L + l] != 0.0f);
    const int n_neg = N - n_pos;
    const double inv_pos = n_pos > 0 ? 1.0 / (double)n_pos : 0.0, inv_neg = n_neg > 0 ? 1.0 / (double)n_neg : 0.0;
    int tps = 0, fps = 0;
    double fpr_prev = 0.0, tpr_prev = 0.0, rec_prev = 0.0, prec_prev = 1.0;
    double auc = 0.0, aupr = 0.0, fdr_rec = 0.0;
    float cur = N > 0 ? s[l] : 0.0f;
    for (int i = 0; i < N; ++i) {
        const bool pos = y[(size_t)i * L + l] != 0.0f;
        tps += pos; fps += !pos;
        const float nxt = (i + 1 < N) ? s[(size_t)(i + 1) * L + l] : 0.0f;
        if (i + 1 == N || nxt != cur) {       // last element of a group of equal scores: one curve point
            const double fpr = (double)fps / (double)(n_neg > 0 ? n_neg : 1), tpr = (double)tps / (double)(n_pos > 0 ? n_pos : 1);
            auc += (fpr - fpr_prev) * (tpr + tpr_prev) * 0.5;
            fpr_prev = fpr; tpr_prev = tpr;
            const double prec = (double)tps / (double)(tps + fps);
            const double rec = n_pos > 0 ? (double)tps / (double)n_pos : 1.0;
            aupr += (rec - rec_prev) * (prec + prec_prev) * 0.5;
            rec_prev = rec; prec_prev = prec;
            if (1.0 - prec <= cutoff) fdr_rec = rec;
        }
        cur = nxt;
    }
    (void)inv_pos; (void)inv_neg;
    out[l] = (n_pos == 0 || n_neg == 0) ? NAN : auc;
    out[L + l] = aupr;
    out[2 * L + l] = fdr_rec;
}

}  // namespace

int launch_label_curves(const float* sorted_scores, const float* sorted_targets, int N, int L, double cutoff, double* out,
                        cudaStream_t stream) {
    label_curves_kernel<<<ceil_div(L, 128), 128, 0, stream>>>(sorted_scores, sorted_targets, N, L, cutoff, out);
    return check_launch("label_curves_kernel");
}

size_t batch_metrics_workspace(int B, int L) { return align_up((size_t)3 * L * sizeof(int), 256) + (size_t)B * 8 * sizeof(int); }

int launch_batch_metrics(const float* prob, const float* y, int B, int L, float thr, double* out, void* ws, cudaStream_t stream) {
    int* counts = static_cast<int*>(ws);
    int* rows = reinterpret_cast<int*>(static_cast<char*>(ws) + align_up((size_t)3 * L * sizeof(int), 256));
    if (cudaMemsetAsync(counts, 0, (size_t)3 * L * sizeof(int), stream) != cudaSuccess) { set_error("cudaMemsetAsync failed"); return 2; }
    label_counts_kernel<<<dim3(ceil_div(L, 256), ceil_div(B, 64)), 256, 0, stream>>>(prob, y, B, L, thr, counts);
    if (int rc = check_launch("label_counts_kernel")) return rc;
    row_stats_kernel<<<B, kRowStatThreads, 0, stream>>>(prob, y, B, L, thr, rows);
    if (int rc = check_launch("row_stats_kernel")) return rc;
    metrics_finalize_kernel<<<1, 256, 0, stream>>>(counts, rows, B, L, out);
    return check_launch("metrics_finalize_kernel");
}

}  // namespace mpv
