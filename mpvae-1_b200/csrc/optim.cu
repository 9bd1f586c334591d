// The tail of the training step (train.py:126-128 / fairsoft_train.py:141-146): clip_grad_norm_ + Adam, as three
// launches over flat buffers instead of the ~40 foreach / elementwise kernels torch issues for it (the fp64
// r_sqrt_sigma alone costs eight fp64 passes there).  Arithmetic follows torch.optim.Adam (L2 weight decay added to the
// gradient, lerp / addcmul moment updates, bias corrections, sqrt(v)/sqrt(bc2) + eps) in the parameter's own dtype.
//
//   grad_norm_kernel + grad_norm_finalize : total L2 norm of the flat gradient bucket (fixed-order, fp64 partials)
//                                           -> {norm, clip coefficient, bias corrections, step size}
//   adam_kernel<P>                        : one pass: p, m, v updated in place; the gradient is read as fp32 (the fp64
//                                           r_sqrt_sigma takes the fp32 g_R of the bucket, no cast pass) and an fp32
//                                           shadow of the new parameter can be written in the same pass
#include "common.cuh"
#include "rows.h"

namespace mpv {
namespace {

constexpr int kNormBlocks = 592;   // 4 per SM

__global__ void __launch_bounds__(256)
grad_norm_kernel(const float* __restrict__ g, size_t n, double* __restrict__ partials) {
    double acc = 0.0;
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    size_t done = 0;
    if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
        const size_t n4 = n / 4;
        const float4* __restrict__ g4 = reinterpret_cast<const float4*>(g);
        for (size_t i = tid; i < n4; i += nthr) {
            const float4 v = g4[i];
            acc += (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
        }
        done = n4 * 4;
    }
    for (size_t i = done + tid; i < n; i += nthr) acc += (double)(g[i] * g[i]);
    acc = warp_sum(acc);
    __shared__ double s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s[w];
        partials[blockIdx.x] = t;
    }
}

// state (fp64): [0] step count t (incremented here)  [1] total norm  [2] gradient multiplier  [3] 1 - beta1^t
// [4] sqrt(1 - beta2^t)  [5] learning rate used.  One thread; the scalars are computed in fp64 like torch's Python side.
__global__ void grad_norm_finalize(const double* __restrict__ partials, int nparts, double max_norm, double grad_scale,
                                   const float* __restrict__ lr_dev, double lr_host, double beta1, double beta2,
                                   double* __restrict__ state) {
    // one warp: lane l sums partials l, l+32, ... ; the 32 lane sums are then added in lane order (fixed order)
    double t = 0.0;
    for (int i = threadIdx.x; i < nparts; i += 32) t += partials[i];
    __shared__ double lanes[32];
    lanes[threadIdx.x] = t;
    __syncwarp();
    if (threadIdx.x != 0) return;
    t = 0.0;
    for (int i = 0; i < 32; ++i) t += lanes[i];
    // the bucket holds the SUM over ranks; grad_scale = 1 / world turns it into the mean the reference clips
    const float norm = (float)(sqrt(t) * grad_scale);   // torch reports the norm in the gradients' dtype
    float coef = 1.0f;
    if (max_norm > 0.0) {
        coef = (float)max_norm / (norm + 1e-6f);     // torch.nn.utils.clip_grad_norm_
        coef = coef > 1.0f ? 1.0f : coef;            // NaN compares false: a NaN norm leaves coef = NaN, as torch does
    }
    const double step = state[0] + 1.0;
    state[0] = step;
    state[1] = (double)norm;
    state[2] = (double)coef * grad_scale;
    state[3] = 1.0 - pow(beta1, step);
    state[4] = sqrt(1.0 - pow(beta2, step));
    state[5] = lr_dev ? (double)*lr_dev : lr_host;
}

template <typename P>
__global__ void __launch_bounds__(256)
adam_kernel(P* __restrict__ p, const float* __restrict__ g, P* __restrict__ m, P* __restrict__ v, float* __restrict__ shadow,
            size_t n, const double* __restrict__ state, double beta1, double beta2, double eps, double weight_decay) {
    const P gscale = (P)state[2], bc1 = (P)state[3], bc2s = (P)state[4], lr = (P)state[5];
    const P step_size = lr / bc1;
    const P b2 = (P)beta2, wd = (P)weight_decay, e = (P)eps;
    const P omb1 = (P)(1.0 - beta1), omb2 = (P)(1.0 - beta2);   // formed in fp64 like torch's Python scalars, then rounded
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const P pi = p[i];
        P gi = (P)g[i] * gscale;
        gi = gi + wd * pi;                                   // grad.add(param, alpha=weight_decay)
        const P mi = m[i] + (gi - m[i]) * omb1;              // exp_avg.lerp_(grad, 1 - beta1)
        const P vi = v[i] * b2 + omb2 * gi * gi;             // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        const P denom = sqrt(vi) / bc2s + e;
        const P pn = pi - step_size * (mi / denom);          // param.addcdiv_(exp_avg, denom, value=-step_size)
        p[i] = pn; m[i] = mi; v[i] = vi;
        if (shadow) shadow[i] = (float)pn;
    }
}

}  // namespace

size_t grad_norm_workspace() { return kNormBlocks * sizeof(double); }

int launch_grad_norm(const float* g, size_t n, void* ws, double max_norm, double grad_scale, const float* lr_dev, double lr_host,
                     double beta1, double beta2, double* state, cudaStream_t stream) {
    double* partials = static_cast<double*>(ws);
    grad_norm_kernel<<<kNormBlocks, 256, 0, stream>>>(g, n, partials);
    if (int rc = check_launch("grad_norm_kernel")) return rc;
    grad_norm_finalize<<<1, 32, 0, stream>>>(partials, kNormBlocks, max_norm, grad_scale, lr_dev, lr_host, beta1, beta2, state);
    return check_launch("grad_norm_finalize");
}

int launch_adam(void* p, int p_is_f64, const float* g, void* m, void* v, float* shadow, size_t n, const double* state,
                double beta1, double beta2, double eps, double weight_decay, cudaStream_t stream) {
    if (n == 0) return 0;
    size_t blocks = (n + 255) / 256;
    if (blocks > (size_t)kNumSMs * 16) blocks = (size_t)kNumSMs * 16;
    if (p_is_f64)
        adam_kernel<double><<<(unsigned)blocks, 256, 0, stream>>>(static_cast<double*>(p), g, static_cast<double*>(m),
                                                                  static_cast<double*>(v), shadow, n, state, beta1, beta2, eps, weight_decay);
    else
        adam_kernel<float><<<(unsigned)blocks, 256, 0, stream>>>(static_cast<float*>(p), g, static_cast<float*>(m),
                                                                 static_cast<float*>(v), shadow, n, state, beta1, beta2, eps, weight_decay);
    return check_launch("adam_kernel");
}

}  // namespace mpv
