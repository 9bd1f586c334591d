// Row math of the probit ELBO forward executed INSIDE the tcgen05 product kernel (contract_tc.cu, FUSE = true).
//
// The product noise.R^T keeps the tensor pipe busy but uses ~13 % of an SM's issue slots; the row forward
// (Phi, log, exp per (sample, row, label) cell, mpvae.py:177-190 / :110-114 / :203-204) is bound by exactly those issue
// slots and leaves the tensor pipe idle.  So the product kernel carries eight extra "math" warps per CTA that follow a
// tile or two behind the tensor pipeline: as soon as a 256 x 256 tile of nr = noise.R^T has been stored (and is still
// in L2), they read it back, add the decoder logits, and reduce
//     over the labels of the tile  -> per-(sample-row, tile) partial log-likelihoods and ranking factors (FusePart)
//     over the samples of a row    -> the predictions mean_s E (written once, final)
// A small finalize kernel (probit_rows.cu) adds the per-tile partials in a fixed order and does the per-row tail.
//
// Rows of nr are ordered b-major (m = b * S + s): the S sample-rows of batch row b are neighbours, so the mean over
// samples is a loop in registers / shared memory of one warp.  A work unit is a "group" = the S rows of one b over the
// 256 columns of one tile; a group belongs to the tile that holds its LAST row and may reach back into the tile above
// (S <= 256), whose completion it then also waits for.  Completion is a device-scope counter per tile that the
// promotion warps of both CTAs of a pair bump after their stores.  Math warps never block the tensor pipeline.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cuda_fp16.h>

#include "common.cuh"
#include "philox.cuh"
#include "probit_math.cuh"

namespace mpv {

constexpr int kFuseMathWarps = 8;          // per CTA: two warpgroups
constexpr int kFuseTileArrivals = 32;      // 16 promotion warps x 2 CTAs bump a tile's counter
constexpr int kFuseMaxS = 256;             // a group reaches back at most one tile

// partial sums of one sample-row over the labels of one 256-column tile, both branches
struct __align__(16) FusePart {
    double lp_l, lp_x;                      // sum_l ll            (label branch, feature branch)
    float pos_l, neg_l, pos_x, neg_x;       // sum_l exp(-5E) [y = 1], sum_l exp(5E) [y = 0]
};

struct FuseFwd {
    int S, B, L, ldn;                       // nr is (B*S, ldn) row-major, rows b-major
    int stable;                             // MPVAE_FLAG_STABLE_CDF
    const float *y, *fe_out, *fx_out;       // (B, L)
    const float* nr;
    float *indiv_prob, *indiv_prob_label;   // (B, L)
    float *E_l, *E_x;                       // (B*S, ldn) clamped probabilities kept for the backward, or nullptr
    FusePart* part;                         // [B*S][tiles_n]
    unsigned int* done;                     // [tiles_m * tiles_n], zeroed before the launch
};

// Flag polling.  An acquire LOAD compiles to LDG.STRONG + CCTL.IVALL (an L1 invalidation) on every poll, and a few
// hundred spinning warps doing that starve the whole SM (measured: the g_R product went from 0.44 to 0.65 ms).  So the
// polls are RELAXED loads and one acquire FENCE follows the successful one.
__device__ __forceinline__ unsigned int ld_relaxed_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_relaxed_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acquire_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void fence_acquire_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }

// whole warp: block until the tile's counter says every promotion warp has stored its part
__device__ __forceinline__ void fuse_wait_tile(const unsigned int* cnt, int lane) {
    if (lane == 0) {
        const long long t0 = clock64();
        while (ld_relaxed_gpu(cnt) < (unsigned)kFuseTileArrivals) {
            __nanosleep(200);
            if (clock64() - t0 > 8000000000LL) __trap();   // ~4 s: a protocol bug must not hang the GPU
        }
        fence_acquire_gpu();
    }
    __syncwarp();
}

// promotion warp, after the stores of its part of tile `idx`
__device__ __forceinline__ void fuse_signal_tile(unsigned int* done, int idx, int lane) {
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicAdd(done + idx, 1u);
}

struct FuseChunk { float y, fe, fx, n0, n1; };

__device__ __forceinline__ void fuse_load(FuseChunk& c, const FuseFwd& f, size_t yoff, size_t r0, bool two, int l) {
    if (l < f.L) {
        c.y = __ldg(f.y + yoff + l);
        c.fe = __ldg(f.fe_out + yoff + l);
        c.fx = __ldg(f.fx_out + yoff + l);
        c.n0 = __ldcg(f.nr + r0 + l);
        c.n1 = two ? __ldcg(f.nr + r0 + f.ldn + l) : 0.0f;
    }
}

// One group: batch row b over the columns [c0, c0 + 256) of tile column tn.  pacc: this warp's [2][256] floats.
template <bool STABLE>
__device__ __forceinline__ void fuse_group(const FuseFwd& f, int b, int tn, int tiles_n, float* __restrict__ pacc, int lane) {
    const int c0 = tn * 256;
    const int ncols = min(256, f.L - c0);
    const int nch = (ncols + 31) >> 5;
    const size_t yoff = (size_t)b * f.L;
#pragma unroll
    for (int c = 0; c < 8; ++c) { pacc[c * 32 + lane] = 0.0f; pacc[256 + c * 32 + lane] = 0.0f; }
    for (int s0 = 0; s0 < f.S; s0 += 2) {
        const bool two = s0 + 1 < f.S;
        const size_t r0 = ((size_t)b * f.S + s0) * f.ldn;
        double lp00 = 0.0, lp01 = 0.0, lp10 = 0.0, lp11 = 0.0;
        float pn0[4] = {0.f, 0.f, 0.f, 0.f}, pn1[4] = {0.f, 0.f, 0.f, 0.f};
        FuseChunk cur, nxt;
        fuse_load(cur, f, yoff, r0, two, c0 + lane);
        for (int c = 0; c < nch; ++c) {
            if (c + 1 < nch) fuse_load(nxt, f, yoff, r0, two, c0 + ((c + 1) << 5) + lane);
            if (c0 + (c << 5) + lane < f.L) {
                // same accumulation order as probit_row_fwd_kernel (pairs of samples), so predictions are bit-equal
                float pl = 0.0f, px = 0.0f;
                {
                    const CellFwd cl = cell_forward<STABLE>(cur.n0 + cur.fe, cur.y);   // mpvae.py:168,177
                    const CellFwd cx = cell_forward<STABLE>(cur.n0 + cur.fx, cur.y);   // mpvae.py:170,180
                    lp00 += (double)cl.ll; lp01 += (double)cx.ll;
                    pn0[0] += cl.epos; pn0[1] += cl.eneg; pn0[2] += cx.epos; pn0[3] += cx.eneg;
                    pl += cl.E; px += cx.E;
                    if (f.E_l) { const size_t o = r0 + c0 + (c << 5) + lane; f.E_l[o] = cl.E; f.E_x[o] = cx.E; }
                }
                if (two) {
                    const CellFwd cl = cell_forward<STABLE>(cur.n1 + cur.fe, cur.y);
                    const CellFwd cx = cell_forward<STABLE>(cur.n1 + cur.fx, cur.y);
                    lp10 += (double)cl.ll; lp11 += (double)cx.ll;
                    pn1[0] += cl.epos; pn1[1] += cl.eneg; pn1[2] += cx.epos; pn1[3] += cx.eneg;
                    pl += cl.E; px += cx.E;
                    if (f.E_l) { const size_t o = r0 + f.ldn + c0 + (c << 5) + lane; f.E_l[o] = cl.E; f.E_x[o] = cx.E; }
                }
                pacc[(c << 5) + lane] += pl;              // a lane only ever touches its own entries
                pacc[256 + (c << 5) + lane] += px;
            }
            cur = nxt;
        }
        lp00 = warp_sum(lp00); lp01 = warp_sum(lp01);
#pragma unroll
        for (int q = 0; q < 4; ++q) pn0[q] = warp_sum(pn0[q]);
        if (two) {
            lp10 = warp_sum(lp10); lp11 = warp_sum(lp11);
#pragma unroll
            for (int q = 0; q < 4; ++q) pn1[q] = warp_sum(pn1[q]);
        }
        if (lane == 0) {
            FusePart p;
            p.lp_l = lp00; p.lp_x = lp01; p.pos_l = pn0[0]; p.neg_l = pn0[1]; p.pos_x = pn0[2]; p.neg_x = pn0[3];
            f.part[((size_t)b * f.S + s0) * tiles_n + tn] = p;
            if (two) {
                p.lp_l = lp10; p.lp_x = lp11; p.pos_l = pn1[0]; p.neg_l = pn1[1]; p.pos_x = pn1[2]; p.neg_x = pn1[3];
                f.part[((size_t)b * f.S + s0 + 1) * tiles_n + tn] = p;
            }
        }
    }
    // predictions: mean over samples (mpvae.py:203-204)
    const float fS = (float)f.S;
    for (int c = 0; c < nch; ++c) {
        const int l = c0 + (c << 5) + lane;
        if (l < f.L) {
            float sl = 0.0f, sx = 0.0f;
            sl += pacc[(c << 5) + lane];
            sx += pacc[256 + (c << 5) + lane];
            f.indiv_prob_label[yoff + l] = sl / fS;
            f.indiv_prob[yoff + l] = sx / fS;
        }
    }
}

// The loop of one math warp of a CTA pair over the pair's tiles (`first`, `first + stride`, ... < num_tiles; tiles are
// numbered tile_m * tiles_n + tile_n).  widx = 0..15 over both CTAs of the pair; groups are dealt round-robin with a
// running counter so that the 25.6 groups of a tile average out over its 16 warps.
template <bool STABLE>
__device__ __forceinline__ void fuse_math_loop(const FuseFwd& f, int first, int stride, int num_tiles, int tiles_n, int widx,
                                               float* __restrict__ pacc, int lane) {
    long long J = 0;
    for (int w = first; w < num_tiles; w += stride) {
        const int tm = w / tiles_n, tn = w % tiles_n;
        const int g_lo = (256 * tm) / f.S;                                   // groups whose last row lies in this tile
        const int g_hi = min(f.B, (256 * (tm + 1)) / f.S);
        const int ng = g_hi - g_lo;
        if (ng > 0) {
            int j = (int)((widx - (J & 15) + 16) & 15);
            if (j < ng) {
                fuse_wait_tile(f.done + w, lane);
                if (g_lo * f.S < 256 * tm) fuse_wait_tile(f.done + w - tiles_n, lane);   // first group starts in the tile above
                for (; j < ng; j += 16) fuse_group<STABLE>(f, g_lo + j, tn, tiles_n, pacc, lane);
            }
            J += ng;
        }
    }
}

// ------------------------------------------------------------------------------------------------ noise just in time
// The other job the product kernel's math warps can take (and the one that pays: pure integer / SFU work with plenty of
// independent instructions per warp): DRAWING the Philox normals of mpvae.py:162 straight into the product's own A
// operand plane while the tensor pipe works on earlier rows.  The plane is cut into blocks of 128 rows (one CTA's share
// of a tile); all math warps of the grid fill block 0, then block 1, ... in the order the tiles consume them, bump the
// block's counter when their share is written, and the TMA producer of a CTA waits for its block's counter before the
// first load of a tile.  Same counter -> element mapping as philox_planes_kernel, so the numbers are identical.
struct FuseNoise {
    __half* plane;                 // [S*B][pitch] halves, rows b-major (b * S + s)
    int S, B, Z, pitch, Bg, row0;  // Bg / row0: the global batch this shard's rows [row0, row0 + B) belong to
    uint2 key, off;
    const unsigned long long* off_dev;
    unsigned int* ready;           // [ceil(S*B / 128)] counters, zeroed before the launch
};

__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// TMA producer thread: block `u` of the plane is complete once every math warp of the grid has signalled it
__device__ __forceinline__ void noise_wait_block(const unsigned int* cnt, unsigned int warps_total) {
    const long long t0 = clock64();
    while (ld_relaxed_gpu(cnt) < warps_total) {
        __nanosleep(100);
        if (clock64() - t0 > 8000000000LL) __trap();   // ~4 s: a protocol bug must not hang the GPU
    }
    fence_acquire_gpu();
    fence_proxy_async_global();                         // generic-proxy stores -> async-proxy (TMA) loads
}

// One math warp's share of the whole plane.  gw = this warp's index among the `warps_total` math warps of the grid.
__device__ __forceinline__ void noise_math_loop(const FuseNoise& f, int gw, int warps_total, int lane) {
    const int M = f.S * f.B;
    const int nblocks = (M + 127) >> 7;
    const unsigned cpr = (unsigned)((f.Z + 3) / 4 + 1);           // counters that can touch one row
    const unsigned units = 128u * cpr;
    const unsigned stride = (unsigned)warps_total * 32u;
    const uint2 off = philox_offset(f.off, f.off_dev);
    for (int u = 0; u < nblocks; ++u) {
        for (unsigned idx = (unsigned)gw * 32u + (unsigned)lane; idx < units; idx += stride) {
            const int m = (u << 7) + (int)(idx / cpr);
            if (m >= M) break;
            const int b = m / f.S, s = m - b * f.S;
            const unsigned long long flat0 = ((unsigned long long)s * f.Bg + f.row0 + b) * f.Z;   // (s, b_global, 0)
            const unsigned long long c = (flat0 >> 2) + (idx % cpr);
            const long long z0 = (long long)(c << 2) - (long long)flat0;                        // -3 .. Z + 3
            if (z0 >= f.Z) continue;
            float n[4];
            philox_normal4(c, f.key, off, n);
            __half* __restrict__ row = f.plane + (size_t)m * f.pitch;
            if (z0 >= 0 && z0 + 4 <= f.Z) {
                __half* dst = row + z0;
                const __half2 p01 = __floats2half2_rn(n[0], n[1]), p23 = __floats2half2_rn(n[2], n[3]);
                const uintptr_t a = reinterpret_cast<uintptr_t>(dst);
                if ((a & 7u) == 0) {
                    uint2 v;
                    v.x = *reinterpret_cast<const uint32_t*>(&p01);
                    v.y = *reinterpret_cast<const uint32_t*>(&p23);
                    *reinterpret_cast<uint2*>(dst) = v;
                } else if ((a & 3u) == 0) {
                    *reinterpret_cast<__half2*>(dst) = p01;
                    *reinterpret_cast<__half2*>(dst + 2) = p23;
                } else {
                    dst[0] = __low2half(p01);
                    *reinterpret_cast<__half2*>(dst + 1) = __halves2half2(__high2half(p01), __low2half(p23));
                    dst[3] = __high2half(p23);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const long long z = z0 + j;
                    if (z >= 0 && z < f.Z) row[z] = __float2half_rn(n[j]);
                }
            }
        }
        fence_proxy_async_global();
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(f.ready + u, 1u);
    }
}

// ------------------------------------------------------------------------------------------------ tile exchange
// The third job for the math warps, in the BACKWARD product g_R = gxs^T . noise of a data-parallel run: the sum of g_R over
// the ranks, tile by tile, while the tensor pipe is still producing later tiles.  Every rank's product writes its
// partial g_R into its own peer-mapped `part` buffer and bumps a per-tile counter (in peer memory too) when a tile is
// stored.  Tile t belongs to rank t mod world: that rank's math warps wait until the tile is complete on EVERY rank
// (remote acquire loads of the counters over NVLink), pull it from all ranks, add the partials in rank order (fixed
// order: every rank ends up with bit-identical sums) and store the result into every rank's g_R -- compute and
// collective in ONE kernel over peer memory.  Counters are monotonic (a tile is complete when its counter reaches
// 32 x the exchange's step number), so nothing is reset between steps.  The K-sliced tail wave of the product (finished by
// tail_fixup_kernel after this kernel) is exchanged afterwards by the stand-alone reduce kernel over that row range.
struct FusePeer {
    int world, rank;
    unsigned int step;                 // flag value of this exchange; a tile's counter target is 32 * (step + *step_dev)
    const unsigned int* step_dev;
    long long timeout_cycles;
    unsigned int* err;                 // this rank's error word (a wait that timed out records the step there)
    float* part[8];                    // every rank's partial g_R (this rank's = the product's C), (Mc, Nc) pitch ldc
    float* g_r[8];                     // every rank's final g_R
    unsigned int* done[8];             // every rank's per-tile counters
    int Mc, Nc, ldc, tiles_n;
    int full_tiles;                    // tiles [0, full_tiles) are exchanged in the kernel (filled in by the launcher)
};

// Publishing a finished tile.  A system-scope fence in each of the 512 promotion threads (MEMBAR.SC.SYS + an L1
// invalidation) stalls the warps the tensor pipe is waiting for, so the promotion warps only count themselves in shared
// memory (CTA-scope release) and ONE otherwise idle control thread per CTA turns sixteen arrivals into one system-scope
// release and one bump of the tile's counter by 16 (release / acquire chains are cumulative).
__device__ __forceinline__ void peer_arrive_tile(volatile int* arrivals, int lane) {
    __threadfence_block();
    __syncwarp();
    if (lane == 0) atomicAdd_block(const_cast<int*>(arrivals), 1);
}
// control warp 3, one thread: the n-th exchangeable tile of this CTA is `t`
__device__ __forceinline__ void peer_publish_tile(const FusePeer& f, volatile int* arrivals, int n, int t) {
    while (*arrivals < 16 * (n + 1)) __nanosleep(100);
    fence_acquire_sys();                                  // acq_rel: the promotion warps' stores, then the counter
    atomicAdd(f.done[f.rank] + t, 16u);
}

// The math warps of one CTA.  Warp 0 is the CTA's WATCHER: it follows this rank's tiles in order, waits until tile k is
// complete on every rank (lane r polls rank r's counter over NVLink: one poller per CTA, not one per warp) and publishes
// the number of exchangeable tiles in shared memory.  The other seven pull rows: unit u = (owned tile u / 256, row
// u % 256), dealt round-robin over all worker warps of the grid.  `ready` is an int in shared memory, zeroed by the
// caller before the CTA-wide barrier that precedes the role split.
__device__ __forceinline__ void peer_math_loop(const FusePeer& f, int cta, int ctas, int mw, int lane, volatile int* ready) {
    const unsigned int target = 32u * (f.step + (f.step_dev ? *f.step_dev : 0u));
    const int owned = (f.full_tiles - f.rank + f.world - 1) / f.world;       // tiles rank, rank + world, ...
    if (mw == 0) {
        // a round looks at the next 32 / world tiles at once (lane = (tile offset, rank)): the remote loads of a round
        // are in flight together, and the ready prefix advances by as many tiles as are complete everywhere
        const int per = 32 / f.world, r = lane % f.world, dk = lane / f.world;
        int k0 = 0;
        const long long t0 = clock64();
        while (k0 < owned) {
            const int k = k0 + dk;
            bool ok = true;
            if (dk < per && k < owned) ok = (int)(ld_relaxed_sys(f.done[r] + f.rank + k * f.world) - target) >= 0;
            const unsigned pending = __ballot_sync(0xffffffffu, !ok);
            // tiles k0 .. k0 + adv - 1 are complete on every rank: the first pending lane bounds the prefix
            int adv = pending ? (__ffs(pending) - 1) / f.world : per;
            if (adv > owned - k0) adv = owned - k0;
            if (adv > 0) {
                k0 += adv;
                if (lane == 0) { fence_acquire_sys(); *ready = k0; }
            } else {
                __nanosleep(500);
                if (clock64() - t0 > f.timeout_cycles) { if (lane == 0) { atomicMax(f.err, f.step); *ready = owned; } return; }
            }
        }
        return;
    }
    const int workers = kFuseMathWarps - 1;
    const int gw = cta * workers + (mw - 1), warps_total = ctas * workers;
    int have = 0;
    for (long long u = gw; u < (long long)owned * 256; u += warps_total) {
        const int k = (int)(u >> 8), row_in = (int)(u & 255);
        const int t = f.rank + k * f.world;
        if (k >= have) {
            while ((have = *ready) <= k) __nanosleep(500);
            fence_acquire_sys();       // the watcher's acquire, then this warp's own: data loads below see the tile
        }
        const int row = (t / f.tiles_n) * 256 + row_in;
        if (row >= f.Mc) continue;
        const int c0 = (t % f.tiles_n) * 256;
        const size_t o = (size_t)row * f.ldc + c0;
#pragma unroll
        for (int jj = 0; jj < 8; jj += 4) {
            float v[4][8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = lane + 32 * (jj + j);
                const bool in = c0 + c < f.Nc;
#pragma unroll
                for (int r = 0; r < 8; ++r) v[j][r] = (in && r < f.world) ? __ldcg(f.part[r] + o + c) : 0.0f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = lane + 32 * (jj + j);
                if (c0 + c >= f.Nc) continue;
                float a = v[j][0];
#pragma unroll
                for (int r = 1; r < 8; ++r)
                    if (r < f.world) a += v[j][r];
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    if (r < f.world) f.g_r[r][o + c] = a;
            }
        }
    }
    __threadfence_system();      // this warp's deliveries are visible before the kernel ends (and the rank signals)
}

}  // namespace mpv
