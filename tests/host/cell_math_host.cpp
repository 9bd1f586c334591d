// Host build of the kernels' per-cell arithmetic (mpvae-1_b200/csrc/probit_math.cuh is __host__ __device__): lets the
// CPU test-suite check the forward cell and the closed-form backward without a GPU.
//   stdin : n, then n records  x y cn cp cq gp   (text)
//   stdout: per record  E ll epos eneg dLdx   for the faithful arithmetic, then the same five for the stable CDF
#define __host__
#define __device__
#define __forceinline__ inline
#include <cstdio>
#include <vector>

#include "../../mpvae-1_b200/csrc/probit_math.cuh"

int main() {
    int n = 0;
    if (std::scanf("%d", &n) != 1) return 1;
    for (int i = 0; i < n; ++i) {
        float x, y, cn, cp, cq, gp;
        if (std::scanf("%f %f %f %f %f %f", &x, &y, &cn, &cp, &cq, &gp) != 6) return 2;
        const mpv::CellFwd f = mpv::cell_forward<false>(x, y);
        const float g = mpv::cell_backward<false>(x, y, cn, cp, cq, gp);
        const mpv::CellFwd fs = mpv::cell_forward<true>(x, y);
        const float gs = mpv::cell_backward<true>(x, y, cn, cp, cq, gp);
        std::printf("%.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g\n", f.E, f.ll, f.epos, f.eneg, g, fs.E, fs.ll, fs.epos,
                    fs.eneg, gs);
    }
    return 0;
}
