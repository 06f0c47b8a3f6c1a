// gf3_fit.cuh -- phase-slope fit of the channel estimate, shared by the fused estimate kernel
// (gf3_rx.cu) and the stage-level equalise() entry point (gf3_stage.cu).
#pragma once
#include <cuda_runtime.h>

namespace gf3 {

// OFDM.py:454-462: phase_diff = unwrap(angle(He)) - unwrap(angle(Hs)) along the bins, then the
// least-squares slope of phase_diff[fit_lo:fit_hi] against 0, 1, 2, ...  (np.polyfit degree 1).
//   phi        : phases in shared memory, row 0 = angle(Hs) at phi[i], row 1 = angle(He) at phi[K + i]
//                (K = row stride; callers that store only the window pass a shifted base); only the
//                entries inside [flo, fhi) are read.  np.unwrap's jumps before the window shift
//                unwrap(He) - unwrap(Hs) by a constant there, which does not change the slope.
//   warp_tot   : int    [NT / 32] shared scratch
//   red        : double [NT / 32] shared scratch
// Called by all NT threads of the CTA (contains __syncthreads); the result is valid in thread 0.
template <int NT>
__device__ __forceinline__ double fit_slope(const double* phi, int K, int flo, int fhi, int* warp_tot, double* red) {
    const int tid = threadIdx.x;
    const int nfit = fhi - flo;
    const int SEG = (nfit + NT - 1) / NT;
    const int i0 = flo + tid * SEG, i1 = min(fhi, i0 + SEG);
    const double PI = 3.14159265358979323846;
    // np.unwrap: a jump dd > pi subtracts 2 pi, dd < -pi adds 2 pi, |dd| == pi is left alone
    int local = 0;
    for (int i = max(i0, flo + 1); i < i1; ++i) {
        const double de = phi[K + i] - phi[K + i - 1], ds = phi[i] - phi[i - 1];
        local += (de > PI ? -1 : de < -PI ? 1 : 0) - (ds > PI ? -1 : ds < -PI ? 1 : 0);
    }
    int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += v;
    }
    if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
    __syncthreads();
    int prefix = incl - local;
    for (int w = 0; w < (tid >> 5); ++w) prefix += warp_tot[w];
    // walk the segment
    const double xbar = 0.5 * (double)(nfit - 1);
    double sxy = 0.0;
    int run = prefix;
    for (int i = i0; i < i1; ++i) {
        if (i >= flo + 1) {
            const double de = phi[K + i] - phi[K + i - 1], ds = phi[i] - phi[i - 1];
            run += (de > PI ? -1 : de < -PI ? 1 : 0) - (ds > PI ? -1 : ds < -PI ? 1 : 0);
        }
        if (i >= flo && i < fhi) {
            const double y = (phi[K + i] - phi[i]) + 2.0 * PI * (double)run;
            sxy += ((double)(i - flo) - xbar) * y;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sxy += __shfl_xor_sync(0xffffffffu, sxy, o);
    if ((tid & 31) == 0) red[tid >> 5] = sxy;
    __syncthreads();
    double tot = 0.0;
    if (tid == 0) {
        for (int w = 0; w < NT / 32; ++w) tot += red[w];
        const double n = (double)nfit;
        const double sxx = n * (n * n - 1.0) / 12.0;
        tot = nfit >= 2 ? tot / sxx : __longlong_as_double(0x7ff8000000000000LL);   // polyfit needs >= 2 points
    }
    return tot;
}

}  // namespace gf3
