// gf3_sync.cu -- chirp synchronisation (OFDM.py:356-372, 393-395).
//
// The reference convolves the whole recording with the time-reversed chirp through one big FFT
// (scipy.signal.convolve, OFDM.py:358), normalises by the global max and walks the result in a
// Python loop.  Here the matched filter is a uniformly-partitioned overlap-save convolution on
// 4096-point real FFT blocks (hop B = 2048) built from the same register-resident FFT engine
// as the receiver:
//   xcorr_fwd_kernel : spectrum of every 50 %-overlapped input block           (4 B read, 8 B written / sample)
//   xcorr_acc_kernel : Y_b = sum_p X_{b-p} H_p, inverse real FFT, last B samples, signed row max
//   peak_pick_kernel : candidate mask + ascending hold-off scan, one CTA per stream
#include <math.h>
#include <stdlib.h>
#include <vector>

#include "gf3_common.cuh"
#include "gf3_fft.cuh"

// persistent grids of the two matched-filter kernels, in CTAs per SM
#ifndef GF3_XC_FWD_CTAS
#define GF3_XC_FWD_CTAS 12
#endif
#ifndef GF3_XC_ACC_CTAS
#define GF3_XC_ACC_CTAS 16
#endif
#ifndef GF3_PEAK_PER
#define GF3_PEAK_PER 24        // positions per thread and round of the one-kernel peak picker (multiple of 4, <= 28)
#endif

namespace gf3 {

using SP = FftPlan<12>;                   // 4096 real samples per block
constexpr int kB = SP::M;                 // hop = partition length = 2048
constexpr int kSyncThreads = 256;

int upload_twiddles(int logN, float2** d_out);   // gf3_lib.cu
int make_chirp(gf3_plan* plan);                  // gf3_tx.cu

// Spectrum layout per block: M complex values; element 0 packs (X[0], X[M]) (both real).

struct FwdArgs {
    const void* r;           // [n_streams, r_stride] float32 / int16 / uint8 samples (template parameter S)
    float2* spec;            // [n_streams, nblk, M]
    const float2* tw;
    float* pmax;             // [n_streams] reset to -inf here (may be null)
    int64_t r_stride, T, n_streams;
    int nblk;                // blocks per stream
    int reverse;             // 1: input is read time-reversed (building the filter spectrum)
    int in_off;              // block b covers input samples [b*B - B + in_off, +2B)
    int valid_len;           // only the first valid_len samples of a block are taken (rest zero)
    float* energy;           // optional [n_streams * nblk, kGroups] (zeroed by the caller): weighted energy of the block's spectrum per bin group
};
// Bin groups of the energy output: the 64 bins k = t + 128 q that one half of the thread group owns in slot q (groups
// 2q, 2q + 1), and their mirror images M - k (groups 16 + 2q, 16 + 2q + 1) -- contiguous runs of the spectrum.  DC and
// k = M/2 go with group 0, the Nyquist bin with group 16.  Weights as in the inverse real transform: 1 for DC and
// Nyquist, 2 for the others.  Input spectra and chirp partitions come from the same kernel, hence the same grouping,
// which is all the Cauchy-Schwarz bound of the detection-only matched filter needs (see xcorr_bound_kernel).
constexpr int kGroups = 32;

// block b of a stream = samples [b*B - B, b*B + B), zero outside [0, T)
// Each 128-thread group transforms one block and untangles it itself: thread t owns the bin pairs
// (k, M-k), k = t + 128 q, so bin, twiddle and addresses need no per-item index arithmetic.  The
// next work item's samples are requested right after the FFT, so they fly during the untangle.
template <class S>
__global__ void __launch_bounds__(kSyncThreads, 2) xcorr_fwd_kernel(const FwdArgs a) {
    using P = SP;
    constexpr int NT = kSyncThreads, T = P::T, R = P::R, M = P::M, N = P::N, MP = P::MP, SF = NT / T;
    constexpr int Q = (M / 2) / T;                     // bin pairs per thread (k = M/2 is one more for t = 0)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* zbuf = reinterpret_cast<float2*>(smem_raw);
    float2* tw = zbuf + SF * MP;
    const int tid = threadIdx.x;
    for (int i = tid; i < P::TW_TOTAL; i += NT) tw[i] = a.tw[i];
    const int g = tid / T, t = tid % T;
    const int64_t total_blocks = a.n_streams * a.nblk;
    const int64_t n_work = (total_blocks + SF - 1) / SF;
    float2 uw[Q];                                      // -j e^{-2 pi i k / N} of this thread's pairs, once per CTA
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        float sn, cs;
        sincospif(2.0f * (float)(t + q * T) / (float)N, &sn, &cs);
        uw[q] = make_float2(-sn, -cs);
    }
    float2 x[R];
    auto load_block = [&](int64_t work) {
        const int64_t blk_global = work * SF + g;
        const bool live = blk_global < total_blocks;
        const int64_t stream = live ? blk_global / a.nblk : 0;
        const int b = (int)(blk_global - stream * a.nblk);
        if (live && b == 0 && t == 0 && a.pmax) a.pmax[stream] = __int_as_float(0xff800000);
        const S* row = reinterpret_cast<const S*>(a.r) + stream * a.r_stride;
        const int64_t s0 = (int64_t)b * kB - kB + a.in_off;
        const bool fast = sizeof(S) == 4 && live && !a.reverse && s0 >= 0 && s0 + 2 * kB <= a.T && a.valid_len >= 2 * kB &&
                          ((reinterpret_cast<uintptr_t>(row + s0) & 7) == 0);
        const bool fast_pcm = sizeof(S) < 4 && live && !a.reverse && s0 >= 0 && s0 + 2 * kB <= a.T && a.valid_len >= 2 * kB &&
                              ((reinterpret_cast<uintptr_t>(row + s0) & (2 * sizeof(S) - 1)) == 0);
        if (fast) {      // interior block, 8-byte aligned: vector loads without bounds checks
#pragma unroll
            for (int i = 0; i < R; ++i) x[i] = __ldg(reinterpret_cast<const float2*>(row + s0) + (t + i * T));
        } else if (fast_pcm) {      // PCM as recorded: one 32- / 16-bit load per sample pair
            if constexpr (sizeof(S) == 2) {
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    const short2 v = __ldg(reinterpret_cast<const short2*>(row + s0) + (t + i * T));
                    x[i] = make_float2((float)v.x, (float)v.y);
                }
            } else if constexpr (sizeof(S) == 1) {
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    const uchar2 v = __ldg(reinterpret_cast<const uchar2*>(row + s0) + (t + i * T));
                    x[i] = make_float2((float)v.x, (float)v.y);
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const int loc = 2 * (t + i * T);
                const int64_t n = s0 + loc;
                float v0 = 0.f, v1 = 0.f;
                if (live) {
                    const bool ok0 = n >= 0 && n < a.T && loc < a.valid_len;
                    const bool ok1 = n + 1 >= 0 && n + 1 < a.T && loc + 1 < a.valid_len;
                    if (!a.reverse) {
                        if (ok0) v0 = (float)row[n];
                        if (ok1) v1 = (float)row[n + 1];
                    } else {
                        if (ok0) v0 = (float)row[a.T - 1 - n];
                        if (ok1) v1 = (float)row[a.T - 2 - n];
                    }
                }
                x[i] = make_float2(v0, v1);
            }
        }
    };
    // persistent: the twiddle table (~30 KB) is staged once per CTA, not once per pair of blocks
    if ((int64_t)blockIdx.x < n_work) load_block(blockIdx.x);
    for (int64_t work = blockIdx.x; work < n_work; work += gridDim.x) {
        __syncthreads();                               // twiddles staged / previous untangle done with zbuf
        float2* zs = zbuf + g * MP;
        fft_forward<P, NT, true>(x, zs, tw, t, g);     // natural order: the mirrored reads below stay conflict-free
        __syncthreads();
        if (work + gridDim.x < n_work) load_block(work + gridDim.x);
        const int64_t bg = work * SF + g;
        if (bg >= total_blocks) continue;
        float2* out = a.spec + bg * M;
        // untangle: X[k] = (s + w2 d)/2, X[M-k] = conj(s - w2 d)/2
        auto pair = [&](int k, float2 w2, float2& x1, float2& x2) {
            const float2 z1 = zs[k], z2 = zs[(M - k) & (M - 1)];
            const float2 s = make_float2(z1.x + z2.x, z1.y - z2.y);
            const float2 d = make_float2(z1.x - z2.x, z1.y + z2.y);
            const float2 tt = cmul(w2, d);
            x1 = make_float2(0.5f * (s.x + tt.x), 0.5f * (s.y + tt.y));
            x2 = make_float2(0.5f * (s.x - tt.x), 0.5f * (tt.y - s.y));
        };
        float e1[Q], e2[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int k = t + q * T;
            float2 x1, x2;
            pair(k, uw[q], x1, x2);
            if (q == 0 && t == 0) {
                out[0] = make_float2(x1.x, x2.x);      // (X[0], X[M]), both real
                e1[q] = x1.x * x1.x;
                e2[q] = x2.x * x2.x;
            } else {
                out[k] = x1;
                out[M - k] = x2;
                e1[q] = 2.f * fmaf(x1.x, x1.x, x1.y * x1.y);
                e2[q] = 2.f * fmaf(x2.x, x2.x, x2.y * x2.y);
            }
        }
        if (t == 0) {                                   // k = M/2 pairs with itself; w2 = -j e^{-j pi/2} = -1
            float2 x1, x2;
            pair(M / 2, make_float2(-1.f, 0.f), x1, x2);
            out[M / 2] = x1;
            e1[0] += 2.f * fmaf(x1.x, x1.x, x1.y * x1.y);
        }
        if (a.energy) {                                 // (uniform)
            // 16 sums over the warp's 32 lanes by a halving exchange (16 shuffles instead of 80): after the steps with
            // lane distances 16, 8, 4, 2 a lane holds ONE slot -- bits 4..1 of its index say which -- summed over 16
            // lanes, the last step adds the other 16.  Two warps share a group: one atomic each.
            static_assert(Q == 8, "energy slots: 8 pairs per thread");
            float v[16];
#pragma unroll
            for (int q = 0; q < Q; ++q) { v[q] = e1[q]; v[8 + q] = e2[q]; }
            const int lane = t & 31;
#pragma unroll
            for (int h = 8; h >= 1; h >>= 1) {           // h = number of slots kept; partner distance 2h
                const bool up = (lane & (2 * h)) != 0;
#pragma unroll
                for (int i = 0; i < h; ++i) {
                    const float give = up ? v[i] : v[i + h];
                    const float got = __shfl_xor_sync(0xffffffffu, give, 2 * h);
                    v[i] = (up ? v[i + h] : v[i]) + got;
                }
            }
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
            if ((lane & 1) == 0) {
                const int slot = lane >> 1;               // bits 4..1: 0..7 = e1[q], 8..15 = e2[q - 8]
                atomicAdd(a.energy + bg * kGroups + (slot < 8 ? 2 * slot : 16 + 2 * (slot - 8)) + (t >= T / 2 ? 1 : 0), v[0]);
            }
        }
    }
}

struct AccArgs {
    const float2* spec;      // [n_streams, nblk_in, M]
    const float2* H;         // [parts, M]
    const float2* tw;
    float* P;                // [n_streams, p_stride]
    float* pmax;             // [n_streams]
    float* blockmax;         // [n_streams, nblk_out] or null: signed maximum of every output block (guides the detection walk)
    const float* bound;      // detection only: [n_streams, nblk_out] upper bound of |P| in every block (xcorr_mac_kernel<.., true>), else null
    float thresh;            // detection threshold (OFDM.py:361)
    int select;              // 0: every block; 1: per stream, the block with the largest bound (n_work = n_streams);
                             // 2: the blocks that can hold a candidate, and their neighbours (see the kernel);
                             // 3: the blocks listed in sel_idx[0 .. *sel_count) (xcorr_select_kernel)
    const int* sel_idx;
    const int* sel_count;
    int64_t p_stride, out_len;   // out_len = T + Lc - 1
    int nblk_in, nblk_out, parts;
    int64_t n_work;          // n_streams * nblk_out
};

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// One CTA (128 threads = one symbol group) per output block.  Thread t owns the bin pairs (k, M-k),
// k = t + 128 q (k = M/2, which pairs with itself, takes the slot of k = 0); thread 0 also carries
// the packed (DC, Nyquist) element, whose products are real.
__global__ void __launch_bounds__(128, 4) xcorr_acc_kernel(const AccArgs a) {
    using P = SP;
    constexpr int NT = 128, T = P::T, R = P::R, M = P::M, N = P::N, MP = P::MP;
    static_assert(T == NT, "one symbol group per CTA");
    constexpr int Q = (M / 2) / NT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* zbuf = reinterpret_cast<float2*>(smem_raw);
    float2* tw = zbuf + MP;
    __shared__ float wmax[NT / 32];
    const int tid = threadIdx.x;
    for (int i = tid; i < P::TW_TOTAL; i += NT) tw[i] = a.tw[i];
    const int k0 = tid == 0 ? M / 2 : tid;             // bin of slot 0; slot q is k0 + q * NT
    float2 iw[Q];                                      // e^{+j 2 pi k / N} of this thread's bin pairs, once per CTA
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        float sn, cs;
        sincospif(2.0f * (float)(q == 0 ? k0 : tid + q * NT) / (float)N, &sn, &cs);
        iw[q] = make_float2(cs, sn);
    }
    // persistent: the twiddle table is staged once per CTA; co-resident CTAs work on neighbouring
    // blocks, so the chirp-partition spectra and the shared input spectra stay hot in L2
    __shared__ int s_sel;
    const int64_t n_work = a.select == 3 ? (int64_t)__ldg(a.sel_count) : a.n_work;
    for (int64_t work = blockIdx.x; work < n_work; work += gridDim.x) {
        const int64_t item = a.select == 3 ? (int64_t)__ldg(a.sel_idx + work) : work;
        int64_t stream = item / a.nblk_out;
        int b = (int)(item - stream * a.nblk_out);
        if (a.select == 1) {
            // the block with the largest bound very likely holds the chirp peak: its maximum seeds pmax[stream], the
            // lower bound of max(P) that the selection of pass 2 needs
            stream = work;
            const float* bd = a.bound + stream * a.nblk_out;
            float best = -1.f;
            int bi = 0;
            for (int i = tid; i < a.nblk_out; i += NT) {
                const float v = bd[i];
                if (v > best) { best = v; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            __shared__ float wbest[NT / 32];
            __shared__ int wsel[NT / 32];
            if ((tid & 31) == 0) { wbest[tid >> 5] = best; wsel[tid >> 5] = bi; }
            __syncthreads();
            if (tid == 0) {
                for (int w = 1; w < NT / 32; ++w)
                    if (wbest[w] > best) { best = wbest[w]; bi = wsel[w]; }
                s_sel = bi;
            }
            __syncthreads();
            b = s_sel;
        } else if (a.select == 2) {
            // |P| <= bound inside a block, and a candidate needs P > thresh * max(P) >= thresh * pmax[stream] (any lower
            // bound of the maximum will do: pass 1 seeded it, the atomics of this pass only raise it).  A block whose
            // bound passes that test -- every block that really holds a candidate or the maximum does, whoever
            // evaluates it and whenever -- is transformed back together with both its neighbours (the detection rule
            // reads one sample either side of a candidate); so are the first and the last block.  One thread decides:
            // pmax moves while the others would read it.
            if (tid == 0) {
                const float* bd = a.bound + stream * a.nblk_out;
                const float lim = a.thresh * a.pmax[stream];
                bool go = b == 0 || b == a.nblk_out - 1 || !(lim > 0.f);
                for (int i = (b > 0 ? b - 1 : 0); i <= b + 1 && i < a.nblk_out; ++i) go = go || bd[i] * 1.001f >= lim;
                s_sel = go ? 1 : 0;
                if (!go && a.blockmax) a.blockmax[stream * a.nblk_out + b] = __int_as_float(0xff800000);
            }
            __syncthreads();
            const int go = s_sel;
            __syncthreads();
            if (!go) continue;
        }
        __syncthreads();

        float2 acc1[Q], acc2[Q];
        float dc = 0.f, ny = 0.f;
#pragma unroll
        for (int q = 0; q < Q; ++q) acc1[q] = acc2[q] = make_float2(0.f, 0.f);
        const int p_lo = b - (a.nblk_in - 1) > 0 ? b - (a.nblk_in - 1) : 0;
        const int p_hi = b < a.parts - 1 ? b : a.parts - 1;
        const float2* X = a.spec + (stream * a.nblk_in + (b - p_lo)) * M;
        const float2* H = a.H + (int64_t)p_lo * M;
        for (int p = p_lo; p <= p_hi; ++p, X -= M, H += M) {
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const int k = q == 0 ? k0 : tid + q * NT, km = M - k;
                const float2 x1 = X[k], h1 = __ldg(H + k), x2 = X[km], h2 = __ldg(H + km);
                acc1[q].x = fmaf(x1.x, h1.x, fmaf(-x1.y, h1.y, acc1[q].x));
                acc1[q].y = fmaf(x1.x, h1.y, fmaf(x1.y, h1.x, acc1[q].y));
                acc2[q].x = fmaf(x2.x, h2.x, fmaf(-x2.y, h2.y, acc2[q].x));
                acc2[q].y = fmaf(x2.x, h2.y, fmaf(x2.y, h2.x, acc2[q].y));
            }
            if (tid == 0) {        // packed (DC, Nyquist): component-wise real products
                const float2 xv = X[0], hv = __ldg(H);
                dc = fmaf(xv.x, hv.x, dc);
                ny = fmaf(xv.y, hv.y, ny);
            }
        }
        // inverse untangle: Z[k] = E + jO, E = Y[k] + conj Y[M-k], O = (Y[k] - conj Y[M-k]) e^{+j theta};
        // the forward engine then runs on conj Z.  Plain indexing: this hand-over and the output below
        // are both conflict-free without padding
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int k = q == 0 ? k0 : tid + q * NT, km = M - k;
            const float2 Y1 = acc1[q], Y2 = acc2[q];
            const float2 E = make_float2(Y1.x + Y2.x, Y1.y - Y2.y);
            const float2 D = make_float2(Y1.x - Y2.x, Y1.y + Y2.y);
            const float2 O = cmul(D, iw[q]);
            zbuf[k] = make_float2(E.x - O.y, -(E.y + O.x));
            zbuf[km] = make_float2(E.x + O.y, E.y - O.x);      // k = M/2: the same value to the same slot
        }
        if (tid == 0) zbuf[0] = make_float2(dc + ny, ny - dc);
        __syncthreads();
        float2 x[R];
#pragma unroll
        for (int i = 0; i < R; ++i) x[i] = zbuf[tid + i * T];
        __syncthreads();
        fft_forward<P, NT, true>(x, zbuf, tw, tid, 0);
        __syncthreads();
        // last B samples of the block: z[m], m in [M/2, M);  x[2m] = Re Y/N, x[2m+1] = -Im Y/N
        const float scale = 1.0f / (float)N;
        float* Prow = a.P + stream * a.p_stride;
        float lmax = __int_as_float(0xff800000);
        const int64_t n0 = (int64_t)b * kB - kB;
        if (n0 + 2 * M <= a.out_len && ((reinterpret_cast<uintptr_t>(Prow + n0) & 7) == 0)) {
#pragma unroll
            for (int i = 0; i < M / 2 / NT; ++i) {
                const int m = M / 2 + tid + i * NT;
                const float2 y = zbuf[m];
                const float2 v = make_float2(y.x * scale, -y.y * scale);
                *reinterpret_cast<float2*>(Prow + n0 + 2 * m) = v;
                lmax = fmaxf(lmax, fmaxf(v.x, v.y));
            }
        } else {
            for (int m = M / 2 + tid; m < M; m += NT) {
                const float2 y = zbuf[m];
                const int64_t n = n0 + 2 * m;
                const float v0 = y.x * scale, v1 = -y.y * scale;
                if (n < a.out_len) { Prow[n] = v0; lmax = fmaxf(lmax, v0); }
                if (n + 1 < a.out_len) { Prow[n + 1] = v1; lmax = fmaxf(lmax, v1); }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        if ((tid & 31) == 0) wmax[tid >> 5] = lmax;
        __syncthreads();
        if (tid == 0) {
            float m = wmax[0];
            for (int w = 1; w < NT / 32; ++w) m = fmaxf(m, wmax[w]);
            if (a.blockmax) a.blockmax[stream * a.nblk_out + b] = m;
            if (m > __int_as_float(0xff800000)) atomic_max_float(a.pmax + stream, m);
        }
    }
}


// ------------------------------------------------------------------------------------------
// Long chirps (the N = 4096 modes: 21 600 - 26 400 taps = 11 - 13 partitions): the partition sum as its own kernel.
// xcorr_acc_kernel above reads `parts` input spectra AND `parts` chirp partitions from L2 for every output block
// (352 KB per block at 11 partitions: L2-bandwidth bound).  Here one 512-thread CTA per SM keeps ALL chirp
// partitions in shared memory (parts x 16 KB <= 208 KB) and computes a run of J consecutive output blocks of one
// stream at a time: every input spectrum of the run's window is read once and used for up to J outputs
// ((J + parts - 1) / J = 2.25 reads per output at J = 8, parts = 11, instead of 11).  Y goes to a scratch array in
// the spectrum layout; xcorr_acc_kernel with ONE all-ones partition then does the inverse transform.  The sum runs
// over p ascending with the fmaf nesting of xcorr_acc_kernel, so P is bit-identical to the two-kernel form (tested).
// ------------------------------------------------------------------------------------------
constexpr int kMacThreads = 1024;        // 32 warps per SM, two bins per thread
constexpr int kMacJ = 8;                  // output blocks per run
constexpr int kMacMinParts = 5, kMacMaxParts = 13;          // up to 13 x 16 KB of shared memory
struct MacArgs {
    const float2* spec;      // [n_streams, nblk_in, M]
    const float2* H;         // [parts, M]
    float2* Y;               // [n_streams, nblk_out, M]
    int nblk_in, nblk_out;
    int runs_per_stream;     // ceil(nblk_out / J)
    int64_t n_units;         // n_streams * runs_per_stream
    const uint8_t* run_sel;  // optional [n_units]: only the runs flagged here are computed (detection only)
};

// The (input block, output) pairs of a run form a fixed band (output j takes input b0 + j - p, p < PARTS): with PARTS a
// template parameter the whole run is straight-line code -- no per-pair index arithmetic or predicates, partitions
// at immediate shared-memory offsets.  Input blocks outside the stream read as zero (adding x * h = +0 leaves the sums
// as they are), outputs past the last block are computed and dropped.
template <int PARTS>
__global__ void __launch_bounds__(kMacThreads, 1) xcorr_mac_kernel(const MacArgs a) {
    constexpr int NT = kMacThreads, M = SP::M, J = kMacJ, NB = M / NT;       // NB = 2 bins per thread: k = tid + NT i
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* Hs = reinterpret_cast<float2*>(smem_raw);                       // [PARTS][M]
    const int tid = threadIdx.x;
    for (int i = tid; i < PARTS * (M / 2); i += NT)
        reinterpret_cast<float4*>(Hs)[i] = __ldg(reinterpret_cast<const float4*>(a.H) + i);
    __syncthreads();
    // Element 0 packs (DC, Nyquist), whose products are component-wise.  With the Nyquist term taken out of the copy
    // in shared memory, the complex multiply-add of slot 0 gives the DC sum exactly (x.x h.x - x.y 0 + acc); warp 0
    // carries the Nyquist sums beside it (lane 0's are the real ones).
    __shared__ float hny[PARTS];
    if (tid < PARTS) { hny[tid] = Hs[tid * M].y; Hs[tid * M].y = 0.f; }
    __syncthreads();
    // (the partitions are re-read from shared memory for every pair on purpose -- volatile, or the compiler keeps the
    //  thread's 4 x PARTS values in registers across the whole kernel and spills the sums)
    const unsigned ht = (unsigned)__cvta_generic_to_shared(Hs + tid);
    auto lds_h = [&](auto offc) -> float2 {
        float2 v;
        asm volatile("ld.volatile.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(v.x), "=f"(v.y) : "r"(ht), "n"(decltype(offc)::value));
        return v;
    };
    uint64_t keep;           // L2 policy of the spectra: every block is read by two or three runs
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
    auto run = [&](auto nyc) {
        constexpr bool NY = decltype(nyc)::value;
        for (int64_t unit = blockIdx.x; unit < a.n_units; unit += gridDim.x) {      // (contiguous ranges per CTA: no faster, 1.25 vs 1.21 ms)
            if (a.run_sel && !a.run_sel[unit]) continue;                              // (uniform over the CTA)
            const int64_t stream = unit / a.runs_per_stream;
            const int b0 = (int)(unit - stream * a.runs_per_stream) * J;
            float2 acc[J][NB];
            [[maybe_unused]] float ny[J];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                if constexpr (NY) ny[j] = 0.f;
#pragma unroll
                for (int i = 0; i < NB; ++i) acc[j][i] = make_float2(0.f, 0.f);
            }
            // running pointer / index of the block being requested (kept opaque: one register each, not a table of addresses)
            int bq = b0 + J - 1;
            const float2* xp = a.spec + (stream * (int64_t)a.nblk_in + bq) * M + tid;
            auto load_x = [&](float2 (&dst)[NB]) {
                if (bq >= 0 && bq < a.nblk_in) {
#pragma unroll
                    for (int i = 0; i < NB; ++i)
                        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(dst[i].x), "=f"(dst[i].y) : "l"(xp + i * NT), "l"(keep));
                } else {
#pragma unroll
                    for (int i = 0; i < NB; ++i) dst[i] = make_float2(0.f, 0.f);
                }
                --bq;
                xp -= M;
                asm volatile("" : "+r"(bq), "+l"(xp));
            };
            float2 xn[NB], xn2[NB];                                           // the next two input blocks are in flight
            load_x(xn);
            load_x(xn2);
            // input blocks newest first: output j then sees p = j - d ascending, as xcorr_acc_kernel sums
            static_for<J + PARTS - 1>([&](auto dc_) {
                constexpr int d = J - 1 - decltype(dc_)::value;               // input block b0 + d, d = J-1 ... -(PARTS-1)
                float2 x[NB];
#pragma unroll
                for (int i = 0; i < NB; ++i) { x[i] = xn[i]; xn[i] = xn2[i]; }
                if constexpr (d > -(PARTS - 1) + 1) load_x(xn2);
                static_for<J>([&](auto jc) {
                    constexpr int j = decltype(jc)::value, p = j - d;
                    if constexpr (p >= 0 && p < PARTS) {
                        static_for<NB>([&](auto ic) {
                            constexpr int i = decltype(ic)::value;
                            const float2 h = lds_h(std::integral_constant<int, (p * M + i * NT) * 8>{});
                            acc[j][i].x = fmaf(x[i].x, h.x, fmaf(-x[i].y, h.y, acc[j][i].x));
                            acc[j][i].y = fmaf(x[i].x, h.y, fmaf(x[i].y, h.x, acc[j][i].y));
                        });
                        if constexpr (NY) ny[j] = fmaf(x[0].y, hny[p], ny[j]);
                    }
                });
            });
            const int nj = a.nblk_out - b0;
            float2* Yo = a.Y + (stream * (int64_t)a.nblk_out + b0) * M + tid;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                if (j < nj) {
                    if constexpr (NY) { if (tid == 0) acc[j][0].y = ny[j]; }
#pragma unroll
                    for (int i = 0; i < NB; ++i)          // streaming store: the sums must not push the spectra out of L2
                        asm volatile("st.global.cs.v2.f32 [%0], {%1, %2};" ::"l"(Yo + (int64_t)j * M + i * NT), "f"(acc[j][i].x), "f"(acc[j][i].y) : "memory");
                }
            }
        }
    };
    if (tid < 32) run(std::true_type{});
    else run(std::false_type{});
}

// Detection only: an upper bound of |P| inside every output block from the group energies of the input spectra and
// of the chirp partitions, without forming the partition sums:
//   |y[n]| <= (1/N) sum_k w_k |Y_b[k]| <= (1/N) sum_p sum_k w_k |X_{b-p}[k]| |H_p[k]|
//          <= (1/N) sum_p sum_groups sqrt(E_{b-p}[g]) sqrt(E_{H_p}[g])            (Cauchy-Schwarz inside every group).
// With 32 groups of 64 neighbouring bins the last step costs a few per cent against the l1 norm itself (a partition of
// a linear chirp is narrow-band and smooth in magnitude); the blocks it rules out are the same.  One warp per block.
__global__ void __launch_bounds__(256) xcorr_bound_kernel(const float* __restrict__ E, const float* __restrict__ HE, float* __restrict__ bound,
                                                          int nblk_in, int nblk_out, int parts, int64_t n_blocks) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t blk = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); blk < n_blocks; blk += nwarps) {
        const int64_t stream = blk / nblk_out;
        const int b = (int)(blk - stream * nblk_out);
        float sum = 0.f;
        for (int p = 0; p < parts; ++p) {
            const int bp = b - p;
            if (bp >= 0 && bp < nblk_in) sum = fmaf(sqrtf(E[(stream * nblk_in + bp) * kGroups + lane]), sqrtf(__ldg(HE + p * kGroups + lane)), sum);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) bound[blk] = sum * (1.0f / (float)SP::N);
    }
}

// Detection only, after pmax[stream] has been seeded with the maximum of the stream's most promising block: the list of
// blocks that are transformed back.  |P| <= bound inside a block and a candidate needs P > thresh * max(P) >= thresh *
// pmax, so a block whose bound passes that test -- every block that really holds a candidate or the maximum does -- is
// listed together with both its neighbours (the detection rule reads one sample either side of a candidate); so are
// the first and the last block of a stream.  All other blocks get block maximum -inf.  run_sel flags the runs of
// kMacJ blocks that hold a listed block (for the partition-sum kernel).
__global__ void __launch_bounds__(256) xcorr_select_kernel(const float* __restrict__ bound, const float* __restrict__ pmax, float thresh,
                                                           int nblk_out, int runs_per_stream, int64_t n_blocks, float* __restrict__ blockmax,
                                                           int* __restrict__ sel_idx, int* __restrict__ sel_count, uint8_t* __restrict__ run_sel) {
    const int64_t blk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (blk >= n_blocks) return;
    const int64_t stream = blk / nblk_out;
    const int b = (int)(blk - stream * nblk_out);
    const float* bd = bound + stream * nblk_out;
    const float lim = thresh * pmax[stream];
    bool go = b == 0 || b == nblk_out - 1 || !(lim > 0.f);
    for (int i = (b > 0 ? b - 1 : 0); i <= b + 1 && i < nblk_out; ++i) go = go || bd[i] * 1.001f >= lim;
    if (go) {
        sel_idx[atomicAdd(sel_count, 1)] = (int)blk;
        run_sel[stream * runs_per_stream + b / kMacJ] = 1;
    } else {
        blockmax[blk] = __int_as_float(0xff800000);
    }
}

__global__ void xcorr_ones_kernel(float2* H1) {       // the unit partition of the inverse stage: (1, 1) packs (DC, Nyquist)
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < SP::M) H1[k] = make_float2(1.f, k == 0 ? 1.f : 0.f);
}


// ------------------------------------------------------------------------------------------
// Fused matched filter: forward FFT + partition multiply-accumulate + inverse FFT of a run of
// consecutive blocks in ONE persistent kernel.  The spectra of the last (parts - 1) input blocks
// stay on chip, so DRAM sees every input sample once and every output sample once (the two-kernel
// form above writes 8 bytes of spectrum per sample and reads them back `parts` times).
//
// One 128-thread CTA = one FFT group.  Thread t owns the bin pairs (k, M-k), k = t + 128 q (k = M/2
// takes the slot of k = 0; thread 0 also carries the real pair (DC, Nyquist)) in every stage: it
// untangles them after the forward FFT, multiplies them with the chirp partitions and turns them
// into the inverse FFT's input.  The ring of past spectra is therefore THREAD-PRIVATE: slot s holds
// the thread's own 8 pairs as float4 (re k, re M-k, im k, im M-k) -- 128-bit conflict-free accesses,
// no barrier around it -- and the chirp partitions are read from an identically laid out global
// table (48 KB for three partitions: L1 / L2 resident), pre-scaled by 1 / (2N) so that neither the
// untangle's 1/2 nor the inverse transform's 1/N costs an instruction.  Both lanes of FFMA2 carry
// the two bins of a pair.  The inverse FFT's last pass stays in registers and goes straight to
// global memory (only the last B samples of a block are output samples).
//
// Work = (stream, output block) pairs in stream-major order; every CTA takes one contiguous range.
// A CTA that starts in the middle of a stream first computes the (parts - 1) spectra before its
// first block (forward FFTs only).
// ------------------------------------------------------------------------------------------
struct FusedArgs {
    const void* r;           // [n_streams, r_stride] samples (float32 / int16 / uint8)
    const float4* Hs;        // [parts][8][128] pair-interleaved chirp partitions, pre-scaled
    const float2* Hdc;       // [parts] (H[0], H[M]) pre-scaled
    const float2* tw;
    float* P;                // [n_streams, p_stride]
    float* pmax;             // [n_streams], preset to -inf
    float* blockmax;         // [n_streams, nblk_out] or null
    int64_t r_stride, p_stride, T, out_len, n_streams;
    int nblk_in, nblk_out, parts;
    int sparse;              // 1: detection-only -- blocks that provably hold no candidate are not transformed back
    float thresh;            // detection threshold (OFDM.py:361), for the sparse rule
    int* counter;            // sparse: next stream to take (zeroed before the launch)
};

template <class S> __device__ __forceinline__ float sample_to_f32(S v) { return (float)v; }

template <class S, int PARTS, int MINB>
__global__ void __launch_bounds__(128, MINB) xcorr_fused_kernel(const FusedArgs a) {
    using P = SP;
    constexpr int NT = 128, T = P::T, R = P::R, M = P::M, N = P::N, MP = P::MP, Q = (M / 2) / NT;
    static_assert(T == NT && Q == 8, "one FFT group per CTA");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* W = reinterpret_cast<float2*>(smem_raw);                       // [MP] FFT exchange / spectrum buffer
    float2* tw = W + MP;                                                   // [TW_TOTAL]
    float4* ring = reinterpret_cast<float4*>(tw + P::TW_TOTAL);            // [parts-1][Q][NT]
    float2* ring_dc = reinterpret_cast<float2*>(ring + (size_t)(PARTS - 1) * Q * NT);     // [parts-1] (thread 0)
    __shared__ float wmax[NT / 32];
    const int tid = threadIdx.x;
    constexpr int R1 = PARTS - 1;
    for (int i = tid; i < P::TW_TOTAL; i += NT) tw[i] = a.tw[i];

    // e^{j 2 pi k / N} of this thread's pairs (cos, sin): every twiddle of the two untangles derives from it
    float2 cs[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int k = (q == 0 && tid == 0) ? M / 2 : tid + q * NT;
        sincospif(2.0f * (float)k / (float)N, &cs[q].y, &cs[q].x);
    }
    const pk64 pm = pk_pack(make_float2(1.f, -1.f));

    const int64_t total = a.n_streams * a.nblk_out;
    const int64_t c_begin = total * blockIdx.x / gridDim.x, c_end = total * (blockIdx.x + 1) / gridDim.x;
    const S* base = reinterpret_cast<const S*>(a.r);

    // ---- samples of block bb of stream `st` in the FFT's first-pass layout x[i] = z[t + i T]
    // HALF = 1: only the second half of the block is fetched (x[R/2 ..]); the first half is the previous block's
    // second half, which the caller kept in registers (consecutive blocks overlap by B samples)
    auto load_block = [&](float2 (&x)[R], int64_t st, int bb, auto halfc) {
        constexpr int I0 = decltype(halfc)::value ? R / 2 : 0;
        const S* row = base + st * a.r_stride;
        const int64_t s0 = (int64_t)bb * kB - kB;
        const bool inside = bb >= 0 && bb < a.nblk_in && s0 >= 0 && s0 + 2 * kB <= a.T;
        if constexpr (sizeof(S) == 4) {
            if (inside && ((reinterpret_cast<uintptr_t>(row + s0) & 7) == 0)) {
#pragma unroll
                for (int i = I0; i < R; ++i) {           // streaming: do not evict the chirp partitions from L1
                    const float* sp = reinterpret_cast<const float*>(row + s0) + 2 * (tid + i * T);
                    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(x[i].x), "=f"(x[i].y) : "l"(sp));
                }
                return;
            }
        } else if constexpr (sizeof(S) == 2) {
            if (inside && ((reinterpret_cast<uintptr_t>(row + s0) & 3) == 0)) {
#pragma unroll
                for (int i = I0; i < R; ++i) {
                    const short2 v = __ldg(reinterpret_cast<const short2*>(row + s0) + (tid + i * T));
                    x[i] = make_float2((float)v.x, (float)v.y);
                }
                return;
            }
        } else {
            if (inside && ((reinterpret_cast<uintptr_t>(row + s0) & 1) == 0)) {
#pragma unroll
                for (int i = I0; i < R; ++i) {
                    const uchar2 v = __ldg(reinterpret_cast<const uchar2*>(row + s0) + (tid + i * T));
                    x[i] = make_float2((float)v.x, (float)v.y);
                }
                return;
            }
        }
#pragma unroll
        for (int i = I0; i < R; ++i) {         // edge block / odd alignment: bounds-checked scalar loads, zero outside [0, T)
            const int64_t n = s0 + 2 * (tid + i * T);
            float v0 = 0.f, v1 = 0.f;
            if (bb >= 0 && bb < a.nblk_in) {
                if (n >= 0 && n < a.T) v0 = sample_to_f32(row[n]);
                if (n + 1 >= 0 && n + 1 < a.T) v1 = sample_to_f32(row[n + 1]);
            }
            x[i] = make_float2(v0, v1);
        }
    };
    // ---- forward FFT of x + untangle: 2 X of this thread's pairs as (re k, re M-k), (im k, im M-k)
    auto forward = [&](float2 (&x)[R], pk64 (&xre)[Q], pk64 (&xim)[Q], float2& dcny) {
        __syncthreads();                                   // W is free (previous block's transforms are done with it)
        fft_forward<P, NT, true>(x, W, tw, tid, 0);
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int k = (q == 0 && tid == 0) ? M / 2 : tid + q * NT;
            const pk64 z1 = *reinterpret_cast<const pk64*>(W + k);
            const pk64 z2 = *reinterpret_cast<const pk64*>(W + (M - k));
            const pk64 sa = p_add(z1, z2);                 // (s.x, d.y)
            const pk64 sd = p_sub(z1, z2);                 // (d.x, s.y)
            // tt = w d, w = -j e^{-j theta} = (-sin, -cos): A = (wr, wi), B = (-wi, wr)
            const pk64 wA = pk_pack(make_float2(-cs[q].y, -cs[q].x)), wB = pk_pack(make_float2(cs[q].x, -cs[q].y));
            const pk64 tt = p_fma(p_bc(p_hi(sa)), wB, p_mul(p_bc(p_lo(sd)), wA));
            xre[q] = p_fma(p_bc(p_lo(tt)), pm, p_bc(p_lo(sa)));      // (s.x + tt.x, s.x - tt.x)
            xim[q] = p_fma(p_bc(p_hi(sd)), pm, p_bc(p_hi(tt)));      // (tt.y + s.y, tt.y - s.y)
        }
        const float2 z0 = W[0];
        dcny = make_float2(2.f * (z0.x + z0.y), 2.f * (z0.x - z0.y));   // 2 X[0], 2 X[M] (both real)
    };
    auto ring_store = [&](int slot, const pk64 (&xre)[Q], const pk64 (&xim)[Q], float2 dcny) {
        float4* rs = ring + (size_t)slot * Q * NT + tid;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const float2 re = pk_unpack(xre[q]), im = pk_unpack(xim[q]);
            rs[q * NT] = make_float4(re.x, re.y, im.x, im.y);
        }
        if (tid == 0) ring_dc[slot] = dcny;
    };
    auto ring_zero = [&]() {
#pragma unroll
        for (int sl = 0; sl < R1; ++sl) {
            float4* rs = ring + (size_t)sl * Q * NT + tid;
#pragma unroll
            for (int q = 0; q < Q; ++q) rs[q * NT] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tid == 0) ring_dc[sl] = make_float2(0.f, 0.f);
        }
    };

    // ---- detection only: group energies.  The 32 bins k = t + 128 q a warp owns in slot q form a group, their mirror
    // images M - k another (two contiguous runs of the spectrum; mixing them would pair the chirp band with bins the
    // chirp does not reach and cost the bound a factor sqrt 2).  With the weights of the inverse real transform (2 per
    // bin; DC and Nyquist are carried separately)
    //   sum_{k in group} 2 |X[k]| |H[k]| <= sqrt(sum 2 |X[k]|^2) sqrt(sum 2 |H[k]|^2)          (Cauchy-Schwarz),
    // so a few square roots per block bound |y[n]| <= sum_k w_k |Y[k]|, Y = sum_p X_{b-p} H_p, BEFORE any partition
    // product is formed.  warp_slots16: 16 per-thread values summed over the warp by a halving exchange (16 shuffles);
    // every lane returns the total of slot lane >> 1 (0..7: bins k of slot q, 8..15: bins M - k).
    auto warp_slots16 = [&](float (&v)[2 * Q]) -> float {
        const int lane = tid & 31;
#pragma unroll
        for (int h = 8; h >= 1; h >>= 1) {
            const bool up = (lane & (2 * h)) != 0;
#pragma unroll
            for (int i = 0; i < h; ++i) {
                const float give = up ? v[i] : v[i + h];
                const float got = __shfl_xor_sync(0xffffffffu, give, 2 * h);
                v[i] = (up ? v[i + h] : v[i]) + got;
            }
        }
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
        return v[0];
    };
    float hs[PARTS];                                       // sqrt of the group energy of partition p, this lane's slot
    float hist[PARTS];                                     // the same of the input blocks b-1, b-2, ... ([0] unused)
#pragma unroll
    for (int p = 0; p < PARTS; ++p) hs[p] = hist[p] = 0.f;
    if (a.sparse) {
#pragma unroll
        for (int p = 0; p < PARTS; ++p) {
            float e[2 * Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const float4 h = __ldg(a.Hs + ((size_t)p * Q + q) * NT + tid);
                e[q] = 2.f * fmaf(h.x, h.x, h.z * h.z);
                e[Q + q] = (q == 0 && tid == 0) ? 0.f : 2.f * fmaf(h.y, h.y, h.w * h.w);     // k = M/2 sits in both lanes of its slot: once
            }
            hs[p] = sqrtf(warp_slots16(e));
        }
    }
    __syncthreads();                                       // twiddles staged
    int64_t cur_stream = -1;
    float2 x[R];
    // ---- sparse (detection-only) mode.  A candidate of chirp_method needs P[i+1] > thresh * max(P) (OFDM.py:361).
    // |y[n]| <= sum_k |Y[k]| for every sample of a block, so a block whose spectrum's l1 norm stays below
    // thresh * (the largest sample seen so far in this stream, a lower bound of max(P)) holds no candidate and the
    // global maximum is not in it: its inverse FFT and its stores are skipped (block maximum = -inf).  What the
    // detection rule reads next to a candidate stays exact: the block after a block that may still turn out hot is
    // always computed, a computed block also writes the sample before its first one (the circular convolution of a
    // 2048-tap partition is valid from sample 2047 on) when its predecessor was skipped, and the first and last
    // block of a stream are always computed.  The detections are identical to the full computation's;
    // how many blocks are skipped depends on the data (chirp peak against the l1 norm of the data blocks' spectra:
    // 88 % at 20 dB and 83 % at 8 dB on the C3 framing, none below ~5 dB).
    float runmax = __int_as_float(0xff800000);
    bool prev_hotish = true, prev_skipped = false;
    __shared__ float wl1[NT / 32], wl2[NT / 32];       // partial bounds of the two tests (separate: no barrier between them)
    // Work distribution.  Dense: one contiguous, equally sized range of blocks per CTA (uniform cost).  Sparse: the
    // cost of a block depends on the data, and the skip rule needs the stream's chirp peak, which lies at the stream's
    // beginning -- so CTAs take WHOLE streams from a global counter (dynamic: a CTA whose streams cannot skip much
    // simply takes fewer of them).
    __shared__ long long s_unit;
#pragma unroll 1
    for (int64_t unit = 0;; ++unit) {
    int64_t c_lo = c_begin, c_hi = c_end;
    if (a.sparse) {
        __syncthreads();
        if (tid == 0) s_unit = (long long)atomicAdd(a.counter, 1);
        __syncthreads();
        const int64_t sidx = s_unit;
        if (sidx >= a.n_streams) break;
        c_lo = sidx * a.nblk_out;
        c_hi = c_lo + a.nblk_out;
    } else if (unit > 0) {
        break;
    }
    cur_stream = -1;
#pragma unroll 1
    for (int64_t c = c_lo; c < c_hi; ++c) {
        const int64_t stream = c / a.nblk_out;
        const int b = (int)(c - stream * a.nblk_out);
        pk64 xre[Q], xim[Q];
        float2 dcny;
        if (stream != cur_stream) {
            // ---- (re)start: the ring holds the spectra of blocks b-1 .. b-(parts-1) (zero before the stream)
            cur_stream = stream;
            runmax = __int_as_float(0xff800000);
            prev_hotish = true;
            prev_skipped = false;
#pragma unroll
            for (int p = 0; p < PARTS; ++p) hist[p] = 0.f;
            ring_zero();
            if constexpr (R1 > 0) {
#pragma unroll 1
                for (int bb = b - R1; bb < b; ++bb) {
                    if (bb < 0 || bb >= a.nblk_in) continue;
                    load_block(x, stream, bb, std::false_type{});
                    forward(x, xre, xim, dcny);
                    ring_store(bb % R1, xre, xim, dcny);
                }
            }
            load_block(x, stream, b, std::false_type{});
        }
        // the second half of this block is the first half of the next one: keep it (the transform overwrites x)
        float2 keep[R / 2];
#pragma unroll
        for (int i = 0; i < R / 2; ++i) keep[i] = x[R / 2 + i];
        // ---- X_b
        if (b < a.nblk_in) {
            forward(x, xre, xim, dcny);
        } else {                                            // past the last input sample: an all-zero block
#pragma unroll
            for (int q = 0; q < Q; ++q) xre[q] = xim[q] = 0ull;
            dcny = make_float2(0.f, 0.f);
        }
        // the next block's samples fly while this one is multiplied and transformed back
        if (c + 1 < c_hi && (c + 1) / a.nblk_out == stream) {
#pragma unroll
            for (int i = 0; i < R / 2; ++i) x[i] = keep[i];
            load_block(x, stream, b + 1, std::true_type{});
        }

        float* Prow = a.P + stream * a.p_stride;
        const int64_t n0 = (int64_t)b * kB - kB;
        if (a.sparse) {
            // bound of |y[n]| over the block from the group energies (see above); the partition sums, the inverse FFT
            // and the stores are skipped when it stays below thresh * (largest sample of the stream so far)
            float e[2 * Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const float2 m2 = pk_unpack(p_fma(xre[q], xre[q], p_mul(xim[q], xim[q])));
                e[q] = 2.f * m2.x;
                e[Q + q] = (q == 0 && tid == 0) ? 0.f : 2.f * m2.y;
            }
            const float eb = sqrtf(warp_slots16(e));
            float part = eb * hs[0];
#pragma unroll
            for (int p = 1; p < PARTS; ++p) part = fmaf(hist[p], hs[p], part);
            if ((tid & 1) != 0) part = 0.f;                                  // two lanes hold every slot
            if (tid == 0) {
                const float2 h0 = __ldg(a.Hdc);
                part += fabsf(dcny.x * h0.x) + fabsf(dcny.y * h0.y);
#pragma unroll
                for (int p = 1; p <= R1; ++p) {
                    const float2 xv = ring_dc[((b - p) % R1 + R1) % R1], hp = __ldg(a.Hdc + p);
                    part += fabsf(xv.x * hp.x) + fabsf(xv.y * hp.y);
                }
            }
#pragma unroll
            for (int p = PARTS - 1; p >= 2; --p) hist[p] = hist[p - 1];
            if constexpr (PARTS > 1) hist[1] = eb;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if ((tid & 31) == 0) wl1[tid >> 5] = part;
            __syncthreads();
            float bound = wl1[0];
#pragma unroll
            for (int w = 1; w < NT / 32; ++w) bound += wl1[w];
            const bool forced = c == c_lo || c + 1 == c_hi || prev_hotish;
            if (!forced && bound * 1.001f < a.thresh * runmax) {             // (runmax = -inf or <= 0: never true)
                if constexpr (R1 > 0) ring_store(b % R1, xre, xim, dcny);
                if (tid == 0 && a.blockmax) a.blockmax[stream * a.nblk_out + b] = __int_as_float(0xff800000);
                prev_hotish = false;
                prev_skipped = true;
                continue;
            }
        }
        // ---- Y_b = sum_p X_{b-p} H_p   (re = pos - neg; all four products are plain FFMA2)
        pk64 arp[Q], arn[Q], ai[Q];
        float dc, ny;
        {
            const float4* Hp = a.Hs + tid;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const float4 h = __ldg(Hp + q * NT);
                const pk64 hre = pk_pack(make_float2(h.x, h.y)), him = pk_pack(make_float2(h.z, h.w));
                arp[q] = p_mul(xre[q], hre);
                arn[q] = p_mul(xim[q], him);
                ai[q] = p_fma(xim[q], hre, p_mul(xre[q], him));
            }
            const float2 h0 = __ldg(a.Hdc);
            dc = dcny.x * h0.x;
            ny = dcny.y * h0.y;
        }
#pragma unroll
        for (int p = 1; p <= R1; ++p) {
            const int slot = ((b - p) % R1 + R1) % R1;
            const float4* rs = ring + (size_t)slot * Q * NT + tid;
            const float4* Hp = a.Hs + (size_t)p * Q * NT + tid;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const float4 xv = rs[q * NT];
                const float4 h = __ldg(Hp + q * NT);
                const pk64 pre = pk_pack(make_float2(xv.x, xv.y)), pim = pk_pack(make_float2(xv.z, xv.w));
                const pk64 hre = pk_pack(make_float2(h.x, h.y)), him = pk_pack(make_float2(h.z, h.w));
                arp[q] = p_fma(pre, hre, arp[q]);
                arn[q] = p_fma(pim, him, arn[q]);
                ai[q] = p_fma(pre, him, ai[q]);
                ai[q] = p_fma(pim, hre, ai[q]);
            }
            if (tid == 0) {
                const float2 xv = ring_dc[slot], h0 = __ldg(a.Hdc + p);
                dc = fmaf(xv.x, h0.x, dc);
                ny = fmaf(xv.y, h0.y, ny);
            }
        }
        if constexpr (R1 > 0) ring_store(b % R1, xre, xim, dcny);   // X_b replaces X_{b-(parts-1)}, which was read just above

        // Y of the thread's pairs
#pragma unroll
        for (int q = 0; q < Q; ++q) arp[q] = p_sub(arp[q], arn[q]);
        if (a.sparse) {
            // second level, for the blocks the energy bound let through: the l1 norm of the block's spectrum itself,
            // sum_k |Y[k]| (bins 1..N/2-1 count twice: Hermitian halves).  |Y| by
            // rsqrt (relative error 2^-22, covered by the 1.001 slack of the comparison below)
            float l1 = 0.f;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const pk64 m2 = p_fma(arp[q], arp[q], p_mul(ai[q], ai[q]));   // |Y|^2 of both bins of the pair
                const float2 e = pk_unpack(m2);
                const float lane0 = e.x * rsqrtf(fmaxf(e.x, 1e-37f)), lane1 = e.y * rsqrtf(fmaxf(e.y, 1e-37f));
                l1 += (q == 0 && tid == 0) ? lane0 : lane0 + lane1;          // k = M/2 sits in both lanes of its slot
            }
            l1 *= 2.f;
            if (tid == 0) l1 += fabsf(dc) + fabsf(ny);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) l1 += __shfl_xor_sync(0xffffffffu, l1, o);
            if ((tid & 31) == 0) wl2[tid >> 5] = l1;
            __syncthreads();
            float bound = wl2[0];
#pragma unroll
            for (int w = 1; w < NT / 32; ++w) bound += wl2[w];
            const bool forced = c == c_lo || c + 1 == c_hi || prev_hotish;      // (as above)
            if (!forced && bound * 1.001f < a.thresh * runmax) {             // (runmax = -inf or <= 0: never true)
                if (tid == 0 && a.blockmax) a.blockmax[stream * a.nblk_out + b] = __int_as_float(0xff800000);
                prev_hotish = false;
                prev_skipped = true;
                continue;
            }
        }
        // ---- inverse untangle: Z[k] = E + jO, E = Y[k] + conj Y[M-k], O = (Y[k] - conj Y[M-k]) e^{+j theta};
        // the forward engine runs on conj Z
        __syncthreads();                                   // every thread has read its part of the forward spectrum in W
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int k = (q == 0 && tid == 0) ? M / 2 : tid + q * NT;
            const float2 yre = pk_unpack(arp[q]), yim = pk_unpack(ai[q]);
            const float2 E = make_float2(yre.x + yre.y, yim.x - yim.y);
            const float2 D = make_float2(yre.x - yre.y, yim.x + yim.y);
            const float2 O = cmul(D, cs[q]);
            W[k] = make_float2(E.x - O.y, -(E.y + O.x));
            W[M - k] = make_float2(E.x + O.y, E.y - O.x);  // k = M/2: the same value to the same slot
        }
        if (tid == 0) W[0] = make_float2(dc + ny, ny - dc);
        __syncthreads();
        float2 y[R];
#pragma unroll
        for (int i = 0; i < R; ++i) y[i] = W[tid + i * T];
        __syncthreads();
        fft_forward_to_regs<P, NT>(y, W, tw, tid, 0);
        // ---- last B samples of the block: z[m], m in [M/2, M): P[n0 + 2m] = Re z, P[n0 + 2m + 1] = -Im z.
        // Last pass (radix 8, stride 256): y[q*8 + i] is z[tid + 128 q + 256 i]
        float lmax = __int_as_float(0xff800000);
        // the sample before this block's first one, when the block that owns it was skipped (z[1023], odd part)
        if (prev_skipped && tid == NT - 1 && b > 0) Prow[n0 + 2 * M / 2 - 1] = -y[8 + 3].y;
        const bool fast = n0 + 2 * M <= a.out_len && ((reinterpret_cast<uintptr_t>(Prow + n0) & 7) == 0);
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int i = 4; i < 8; ++i) {
                const int m = tid + q * T + 256 * i;
                const float2 v = make_float2(y[q * 8 + i].x, -y[q * 8 + i].y);
                const int64_t n = n0 + 2 * m;
                if (fast) {
                    *reinterpret_cast<float2*>(Prow + n) = v;
                    lmax = fmaxf(lmax, fmaxf(v.x, v.y));
                } else {
                    if (n < a.out_len) { Prow[n] = v.x; lmax = fmaxf(lmax, v.x); }
                    if (n + 1 < a.out_len) { Prow[n + 1] = v.y; lmax = fmaxf(lmax, v.y); }
                }
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        if ((tid & 31) == 0) wmax[tid >> 5] = lmax;
        __syncthreads();
        if (tid == 0) {
            float mx = wmax[0];
            for (int w = 1; w < NT / 32; ++w) mx = fmaxf(mx, wmax[w]);
            if (a.blockmax) a.blockmax[stream * a.nblk_out + b] = mx;
            if (mx > __int_as_float(0xff800000)) atomic_max_float(a.pmax + stream, mx);
            wmax[0] = mx;
        }
        if (a.sparse) {                                    // (uniform: every thread takes the same decisions)
            __syncthreads();
            const float mx = wmax[0];
            runmax = fmaxf(runmax, mx);
            prev_hotish = mx * 1.001f >= a.thresh * runmax;
            prev_skipped = false;
        }
    }
    }   // units
}

// chirp-partition spectra [parts][M] (element 0 packs (H[0], H[M])) -> the fused kernel's pair layout, scaled
__global__ void xcorr_pack_h_kernel(const float2* __restrict__ H, int parts, float scale, float4* __restrict__ Hs, float2* __restrict__ Hdc) {
    constexpr int M = SP::M, NT = 128, Q = (M / 2) / NT;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= parts * Q * NT) return;
    const int p = i / (Q * NT), q = (i / NT) % Q, t = i % NT;
    const int k = (q == 0 && t == 0) ? M / 2 : t + q * NT;
    const float2 h1 = H[(size_t)p * M + k], h2 = H[(size_t)p * M + (M - k)];
    Hs[i] = make_float4(h1.x * scale, h2.x * scale, h1.y * scale, h2.y * scale);
    if (q == 0 && t == 0) Hdc[p] = make_float2(H[(size_t)p * M].x * scale, H[(size_t)p * M].y * scale);
}

__global__ void fill_f32_kernel(float* p, int64_t n, float v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ------------------------------------------------------------------------------------------
// chirp_method's detection rule + get_symbols' bookkeeping, one CTA per stream.
//   zeros[i] = (D[i]*D[i+1] <= 0) & (Pn[i+1] > thresh),  Pn = P/pmax, D = diff(Pn), i in [0, len-2)
//   ascending scan: a surviving detection at i clears zeros[i+1 .. i+Lc]; if i + Lc >= len(zeros)
//   the reference's except-branch wipes every detection (OFDM.py:366-370).
// ------------------------------------------------------------------------------------------
struct PeakArgs {
    const float* P;
    const float* pmax;
    uint32_t* mask;              // [n_streams, nw] candidate bits (work buffer)
    int64_t* peaks;
    int32_t* count;
    int64_t p_stride, plen;      // plen = T + Lc - 1
    int64_t n_streams, nw;       // nw = 32-bit words per stream = ceil((plen - 2) / 32)
    int32_t max_peaks, Lc;
    float thresh;
    const float* blockmax;       // optional [n_streams, nblk]: maximum of every 2048-sample block of P (fused matched filter)
    int32_t nblk;
    int32_t sparse;              // P holds valid samples only in (and next to) blocks whose maximum passes the threshold
};

// Phase 1, fully parallel: one candidate bit per position of `zeros` (OFDM.py:360-361).  A CTA takes
// chunks of 64 mask words (2048 positions) of one stream; each warp evaluates 32 consecutive positions
// per step with coalesced loads and stores one word of the bit mask.
__global__ void __launch_bounds__(256) peak_mark_kernel(const PeakArgs a) {
    constexpr int WPC = 64;                                        // mask words per chunk
    const int64_t nz = a.plen - 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t cps = (a.nw + WPC - 1) / WPC;                    // chunks per stream
    const int64_t n_chunks = a.n_streams * cps;
    for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const int64_t stream = c / cps, w0 = (c - stream * cps) * WPC + warp * (WPC / 8);
        const float* Prow = a.P + stream * a.p_stride;
        uint32_t* mrow = a.mask + stream * a.nw;
        const float inv = 1.0f / a.pmax[stream];
#pragma unroll
        for (int j = 0; j < WPC / 8; ++j) {
            const int64_t w = w0 + j;
            if (w >= a.nw) break;
            const int64_t i = w * 32 + lane;
            bool cand = false;
            if (i < nz) {
                const float prev = Prow[i] * inv, cur = Prow[i + 1] * inv, nxt = Prow[i + 2] * inv;
                const float d0 = cur - prev, d1 = nxt - cur;
                cand = d0 * d1 <= 0.f && cur > a.thresh;
            }
            const unsigned m = __ballot_sync(0xffffffffu, cand);
            if (lane == 0) mrow[w] = m;
        }
    }
}

// Phase 2, one CTA per stream: ascending walk over the candidate bits with the hold-off of one chirp
// length (OFDM.py:364-366) and the end-of-signal wipe-out (OFDM.py:366-370).  The walk reads 8192
// positions per step and jumps over every hold-off window.
__global__ void __launch_bounds__(256) peak_scan_kernel(const PeakArgs a) {
    constexpr int NT = 256;
    __shared__ long long s_first;
    const int tid = threadIdx.x;
    const int64_t stream = blockIdx.x;
    const uint32_t* mrow = a.mask + stream * a.nw;
    const int64_t nz = a.plen - 2;
    int32_t found = 0;
    bool wiped = false;
    int64_t pos = 0;
    while (pos < nz) {
        if (tid == 0) s_first = 0x7fffffffffffffffLL;
        __syncthreads();
        const int64_t w0 = pos >> 5, w = w0 + tid;
        if (w < a.nw) {
            uint32_t m = mrow[w];
            if (w == w0) m &= 0xffffffffu << (pos & 31);          // positions before pos are behind us
            if (m) atomicMin(&s_first, (long long)(w * 32 + __ffs(m) - 1));
        }
        __syncthreads();
        const long long first = s_first;
        __syncthreads();
        if (first == 0x7fffffffffffffffLL) { pos = (w0 + NT) * 32; continue; }
        if (first + a.Lc >= nz) { wiped = true; break; }
        if (tid == 0 && found < a.max_peaks) a.peaks[stream * a.max_peaks + found] = first;
        ++found;
        pos = first + a.Lc + 1;
    }
    if (tid == 0) a.count[stream] = wiped ? 0 : found;
}

// Single-kernel variant, one CTA per stream walking 6144-position tiles: with hundreds of streams in a
// batch the streams themselves supply the parallelism, and one pass over P is cheaper than mark + scan.
// A round costs one DRAM latency, so it is made wide: every thread takes 24 positions from six
// 128-bit loads (tiles start on a 16-byte boundary at or below the walk position; positions behind it
// are masked), evaluates them without branches, and the CTA agrees on the first candidate with one
// barrier per round (three result slots in rotation: the slot of round r+2 is cleared in round r).
__global__ void __launch_bounds__(256) peak_pick_kernel(const PeakArgs a) {
    constexpr int NT = 256, PER = GF3_PEAK_PER, TILE = NT * PER;
    constexpr long long kNone = 0x7fffffffffffffffLL;
    __shared__ long long s_first[3];
    const int tid = threadIdx.x;
    const int64_t stream = blockIdx.x;
    const float* Prow = a.P + stream * a.p_stride;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(Prow) & 15) == 0;
    const float inv = 1.0f / a.pmax[stream];
    const int64_t nz = a.plen - 2;
    int32_t found = 0;
    bool wiped = false;
    int64_t pos = 0;
    if (tid < 3) s_first[tid] = kNone;
    __syncthreads();
    __shared__ int s_blk;
    const float* bm = a.blockmax ? a.blockmax + stream * a.nblk : nullptr;
    for (int round = 0; pos < nz; ++round) {
        if (bm) {
            // A candidate at position i needs P[i+1] / pmax > thresh, so it lies in a block whose maximum passes
            // the same test (x -> x * inv is monotonic for inv > 0; for inv <= 0 or NaN nothing is skipped).
            // Jump to the first such block at or after the walk position: with one packet per stream all but a
            // handful of the ~120 blocks are skipped, and P is read only around the chirp peaks.
            if (inv > 0.f) {
                int blk0 = (int)((pos + 1) / kB);
                int first_hot;
                for (;;) {
                    if (tid == 0) s_blk = 0x7fffffff;
                    __syncthreads();
                    const int blk = blk0 + tid;
                    if (blk < a.nblk && bm[blk] * inv > a.thresh) atomicMin(&s_blk, blk);
                    __syncthreads();
                    first_hot = s_blk;
                    __syncthreads();
                    if (first_hot != 0x7fffffff || blk0 + NT >= a.nblk) break;
                    blk0 += NT;
                }
                if (first_hot == 0x7fffffff) break;                              // no block left that can hold a candidate
                const int64_t p0 = (int64_t)first_hot * kB - 1;                  // position whose P[i+1] is the block's first sample
                if (p0 > pos) pos = p0;
            }
        }
        const int slot = round % 3;
        const int64_t tb = pos & ~(int64_t)3;
        const int64_t base = tb + (int64_t)tid * PER;
        if (base < nz) {
            float v[PER + 2];
            if (vec_ok && base + PER + 2 <= a.plen) {
#pragma unroll
                for (int e = 0; e < PER / 4; ++e) {
                    const float4 q = __ldg(reinterpret_cast<const float4*>(Prow + base) + e);
                    v[4 * e] = q.x; v[4 * e + 1] = q.y; v[4 * e + 2] = q.z; v[4 * e + 3] = q.w;
                }
                const float2 q2 = __ldg(reinterpret_cast<const float2*>(Prow + base + PER));
                v[PER] = q2.x; v[PER + 1] = q2.y;
            } else {
#pragma unroll
                for (int e = 0; e < PER + 2; ++e) v[e] = base + e < a.plen ? Prow[base + e] : 0.f;
            }
            unsigned m = 0;
#pragma unroll
            for (int e = 0; e < PER; ++e) {
                const float prev = v[e] * inv, cur = v[e + 1] * inv, nxt = v[e + 2] * inv;
                const float d0 = cur - prev, d1 = nxt - cur;
                if (d0 * d1 <= 0.f && cur > a.thresh) m |= 1u << e;
            }
            if (a.sparse && m) {
                // a candidate at position i is the sample P[i+1]: it counts only inside a block that can hold one (the
                // other blocks were not computed: their memory is not P).  24 positions touch at most two blocks
                const int64_t blo = (base + 1) / kB;
                const bool hot_lo = blo < a.nblk && bm[blo] * inv > a.thresh;
                const bool hot_hi = blo + 1 < a.nblk && bm[blo + 1] * inv > a.thresh;
                const int split = (int)((blo + 1) * kB - (base + 1));            // first e whose sample lies in the next block
                const unsigned lo_mask = split >= 32 ? 0xffffffffu : ((1u << split) - 1u);
                m &= (hot_lo ? lo_mask : 0u) | (hot_hi ? ~lo_mask : 0u);
            }
            if (base < pos) m &= 0xffffffffu << (int)(pos - base);       // positions behind the walk
            if (base + PER > nz) m &= 0xffffffffu >> (32 - (int)(nz - base)); // positions past the end
            if (m) atomicMin(&s_first[slot], (long long)(base + __ffs(m) - 1));
        }
        __syncthreads();
        const long long first = s_first[slot];
        if (tid == 0) s_first[(round + 2) % 3] = kNone;
        if (first == kNone) { pos = tb + TILE; continue; }
        if (first + a.Lc >= nz) { wiped = true; break; }
        if (tid == 0 && found < a.max_peaks) a.peaks[stream * a.max_peaks + found] = first;
        ++found;
        pos = first + a.Lc + 1;
    }
    if (tid == 0) a.count[stream] = wiped ? 0 : found;
}

// get_symbols' index bookkeeping (OFDM.py:393-397) for a batch of streams on the device: detections ->
// packet start offsets into the flat sample array.  zero_indicies = where(zeros) + 2, the last one (the
// terminating chirp) dropped.  A stream is "ok" when it holds exactly pk_expected packets and the last
// one ends inside the stream; otherwise its offsets are clamped into the stream (the packets decode to
// garbage instead of reading out of bounds) and ok = 0, so the caller can discount it.
__global__ void peaks_to_offsets_kernel(const int64_t* __restrict__ peaks, const int32_t* __restrict__ count, int64_t n_streams,
                                        int32_t max_peaks, int64_t r_stride, int64_t T, int32_t pk_expected, int64_t pkt_samples,
                                        int64_t* __restrict__ pkt_offset, uint8_t* __restrict__ ok) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_streams * pk_expected) return;
    const int64_t s = i / pk_expected;
    const int j = (int)(i - s * pk_expected);
    const int det = count[s] < max_peaks ? count[s] : max_peaks;
    const int64_t last = T - pkt_samples;                                 // last start that keeps the packet inside the stream
    bool good = count[s] == pk_expected + 1 && count[s] <= max_peaks && last >= 0;
    if (good) good = peaks[s * max_peaks + pk_expected - 1] + 2 <= last;
    int64_t st = j < det ? peaks[s * max_peaks + j] + 2 : 0;
    st = st > last ? last : st;
    st = st < 0 ? 0 : st;
    pkt_offset[i] = s * r_stride + st;
    if (j == 0 && ok) ok[s] = good ? 1 : 0;
}


// ------------------------------------------------------------------------------------------
// Schmidl & Cox timing metric (OFDM.py:376-387; unused by receive() since the chirp became the standard, kept as a
// public method).  The reference runs the recursion  P[d+1] = P[d] + r[d+L] r[d+2L] - r[d] r[d+L]  over the first
// `search` samples in a Python loop and returns argmax |P| + N - 1.  The recursion is a prefix sum: one CTA per
// stream scans tiles of 2048 terms in double precision (thread-local scan of 8 terms, warp shuffle scan, carry
// across tiles) and keeps the first index of the largest |P|.
// ------------------------------------------------------------------------------------------
struct ScArgs {
    const void* r;
    int64_t r_stride, search;     // P has `search` entries: P[0] = 0, P[d+1] from d = 0 .. search-2
    int64_t* index;               // [n_streams] argmax |P| (first occurrence), WITHOUT the + N - 1
    double* value;                // [n_streams] |P| there (may be null)
    int L;
};

template <class S>
__global__ void __launch_bounds__(256) schmidlcox_kernel(const ScArgs a) {
    constexpr int NT = 256, PER = 8, TILE = NT * PER;
    __shared__ double wsum[NT / 32];
    __shared__ double s_best[NT / 32];
    __shared__ long long s_idx[NT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const S* r = reinterpret_cast<const S*>(a.r) + (int64_t)blockIdx.x * a.r_stride;
    const int64_t nterms = a.search - 1;
    double carry = 0.0, best = 0.0;                       // P[0] = 0 at index 0
    long long best_i = 0;
    for (int64_t t0 = 0; t0 < nterms; t0 += TILE) {
        const int64_t d0 = t0 + (int64_t)tid * PER;
        double v[PER];
        double run = 0.0;
#pragma unroll
        for (int e = 0; e < PER; ++e) {
            const int64_t d = d0 + e;
            double term = 0.0;
            if (d < nterms) {
                const double x0 = (double)r[d], x1 = (double)r[d + a.L], x2 = (double)r[d + 2 * (int64_t)a.L];
                term = x1 * x2 - x0 * x1;
            }
            run += term;
            v[e] = run;                                   // inclusive scan inside the thread
        }
        double incl = run;                                // warp scan of the thread totals
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        double before = carry + (incl - run);
        double tile_total = 0.0;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) {
            if (w < warp) before += wsum[w];
            tile_total += wsum[w];
        }
        // P[d + 1] = before + v[e]: candidates at indices d0 + e + 1 (ascending inside the thread: ">" keeps the first)
#pragma unroll
        for (int e = 0; e < PER; ++e) {
            const int64_t d = d0 + e;
            if (d < nterms) {
                const double m = fabs(before + v[e]);
                if (m > best) { best = m; best_i = d + 1; }
            }
        }
        carry += tile_total;
        __syncthreads();
    }
    // first index of the maximum over the CTA
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
    }
    if (lane == 0) { s_best[warp] = best; s_idx[warp] = best_i; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < NT / 32; ++w)
            if (s_best[w] > best || (s_best[w] == best && s_idx[w] < best_i)) { best = s_best[w]; best_i = s_idx[w]; }
        a.index[blockIdx.x] = best_i;
        if (a.value) a.value[blockIdx.x] = best;
    }
}

// ------------------------------------------------------------------------------------------ host side
static int run_fwd(const gf3_plan* plan, const void* r, int fmt, int64_t r_stride, int64_t n_streams, int64_t T, int nblk, int reverse,
                   int in_off, int valid_len, float2* spec, const float2* tw, float* pmax, cudaStream_t st, float* energy = nullptr) {
    constexpr int SF = kSyncThreads / SP::T;
    FwdArgs f;
    f.r = r; f.spec = spec; f.tw = tw; f.pmax = pmax; f.r_stride = r_stride; f.T = T; f.n_streams = n_streams;
    f.nblk = nblk; f.reverse = reverse; f.in_off = in_off; f.valid_len = valid_len; f.energy = energy;
    if (energy) GF3_CHECK_CUDA(cudaMemsetAsync(energy, 0, (size_t)n_streams * nblk * kGroups * sizeof(float), st));
    const size_t smem = (size_t)(SF * SP::MP + SP::TW_TOTAL) * sizeof(float2);
    int64_t gx = (n_streams * nblk + SF - 1) / SF;
    const int sms = plan->sm_count;
    if (gx > (int64_t)sms * GF3_XC_FWD_CTAS) gx = (int64_t)sms * GF3_XC_FWD_CTAS;   // 2 CTAs / SM resident, several rounds of them
    if (fmt == GF3_SAMPLE_F32) {
        GF3_CHECK_CUDA(cudaFuncSetAttribute(xcorr_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        xcorr_fwd_kernel<float><<<(unsigned)gx, kSyncThreads, smem, st>>>(f);
    } else if (fmt == GF3_SAMPLE_I16) {
        GF3_CHECK_CUDA(cudaFuncSetAttribute(xcorr_fwd_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        xcorr_fwd_kernel<int16_t><<<(unsigned)gx, kSyncThreads, smem, st>>>(f);
    } else {
        GF3_CHECK_CUDA(cudaFuncSetAttribute(xcorr_fwd_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        xcorr_fwd_kernel<uint8_t><<<(unsigned)gx, kSyncThreads, smem, st>>>(f);
    }
    GF3_LAUNCH_CHECK();
    return GF3_OK;
}

int sync_plan_init(gf3_plan* plan) {
    const gf3_params& p = plan->p;
    int rc = make_chirp(plan);
    if (rc) return rc;
    plan->sync_logN = SP::LOGN;
    plan->sync_parts = (p.chirp_len + kB - 1) / kB;
    if (plan->logN == SP::LOGN) plan->d_sync_tw = plan->d_tw;
    else { rc = upload_twiddles(SP::LOGN, &plan->d_sync_tw); if (rc) return rc; }
    GF3_CHECK_CUDA(cudaMalloc(&plan->d_chirp_spec, (size_t)plan->sync_parts * SP::M * sizeof(float2)));
    // H_p = rfft_{2B}([h[pB .. pB+B), 0 ... 0]),  h[m] = chirp[Lc-1-m]  (fsweep of OFDM.py:357)
    GF3_CHECK_CUDA(cudaMalloc(&plan->d_chirp_energy, (size_t)plan->sync_parts * kGroups * sizeof(float)));
    rc = run_fwd(plan, plan->d_chirp, GF3_SAMPLE_F32, 0, 1, p.chirp_len, plan->sync_parts, 1, kB, kB, plan->d_chirp_spec,
                 plan->d_sync_tw, nullptr, 0, plan->d_chirp_energy);
    if (rc) return rc;
    {
        constexpr int QN = (SP::M / 2);
        GF3_CHECK_CUDA(cudaMalloc(&plan->d_chirp_pairs, (size_t)plan->sync_parts * QN * sizeof(float4)));
        GF3_CHECK_CUDA(cudaMalloc(&plan->d_chirp_dc, (size_t)plan->sync_parts * sizeof(float2)));
        const int n = plan->sync_parts * QN;
        xcorr_pack_h_kernel<<<(n + 255) / 256, 256>>>(plan->d_chirp_spec, plan->sync_parts, 0.5f / (float)SP::N,
                                                      reinterpret_cast<float4*>(plan->d_chirp_pairs), plan->d_chirp_dc);
        GF3_LAUNCH_CHECK();
    }
    GF3_CHECK_CUDA(cudaMalloc(&plan->d_chirp_one, (size_t)SP::M * sizeof(float2)));
    xcorr_ones_kernel<<<(SP::M + 255) / 256, 256>>>(plan->d_chirp_one);
    GF3_LAUNCH_CHECK();
    GF3_CHECK_CUDA(cudaDeviceSynchronize());
    return GF3_OK;
}

void sync_plan_free(gf3_plan* plan) {
    if (plan->d_chirp_one) cudaFree(plan->d_chirp_one);
    if (plan->d_chirp_energy) cudaFree(plan->d_chirp_energy);
    plan->d_chirp_one = nullptr; plan->d_chirp_energy = nullptr;
    if (plan->d_chirp_pairs) cudaFree(plan->d_chirp_pairs);
    if (plan->d_chirp_dc) cudaFree(plan->d_chirp_dc);
    plan->d_chirp_pairs = nullptr; plan->d_chirp_dc = nullptr;
    if (plan->d_sync_tw && plan->d_sync_tw != plan->d_tw) cudaFree(plan->d_sync_tw);
    if (plan->d_chirp_spec) cudaFree(plan->d_chirp_spec);
    if (plan->d_chirp) cudaFree(plan->d_chirp);
    plan->d_sync_tw = nullptr; plan->d_chirp_spec = nullptr; plan->d_chirp = nullptr;
}

struct XcorrGeom { int nblk_out, nblk_in; int64_t out_len; size_t per_stream; int64_t tile; };
static XcorrGeom xcorr_geom(const gf3_plan* plan, int64_t n_streams, int64_t T) {
    XcorrGeom g;
    g.out_len = T + plan->p.chirp_len - 1;
    g.nblk_out = (int)((g.out_len + kB - 1) / kB);
    int64_t nin = (T - 1) / kB + 2;                      // blocks that see at least one real sample
    g.nblk_in = (int)(nin < g.nblk_out ? nin : g.nblk_out);
    g.per_stream = (size_t)g.nblk_in * SP::M * sizeof(float2);
    size_t cap = (size_t)2 << 30;                         // bound the scratch to 2 GiB: streams are tiled
    if (const char* e = getenv("GF3_XC_TILE_MB")) { const long mb = atol(e); if (mb > 0) cap = (size_t)mb << 20; }
    int64_t tile = (int64_t)(cap / g.per_stream);
    if (tile < 1) tile = 1;
    if (tile > n_streams) tile = n_streams;
    g.tile = tile;
    return g;
}

#ifndef GF3_XC_FUSED_MAX_PARTS
#define GF3_XC_FUSED_MAX_PARTS 4      // ring of (parts - 1) x 16 KB spectra per CTA next to 33 KB of FFT buffer + twiddles
#endif
#ifndef GF3_XC_FUSED_MINB
#define GF3_XC_FUSED_MINB 2
#endif
static size_t fused_smem(int parts) {
    return (size_t)(SP::MP + SP::TW_TOTAL) * sizeof(float2) + (size_t)(parts - 1) * ((SP::M / 2) * sizeof(float4) + sizeof(float2)) + 16;
}
// CTAs per SM of the fused kernel: 2 (<= 255 registers; the shared-memory carve-out then leaves ~120 KB of L1,
// enough to keep the 48 KB of chirp partitions resident: 1.81 ms per 1024 C3 streams) or 3 (<= 168 registers, spills,
// 28 KB of L1: 2.3 ms).  GF3_XC_MINB=3 selects the latter (experiments).
static int fused_minb() {
    if (const char* e = getenv("GF3_XC_MINB")) return atoi(e) == 3 ? 3 : 2;
    return GF3_XC_FUSED_MINB;
}
// The fused kernel pays (parts - 1) extra forward FFTs per CTA (the spectra before its first block), so it needs
// runs of blocks that are long against the partition count; otherwise (one long recording with a 21 600-tap
// chirp: 11 partitions) the two-kernel form stays.  GF3_XCORR_PATH=fused|split overrides (experiments).
// Long chirps: partition sum in its own kernel (xcorr_mac_kernel), inverse transform by xcorr_acc_kernel with the unit
// partition.  GF3_XCORR_MAC=0 keeps the two-kernel form (experiments / the bit-identity test).
static bool mac_applies(const gf3_plan* plan) {
    const bool can = plan->sync_parts >= kMacMinParts && plan->sync_parts <= kMacMaxParts;
    if (const char* e = getenv("GF3_XCORR_MAC")) return atoi(e) != 0 && can;
    return can;
}
static size_t mac_y_bytes(const XcorrGeom& g) { return (((size_t)g.tile * (size_t)g.nblk_out * SP::M * sizeof(float2)) + 255) & ~(size_t)255; }
// detection only: group energies [tile * nblk_in, kGroups] | bounds [tile * nblk_out] | listed blocks [tile * nblk_out] |
// their count (256 B) | run flags [tile * ceil(nblk_out / kMacJ)]
static size_t detect_scratch_bytes(const XcorrGeom& g) {
    const size_t nb = (size_t)g.tile * g.nblk_out, runs = (size_t)g.tile * ((g.nblk_out + kMacJ - 1) / kMacJ);
    return (size_t)g.tile * g.nblk_in * kGroups * sizeof(float) + nb * sizeof(float) + nb * sizeof(int) + 256 + ((runs + 255) & ~(size_t)255);
}

static bool fused_applies(const gf3_plan* plan, int64_t n_streams, const XcorrGeom& g) {
    if (plan->sync_parts > GF3_XC_FUSED_MAX_PARTS) return false;
    if (const char* e = getenv("GF3_XCORR_PATH")) {
        if (!strcmp(e, "split")) return false;
        if (!strcmp(e, "fused")) return true;
    }
    const int64_t total = n_streams * g.nblk_out;
    const int64_t ctas = (int64_t)plan->sm_count * fused_minb();
    return total / ctas >= 8 * (int64_t)(plan->sync_parts - 1) || total < ctas;
}

template <class S, int PARTS, int MINB>
static int launch_fused_t(const gf3_plan* plan, const FusedArgs& a, cudaStream_t st) {
    auto kern = xcorr_fused_kernel<S, PARTS, MINB>;
    const size_t smem = fused_smem(PARTS);
    int per_sm = 0;
    GF3_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GF3_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > MINB) per_sm = MINB;
    const int64_t total = a.n_streams * a.nblk_out;
    int64_t grid = (int64_t)plan->sm_count * per_sm;
    if (grid > total) grid = total;
    fill_f32_kernel<<<(unsigned)((a.n_streams + 255) / 256), 256, 0, st>>>(a.pmax, a.n_streams, -INFINITY);
    GF3_LAUNCH_CHECK();
    if (getenv("GF3_DEBUG"))
        fprintf(stderr, "[gf3] xcorr fused: parts=%d smem=%zu B grid=%lld (%d CTAs/SM) blocks=%lld\n", PARTS, smem,
                (long long)grid, per_sm, (long long)total);
    kern<<<(unsigned)grid, 128, smem, st>>>(a);
    GF3_LAUNCH_CHECK();
    return GF3_OK;
}
template <class S, int MINB>
static int launch_fused_p(const gf3_plan* plan, const FusedArgs& a, cudaStream_t st) {
    switch (a.parts) {
        case 1: return launch_fused_t<S, 1, MINB>(plan, a, st);
        case 2: return launch_fused_t<S, 2, MINB>(plan, a, st);
        case 3: return launch_fused_t<S, 3, MINB>(plan, a, st);
        case 4: return launch_fused_t<S, 4, MINB>(plan, a, st);
        default: gf3::set_error("xcorr fused: %d partitions not supported", a.parts); return GF3_ERR_INVALID;
    }
}
template <class S>
static int launch_fused(const gf3_plan* plan, const FusedArgs& a, cudaStream_t st) {
    if constexpr (sizeof(S) == 4) {
        if (fused_minb() == 3) return launch_fused_p<S, 3>(plan, a, st);      // experiment: 3 CTAs / SM, <= 168 registers
    }
    return launch_fused_p<S, 2>(plan, a, st);
}

// chirp_method's convolution for a batch of streams (OFDM.py:357-358); blockmax is optional
static int xcorr_common(const gf3_plan* plan, const void* r, int fmt, int64_t r_stride, int64_t n_streams, int64_t T,
                        float* P, int64_t p_stride, float* pmax, float* blockmax, void* work, cudaStream_t st, bool sparse = false,
                        int* counter = nullptr) {
    const XcorrGeom g = xcorr_geom(plan, n_streams, T);
    GF3_REQUIRE(p_stride >= g.out_len, "xcorr: p_stride %lld < T + chirp_len - 1 = %lld", (long long)p_stride, (long long)g.out_len);
    GF3_REQUIRE(fmt == GF3_SAMPLE_F32 || fmt == GF3_SAMPLE_I16 || fmt == GF3_SAMPLE_U8, "xcorr: unknown sample format %d", fmt);
    if (fused_applies(plan, n_streams, g)) {
        FusedArgs a;
        a.r = r; a.Hs = reinterpret_cast<const float4*>(plan->d_chirp_pairs); a.Hdc = plan->d_chirp_dc; a.tw = plan->d_sync_tw;
        a.P = P; a.pmax = pmax; a.blockmax = blockmax; a.r_stride = r_stride; a.p_stride = p_stride; a.T = T; a.out_len = g.out_len;
        a.n_streams = n_streams; a.nblk_in = g.nblk_in; a.nblk_out = g.nblk_out; a.parts = plan->sync_parts;
        a.sparse = (sparse && blockmax && counter) ? 1 : 0; a.thresh = plan->p.thresh; a.counter = counter;
        if (a.sparse) {
            fill_f32_kernel<<<1, 32, 0, st>>>(reinterpret_cast<float*>(counter), 1, 0.0f);      // int 0
            GF3_LAUNCH_CHECK();
        }
        if (fmt == GF3_SAMPLE_F32) return launch_fused<float>(plan, a, st);
        if (fmt == GF3_SAMPLE_I16) return launch_fused<int16_t>(plan, a, st);
        return launch_fused<uint8_t>(plan, a, st);
    }
    GF3_REQUIRE(work != nullptr, "xcorr: null work buffer");
    float2* spec = reinterpret_cast<float2*>(work);
    const size_t smem = (size_t)(SP::MP + SP::TW_TOTAL) * sizeof(float2);
    GF3_CHECK_CUDA(cudaFuncSetAttribute(xcorr_acc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t esz = fmt == GF3_SAMPLE_F32 ? 4 : fmt == GF3_SAMPLE_I16 ? 2 : 1;
    const bool mac = mac_applies(plan);
    float2* Ysum = reinterpret_cast<float2*>(reinterpret_cast<char*>(work) + ((g.per_stream * (size_t)g.tile + 255) & ~(size_t)255));
    const size_t mac_smem = (size_t)plan->sync_parts * SP::M * sizeof(float2);
    void (*mac_kern)(const MacArgs) = nullptr;
    const bool bound_only = sparse && blockmax;             // detection only: bounds of |P| per block, two selective inverse passes
    if (mac) {
        switch (plan->sync_parts) {
#define GF3_MAC_CASE(N_) case N_: mac_kern = xcorr_mac_kernel<N_>; break;
            GF3_MAC_CASE(5) GF3_MAC_CASE(6) GF3_MAC_CASE(7) GF3_MAC_CASE(8) GF3_MAC_CASE(9) GF3_MAC_CASE(10) GF3_MAC_CASE(11)
            GF3_MAC_CASE(12) GF3_MAC_CASE(13)
#undef GF3_MAC_CASE
            default: gf3::set_error("xcorr: internal: no partition-sum kernel for %d partitions", plan->sync_parts); return GF3_ERR_INVALID;
        }
        GF3_CHECK_CUDA(cudaFuncSetAttribute(mac_kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mac_smem));
    }
    for (int64_t s0 = 0; s0 < n_streams; s0 += g.tile) {
        const int64_t ns = (n_streams - s0 < g.tile) ? n_streams - s0 : g.tile;
        // detection only: group energies, bounds, listed blocks, their count, run flags (detect_scratch_bytes)
        float* energy = reinterpret_cast<float*>(reinterpret_cast<char*>(Ysum) + (mac ? mac_y_bytes(g) : 0));
        float* bound = energy + (size_t)ns * g.nblk_in * kGroups;
        int* sel_idx = reinterpret_cast<int*>(bound + (size_t)ns * g.nblk_out);
        int* sel_count = sel_idx + (size_t)ns * g.nblk_out;
        uint8_t* run_sel = reinterpret_cast<uint8_t*>(sel_count) + 256;
        int rc = run_fwd(plan, reinterpret_cast<const char*>(r) + (size_t)(s0 * r_stride) * esz, fmt, r_stride, ns, T, g.nblk_in, 0, 0, 2 * kB,
                         spec, plan->d_sync_tw, pmax + s0, st, bound_only ? energy : nullptr);
        if (rc) return rc;
        AccArgs a;
        a.spec = spec; a.H = plan->d_chirp_spec; a.tw = plan->d_sync_tw; a.P = P + s0 * p_stride; a.pmax = pmax + s0;
        a.blockmax = blockmax ? blockmax + s0 * g.nblk_out : nullptr;
        a.p_stride = p_stride; a.out_len = g.out_len; a.nblk_in = g.nblk_in; a.nblk_out = g.nblk_out; a.parts = plan->sync_parts;
        a.n_work = ns * g.nblk_out;
        a.bound = nullptr; a.thresh = plan->p.thresh; a.select = 0; a.sel_idx = nullptr; a.sel_count = nullptr;
        if (bound_only) {
            // detection only (gf3_sync_detect): 1. bounds of |P| per block from the group energies the forward kernel
            // recorded; 2. per stream, the block with the largest bound is transformed back (partition sum + inverse in
            // xcorr_acc_kernel) and seeds pmax; 3. the same for every block that can hold a candidate, and its
            // neighbours.  The detections are those of the full computation (tested).
            const int64_t nb = ns * g.nblk_out;
            int64_t bgrid = (nb + 7) / 8;
            if (bgrid > (int64_t)plan->sm_count * 8) bgrid = (int64_t)plan->sm_count * 8;
            xcorr_bound_kernel<<<(unsigned)bgrid, 256, 0, st>>>(energy, plan->d_chirp_energy, bound, g.nblk_in, g.nblk_out, plan->sync_parts, nb);
            GF3_LAUNCH_CHECK();
            a.bound = bound;
            a.select = 1; a.n_work = ns;
            xcorr_acc_kernel<<<(unsigned)(ns < (int64_t)plan->sm_count * GF3_XC_ACC_CTAS ? ns : (int64_t)plan->sm_count * GF3_XC_ACC_CTAS), 128, smem, st>>>(a);
            GF3_LAUNCH_CHECK();
            a.select = 2; a.n_work = ns * g.nblk_out;
            if (mac) {
                // long chirps: the listed blocks go through the partition-sum kernel (runs that hold one) and the inverse
                // stage with the unit partition, as in the full computation -- the same P, a third of the time per block
                const int runs = (g.nblk_out + kMacJ - 1) / kMacJ;
                GF3_CHECK_CUDA(cudaMemsetAsync(sel_count, 0, 256 + (size_t)ns * runs, st));
                xcorr_select_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(bound, a.pmax, a.thresh, g.nblk_out, runs, nb, a.blockmax,
                                                                                 sel_idx, sel_count, run_sel);
                GF3_LAUNCH_CHECK();
                MacArgs m;
                m.spec = spec; m.H = plan->d_chirp_spec; m.Y = Ysum; m.nblk_in = g.nblk_in; m.nblk_out = g.nblk_out;
                m.runs_per_stream = runs; m.n_units = ns * runs; m.run_sel = run_sel;
                const int64_t mg = m.n_units < plan->sm_count ? m.n_units : plan->sm_count;
                mac_kern<<<(unsigned)mg, kMacThreads, mac_smem, st>>>(m);
                GF3_LAUNCH_CHECK();
                a.spec = Ysum; a.H = plan->d_chirp_one; a.nblk_in = g.nblk_out; a.parts = 1;
                a.select = 3; a.sel_idx = sel_idx; a.sel_count = sel_count;
            }
        } else if (mac) {
            MacArgs m;
            m.spec = spec; m.H = plan->d_chirp_spec; m.Y = Ysum; m.nblk_in = g.nblk_in; m.nblk_out = g.nblk_out;
            m.runs_per_stream = (g.nblk_out + kMacJ - 1) / kMacJ;
            m.n_units = ns * m.runs_per_stream; m.run_sel = nullptr;
            int64_t mg = m.n_units < plan->sm_count ? m.n_units : plan->sm_count;          // one CTA per SM: the partitions fill its shared memory
            mac_kern<<<(unsigned)mg, kMacThreads, mac_smem, st>>>(m);
            GF3_LAUNCH_CHECK();
            a.spec = Ysum; a.H = plan->d_chirp_one; a.nblk_in = g.nblk_out; a.parts = 1;    // inverse stage: Y_b times the unit partition
        }
        int64_t grid = a.n_work;
        if (grid > (int64_t)plan->sm_count * GF3_XC_ACC_CTAS) grid = (int64_t)plan->sm_count * GF3_XC_ACC_CTAS;   // 4 CTAs / SM resident
        xcorr_acc_kernel<<<(unsigned)grid, 128, smem, st>>>(a);
        GF3_LAUNCH_CHECK();
    }
    return GF3_OK;
}

// blockmax + one stream per CTA: is the single-pass picker used?  (the sparse matched filter needs it: the mark + scan
// form reads all of P)
static bool single_pass_picker(const gf3_plan* plan, int64_t n_streams, bool have_blockmax) {
    return n_streams >= 2 * (int64_t)plan->sm_count || (have_blockmax && n_streams >= plan->sm_count / 2);
}

static int peak_pick_common(const gf3_plan* plan, const float* P, int64_t p_stride, int64_t n_streams, int64_t T, const float* pmax,
                            const float* blockmax, int nblk, int64_t* peaks, int32_t max_peaks, int32_t* count, void* work,
                            cudaStream_t st, bool sparse = false) {
    PeakArgs a;
    a.P = P; a.pmax = pmax; a.peaks = peaks; a.count = count; a.p_stride = p_stride;
    a.plen = T + plan->p.chirp_len - 1; a.max_peaks = max_peaks; a.Lc = plan->p.chirp_len; a.thresh = plan->p.thresh;
    GF3_REQUIRE(a.plen >= 3, "peak_pick: signal too short");
    a.mask = reinterpret_cast<uint32_t*>(work);
    a.n_streams = n_streams;
    a.nw = (a.plen - 2 + 31) / 32;
    a.blockmax = blockmax; a.nblk = nblk; a.sparse = sparse ? 1 : 0;
    GF3_REQUIRE(!sparse || (blockmax && single_pass_picker(plan, n_streams, true)), "peak_pick: internal: sparse P needs the block maxima");
    int64_t blocks = a.n_streams * ((a.nw + 63) / 64);               // chunks of 64 mask words
    const int64_t cap = (int64_t)plan->sm_count * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (single_pass_picker(plan, n_streams, blockmax != nullptr)) {
        // enough streams to fill the GPU: one pass, one CTA each (with block maxima even a few dozen streams are
        // cheaper this way: the walk touches only the blocks around the chirp peaks)
        peak_pick_kernel<<<(unsigned)n_streams, 256, 0, st>>>(a);
        GF3_LAUNCH_CHECK();
        return GF3_OK;
    }
    GF3_REQUIRE(work != nullptr, "peak_pick: null work buffer (gf3_peak_pick_work_bytes)");
    peak_mark_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
    GF3_LAUNCH_CHECK();
    peak_scan_kernel<<<(unsigned)n_streams, 256, 0, st>>>(a);
    GF3_LAUNCH_CHECK();
    return GF3_OK;
}

}  // namespace gf3

using namespace gf3;

extern "C" size_t gf3_xcorr_work_bytes(const gf3_plan* plan, int64_t n_streams, int64_t T) {
    if (!plan || n_streams <= 0 || T <= 0) return 0;
    const XcorrGeom g = xcorr_geom(plan, n_streams, T);
    if (fused_applies(plan, n_streams, g)) return 16;                 // the fused kernel keeps its spectra on chip
    const size_t xs = (g.per_stream * (size_t)g.tile + 255) & ~(size_t)255;
    return xs + (mac_applies(plan) ? mac_y_bytes(g) : 0) + detect_scratch_bytes(g);
}

extern "C" int gf3_xcorr(const gf3_plan* plan, const float* r, int64_t r_stride, int64_t n_streams,
                         int64_t T, float* P, int64_t p_stride, float* pmax, void* work, void* stream) {
    GF3_REQUIRE(plan && r && P && pmax && work, "xcorr: null argument");
    GF3_REQUIRE(n_streams >= 0 && T >= 1, "xcorr: bad sizes");
    if (n_streams == 0) return GF3_OK;
    return xcorr_common(plan, r, GF3_SAMPLE_F32, r_stride, n_streams, T, P, p_stride, pmax, nullptr, work,
                        reinterpret_cast<cudaStream_t>(stream));
}

// layout of gf3_sync_streams' work buffer: [block maxima | matched-filter scratch | candidate bit mask]
static size_t sync_off_blockmax(const XcorrGeom& g, int64_t n_streams) { return ((((size_t)n_streams * g.nblk_out * sizeof(float)) + 255) & ~(size_t)255) + 256; }   // + the stream counter

extern "C" size_t gf3_sync_work_bytes(const gf3_plan* plan, int64_t n_streams, int64_t T) {
    if (!plan || n_streams <= 0 || T <= 0) return 0;
    const XcorrGeom g = xcorr_geom(plan, n_streams, T);
    const size_t xw = (gf3_xcorr_work_bytes(plan, n_streams, T) + 255) & ~(size_t)255;
    return sync_off_blockmax(g, n_streams) + xw + gf3_peak_pick_work_bytes(plan, n_streams, T) + 256;
}

static int sync_common(const gf3_plan* plan, const void* r, int32_t sample_format, int64_t r_stride, int64_t n_streams,
                       int64_t T, float* P, int64_t p_stride, float* pmax, int64_t* peaks, int32_t max_peaks,
                       int32_t* count, void* work, void* stream, bool detect_only) {
    GF3_REQUIRE(plan && r && P && pmax && peaks && count && work, "sync_streams: null argument");
    GF3_REQUIRE(max_peaks >= 1 && n_streams >= 0 && n_streams <= 0x7fffffff && T >= 1, "sync_streams: bad sizes");
    if (n_streams == 0) return GF3_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const XcorrGeom g = xcorr_geom(plan, n_streams, T);
    char* w = reinterpret_cast<char*>(work);
    const bool fused = fused_applies(plan, n_streams, g);
    float* blockmax = reinterpret_cast<float*>(w);                     // both forms of the matched filter record the block maxima
    void* xwork = w + sync_off_blockmax(g, n_streams);
    void* pwork = reinterpret_cast<char*>(xwork) + ((gf3_xcorr_work_bytes(plan, n_streams, T) + 255) & ~(size_t)255);
    // detection only: blocks of P that provably hold no candidate are not computed (the detections are the same)
    const bool sparse = detect_only && single_pass_picker(plan, n_streams, true) && !getenv("GF3_SYNC_DENSE");
    int* counter = reinterpret_cast<int*>(w + sync_off_blockmax(g, n_streams) - 256);
    int rc = xcorr_common(plan, r, sample_format, r_stride, n_streams, T, P, p_stride, pmax, blockmax, xwork, st, sparse, counter);
    if (rc) return rc;
    return peak_pick_common(plan, P, p_stride, n_streams, T, pmax, blockmax, g.nblk_out, peaks, max_peaks, count, pwork, st, sparse);
}

extern "C" int gf3_sync_streams(const gf3_plan* plan, const void* r, int32_t sample_format, int64_t r_stride, int64_t n_streams,
                                int64_t T, float* P, int64_t p_stride, float* pmax, int64_t* peaks, int32_t max_peaks,
                                int32_t* count, void* work, void* stream) {
    return sync_common(plan, r, sample_format, r_stride, n_streams, T, P, p_stride, pmax, peaks, max_peaks, count, work, stream, false);
}

extern "C" int gf3_sync_detect(const gf3_plan* plan, const void* r, int32_t sample_format, int64_t r_stride, int64_t n_streams,
                               int64_t T, float* P_scratch, int64_t p_stride, float* pmax, int64_t* peaks, int32_t max_peaks,
                               int32_t* count, void* work, void* stream) {
    return sync_common(plan, r, sample_format, r_stride, n_streams, T, P_scratch, p_stride, pmax, peaks, max_peaks, count, work, stream, true);
}

extern "C" size_t gf3_peak_pick_work_bytes(const gf3_plan* plan, int64_t n_streams, int64_t T) {
    if (!plan || n_streams <= 0 || T <= 0) return 0;
    const int64_t nz = T + plan->p.chirp_len - 3;
    return (size_t)n_streams * (size_t)((nz + 31) / 32) * sizeof(uint32_t);
}

extern "C" int gf3_peak_pick(const gf3_plan* plan, const float* P, int64_t p_stride, int64_t n_streams,
                             int64_t T, const float* pmax, int64_t* peaks, int32_t max_peaks, int32_t* count,
                             void* work, void* stream) {
    GF3_REQUIRE(plan && P && pmax && peaks && count, "peak_pick: null argument");
    GF3_REQUIRE(max_peaks >= 1 && n_streams >= 0 && n_streams <= 0x7fffffff, "peak_pick: bad sizes");
    if (n_streams == 0) return GF3_OK;
    GF3_REQUIRE(work != nullptr, "peak_pick: null work buffer (gf3_peak_pick_work_bytes)");
    return peak_pick_common(plan, P, p_stride, n_streams, T, pmax, nullptr, 0, peaks, max_peaks, count, work,
                            reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int gf3_peaks_to_offsets(const gf3_plan* plan, const int64_t* peaks, const int32_t* count, int64_t n_streams,
                                    int32_t max_peaks, int64_t r_stride, int64_t T, int32_t pk_expected,
                                    int64_t* pkt_offset, uint8_t* ok, void* stream) {
    GF3_REQUIRE(plan && peaks && count && pkt_offset, "peaks_to_offsets: null argument");
    GF3_REQUIRE(n_streams >= 0 && max_peaks >= 1 && pk_expected >= 1 && pk_expected < max_peaks,
                "peaks_to_offsets: need 1 <= pk_expected < max_peaks");
    if (n_streams == 0) return GF3_OK;
    const gf3_params& p = plan->p;
    const int64_t pkt_samples = (int64_t)(2 * p.n_pilots + p.packet_len) * (p.N + p.cp);
    const int64_t n = n_streams * pk_expected;
    peaks_to_offsets_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        peaks, count, n_streams, max_peaks, r_stride, T, pk_expected, pkt_samples, pkt_offset, ok);
    GF3_LAUNCH_CHECK();
    return GF3_OK;
}

extern "C" int gf3_schmidlcox(const gf3_plan* plan, const void* r, int32_t sample_format, int64_t r_stride, int64_t n_streams,
                              int64_t T, int64_t search, int64_t* index, double* value, void* stream) {
    GF3_REQUIRE(plan && r && index, "schmidlcox: null argument");
    GF3_REQUIRE(n_streams >= 0 && n_streams <= 0x7fffffff && search >= 1, "schmidlcox: bad sizes");
    const int L = plan->p.N / 2;                               // CamG.L = K + 1 (OFDM.py:54)
    // the recursion reads r[d + 2L] for d <= search - 2: the reference raises IndexError on a shorter recording
    GF3_REQUIRE(T >= search - 1 + 2 * (int64_t)L, "schmidlcox: %lld samples, %lld needed (search - 1 + 2L)", (long long)T,
                (long long)(search - 1 + 2 * (int64_t)L));
    if (n_streams == 0) return GF3_OK;
    ScArgs a;
    a.r = r; a.r_stride = r_stride; a.search = search; a.index = index; a.value = value; a.L = L;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (sample_format == GF3_SAMPLE_F32) schmidlcox_kernel<float><<<(unsigned)n_streams, 256, 0, st>>>(a);
    else if (sample_format == GF3_SAMPLE_I16) schmidlcox_kernel<int16_t><<<(unsigned)n_streams, 256, 0, st>>>(a);
    else if (sample_format == GF3_SAMPLE_U8) schmidlcox_kernel<uint8_t><<<(unsigned)n_streams, 256, 0, st>>>(a);
    else { gf3::set_error("schmidlcox: unknown sample format %d", sample_format); return GF3_ERR_INVALID; }
    GF3_LAUNCH_CHECK();
    return GF3_OK;
}
