// detect.cu -- B200 formulation of flypylib's voxel2obj (flypylib/fplobjdetect.py:132-257).
//
// Data layout in HBM
//   pred / smooth  : (Z,Y,X) float32, C order, interior only.  The reference pads the map with
//                    r = obj_min_dist zeros on every face, filters, then zeroes the r-wide border
//                    (fplobjdetect.py:158-175).  Because the Gaussian is separable and the pad is
//                    zero, interior outputs never consume a non-zero pad-zone intermediate, and the
//                    final border is exactly 0 -- so the padded map is never materialised; its
//                    (PZ*PY*PX - Z*Y*X) border zeros only enter the percentile as a count.
//   suppressed     : 1 bit / interior voxel ("is_valid" of the reference, inverted).
//   candidates     : SoA (uint64 flat interior index, float32 value), ~3 % of the padded volume.
//
// Kernels (all HBM-bound integer / fp64 streaming work, no tensor cores):
//   gauss_strided / gauss_contig  exact restatement of SciPy's symmetric correlate1d: double
//                                 accumulate, separate mul and add (__dmul_rn/__dadd_rn, never
//                                 contracted), float32 store after each axis.
//   select_hist / select_scan     3-pass (11/11/10 bit) radix select of an order statistic on the
//                                 monotone uint32 key of the float32 values.
//   compact_candidates            smooth > threshold (strict) and > 0, block-aggregated append.
//   nms_filter / nms_ballcheck /  "rounds" formulation of the greedy loop: a valid candidate with no
//   nms_suppress                  better valid voxel inside its ball is selected this round; balls of
//                                 selected points are suppressed; repeat.  Greedy selection under the
//                                 strict total order (value desc, flat index asc) is the unique
//                                 lexicographically-first maximal independent set, so this yields the
//                                 reference's detections; sorting by that order gives its emission order.
//   bitonic_*/finish_rows         order, un-pad, buffer crop, offset (fplobjdetect.py:233-253).
#include "common.cuh"
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

namespace fpl {
namespace v2o {

constexpr int kMaxLw = 32;                 // sigma <= 16 (truncate 2.0)
struct Taps { double w[2 * kMaxLw + 1]; }; // correlate-order taps, centre at w[lw]

// monotone uint32 key of a float32 (radix select, cut-offs): a < b  <=>  key(a) < key(b)
__device__ __forceinline__ unsigned f2key(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) {
    unsigned u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

struct SelectState {                 // lives in device memory
    unsigned long long rank;         // rank still to locate inside the current prefix class
    unsigned prefix;                 // key bits fixed so far (high bits)
    unsigned prefix_mask;            // which bits are fixed
    unsigned long long extra_zeros;  // implicit +0.0 values of the never-materialised border
    unsigned long long nan_count;
    unsigned long long n_above;      // interior values whose key lies above the current prefix class
    unsigned long long n_bin;        // interior values inside the current prefix class
    unsigned long long hist[16][2048];   // kHistSlots partial histograms: blocks spread their flushes over the slots
};
constexpr int kHistSlots = 16;

// run-length compressed feed of a block-level shared-memory histogram
struct HistFeed {
    unsigned last_bin = 0xffffffffu, run = 0;
    __device__ __forceinline__ void add(unsigned *h, unsigned b) {
        if (b == last_bin) { ++run; }
        else { if (run) atomicAdd(&h[last_bin], run); last_bin = b; run = 1; }
    }
    __device__ __forceinline__ void flush(unsigned *h) { if (run) atomicAdd(&h[last_bin], run); run = 0; last_bin = 0xffffffffu; }
};

// ------------------------------------------------------------------------------------------------
// index map of one filtered line: interior coordinate p (may lie outside [0,n)) -> interior source
// index, or -1 when the padded-and-reflected line holds a pad zero there.
// padded line: r zeros | n values | r zeros, extended by SciPy 'reflect' (d c b a | a b c d | d c b a)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long line_src(long long p, long long n, int r) {
    long long np_ = n + 2LL * r;
    long long pp = p + r;
    if (pp < 0 || pp >= np_) {
        long long period = 2 * np_;
        long long m = pp % period;
        if (m < 0) m += period;
        pp = (m >= np_) ? (period - 1 - m) : m;
    }
    long long q = pp - r;
    return (q >= 0 && q < n) ? q : -1;
}

// ------------------------------------------------------------------------------------------------
// generic pass: one thread per output, runtime lw.  Correctness anchor + fallback for unusual sigma.
// volume viewed as (outer, n, inner), filtering along n.
// ------------------------------------------------------------------------------------------------
__global__ void gauss_generic_kernel(const float *__restrict__ in, float *__restrict__ out,
                                     long long outer, long long n, long long inner, int r, int lw,
                                     Taps taps) {
    long long total = outer * n * inner;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        long long i = idx % inner;
        long long c = (idx / inner) % n;
        long long o = idx / (inner * n);
        const float *line = in + o * n * inner + i;
        double tmp = __dmul_rn((double)line[c * inner], taps.w[lw]);
        for (int j = -lw; j < 0; ++j) {
            long long qa = line_src(c + j, n, r), qb = line_src(c - j, n, r);
            double a = qa >= 0 ? (double)line[qa * inner] : 0.0;
            double b = qb >= 0 ? (double)line[qb * inner] : 0.0;
            tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(a, b), taps.w[lw + j]));
        }
        out[idx] = (float)tmp;
    }
}

// ------------------------------------------------------------------------------------------------
// fast passes, compile-time half width LW.  A block owns 64 independent lines x TN outputs along
// the filtered axis; the (TN+2LW) x 64 source tile is staged once in shared memory (float), each
// thread then produces runs of RUN consecutive outputs from a register window of RUN+2LW doubles
// (every source value is converted to double once per run).
// ------------------------------------------------------------------------------------------------
constexpr int kLines = 64;
constexpr int kGroups = 2;          // 128-thread blocks: more, finer-grained blocks per SM overlap load and FP64 phases
constexpr int kRun = 16;
constexpr int kRunsPerThread = 2;
constexpr int kTN = kGroups * kRunsPerThread * kRun;   // 64 outputs along the axis per block

// Exact SciPy chain of one output: tmp = x0*w0; tmp += (x[-j] + x[+j]) * w[j], every operation rounded.
template <int LW>
__device__ __forceinline__ float exact_output(const double (&x)[kRun + 2 * LW], const Taps &taps, int c) {
    double tmp = __dmul_rn(x[c], taps.w[LW]);
#pragma unroll
    for (int j = -LW; j < 0; ++j)
        tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(x[c + j], x[c - j]), taps.w[LW + j]));
    return (float)tmp;
}

// kRun outputs from a register window.
//
// Certified fast chain (cert > 0, all window inputs and all taps non-negative): the FP64 pipe is the
// bound of this kernel (31 non-fused DP operations per output), so the chain is first evaluated with
// the product and the accumulate FUSED (t' = fma(x[-j]+x[+j], w[j], t'): 21 DP operations).  With
// non-negative terms every partial sum is bounded by the result, and the two chains differ by at most
// (1 + 2*LW) * 2^-53 relative, i.e. < 22 ulp(t') for LW <= 10.  float32(t') therefore equals the
// reference's float32(t) unless t' lies within that distance of a float32 rounding boundary (the low
// 29 mantissa bits of t' within `cert` >= 64 of 2^28), or outside the float32 normal range.  Those
// outputs (about 2.4e-7 of them at cert = 64) are recomputed with the exact chain; all others are proven
// to round identically.  Exactness never rests on the fast chain.
template <int LW>
__device__ __forceinline__ void run_from_window(const double (&x)[kRun + 2 * LW], const Taps &taps,
                                                float (&res)[kRun], unsigned cert,
                                                const float *win, int wstride) {
    bool nonneg = false;
    if (cert) {                      // the certificate needs non-negative inputs: check the sign bits of the window
        unsigned sgn = 0;
#pragma unroll
        for (int k = 0; k < kRun + 2 * LW; ++k) sgn |= __float_as_uint(win[k * wstride]);
        nonneg = (sgn >> 31) == 0u;
    }
    if (cert && nonneg) {
        unsigned bad = 0;
#pragma unroll
        for (int q = 0; q < kRun; ++q) {
            const int c = q + LW;
            double t = __dmul_rn(x[c], taps.w[LW]);
#pragma unroll
            for (int j = -LW; j < 0; ++j) t = __fma_rn(__dadd_rn(x[c + j], x[c - j]), taps.w[LW + j], t);
            const unsigned long long bits = (unsigned long long)__double_as_longlong(t);
            const unsigned lo = (unsigned)bits, hi = (unsigned)(bits >> 32);
            const unsigned dist = (lo & 0x1fffffffu) - (0x10000000u - cert);      // <= 2*cert near a boundary
            const unsigned e = (hi >> 20) & 0x7ffu;                                // sign bit is clear
            // safe exponent range: 2^-125 <= t < 2^127 (float32 normal, no overflow); exact zero is safe
            const bool unsafe = (dist <= 2u * cert) || ((e - 898u) >= 252u && bits != 0ULL);
            res[q] = (float)t;
            bad |= (unsafe ? 1u : 0u) << q;
        }
        // rare: recompute the flagged outputs with the exact chain, re-reading the window from shared
        // memory (kept out of the unrolled register code on purpose)
#pragma unroll 1
        for (int q = 0; bad >> q; ++q)
            if ((bad >> q) & 1u) {
                const float *wp = win + (q + LW) * wstride;
                double tmp = __dmul_rn((double)wp[0], taps.w[LW]);
#pragma unroll
                for (int j = -LW; j < 0; ++j)
                    tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn((double)wp[j * wstride], (double)wp[-j * wstride]), taps.w[LW + j]));
                const float rv = (float)tmp;
#pragma unroll
                for (int qq = 0; qq < kRun; ++qq)
                    if (qq == q) res[qq] = rv;            // predicated moves: res stays in registers
            }
        return;
    }
#pragma unroll
    for (int q = 0; q < kRun; ++q) res[q] = exact_output<LW>(x, taps, q + LW);
}

// filtered axis has stride `inner` (> 1 in practice: z and y passes); lines are contiguous in memory
template <int LW>
__global__ void __launch_bounds__(kLines * kGroups)
gauss_strided_kernel(const float *__restrict__ in, float *__restrict__ out, long long n,
                     long long inner, int r, Taps taps, unsigned cert) {
    __shared__ __align__(16) float tile[kTN + 2 * LW][kLines];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const unsigned nchunks = (unsigned)((n + kTN - 1) / kTN);
    const long long n0 = (long long)(blockIdx.x % nchunks) * kTN;      // n-chunk fastest: blocks that
    const long long i = (long long)(blockIdx.x / nchunks) * kLines + tx; // share halo rows are co-resident
    const long long o = blockIdx.y;
    const float *src = in + o * n * inner;
    float *dst = out + o * n * inner;
    const bool live = i < inner;
    // no reflection can occur when the tile (with halo) stays inside the zero-padded line
    const bool simple = (n0 - LW + r >= 0) && (n0 + kTN + LW - 1 < n + r);
    const long long i0 = i - tx;
    if (simple && (inner & 3) == 0 && i0 + kLines <= inner && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        // vector staging (the passes are instruction-issue bound, profiles/): the 64 lines of a tile row are
        // contiguous in memory -> 16 float4 per row, 8 rows per step, every load of the thread in flight at once
        constexpr int kW = kTN + 2 * LW, kSteps = (kW + 7) / 8;
        const int tid = ty * kLines + tx, c4 = tid & 15, r0 = tid >> 4;
        const float4 *s4 = reinterpret_cast<const float4 *>(src + i0) + c4;
        float4 tmp[kSteps];
#pragma unroll
        for (int sidx = 0; sidx < kSteps; ++sidx) {
            const int row = r0 + 8 * sidx;
            const long long p = n0 - LW + row;
            tmp[sidx] = (row < kW && p >= 0 && p < n) ? __ldg(s4 + p * (inner >> 2)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int sidx = 0; sidx < kSteps; ++sidx) {
            const int row = r0 + 8 * sidx;
            if (row < kW) *reinterpret_cast<float4 *>(&tile[row][4 * c4]) = tmp[sidx];
        }
    } else {
        // stage the (kTN + 2 LW) x 64 source tile: loads are issued in batches of kBatch rows before any is
        // stored, so that kBatch global loads per thread are in flight (the pass is latency bound otherwise)
        constexpr int kRowsT = (kTN + 2 * LW + kGroups - 1) / kGroups;
        constexpr int kBatch = 14;
#pragma unroll 1
        for (int b0 = 0; b0 < kRowsT; b0 += kBatch) {
            float tmp[kBatch];
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                const int row = ty + (b0 + b) * kGroups;
                float vv = 0.f;
                if (b0 + b < kRowsT && row < kTN + 2 * LW && live) {
                    const long long p = n0 - LW + row;
                    const long long q = simple ? ((p >= 0 && p < n) ? p : -1) : line_src(p, n, r);
                    if (q >= 0) vv = __ldg(src + q * inner + i);
                }
                tmp[b] = vv;
            }
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                const int row = ty + (b0 + b) * kGroups;
                if (b0 + b < kRowsT && row < kTN + 2 * LW) tile[row][tx] = tmp[b];
            }
        }
    }
    __syncthreads();
#pragma unroll 1
    for (int run = 0; run < kRunsPerThread; ++run) {
        const int base = (ty * kRunsPerThread + run) * kRun;
        if (n0 + base >= n) break;
        double x[kRun + 2 * LW];
#pragma unroll
        for (int k = 0; k < kRun + 2 * LW; ++k) x[k] = (double)tile[base + k][tx];
        float res[kRun];
        run_from_window<LW>(x, taps, res, cert, &tile[base][tx], kLines);
        if (live) {
#pragma unroll
            for (int q = 0; q < kRun; ++q)
                if (n0 + base + q < n) dst[(n0 + base + q) * inner + i] = res[q];
        }
    }
}

// filtered axis is the contiguous one (x pass).  64 rows x kTN outputs; the tile is transposed on the
// way in (pitch 65 words -> conflict-free both ways) and the results are transposed on the way out so
// that both global read and write are coalesced.
template <int LW>
__global__ void __launch_bounds__(kLines * kGroups)
gauss_contig_kernel(const float *__restrict__ in, float *__restrict__ out, long long rows,
                    long long n, int r, Taps taps, unsigned cert) {
    constexpr int kPitch = kLines + 1;
    __shared__ float tile[(kTN + 2 * LW) * kPitch];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * kLines + tx;
    const unsigned nchunks = (unsigned)((n + kTN - 1) / kTN);
    const long long n0 = (long long)(blockIdx.x % nchunks) * kTN;
    const long long row0 = (long long)(blockIdx.x / nchunks) * kLines;
    constexpr int kW = kTN + 2 * LW;
    const bool simple_t = (n0 - LW + r >= 0) && (n0 + kTN + LW - 1 < n + r);
    const bool vec_rows = (n & 3) == 0 && row0 + kLines <= rows;
    if (simple_t && vec_rows && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
        // vector staging: a warp covers 4 rows x 8 float4 (128 contiguous bytes per row); with the 65-word
        // pitch the transposed scalar stores of such a group hit 32 different banks
        constexpr int kLWa = (LW + 3) & ~3, kShift = kLWa - LW;       // tile origin rounded down to 4 floats
        constexpr int kCG = (kTN + 2 * kLWa + 31) / 32;                // column groups of 32 floats
        constexpr int kIters = (16 * kCG + 3) / 4;                     // per warp (4 warps)
        const int w = tid >> 5, lane = tid & 31, c4l = lane & 7, rl = lane >> 3;
        float4 tmp[kIters];
#pragma unroll
        for (int j = 0; j < kIters; ++j) {
            const int it = w + 4 * j, rg = it / kCG, cg = it - rg * kCG;
            const long long xg = n0 - kLWa + 4 * (cg * 8 + c4l);
            tmp[j] = (it < 16 * kCG && xg >= 0 && xg < n)
                         ? __ldg(reinterpret_cast<const float4 *>(in + (row0 + rg * 4 + rl) * n + xg))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < kIters; ++j) {
            const int it = w + 4 * j, rg = it / kCG, cg = it - rg * kCG;
            if (it < 16 * kCG) {
                const int xx0 = 4 * (cg * 8 + c4l) - kShift, rr = rg * 4 + rl;
                const float e4[4] = {tmp[j].x, tmp[j].y, tmp[j].z, tmp[j].w};
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (xx0 + e >= 0 && xx0 + e < kW) tile[(xx0 + e) * kPitch + rr] = e4[e];
            }
        }
    } else {
        // thread -> (row group, x lane): each warp reads 32 consecutive x of one row (coalesced); loads are
        // batched (kBatch in flight per thread) before the transposed stores
        constexpr int kThreadsB = kLines * kGroups;
        constexpr int kXL = 32;                              // x lanes
        constexpr int kRG = kThreadsB / kXL;                 // rows handled concurrently
        constexpr int kXSteps = (kW + kXL - 1) / kXL;
        constexpr int kBatch = 8;
        const int xl = tid % kXL, rg = tid / kXL;
        const bool simple = (n0 - LW + r >= 0) && (n0 + kTN + LW - 1 < n + r);
#pragma unroll 1
        for (int rr0 = rg; rr0 < kLines; rr0 += kRG * kBatch) {
#pragma unroll 1
            for (int xs = 0; xs < kXSteps; ++xs) {
                const int xx = xs * kXL + xl;
                const long long p = n0 - LW + xx;
                long long q = -1;
                if (xx < kW) q = simple ? ((p >= 0 && p < n) ? p : -1) : line_src(p, n, r);
                float tmp[kBatch];
#pragma unroll
                for (int b = 0; b < kBatch; ++b) {
                    const int rr = rr0 + b * kRG;
                    tmp[b] = (rr < kLines && row0 + rr < rows && q >= 0) ? __ldg(in + (row0 + rr) * n + q) : 0.f;
                }
#pragma unroll
                for (int b = 0; b < kBatch; ++b) {
                    const int rr = rr0 + b * kRG;
                    if (rr < kLines && xx < kW) tile[xx * kPitch + rr] = tmp[b];
                }
            }
        }
    }
    __syncthreads();
    float res[kRunsPerThread][kRun];
#pragma unroll
    for (int run = 0; run < kRunsPerThread; ++run) {
        const int base = (ty * kRunsPerThread + run) * kRun;
        double x[kRun + 2 * LW];
#pragma unroll
        for (int k = 0; k < kRun + 2 * LW; ++k) x[k] = (double)tile[(base + k) * kPitch + tx];
        run_from_window<LW>(x, taps, res[run], cert, &tile[base * kPitch + tx], kPitch);
    }
    __syncthreads();
#pragma unroll
    for (int run = 0; run < kRunsPerThread; ++run) {
        const int base = (ty * kRunsPerThread + run) * kRun;
#pragma unroll
        for (int q = 0; q < kRun; ++q) tile[(base + q) * kPitch + tx] = res[run][q];
    }

    __syncthreads();
    if (vec_rows && n0 + kTN <= n && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        const int w = tid >> 5, lane = tid & 31, c4l = lane & 7, rl = lane >> 3;
#pragma unroll
        for (int j = 0; j < 8; ++j) {                       // 16 row groups x 2 column groups over 4 warps
            const int it = w + 4 * j, rg = it >> 1, cg = it & 1;
            const int xx0 = 4 * (cg * 8 + c4l), rr = rg * 4 + rl;
            const float4 o4 = make_float4(tile[xx0 * kPitch + rr], tile[(xx0 + 1) * kPitch + rr],
                                          tile[(xx0 + 2) * kPitch + rr], tile[(xx0 + 3) * kPitch + rr]);
            *reinterpret_cast<float4 *>(out + (row0 + rr) * n + n0 + xx0) = o4;
        }
        return;
    }
    for (int idx = tid; idx < kLines * kTN; idx += kLines * kGroups) {
        int rr = idx / kTN, xx = idx - rr * kTN;
        if (row0 + rr < rows && n0 + xx < n) out[(row0 + rr) * n + n0 + xx] = tile[xx * kPitch + rr];
    }
}

// certified-FMA margin (double ulps); 0 = exact chain only.  Off by default: the passes turned out to be
// instruction-issue bound, not FP64-pipe bound (ncu: issue slots 78 % busy, FP64 pipe 32 %), and the
// certificate's integer work costs more issue slots than the 10 fused operations save (profiles/).
static unsigned g_gauss_cert = 0;

template <int LW>
static int launch_pass_fast(fpl_ctx *ctx, const float *in, float *out, long long outer, long long n,
                            long long inner, int r, const Taps &taps, cudaStream_t st) {
    dim3 block(kLines, kGroups);
    const long long nchunks = (n + kTN - 1) / kTN;
    // the certificate needs non-negative taps (Gaussian taps always are) and its error bound covers LW <= 10
    static const int env_cert = getenv("FPL_GAUSS_CERT") ? atoi(getenv("FPL_GAUSS_CERT")) : -1;   // experiments
    unsigned cert = LW <= 10 ? (env_cert >= 0 ? (unsigned)env_cert : g_gauss_cert) : 0u;
    for (int i = 0; i < 2 * LW + 1; ++i) if (!(taps.w[i] >= 0.0)) cert = 0u;
    if (inner == 1) {
        const long long rows = outer;
        const long long gx = nchunks * ((rows + kLines - 1) / kLines);
        FPL_REQUIRE(gx < 2147483647LL, "gauss pass: volume too large for one launch");
        gauss_contig_kernel<LW><<<dim3((unsigned)gx), block, 0, st>>>(in, out, rows, n, r, taps, cert);
        FPL_LAUNCH_CHECK(ctx);
        return FPL_OK;
    }
    const long long gx = nchunks * ((inner + kLines - 1) / kLines);
    FPL_REQUIRE(gx < 2147483647LL && outer <= 65535, "gauss pass: volume too large for one launch");
    gauss_strided_kernel<LW><<<dim3((unsigned)gx, (unsigned)outer), block, 0, st>>>(in, out, n, inner, r, taps, cert);
    FPL_LAUNCH_CHECK(ctx);
    return FPL_OK;
}

static int launch_pass(fpl_ctx *ctx, const float *in, float *out, long long outer, long long n,
                       long long inner, int r, int lw, const Taps &taps, cudaStream_t st,
                       bool force_generic) {
    if (!force_generic) {
        switch (lw) {
            case 2:  return launch_pass_fast<2>(ctx, in, out, outer, n, inner, r, taps, st);
            case 3:  return launch_pass_fast<3>(ctx, in, out, outer, n, inner, r, taps, st);
            case 4:  return launch_pass_fast<4>(ctx, in, out, outer, n, inner, r, taps, st);
            case 8:  return launch_pass_fast<8>(ctx, in, out, outer, n, inner, r, taps, st);
            case 10: return launch_pass_fast<10>(ctx, in, out, outer, n, inner, r, taps, st);
            default: break;
        }
    }
    long long total = outer * n * inner;
    long long blocks = (total + 255) / 256;
    long long cap = (long long)ctx->sm_count * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    gauss_generic_kernel<<<(unsigned)blocks, 256, 0, st>>>(in, out, outer, n, inner, r, lw, taps);
    FPL_LAUNCH_CHECK(ctx);
    return FPL_OK;
}

// ------------------------------------------------------------------------------------------------
// radix select
// ------------------------------------------------------------------------------------------------
constexpr int kHistBits0 = 11, kHistBits1 = 11, kHistBits2 = 10;

__global__ void select_init_kernel(SelectState *s, unsigned long long rank,
                                   unsigned long long extra_zeros) {
    int t = threadIdx.x;
    for (int b = t; b < kHistSlots * 2048; b += blockDim.x) (&s->hist[0][0])[b] = 0;
    if (t == 0) { s->rank = rank; s->prefix = 0; s->prefix_mask = 0; s->extra_zeros = extra_zeros; s->nan_count = 0; s->n_above = 0; s->n_bin = 0; }
}

// histogram of ((key >> shift) & (bins-1)) over values whose key matches the current prefix
__global__ void __launch_bounds__(512)
select_hist_kernel(const float *__restrict__ v, long long n, SelectState *s, int shift, int bins,
                   int count_nan) {
    __shared__ unsigned h[2048];
    __shared__ unsigned nan_local;
    for (int b = threadIdx.x; b < bins; b += blockDim.x) h[b] = 0;
    if (threadIdx.x == 0) nan_local = 0;
    __syncthreads();
    const unsigned prefix = s->prefix, pmask = s->prefix_mask;
    const unsigned bmask = (unsigned)bins - 1u;
    long long n4 = ((reinterpret_cast<uintptr_t>(v) & 15) == 0) ? (n >> 2) : 0;
    const float4 *v4 = reinterpret_cast<const float4 *>(v);
    unsigned last_bin = 0xffffffffu, run = 0, my_nan = 0;
    auto feed = [&](float f) {
        if (count_nan && f != f) { ++my_nan; }
        unsigned k = f2key(f);
        if ((k & pmask) == prefix) {
            unsigned b = (k >> shift) & bmask;
            if (b == last_bin) { ++run; }
            else { if (run) atomicAdd(&h[last_bin], run); last_bin = b; run = 1; }
        }
    };
    {
        const long long stride = (long long)gridDim.x * blockDim.x;
        long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
        for (; i + 3 * stride < n4; i += 4 * stride) {      // four 16-byte loads in flight per thread
            const float4 q0 = __ldg(v4 + i), q1 = __ldg(v4 + i + stride), q2 = __ldg(v4 + i + 2 * stride),
                         q3 = __ldg(v4 + i + 3 * stride);
            feed(q0.x); feed(q0.y); feed(q0.z); feed(q0.w);
            feed(q1.x); feed(q1.y); feed(q1.z); feed(q1.w);
            feed(q2.x); feed(q2.y); feed(q2.z); feed(q2.w);
            feed(q3.x); feed(q3.y); feed(q3.z); feed(q3.w);
        }
        for (; i < n4; i += stride) {
            const float4 q = __ldg(v4 + i);
            feed(q.x); feed(q.y); feed(q.z); feed(q.w);
        }
    }
    // tail
    for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        feed(__ldg(v + i));
    if (run) atomicAdd(&h[last_bin], run);
    if (my_nan) atomicAdd(&nan_local, my_nan);
    __syncthreads();
    for (int b = threadIdx.x; b < bins; b += blockDim.x)
        if (h[b]) atomicAdd(&s->hist[blockIdx.x % kHistSlots][b], (unsigned long long)h[b]);
    if (threadIdx.x == 0 && nan_local) atomicAdd(&s->nan_count, (unsigned long long)nan_local);
}

// one block: add the implicit zeros, locate the bin holding `rank` (parallel prefix over <= 2048 bins),
// extend the prefix, clear the histogram
__global__ void __launch_bounds__(1024)
select_scan_kernel(SelectState *s, int shift, int bins) {
    __shared__ unsigned long long cum[2048];
    __shared__ int sel_bin;
    const int t = threadIdx.x;
    const unsigned zkey = 0x80000000u;     // key of +0.0f
    const bool zero_here = (zkey & s->prefix_mask) == s->prefix;
    const int zbin = (int)((zkey >> shift) & (unsigned)(bins - 1));
    for (int b = t; b < 2048; b += 1024) {
        unsigned long long c = 0ULL;
        if (b < bins)
            for (int sl = 0; sl < kHistSlots; ++sl) c += s->hist[sl][b];
        if (zero_here && b == zbin) c += s->extra_zeros;
        cum[b] = c;
    }
    if (t == 0) sel_bin = bins - 1;
    __syncthreads();
    for (int off = 1; off < 2048; off <<= 1) {          // Hillis-Steele inclusive scan, two entries per thread
        unsigned long long v0 = 0, v1 = 0;
        const int b0 = t, b1 = t + 1024;
        if (b0 >= off) v0 = cum[b0 - off];
        if (b1 >= off) v1 = cum[b1 - off];
        __syncthreads();
        cum[b0] += v0; cum[b1] += v1;
        __syncthreads();
    }
    const unsigned long long rank = s->rank;
    for (int b = t; b < bins; b += 1024) {
        const unsigned long long before = b ? cum[b - 1] : 0ULL;
        if (rank >= before && rank < cum[b]) sel_bin = b;     // exactly one bin satisfies this
    }
    __syncthreads();
    const int sel = sel_bin;
    const unsigned long long before = sel ? cum[sel - 1] : 0ULL;
    const unsigned long long upto = cum[sel], total = cum[bins - 1];
    __syncthreads();
    for (int b = t; b < kHistSlots * 2048; b += 1024) (&s->hist[0][0])[b] = 0;
    if (t == 0) {
        // interior (materialised) values above / inside the selected bin: the implicit zeros are not interior
        s->n_above += (total - upto) - ((zero_here && zbin > sel) ? s->extra_zeros : 0ULL);
        s->n_bin = (upto - before) - ((zero_here && zbin == sel) ? s->extra_zeros : 0ULL);
        s->rank = rank - before;
        s->prefix |= ((unsigned)sel) << shift;
        s->prefix_mask |= ((unsigned)(bins - 1)) << shift;
    }
}

struct ThreshOut {            // device
    double thresh;
    float v_lo, v_hi;
    unsigned long long nan_count;
};

// NumPy _lerp in float32 + np.maximum(., thd)
__global__ void threshold_kernel(const SelectState *lo, const SelectState *hi, float gamma, double thd,
                                 ThreshOut *out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float a = key2f(lo->prefix), b = key2f(hi->prefix);
    float d = __fsub_rn(b, a);
    float res = __fadd_rn(a, __fmul_rn(d, gamma));
    if (gamma >= 0.5f) res = __fsub_rn(b, __fmul_rn(d, __fsub_rn(1.0f, gamma)));
    double t;
    if (lo->nan_count) {
        t = nan("");
    } else {
        double p = (double)res;
        t = (p != p || thd != thd) ? nan("") : (p > thd ? p : thd);
    }
    out->thresh = t; out->v_lo = a; out->v_hi = b; out->nan_count = lo->nan_count;
}

// ------------------------------------------------------------------------------------------------
// candidate compaction
// ------------------------------------------------------------------------------------------------
struct Counters {                       // device
    unsigned long long n_cand;          // list A
    unsigned long long n_next;          // list B
    unsigned long long n_work;          // worklist
    unsigned long long n_sel_round;     // newly selected this round
    unsigned long long n_det;           // all selected
    unsigned long long overflow;
    unsigned long long ball_checks;
    unsigned long long n_alive_owned;   // slab sessions: valid candidates inside the owned index range (this round)
};

// every thread takes 8 consecutive voxels (two 16-byte loads in flight), a warp reserves its output
// range with one atomic; candidate order in the list is irrelevant to the result
__global__ void __launch_bounds__(256)
compact_candidates_kernel(const float *__restrict__ v, long long n, double thresh,
                          unsigned long long *cand_idx, float *cand_val, long long capacity,
                          Counters *cnt) {
    const unsigned lane = threadIdx.x & 31;
    const bool aligned = (reinterpret_cast<uintptr_t>(v) & 15) == 0;
    const long long n8 = (n + 7) / 8;
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c - threadIdx.x < n8;
         c += (long long)gridDim.x * blockDim.x) {
        const long long i0 = c * 8;
        float f[8];
        if (c < n8 && aligned && i0 + 8 <= n) {
            const float4 a4 = __ldg(reinterpret_cast<const float4 *>(v + i0));
            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(v + i0) + 1);
            f[0] = a4.x; f[1] = a4.y; f[2] = a4.z; f[3] = a4.w; f[4] = b4.x; f[5] = b4.y; f[6] = b4.z; f[7] = b4.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = (c < n8 && i0 + j < n) ? __ldg(v + i0 + j) : 0.f;
        }
        unsigned mask = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (((double)f[j] > thresh) && (f[j] > 0.f)) mask |= 1u << j;
        const unsigned cnt_me = __popc(mask);
        // exclusive prefix over the warp
        unsigned pre = cnt_me;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, pre, o);
            if (lane >= (unsigned)o) pre += t;
        }
        const unsigned total = __shfl_sync(0xffffffffu, pre, 31);
        // one global atomic per block iteration (2048 voxels): warp totals -> shared -> block base
        __shared__ unsigned s_wtot[8];
        __shared__ unsigned long long s_base;
        const unsigned wid = threadIdx.x >> 5;
        if (lane == 31) s_wtot[wid] = total;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned t = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) t += s_wtot[k];
            s_base = t ? atomicAdd(&cnt->n_cand, (unsigned long long)t) : 0ULL;
        }
        __syncthreads();
        unsigned wbefore = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) if ((unsigned)k < wid) wbefore += s_wtot[k];
        unsigned long long pos = s_base + wbefore + (pre - cnt_me);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (mask & (1u << j)) {
                if ((long long)pos < capacity) { cand_idx[pos] = (unsigned long long)(i0 + j); cand_val[pos] = f[j]; }
                else atomicAdd(&cnt->overflow, 1ULL);
                ++pos;
            }
    }
}

// ------------------------------------------------------------------------------------------------
// NMS rounds
// ------------------------------------------------------------------------------------------------
struct Dims { long long Z, Y, X; };

// flat index -> (z,y,x); 32-bit arithmetic whenever the volume has fewer than 2^32 voxels
__device__ __forceinline__ void decode_idx(unsigned long long idx, const Dims &d, long long &z, long long &y, long long &x) {
    if ((unsigned long long)d.Z * d.Y * d.X <= 0xffffffffULL) {
        const unsigned i = (unsigned)idx, X = (unsigned)d.X, Y = (unsigned)d.Y;
        const unsigned row = i / X;
        x = i - row * X; z = row / Y; y = row - (unsigned)z * Y;
    } else {
        x = (long long)(idx % (unsigned long long)d.X);
        y = (long long)((idx / (unsigned long long)d.X) % (unsigned long long)d.Y);
        z = (long long)(idx / ((unsigned long long)d.X * d.Y));
    }
}

__device__ __forceinline__ bool is_suppressed(const unsigned *sup, unsigned long long idx) {
    return (sup[idx >> 5] >> (idx & 31)) & 1u;
}

// A -> B (still valid) and worklist (valid and no better valid voxel among the 26 neighbours)
__global__ void __launch_bounds__(256)
nms_filter_kernel(const float *__restrict__ v, const unsigned *__restrict__ sup, Dims d,
                  const unsigned long long *__restrict__ a_idx, const float *__restrict__ a_val,
                  unsigned long long *b_idx, float *b_val, unsigned long long *w_idx, float *w_val,
                  long long w_capacity, Counters *cnt, unsigned long long own_lo = 0ULL,
                  unsigned long long own_hi = ~0ULL) {
    // [own_lo, own_hi): flat index range whose voxels this call may put on the worklist (slab sessions: the
    // planes a rank owns; halo candidates stay in the list but are decided by their owner)
    const unsigned long long nA = cnt->n_cand;
    const unsigned lane = threadIdx.x & 31;
    for (unsigned long long i0 = (unsigned long long)blockIdx.x * blockDim.x; i0 < nA;
         i0 += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long i = i0 + threadIdx.x;
        bool alive = false, is_work = false, owned_alive = false;
        unsigned long long idx = 0; float val = 0.f;
        if (i < nA) {
            idx = a_idx[i]; val = a_val[i];
            alive = !is_suppressed(sup, idx);
        }
        if (alive) {
            long long x, y, z;
            decode_idx(idx, d, z, y, x);
            bool better = false;
            for (int dz = -1; dz <= 1 && !better; ++dz) {
                long long zz = z + dz; if (zz < 0 || zz >= d.Z) continue;
                for (int dy = -1; dy <= 1 && !better; ++dy) {
                    long long yy = y + dy; if (yy < 0 || yy >= d.Y) continue;
                    for (int dx = -1; dx <= 1; ++dx) {
                        long long xx = x + dx; if (xx < 0 || xx >= d.X) continue;
                        if (!(dz | dy | dx)) continue;
                        unsigned long long q = ((unsigned long long)zz * d.Y + yy) * d.X + xx;
                        float vq = __ldg(v + q);
                        if ((vq > val || (vq == val && q < idx)) && !is_suppressed(sup, q)) {
                            better = true; break;
                        }
                    }
                }
            }
            owned_alive = idx >= own_lo && idx < own_hi;
            is_work = !better && owned_alive;
        }
        if (own_hi != ~0ULL) {                           // slab sessions: valid owned candidates, one atomic per warp
            const unsigned m_own = __ballot_sync(0xffffffffu, owned_alive);
            if (lane == 0 && m_own) atomicAdd(&cnt->n_alive_owned, (unsigned long long)__popc(m_own));
        }
        unsigned m_alive = __ballot_sync(0xffffffffu, alive);
        unsigned m_work = __ballot_sync(0xffffffffu, is_work);
        unsigned long long base_b = 0, base_w = 0;
        if (lane == 0) {
            if (m_alive) base_b = atomicAdd(&cnt->n_next, (unsigned long long)__popc(m_alive));
            if (m_work) base_w = atomicAdd(&cnt->n_work, (unsigned long long)__popc(m_work));
        }
        base_b = __shfl_sync(0xffffffffu, base_b, 0);
        base_w = __shfl_sync(0xffffffffu, base_w, 0);
        unsigned below = (1u << lane) - 1u;
        if (alive) {
            unsigned long long p = base_b + __popc(m_alive & below);
            b_idx[p] = idx; b_val[p] = val;
        }
        if (is_work) {
            unsigned long long p = base_w + __popc(m_work & below);
            if ((long long)p < w_capacity) { w_idx[p] = idx; w_val[p] = val; }
            else atomicAdd(&cnt->overflow, 1ULL);
        }
    }
}

__device__ __forceinline__ int isqrt_floor(int v) {
    int s = (int)sqrtf((float)v);
    while (s * s > v) --s;
    while ((s + 1) * (s + 1) <= v) ++s;
    return s;
}

// ---- brick maxima: G[bz][by][bx] = max of the smoothed map over an 8^3 brick (validity-agnostic) -----
constexpr int kBrick = 8;
__global__ void __launch_bounds__(256)
brick_max_kernel(const float *__restrict__ v, Dims d, int gy, int gx, float *__restrict__ g) {
    // one block per (bz, by) brick row.  Warp w owns 8 of the 64 (z,y) rows; a lane owns one brick
    // (8 consecutive x) per 256-wide chunk and keeps all its 8 rows' loads in flight.
    extern __shared__ float s_max[];                 // [gx]
    const int bz = blockIdx.x / gy, by = blockIdx.x % gy;
    for (int i = threadIdx.x; i < gx; i += blockDim.x) s_max[i] = -INFINITY;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool vec_ok = (d.X % 4 == 0) && ((reinterpret_cast<uintptr_t>(v) & 15) == 0);
    for (long long xc = 0; xc < d.X; xc += 256) {
        const long long x = xc + 8 * lane;
        float m = -INFINITY;
        if (x < d.X) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int row = warp + 8 * k;
                const long long z = (long long)bz * kBrick + row / kBrick, y = (long long)by * kBrick + row % kBrick;
                if (z >= d.Z || y >= d.Y) continue;
                const float *rp = v + (z * d.Y + y) * d.X + x;
                if (vec_ok && x + 8 <= d.X) {
                    const float4 a4 = __ldg(reinterpret_cast<const float4 *>(rp));
                    const float4 b4 = __ldg(reinterpret_cast<const float4 *>(rp) + 1);
                    m = fmaxf(m, fmaxf(fmaxf(fmaxf(a4.x, a4.y), fmaxf(a4.z, a4.w)), fmaxf(fmaxf(b4.x, b4.y), fmaxf(b4.z, b4.w))));
                } else {
                    for (int e = 0; e < 8 && x + e < d.X; ++e) m = fmaxf(m, __ldg(rp + e));
                }
            }
            // float max via integer atomics on the monotone key of non-negative values
            atomicMax(reinterpret_cast<int *>(&s_max[x / kBrick]), m >= 0.f ? __float_as_int(m) : (int)0x80000000);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < gx; i += blockDim.x) g[((size_t)bz * gy + by) * gx + i] = s_max[i];
}

// one block per worklist entry: is there a better valid voxel inside the ball (d2 <= r^2)?
// A voxel better than p has value >= S(p), so it can only sit in a brick whose maximum is >= S(p).
// Phase 1 lists those bricks (of the <= 8^3 bricks touching the ball; the brick grid is a few MB and
// stays in L2); phase 2 scans only the listed bricks densely, testing distance, order and validity.
// For an isolated peak the list is empty apart from its own blob.  Exactness does not depend on the
// grid ignoring validity: it only over-approximates the set of bricks to scan.
__global__ void __launch_bounds__(256)
nms_ballcheck_kernel(const float *__restrict__ v, const unsigned *__restrict__ sup, Dims d, int r,
                     const float *__restrict__ g, int gz, int gy, int gx,
                     const unsigned long long *__restrict__ w_idx, const float *__restrict__ w_val,
                     unsigned long long *det_idx, float *det_val, unsigned long long *sel_idx,
                     long long det_capacity, Counters *cnt, const SelectState *cut) {
    __shared__ int s_list[1024];
    __shared__ int s_n, found;
    const unsigned long long nW = cnt->n_work;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int r2 = r * r;
    // fast path, first round: the worklist holds every local maximum above the level-1 cut-off; entries
    // below the (meanwhile known) level-2 cut-off are not part of the candidate superset
    const unsigned cutoff = cut ? (cut->prefix > 0x80000001u ? cut->prefix : 0x80000001u) : 0u;
    for (unsigned long long w = blockIdx.x; w < nW; w += gridDim.x) {
        const unsigned long long idx = w_idx[w];
        const float val = w_val[w];
        if (cut && f2key(val) < cutoff) continue;          // block-uniform
        if (threadIdx.x == 0) { found = 0; s_n = 0; }
        __syncthreads();
        const int x = (int)(idx % (unsigned long long)d.X);
        const int y = (int)((idx / (unsigned long long)d.X) % (unsigned long long)d.Y);
        const int z = (int)(idx / ((unsigned long long)d.X * d.Y));
        const int bz0 = max(z - r, 0) / kBrick, bz1 = min((long long)z + r, d.Z - 1) / kBrick;
        const int by0 = max(y - r, 0) / kBrick, by1 = min((long long)y + r, d.Y - 1) / kBrick;
        const int bx0 = max(x - r, 0) / kBrick, bx1 = min((long long)x + r, d.X - 1) / kBrick;
        const int nbz = bz1 - bz0 + 1, nby = by1 - by0 + 1, nbx = bx1 - bx0 + 1;
        for (int i = threadIdx.x; i < nbz * nby * nbx; i += blockDim.x) {
            const int bx = bx0 + i % nbx, by = by0 + (i / nbx) % nby, bz = bz0 + i / (nbx * nby);
            const float gm = __ldg(g + ((size_t)bz * gy + by) * gx + bx);
            if (!(gm >= val)) continue;
            // closest point of the brick to p
            const int cz = min(max(z, bz * kBrick), bz * kBrick + kBrick - 1);
            const int cy = min(max(y, by * kBrick), by * kBrick + kBrick - 1);
            const int cx = min(max(x, bx * kBrick), bx * kBrick + kBrick - 1);
            const int dd = (cz - z) * (cz - z) + (cy - y) * (cy - y) + (cx - x) * (cx - x);
            if (dd > r2) continue;
            const int slot = atomicAdd(&s_n, 1);
            if (slot < 1024) s_list[slot] = (bz << 20) | (by << 10) | bx;
        }
        __syncthreads();
        const int nlist = min(s_n, 1024);           // <= 8^3 bricks can touch a ball of radius <= 27
        for (int li = warp; li < nlist; li += nwarps) {
            if (*(volatile int *)&found) break;
            const int code = s_list[li];
            const int bz = code >> 20, by = (code >> 10) & 1023, bx = code & 1023;
            bool hit = false;
            // 64 rows of 8 voxels: lane -> rows (lane, lane + 32), 8 voxels each as two float4
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int row = lane + 32 * h;
                const long long zz = (long long)bz * kBrick + row / kBrick, yy = (long long)by * kBrick + row % kBrick;
                if (zz >= d.Z || yy >= d.Y) continue;
                const int dzy = (int)((zz - z) * (zz - z) + (yy - y) * (yy - y));
                if (dzy > r2) continue;
                const unsigned long long rowbase = ((unsigned long long)zz * d.Y + yy) * d.X;
#pragma unroll
                for (int e = 0; e < kBrick; ++e) {
                    const long long xx = (long long)bx * kBrick + e;
                    if (xx >= d.X) break;
                    const int ddx = (int)(xx - x);
                    if (dzy + ddx * ddx > r2) continue;
                    const unsigned long long q = rowbase + xx;
                    const float vq = __ldg(v + q);
                    if ((vq > val || (vq == val && q < idx)) && !is_suppressed(sup, q)) hit = true;
                }
            }
            if (__any_sync(0xffffffffu, hit)) { if (lane == 0) found = 1; }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (!found) {
                unsigned long long p = atomicAdd(&cnt->n_det, 1ULL);
                unsigned long long s = atomicAdd(&cnt->n_sel_round, 1ULL);
                if ((long long)p < det_capacity) { det_idx[p] = idx; det_val[p] = val; sel_idx[s] = idx; }
                else atomicAdd(&cnt->overflow, 1ULL);
            }
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&cnt->ball_checks, nW);
}

// one block per newly selected point: set the suppressed bit of every voxel in its ball
__global__ void __launch_bounds__(256)
nms_suppress_kernel(unsigned *sup, Dims d, int r, const unsigned long long *__restrict__ sel_idx,
                    const Counters *cnt) {
    const unsigned long long nS = cnt->n_sel_round;
    const int side = 2 * r + 1;
    for (unsigned long long s = blockIdx.x; s < nS; s += gridDim.x) {
        const unsigned long long idx = sel_idx[s];
        const long long x = (long long)(idx % (unsigned long long)d.X);
        const long long y = (long long)((idx / (unsigned long long)d.X) % (unsigned long long)d.Y);
        const long long z = (long long)(idx / ((unsigned long long)d.X * d.Y));
        for (int row = threadIdx.x; row < side * side; row += blockDim.x) {
            int dz = row / side - r, dy = row % side - r;
            int rem = r * r - dz * dz - dy * dy;
            if (rem < 0) continue;
            long long zz = z + dz, yy = y + dy;
            if (zz < 0 || zz >= d.Z || yy < 0 || yy >= d.Y) continue;
            int hw = isqrt_floor(rem);
            long long x0 = x - hw < 0 ? 0 : x - hw;
            long long x1 = x + hw >= d.X ? d.X - 1 : x + hw;
            unsigned long long q0 = ((unsigned long long)zz * d.Y + yy) * d.X + x0;
            unsigned long long q1 = ((unsigned long long)zz * d.Y + yy) * d.X + x1;
            // 64-bit words: a row of <= 2r+1 voxels touches at most two of them (the bitmap is 8-byte aligned
            // and padded to a multiple of 64 bits)
            unsigned long long *sup64 = reinterpret_cast<unsigned long long *>(sup);
            for (unsigned long long wd = q0 >> 6; wd <= (q1 >> 6); ++wd) {
                unsigned long long lo = wd << 6;
                unsigned b0 = q0 > lo ? (unsigned)(q0 - lo) : 0u;
                unsigned b1 = q1 < lo + 63 ? (unsigned)(q1 - lo) : 63u;
                unsigned long long mask = (b1 == 63u ? ~0ULL : ((1ULL << (b1 + 1)) - 1ULL)) & ~((1ULL << b0) - 1ULL);
                atomicOr(&sup64[wd], mask);
            }
        }
    }
}

__global__ void round_reset_kernel(Counters *cnt) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        cnt->n_cand = cnt->n_next; cnt->n_next = 0; cnt->n_work = 0; cnt->n_sel_round = 0; cnt->n_alive_owned = 0;
    }
}

// one block per point given by coordinates (may lie outside the volume: only the part of its ball inside counts)
__global__ void __launch_bounds__(256)
nms_suppress_zyx_kernel(unsigned *sup, Dims d, int r, const long long *__restrict__ zyx, long long n_pts) {
    const int side = 2 * r + 1;
    for (long long s = blockIdx.x; s < n_pts; s += gridDim.x) {
        const long long z = zyx[3 * s], y = zyx[3 * s + 1], x = zyx[3 * s + 2];
        for (int row = threadIdx.x; row < side * side; row += blockDim.x) {
            int dz = row / side - r, dy = row % side - r;
            int rem = r * r - dz * dz - dy * dy;
            if (rem < 0) continue;
            long long zz = z + dz, yy = y + dy;
            if (zz < 0 || zz >= d.Z || yy < 0 || yy >= d.Y) continue;
            int hw = isqrt_floor(rem);
            long long x0 = x - hw < 0 ? 0 : x - hw;
            long long x1 = x + hw >= d.X ? d.X - 1 : x + hw;
            if (x1 < x0) continue;
            unsigned long long q0 = ((unsigned long long)zz * d.Y + yy) * d.X + x0;
            unsigned long long q1 = ((unsigned long long)zz * d.Y + yy) * d.X + x1;
            // 64-bit words: a row of <= 2r+1 voxels touches at most two of them (the bitmap is 8-byte aligned
            // and padded to a multiple of 64 bits)
            unsigned long long *sup64 = reinterpret_cast<unsigned long long *>(sup);
            for (unsigned long long wd = q0 >> 6; wd <= (q1 >> 6); ++wd) {
                unsigned long long lo = wd << 6;
                unsigned b0 = q0 > lo ? (unsigned)(q0 - lo) : 0u;
                unsigned b1 = q1 < lo + 63 ? (unsigned)(q1 - lo) : 63u;
                unsigned long long mask = (b1 == 63u ? ~0ULL : ((1ULL << (b1 + 1)) - 1ULL)) & ~((1ULL << b0) - 1ULL);
                atomicOr(&sup64[wd], mask);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// ordering of the detections: (value desc, flat index asc)  -> bitonic sort on (key, idx) pairs
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool det_before(unsigned ka, unsigned long long ia, unsigned kb,
                                           unsigned long long ib) {
    // ka/kb are ~f2key(value): ascending in this key == descending in value
    return ka < kb || (ka == kb && ia < ib);
}

__global__ void sort_prepare_kernel(const unsigned long long *det_idx, const float *det_val,
                                    long long n, long long n_pow2, unsigned *skey,
                                    unsigned long long *sidx) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_pow2;
         i += (long long)gridDim.x * blockDim.x) {
        if (i < n) { skey[i] = ~f2key(det_val[i]); sidx[i] = det_idx[i]; }
        else { skey[i] = 0xffffffffu; sidx[i] = 0xffffffffffffffffULL; }
    }
}

__global__ void bitonic_step_kernel(unsigned *skey, unsigned long long *sidx, long long n_pow2,
                                    long long j, long long k) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_pow2;
         i += (long long)gridDim.x * blockDim.x) {
        long long l = i ^ j;
        if (l > i) {
            unsigned ka = skey[i], kb = skey[l];
            unsigned long long ia = sidx[i], ib = sidx[l];
            bool up = (i & k) == 0;
            bool swap = up ? det_before(kb, ib, ka, ia) : det_before(ka, ia, kb, ib);
            if (swap) { skey[i] = kb; skey[l] = ka; sidx[i] = ib; sidx[l] = ia; }
        }
    }
}

// all steps with j < 2048 of the bitonic network, for k in [k_first, k_last], on 2048-element chunks held in shared
// memory (one launch instead of up to 66 tiny ones; the chunks are aligned, so partners i ^ j stay inside a chunk)
constexpr int kBitonicChunk = 2048;
__global__ void __launch_bounds__(1024)
bitonic_local_kernel(unsigned *skey, unsigned long long *sidx, long long n_pow2, long long k_first, long long k_last) {
    __shared__ unsigned s_key[kBitonicChunk];
    __shared__ unsigned long long s_idx[kBitonicChunk];
    for (long long base = (long long)blockIdx.x * kBitonicChunk; base < n_pow2; base += (long long)gridDim.x * kBitonicChunk) {
        for (int i = threadIdx.x; i < kBitonicChunk; i += blockDim.x)
            if (base + i < n_pow2) { s_key[i] = skey[base + i]; s_idx[i] = sidx[base + i]; }
        __syncthreads();
        for (long long k = k_first; k <= k_last; k <<= 1) {
            long long j0 = k >> 1;
            if (j0 >= kBitonicChunk) j0 = kBitonicChunk >> 1;
            for (long long j = j0; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < kBitonicChunk; i += blockDim.x) {
                    const int l = i ^ (int)j;
                    if (l > i && base + l < n_pow2) {
                        const unsigned ka = s_key[i], kb = s_key[l];
                        const unsigned long long ia = s_idx[i], ib = s_idx[l];
                        const bool up = ((base + i) & k) == 0;
                        const bool swap = up ? det_before(kb, ib, ka, ia) : det_before(ka, ia, kb, ib);
                        if (swap) { s_key[i] = kb; s_key[l] = ka; s_idx[i] = ib; s_idx[l] = ia; }
                    }
                }
                __syncthreads();
            }
        }
        for (int i = threadIdx.x; i < kBitonicChunk; i += blockDim.x)
            if (base + i < n_pow2) { skey[base + i] = s_key[i]; sidx[base + i] = s_idx[i]; }
        __syncthreads();
    }
}

// fplobjdetect.py:233-253: columns (x,y,z,conf) float64; coordinates are already "un-padded"
// (interior indices); keep rows with b <= coord < size-b (buffer in x,y,z order); add the offset.
// Order-preserving compaction by one block.
__global__ void __launch_bounds__(1024)
finish_rows_kernel(const unsigned *skey, const unsigned long long *sidx, long long n, Dims d,
                   int bx, int by, int bz, double ox, double oy, double oz, double *rows,
                   long long capacity, unsigned long long *n_out, unsigned long long *overflow,
                   const ThreshOut *tout) {
    __shared__ unsigned warp_tot[32];
    __shared__ unsigned long long base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long i0 = 0; i0 < n; i0 += blockDim.x) {
        long long i = i0 + threadIdx.x;
        bool keep = false; double x = 0, y = 0, z = 0, c = 0;
        if (i < n) {
            unsigned long long idx = sidx[i];
            long long xi = (long long)(idx % (unsigned long long)d.X);
            long long yi = (long long)((idx / (unsigned long long)d.X) % (unsigned long long)d.Y);
            long long zi = (long long)(idx / ((unsigned long long)d.X * d.Y));
            keep = xi >= bx && yi >= by && zi >= bz && xi < d.X - bx && yi < d.Y - by && zi < d.Z - bz;
            x = (double)xi + ox; y = (double)yi + oy; z = (double)zi + oz;
            const float cf = key2f(~skey[i]);
            c = (double)cf;
            // fast path: NMS ran on a superset (cut-off below the threshold); selected points that are not
            // candidates of the reference (pred > thresh, and the loop stops at max_val <= 0) are dropped here
            if (tout) keep = keep && (c > tout->thresh) && (cf > 0.f);
        }
        unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_tot[warp] = __popc(m);
        __syncthreads();
        unsigned before = 0, total = 0;
        for (unsigned w = 0; w < (blockDim.x >> 5); ++w) { if (w < warp) before += warp_tot[w]; total += warp_tot[w]; }
        if (keep) {
            unsigned long long p = base + before + __popc(m & ((1u << lane) - 1u));
            if ((long long)p < capacity) { rows[4 * p] = x; rows[4 * p + 1] = y; rows[4 * p + 2] = z; rows[4 * p + 3] = c; }
            else atomicAdd(overflow, 1ULL);
        }
        __syncthreads();
        if (threadIdx.x == 0) base += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = base;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int check_params(const fpl_v2o_params *p, int64_t Z, int64_t Y, int64_t X) {
    FPL_REQUIRE(p != nullptr, "voxel2obj: params is NULL");
    FPL_REQUIRE(Z > 0 && Y > 0 && X > 0, "voxel2obj: empty volume (%lld,%lld,%lld)", (long long)Z,
                (long long)Y, (long long)X);
    FPL_REQUIRE(p->obj_min_dist >= 0 && p->obj_min_dist <= 64, "voxel2obj: obj_min_dist must be in 0..64");
    FPL_REQUIRE(p->lw >= -1 && p->lw <= kMaxLw, "voxel2obj: Gaussian half width %d unsupported (max %d)",
                p->lw, kMaxLw);
    FPL_REQUIRE(p->lw < 0 || p->h_weights != nullptr, "voxel2obj: weights missing");
    return FPL_OK;
}

static int smooth_impl(fpl_ctx *ctx, const float *d_pred, int64_t Z, int64_t Y, int64_t X,
                       const fpl_v2o_params *p, float *d_smooth, float *d_tmp, cudaStream_t st,
                       bool force_generic) {
    const long long n = Z * Y * X;
    const int r = p->obj_min_dist;
    if (r == 0) {   // the reference's  pred[-0:,:,:] = 0  clears the whole map
        FPL_CUDA_CHECK(cudaMemsetAsync(d_smooth, 0, sizeof(float) * n, st));
        return FPL_OK;
    }
    if (p->lw < 0) {
        FPL_CUDA_CHECK(cudaMemcpyAsync(d_smooth, d_pred, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
        return FPL_OK;
    }
    fpl::ProfScope prof(ctx, st, fpl::PROF_GAUSS, 3.0 * 8.0 * (double)n);
    Taps taps;
    memset(&taps, 0, sizeof(taps));
    for (int i = 0; i < 2 * p->lw + 1; ++i) taps.w[i] = p->h_weights[i];
    // axis 0 (z): (1, Z, Y*X)   pred -> smooth
    FPL_TRY(launch_pass(ctx, d_pred, d_smooth, 1, Z, Y * X, r, p->lw, taps, st, force_generic));
    // axis 1 (y): (Z, Y, X)     smooth -> tmp
    FPL_TRY(launch_pass(ctx, d_smooth, d_tmp, Z, Y, X, r, p->lw, taps, st, force_generic));
    // axis 2 (x): (Z*Y, X, 1)   tmp -> smooth
    FPL_TRY(launch_pass(ctx, d_tmp, d_smooth, Z * Y, X, 1, r, p->lw, taps, st, force_generic));
    return FPL_OK;
}

static int select_rank(fpl_ctx *ctx, const float *d_v, long long n, unsigned long long extra_zeros,
                       unsigned long long rank, SelectState *d_state, cudaStream_t st) {
    int blocks = ctx->sm_count * 4;
    select_init_kernel<<<1, 256, 0, st>>>(d_state, rank, extra_zeros);
    FPL_LAUNCH_CHECK(ctx);
    const int shifts[3] = {21, 10, 0};
    const int bins[3] = {1 << kHistBits0, 1 << kHistBits1, 1 << kHistBits2};
    for (int pass = 0; pass < 3; ++pass) {
        select_hist_kernel<<<blocks, 512, 0, st>>>(d_v, n, d_state, shifts[pass], bins[pass], pass == 0);
        FPL_LAUNCH_CHECK(ctx);
        select_scan_kernel<<<1, 1024, 0, st>>>(d_state, shifts[pass], bins[pass]);
        FPL_LAUNCH_CHECK(ctx);
    }
    return FPL_OK;
}

static int threshold_impl(fpl_ctx *ctx, const float *d_smooth, int64_t Z, int64_t Y, int64_t X,
                          const fpl_v2o_params *p, ThreshOut *d_out, SelectState *d_states,
                          cudaStream_t st) {
    const long long n = Z * Y * X;
    const int r = p->obj_min_dist;
    const unsigned long long n_pad = (unsigned long long)(Z + 2 * r) * (Y + 2 * r) * (X + 2 * r);
    FPL_REQUIRE(p->rank_lo >= 0 && (unsigned long long)p->rank_lo < n_pad && p->rank_hi >= 0 &&
                (unsigned long long)p->rank_hi < n_pad, "voxel2obj: percentile ranks out of range");
    const unsigned long long extra = n_pad - (unsigned long long)n;
    fpl::ProfScope prof(ctx, st, fpl::PROF_SELECT, 3.0 * 4.0 * (double)n);
    FPL_TRY(select_rank(ctx, d_smooth, n, extra, (unsigned long long)p->rank_lo, d_states, st));
    SelectState *hi = d_states;
    if (p->rank_hi != p->rank_lo) {
        hi = d_states + 1;
        FPL_TRY(select_rank(ctx, d_smooth, n, extra, (unsigned long long)p->rank_hi, hi, st));
    }
    threshold_kernel<<<1, 32, 0, st>>>(d_states, hi, p->gamma, p->thd, d_out);
    FPL_LAUNCH_CHECK(ctx);
    return FPL_OK;
}

// ------------------------------------------------------------------------------------------------
// fused ("fast") detection path: two dense passes over the smoothed map instead of five
//   level-1 select histogram  : taken by the Gaussian x pass (gauss_contig_kernel) on its way out
//   dense pass 1              : level-2 histogram + 8^3 brick maxima + every 26-neighbourhood local
//                               maximum above the level-1 cut-off (= worklist of the first NMS round:
//                               with nothing suppressed yet, "no better valid neighbour" is exactly
//                               "local maximum", decided once from data that is being streamed anyway)
//   [round 1: ball check + suppression of the local maxima above the level-2 cut-off]
//   dense pass 2              : level-3 histogram + compaction of the not yet suppressed voxels
//                               above the level-2 cut-off (the survivors that later rounds work on)
// The NMS runs on a SUPERSET of the reference's candidates: cut-off c <= threshold (the lower edge of
// the 22-bit radix class that holds the percentile).  Greedy selection in descending (value, -index)
// order reaches every true candidate before any extra element, and an extra element is only selected
// when its ball holds no better valid voxel, so it can only suppress voxels that are not candidates
// either: the selection among true candidates is unchanged, and selected extras (value <= threshold)
// are dropped when the rows are written (finish_rows_kernel).
// ------------------------------------------------------------------------------------------------
// level-1 histogram of the fused path: bin = top 11 bits of the monotone key.  Neighbouring voxels of the
// smoothed map share their top bits, so a thread compares the RAW top bits with its previous voxel (2
// instructions) and only converts / flushes a run when they change.
__global__ void __launch_bounds__(512)
hist1_kernel(const float *__restrict__ v, long long n, SelectState *s) {
    __shared__ unsigned h[2048];
    __shared__ unsigned nan_local;
    for (int b = threadIdx.x; b < 2048; b += blockDim.x) h[b] = 0;
    if (threadIdx.x == 0) nan_local = 0;
    __syncthreads();
    unsigned last = 0xffffffffu, run = 0, my_nan = 0;
    auto flush = [&]() {
        if (run) atomicAdd(&h[(last & 0x400u) ? ((~last) & 0x7ffu) : (last | 0x400u)], run);
    };
    auto feed = [&](float f) {
        const unsigned top = __float_as_uint(f) >> 21;
        if (f != f) ++my_nan;
        if (top == last) { ++run; }
        else { flush(); last = top; run = 1; }
    };
    const long long n4 = ((reinterpret_cast<uintptr_t>(v) & 15) == 0) ? (n >> 2) : 0;
    const float4 *v4 = reinterpret_cast<const float4 *>(v);
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        const float4 q0 = __ldg(v4 + i), q1 = __ldg(v4 + i + stride), q2 = __ldg(v4 + i + 2 * stride), q3 = __ldg(v4 + i + 3 * stride);
        feed(q0.x); feed(q0.y); feed(q0.z); feed(q0.w);
        feed(q1.x); feed(q1.y); feed(q1.z); feed(q1.w);
        feed(q2.x); feed(q2.y); feed(q2.z); feed(q2.w);
        feed(q3.x); feed(q3.y); feed(q3.z); feed(q3.w);
    }
    for (; i < n4; i += stride) {
        const float4 q = __ldg(v4 + i);
        feed(q.x); feed(q.y); feed(q.z); feed(q.w);
    }
    for (long long j = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n; j += stride) feed(__ldg(v + j));
    flush();
    if (my_nan) atomicAdd(&nan_local, my_nan);
    __syncthreads();
    for (int b = threadIdx.x; b < 2048; b += blockDim.x)
        if (h[b]) atomicAdd(&s->hist[blockIdx.x % kHistSlots][b], (unsigned long long)h[b]);
    if (threadIdx.x == 0 && nan_local) atomicAdd(&s->nan_count, (unsigned long long)nan_local);
}

__global__ void select_copy_kernel(const SelectState *src, SelectState *dst) {
    for (int b = threadIdx.x; b < 2048; b += blockDim.x) {       // dst was cleared by select_init_kernel
        unsigned long long c = 0ULL;
        for (int sl = 0; sl < kHistSlots; ++sl) c += src->hist[sl][b];
        dst->hist[0][b] = c;
    }
    if (threadIdx.x == 0) dst->nan_count = src->nan_count;
}

__device__ __forceinline__ unsigned cutoff_key(const SelectState *s) {
    // lower edge of the current prefix class, but never below the smallest positive value
    return s->prefix > 0x80000001u ? s->prefix : 0x80000001u;
}

// (z,y,x) has no better voxel in the rows (z+dz, y+dy), (dz,dy) != (0,0), at x-1..x+1.  "Better" is the
// order of the greedy loop: larger value, or equal value and lower flat index.  All 24 loads are issued
// before the first comparison (one memory latency instead of a chain of dependent ones).
__device__ __forceinline__ bool no_better_in_other_rows(const float *__restrict__ v, Dims d, long long z, long long y,
                                                        long long x, float val) {
    float q[8][3];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
        const int o = nb < 4 ? nb : nb + 1;                      // (dz,dy) in row-major order, skipping (0,0)
        const long long zz = z + (o / 3 - 1), yy = y + (o % 3 - 1);
        const bool rok = zz >= 0 && zz < d.Z && yy >= 0 && yy < d.Y;
        const float *rp = v + ((unsigned long long)(rok ? zz : z) * d.Y + (rok ? yy : y)) * d.X;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const long long xx = x + dx - 1;
            q[nb][dx] = (rok && xx >= 0 && xx < d.X) ? __ldg(rp + xx) : -INFINITY;
        }
    }
    bool ok = true;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
            ok = ok && !(nb < 4 ? (q[nb][dx] >= val) : (q[nb][dx] > val));   // rows before (z,y) hold lower flat indices
    return ok;
}

// Persistent blocks over (bz, by) brick rows (8 z x 8 y rows x all x); warp w owns y offset w, a lane owns 8
// consecutive x per 256-wide chunk for all 8 z (32-byte vector loads, 8 rows in flight).  Stage 1 (unrolled,
// registers only): brick maximum, level-2 histogram, candidates that survive the x-neighbour test.  Stage 2
// (rare, rolled, out of line): the remaining 24 neighbours of those few voxels, re-read through L1/L2.
// FAST: one order statistic whose level-1 class holds non-negative floats (the usual case) -- the class test is
// two instructions on the raw bits; the generic variant handles two statistics and negative classes.
template <bool FAST>
__device__ __forceinline__ void dense_pass1_body(const float *__restrict__ v, Dims d, int gz, int gy, int gx,
                                                 float *__restrict__ g, SelectState *st, int ns,
                                                 unsigned long long *w_idx, float *w_val, long long w_cap,
                                                 Counters *cnt) {
    extern __shared__ unsigned dsm[];
    float *s_max = reinterpret_cast<float *>(dsm);          // [gx]
    unsigned *h2 = dsm + ((gx + 31) & ~31);                 // [ns][2048]
    for (int i = threadIdx.x; i < ns * 2048; i += blockDim.x) h2[i] = 0;
    const unsigned pre0 = st[0].prefix, msk0 = st[0].prefix_mask;
    const unsigned pre1 = ns > 1 ? st[1].prefix : pre0, msk1 = ns > 1 ? st[1].prefix_mask : 0xffffffffu;
    const bool two = ns > 1;
    const float cutoff_f = key2f(cutoff_key(&st[0]));          // > 0: `f >= cutoff_f` == `key(f) >= cutoff`, NaN excluded
    // a prefix class of non-negative floats (key = bits | 0x80000000) can be matched on the raw bits: two
    // instructions per voxel instead of five (these passes are close to instruction-issue bound)
    const bool raw0 = (pre0 >> 31) != 0, raw1 = (pre1 >> 31) != 0;
    const unsigned rpre0 = pre0 & 0x7fffffffu, rpre1 = pre1 & 0x7fffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool vec_ok = (d.X % 4 == 0) && ((reinterpret_cast<uintptr_t>(v) & 15) == 0);
    for (int br = blockIdx.x; br < gz * gy; br += gridDim.x) {
        const int bz = br / gy, by = br % gy;
        for (int i = threadIdx.x; i < gx; i += blockDim.x) s_max[i] = -INFINITY;
        __syncthreads();
        const long long y = (long long)by * kBrick + warp;
        for (long long xc = 0; xc < d.X; xc += 256) {
            const long long x = xc + 8 * lane;
            float f[8][8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const long long z = (long long)bz * kBrick + k;
                const bool rok = z < d.Z && y < d.Y && x < d.X;
                const float *rp = v + (z * d.Y + y) * d.X + x;
                if (rok && vec_ok && x + 8 <= d.X) {
                    const float4 a4 = __ldg(reinterpret_cast<const float4 *>(rp));
                    const float4 b4 = __ldg(reinterpret_cast<const float4 *>(rp) + 1);
                    f[k][0] = a4.x; f[k][1] = a4.y; f[k][2] = a4.z; f[k][3] = a4.w;
                    f[k][4] = b4.x; f[k][5] = b4.y; f[k][6] = b4.z; f[k][7] = b4.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[k][e] = (rok && x + e < d.X) ? __ldg(rp + e) : -INFINITY;
                }
            }
            float m = -INFINITY;
            unsigned long long cmask = 0ULL;
            // x neighbours across the lane boundary (all lanes take part in the shuffles)
            float lf[8], rt[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const long long z = (long long)bz * kBrick + k;
                lf[k] = __shfl_up_sync(0xffffffffu, f[k][7], 1); rt[k] = __shfl_down_sync(0xffffffffu, f[k][0], 1);
                if (lane == 0 || lane == 31) {
                    const bool rok = z < d.Z && y < d.Y;
                    const float *rp = v + (z * d.Y + y) * d.X;
                    if (lane == 0) lf[k] = (rok && x > 0 && x - 1 < d.X) ? __ldg(rp + x - 1) : -INFINITY;
                    else rt[k] = (rok && x + 8 < d.X) ? __ldg(rp + x + 8) : -INFINITY;
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                // branch-free per-voxel part: brick maximum, "in the level-2 class" mask, x-neighbour test mask
                unsigned inb = 0, cx = 0;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float fe = f[k][e];                      // -inf outside the volume
                    m = fmaxf(m, fe);
                    const unsigned bits = __float_as_uint(fe);
                    const bool hit = FAST ? ((bits & msk0) == rpre0)
                                          : ((raw0 ? ((bits & msk0) == rpre0) : ((f2key(fe) & msk0) == pre0)) ||
                                             (two && (raw1 ? ((bits & msk1) == rpre1) : ((f2key(fe) & msk1) == pre1))));
                    inb |= (hit ? 1u : 0u) << e;
                    const float le = e ? f[k][e - 1] : lf[k], re = e < 7 ? f[k][e + 1] : rt[k];
                    // lower flat index wins ties: the left neighbour beats an equal value, the right one does not
                    const bool c = fe >= cutoff_f && fe > le && fe >= re;        // (NaN maps never get here: nan_count)
                    cx |= (c ? 1u : 0u) << e;
                }
                if (inb | cx) {                                    // rare: one branch per row of 8 voxels
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float fe = f[k][e];
                        if (inb & (1u << e)) {
                            const unsigned bits = __float_as_uint(fe);
                            if (FAST) atomicAdd(&h2[(bits >> 10) & 2047u], 1u);
                            else {
                                if (raw0 ? ((bits & msk0) == rpre0) : ((f2key(fe) & msk0) == pre0))
                                    atomicAdd(&h2[((raw0 ? bits : f2key(fe)) >> 10) & 2047u], 1u);
                                if (two && (raw1 ? ((bits & msk1) == rpre1) : ((f2key(fe) & msk1) == pre1)))
                                    atomicAdd(&h2[2048 + (((raw1 ? bits : f2key(fe)) >> 10) & 2047u)], 1u);
                            }
                        }
                        if (cx & (1u << e)) {
                            // the z neighbours inside the brick are in this thread's registers: plane below first (lower index)
                            bool c = true;
                            if (k > 0) {
                                const float a0 = e ? f[k - 1][e - 1] : lf[k - 1], a1 = f[k - 1][e], a2 = e < 7 ? f[k - 1][e + 1] : rt[k - 1];
                                c = !(a0 >= fe) && !(a1 >= fe) && !(a2 >= fe);
                            }
                            if (k < 7) {
                                const float a0 = e ? f[k + 1][e - 1] : lf[k + 1], a1 = f[k + 1][e], a2 = e < 7 ? f[k + 1][e + 1] : rt[k + 1];
                                c = c && !(a0 > fe) && !(a1 > fe) && !(a2 > fe);
                            }
                            if (c) cmask |= 1ULL << (k * 8 + e);
                        }
                    }
                }
            }
            if (x < d.X)
                atomicMax(reinterpret_cast<int *>(&s_max[x / kBrick]), m >= 0.f ? __float_as_int(m) : (int)0x80000000);
#pragma unroll 1
            while (cmask) {
                const int b = __ffsll((long long)cmask) - 1;
                cmask &= cmask - 1ULL;
                const long long z = (long long)bz * kBrick + (b >> 3), xx = x + (b & 7);
                const unsigned long long idx = ((unsigned long long)z * d.Y + y) * d.X + xx;
                const float val = __ldg(v + idx);
                if (no_better_in_other_rows(v, d, z, y, xx, val)) {
                    const unsigned long long pos = atomicAdd(&cnt->n_work, 1ULL);
                    if ((long long)pos < w_cap) { w_idx[pos] = idx; w_val[pos] = val; }
                    else atomicAdd(&cnt->overflow, 1ULL);
                }
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < gx; i += blockDim.x) g[((size_t)bz * gy + by) * gx + i] = s_max[i];
        __syncthreads();
    }
    for (int i = threadIdx.x; i < ns * 2048; i += blockDim.x)
        if (h2[i]) atomicAdd(&st[i >> 11].hist[blockIdx.x % kHistSlots][i & 2047], (unsigned long long)h2[i]);
}

__global__ void __launch_bounds__(256)
dense_pass1_kernel(const float *__restrict__ v, Dims d, int gz, int gy, int gx, float *__restrict__ g,
                   SelectState *st, int ns, unsigned long long *w_idx, float *w_val, long long w_cap,
                   Counters *cnt) {
    if (ns == 1 && (st[0].prefix >> 31)) dense_pass1_body<true>(v, d, gz, gy, gx, g, st, ns, w_idx, w_val, w_cap, cnt);
    else dense_pass1_body<false>(v, d, gz, gy, gx, g, st, ns, w_idx, w_val, w_cap, cnt);
}

// level-3 histogram + compaction of the still valid voxels above the level-2 cut-off (list A).
// A thread takes kCh chunks of 8 consecutive voxels per iteration (8 x 16-byte loads in flight); the block
// reserves its output range with one global atomic per iteration (list order is irrelevant to the result).
__global__ void __launch_bounds__(256)
dense_pass2_kernel(const float *__restrict__ v, long long n, const unsigned *__restrict__ sup, SelectState *st,
                   int ns, unsigned long long *cand_idx, float *cand_val, long long capacity, Counters *cnt) {
    constexpr int kCh = 4;
    __shared__ unsigned h3[2 * 1024];
    for (int i = threadIdx.x; i < ns * 1024; i += blockDim.x) h3[i] = 0;
    __syncthreads();
    const unsigned pre0 = st[0].prefix, msk0 = st[0].prefix_mask;
    const unsigned pre1 = ns > 1 ? st[1].prefix : pre0, msk1 = ns > 1 ? st[1].prefix_mask : msk0;
    const float cutoff_f = key2f(cutoff_key(&st[0]));          // see dense_pass1_kernel
    const bool raw0 = (pre0 >> 31) != 0, raw1 = (pre1 >> 31) != 0, two = ns > 1;
    const unsigned rpre0 = pre0 & 0x7fffffffu, rpre1 = pre1 & 0x7fffffffu;
    const unsigned lane = threadIdx.x & 31;
    const bool aligned = (reinterpret_cast<uintptr_t>(v) & 15) == 0;
    const long long n8 = (n + 7) / 8;
    for (long long c0 = (long long)blockIdx.x * (kCh * 256); c0 < n8; c0 += (long long)gridDim.x * (kCh * 256)) {
        float f[kCh][8];
        unsigned valid[kCh];
#pragma unroll
        for (int k = 0; k < kCh; ++k) {
            const long long c = c0 + k * 256 + threadIdx.x, i0 = c * 8;
            valid[k] = 0;
            if (c < n8 && aligned && i0 + 8 <= n) {
                const float4 a4 = __ldg(reinterpret_cast<const float4 *>(v + i0));
                const float4 b4 = __ldg(reinterpret_cast<const float4 *>(v + i0) + 1);
                f[k][0] = a4.x; f[k][1] = a4.y; f[k][2] = a4.z; f[k][3] = a4.w;
                f[k][4] = b4.x; f[k][5] = b4.y; f[k][6] = b4.z; f[k][7] = b4.w;
                valid[k] = 0xffu;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const bool ok = c < n8 && i0 + j < n;
                    f[k][j] = ok ? __ldg(v + i0 + j) : 0.f;
                    if (ok) valid[k] |= 1u << j;
                }
            }
        }
        unsigned mask[kCh], cnt_me = 0;
#pragma unroll
        for (int k = 0; k < kCh; ++k) {
            const long long i0 = (c0 + k * 256 + threadIdx.x) * 8;
            unsigned m = 0;
            if (valid[k] == 0xffu && raw0 && !two) {           // the common case: 5 instructions per voxel
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float fe = f[k][j];
                    const unsigned bits = __float_as_uint(fe);
                    if ((bits & msk0) == rpre0) atomicAdd(&h3[bits & 1023u], 1u);      // rare (22-bit class)
                    if (fe >= cutoff_f) m |= 1u << j;
                }
            } else {
#pragma unroll 1
                for (int j = 0; j < 8; ++j)
                    if (valid[k] & (1u << j)) {
                        const float fe = __ldg(v + i0 + j);        // re-read: no dynamic indexing of the register array
                        const unsigned bits = __float_as_uint(fe);
                        if (raw0 ? ((bits & msk0) == rpre0) : ((f2key(fe) & msk0) == pre0)) atomicAdd(&h3[(raw0 ? bits : f2key(fe)) & 1023u], 1u);
                        if (two && (raw1 ? ((bits & msk1) == rpre1) : ((f2key(fe) & msk1) == pre1)))
                            atomicAdd(&h3[1024 + ((raw1 ? bits : f2key(fe)) & 1023u)], 1u);
                        if (fe >= cutoff_f) m |= 1u << j;
                    }
            }
            if (m) m &= ~((sup[i0 >> 5] >> (i0 & 31)) & 0xffu);     // i0 is a multiple of 8: one byte of one word
            mask[k] = m;
            cnt_me += __popc(m);
        }
        unsigned pre = cnt_me;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, pre, o);
            if (lane >= (unsigned)o) pre += t;
        }
        const unsigned total = __shfl_sync(0xffffffffu, pre, 31);
        // one global atomic per warp and iteration, and only where there is something to append (list order
        // is irrelevant to the result); no block-level barrier in the streaming loop
        unsigned long long wbase = 0;
        if (total) {
            if (lane == 0) wbase = atomicAdd(&cnt->n_cand, (unsigned long long)total);
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
        }
        unsigned long long pos = wbase + (pre - cnt_me);
#pragma unroll
        for (int k = 0; k < kCh; ++k) {
            const long long i0 = (c0 + k * 256 + threadIdx.x) * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (mask[k] & (1u << j)) {
                    if ((long long)pos < capacity) { cand_idx[pos] = (unsigned long long)(i0 + j); cand_val[pos] = f[k][j]; }
                    else atomicAdd(&cnt->overflow, 1ULL);
                    ++pos;
                }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ns * 1024; i += blockDim.x)
        if (h3[i]) atomicAdd(&st[i >> 10].hist[blockIdx.x % kHistSlots][i & 1023], (unsigned long long)h3[i]);
}

// between round 1 and dense pass 2: the worklist of round 1 is spent, list A starts empty
__global__ void fast_reset_kernel(Counters *cnt) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { cnt->n_cand = 0; cnt->n_next = 0; cnt->n_work = 0; cnt->n_sel_round = 0; }
}

__global__ void select_set_prefix_kernel(SelectState *s, unsigned prefix, unsigned mask) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { s->prefix = prefix; s->prefix_mask = mask; }
}
__global__ void hist_fold_kernel(const SelectState *s, unsigned long long *dst) {
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < 2048; b += gridDim.x * blockDim.x) {
        unsigned long long c = 0ULL;
        for (int sl = 0; sl < kHistSlots; ++sl) c += s->hist[sl][b];
        dst[b] = c;
    }
}
__global__ void idx_to_zyx_kernel(const unsigned long long *idx, long long n, Dims d, long long *zyx) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long q = idx[i];
        zyx[3 * i] = (long long)(q / ((unsigned long long)d.X * d.Y));
        zyx[3 * i + 1] = (long long)((q / (unsigned long long)d.X) % (unsigned long long)d.Y);
        zyx[3 * i + 2] = (long long)(q % (unsigned long long)d.X);
    }
}
__global__ void det_rows_kernel(const unsigned long long *idx, const float *val, long long n, Dims d, double *rows) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long q = idx[i];
        rows[4 * i] = (double)(q / ((unsigned long long)d.X * d.Y));
        rows[4 * i + 1] = (double)((q / (unsigned long long)d.X) % (unsigned long long)d.Y);
        rows[4 * i + 2] = (double)(q % (unsigned long long)d.X);
        rows[4 * i + 3] = (double)val[i];
    }
}

// ---- exchange blocks of the exact multi-GPU rounds: row 0 = (n_selected, alive_owned, overflow), rows 1.. = (z,y,x) of the
// points this rank selected in the round, z global.  One fixed-size block per rank -> ONE all-gather per round, no host
// round trip between decision and suppression.
__global__ void slab_pack_kernel(const unsigned long long *sel_idx, const Counters *cnt, Dims d, long long z_offset,
                                 long long *block, long long sel_cap, int first, unsigned long long n_first) {
    const unsigned long long n = cnt->n_sel_round;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        block[0] = (long long)n;
        block[1] = first ? (long long)cnt->n_work : (long long)cnt->n_alive_owned;      // round 1: owned local maxima
        block[2] = (long long)cnt->overflow + ((long long)n > sel_cap ? 1 : 0);
    }
    const unsigned long long m = n < (unsigned long long)sel_cap ? n : (unsigned long long)sel_cap;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < m; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long q = sel_idx[i];
        block[3 * (1 + i)] = (long long)(q / ((unsigned long long)d.X * d.Y)) + z_offset;
        block[3 * (1 + i) + 1] = (long long)((q / (unsigned long long)d.X) % (unsigned long long)d.Y);
        block[3 * (1 + i) + 2] = (long long)(q % (unsigned long long)d.X);
    }
}

// suppress the balls of the points of all ranks' blocks (blocks are (1 + sel_cap) rows apart); z_offset: global -> slab
__global__ void __launch_bounds__(256)
nms_suppress_blocks_kernel(unsigned *sup, Dims d, int r, const long long *__restrict__ blocks, int n_blocks, long long sel_cap,
                           long long z_offset) {
    const int side = 2 * r + 1;
    for (int b = 0; b < n_blocks; ++b) {
        const long long *blk = blocks + (long long)b * 3 * (1 + sel_cap);
        long long n_pts = blk[0];
        if (n_pts > sel_cap) n_pts = sel_cap;
        for (long long s = blockIdx.x; s < n_pts; s += gridDim.x) {
            const long long z = blk[3 * (1 + s)] - z_offset, y = blk[3 * (1 + s) + 1], x = blk[3 * (1 + s) + 2];
            if (z + r < 0 || z - r >= d.Z) continue;             // ball does not reach this slab
            for (int row = threadIdx.x; row < side * side; row += blockDim.x) {
                int dz = row / side - r, dy = row % side - r;
                int rem = r * r - dz * dz - dy * dy;
                if (rem < 0) continue;
                long long zz = z + dz, yy = y + dy;
                if (zz < 0 || zz >= d.Z || yy < 0 || yy >= d.Y) continue;
                int hw = isqrt_floor(rem);
                long long x0 = x - hw < 0 ? 0 : x - hw;
                long long x1 = x + hw >= d.X ? d.X - 1 : x + hw;
                if (x1 < x0) continue;
                unsigned long long q0 = ((unsigned long long)zz * d.Y + yy) * d.X + x0;
                unsigned long long q1 = ((unsigned long long)zz * d.Y + yy) * d.X + x1;
                unsigned long long *sup64 = reinterpret_cast<unsigned long long *>(sup);
                for (unsigned long long wd = q0 >> 6; wd <= (q1 >> 6); ++wd) {
                    unsigned long long lo = wd << 6;
                    unsigned b0 = q0 > lo ? (unsigned)(q0 - lo) : 0u;
                    unsigned b1 = q1 < lo + 63 ? (unsigned)(q1 - lo) : 63u;
                    unsigned long long mask = (b1 == 63u ? ~0ULL : ((1ULL << (b1 + 1)) - 1ULL)) & ~((1ULL << b0) - 1ULL);
                    atomicOr(&sup64[wd], mask);
                }
            }
        }
    }
}

// warp-per-entry ball check (defined in detect_approx.cuh, which is included further down); EXACT = the reference's order
struct ApproxState;
template <bool EXACT>
__global__ void approx_ballcheck_kernel(const float *__restrict__ v, const unsigned *__restrict__ sup, Dims d, int r,
                                        const float *__restrict__ g, int gy, int gx, const unsigned long long *__restrict__ w_idx,
                                        const float *__restrict__ w_val, unsigned long long *det_idx, float *det_val,
                                        unsigned long long *sel_idx, long long det_capacity, Counters *cnt, ApproxState *S,
                                        unsigned long long *amb_idx, unsigned long long own_lo, unsigned long long own_hi);

struct DetectBuffers {
    unsigned *sup; size_t sup_words;
    unsigned long long *a_idx, *b_idx, *w_idx, *det_idx, *sel_idx, *sidx;
    float *a_val, *b_val, *w_val, *det_val;
    unsigned *skey;
    Counters *cnt;
    unsigned long long *n_out;
    float *grid;                    // brick maxima
    int gz, gy, gx;
    long long list_cap, det_cap;
};

static long long next_pow2(long long v) { long long p = 1; while (p < v) p <<= 1; return p; }

static size_t brick_grid_bytes(int64_t Z, int64_t Y, int64_t X) {
    return (size_t)((Z + kBrick - 1) / kBrick) * ((Y + kBrick - 1) / kBrick) * ((X + kBrick - 1) / kBrick) * 4;
}

static size_t detect_workspace_bytes(int64_t Z, int64_t Y, int64_t X, long long list_cap, long long det_cap) {
    const long long n = Z * Y * X;
    size_t b = 0;
    auto add = [&](size_t x) { b += (x + 255) & ~size_t(255); };
    add(((size_t)n + 63) / 64 * 8);
    add(list_cap * 8); add(list_cap * 8); add(list_cap * 4); add(list_cap * 4);   // A, B
    add(list_cap * 8); add(list_cap * 4);                                          // worklist
    add(det_cap * 8); add(det_cap * 4); add(det_cap * 8);                          // det, sel
    long long sp = next_pow2(det_cap);
    add(sp * 4); add(sp * 8);
    add(sizeof(Counters)); add(64);
    add(brick_grid_bytes(Z, Y, X));
    return b + 4096;
}

static int take_detect_buffers(fpl_ctx *ctx, int64_t Z, int64_t Y, int64_t X, long long list_cap, long long det_cap,
                               DetectBuffers &B, cudaStream_t st) {
    const long long n = Z * Y * X;
    fpl::Arena &A = ctx->arena;
    B.list_cap = list_cap; B.det_cap = det_cap;
    B.sup_words = ((size_t)n + 63) / 64 * 2;          // whole 64-bit words (nms_suppress uses 64-bit atomics)
    B.sup = (unsigned *)A.take(B.sup_words * 4);
    B.a_idx = (unsigned long long *)A.take(list_cap * 8);
    B.b_idx = (unsigned long long *)A.take(list_cap * 8);
    B.a_val = (float *)A.take(list_cap * 4);
    B.b_val = (float *)A.take(list_cap * 4);
    B.w_idx = (unsigned long long *)A.take(list_cap * 8);
    B.w_val = (float *)A.take(list_cap * 4);
    B.det_idx = (unsigned long long *)A.take(det_cap * 8);
    B.det_val = (float *)A.take(det_cap * 4);
    B.sel_idx = (unsigned long long *)A.take(det_cap * 8);
    long long sp = next_pow2(det_cap);
    B.skey = (unsigned *)A.take(sp * 4);
    B.sidx = (unsigned long long *)A.take(sp * 8);
    B.cnt = (Counters *)A.take(sizeof(Counters));
    B.n_out = (unsigned long long *)A.take(64);
    B.gz = (int)((Z + kBrick - 1) / kBrick); B.gy = (int)((Y + kBrick - 1) / kBrick); B.gx = (int)((X + kBrick - 1) / kBrick);
    B.grid = (float *)A.take(brick_grid_bytes(Z, Y, X));
    if (!B.sup || !B.a_idx || !B.b_idx || !B.a_val || !B.b_val || !B.w_idx || !B.w_val || !B.det_idx ||
        !B.det_val || !B.sel_idx || !B.skey || !B.sidx || !B.cnt || !B.n_out || !B.grid) {
        fpl::set_error("voxel2obj: internal workspace sizing error");
        return FPL_ENOMEM;
    }
    FPL_REQUIRE(B.gz < 2048 && B.gy < 1024 && B.gx < 1024 && B.gx * sizeof(float) <= 16384,
                "voxel2obj: volume outside the brick-grid limits");
    FPL_CUDA_CHECK(cudaMemsetAsync(B.sup, 0, B.sup_words * 4, st));
    FPL_CUDA_CHECK(cudaMemsetAsync(B.cnt, 0, sizeof(Counters), st));
    FPL_CUDA_CHECK(cudaMemsetAsync(B.n_out, 0, 64, st));
    return FPL_OK;
}

// NMS rounds on candidate list A (count `remaining`, also in cnt->n_cand) until no valid candidate is left.
// *h_cnt (pinned) holds the counters of the last completed round on return.
static int run_rounds(fpl_ctx *ctx, DetectBuffers &B, const float *d_smooth, Dims d, int r, unsigned long long remaining,
                      long long *rounds, Counters *h_cnt, int64_t capacity, cudaStream_t st) {
    const int grid_stream = ctx->sm_count * 8;
    unsigned long long *a_idx = B.a_idx, *b_idx = B.b_idx;
    float *a_val = B.a_val, *b_val = B.b_val;
    while (remaining > 0) {
        ++*rounds;
        long long fblocks = (long long)((remaining + 255) / 256);
        if (fblocks > grid_stream) fblocks = grid_stream;
        nms_filter_kernel<<<(unsigned)fblocks, 256, 0, st>>>(d_smooth, B.sup, d, a_idx, a_val, b_idx,
                                                            b_val, B.w_idx, B.w_val, B.list_cap, B.cnt);
        FPL_LAUNCH_CHECK(ctx);
        // (block per entry: on dense maps a ball check scans hundreds of bricks -- eight warps with early exit beat the
        // warp-per-entry kernel of the two-tier path, which wins on sparse maps; measured 5.9 vs 7.8 ms on the bench map)
        nms_ballcheck_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(d_smooth, B.sup, d, r, B.grid, B.gz, B.gy, B.gx, B.w_idx, B.w_val,
                                                               B.det_idx, B.det_val, B.sel_idx, B.det_cap,
                                                               B.cnt, nullptr);
        FPL_LAUNCH_CHECK(ctx);
        nms_suppress_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(B.sup, d, r, B.sel_idx, B.cnt);
        FPL_LAUNCH_CHECK(ctx);
        FPL_CUDA_CHECK(cudaMemcpyAsync(h_cnt, B.cnt, sizeof(Counters), cudaMemcpyDeviceToHost, st));
        round_reset_kernel<<<1, 32, 0, st>>>(B.cnt);
        FPL_LAUNCH_CHECK(ctx);
        FPL_CUDA_CHECK(cudaStreamSynchronize(st));
        if (h_cnt->overflow) {
            fpl::set_error("voxel2obj: detection capacity %lld too small (need > %llu)",
                           (long long)capacity, h_cnt->n_det);
            return FPL_EOVERFLOW;
        }
        if (h_cnt->n_sel_round == 0 && h_cnt->n_next > 0) {
            fpl::set_error("voxel2obj: NMS round made no progress (internal error)");
            return FPL_ECUDA;
        }
        // the points selected this round are still in list B (they are only now suppressed);
        // they disappear in the next filter pass.
        remaining = h_cnt->n_next;
        unsigned long long *ti = a_idx; a_idx = b_idx; b_idx = ti;
        float *tv = a_val; a_val = b_val; b_val = tv;
        if (h_cnt->n_next == h_cnt->n_sel_round) remaining = 0;     // everything left was selected this round
    }
    return FPL_OK;
}

// order (conf desc, flat index asc), un-pad, buffer crop, offset -> d_dets; d_tout != nullptr additionally
// drops selected points that are not above the threshold (fast path)
static int finish_detections(fpl_ctx *ctx, DetectBuffers &B, unsigned long long n_det, Dims d, const fpl_v2o_params *p,
                             double *d_dets, int64_t capacity, const ThreshOut *d_tout, unsigned long long *n_rows,
                             cudaStream_t st) {
    *n_rows = 0;
    if (n_det == 0) return FPL_OK;
    const int grid_stream = ctx->sm_count * 8;
    long long np2 = next_pow2((long long)n_det);
    int sblocks = (int)((np2 + 255) / 256); if (sblocks > grid_stream) sblocks = grid_stream;
    sort_prepare_kernel<<<sblocks, 256, 0, st>>>(B.det_idx, B.det_val, (long long)n_det, np2, B.skey, B.sidx);
    FPL_LAUNCH_CHECK(ctx);
    // bitonic network: everything up to k = 2048 in one launch (chunks in shared memory); for larger k the steps with
    // j >= 2048 run on global memory, the rest of that k again in one shared-memory launch
    int lblocks = (int)((np2 + kBitonicChunk - 1) / kBitonicChunk); if (lblocks > grid_stream) lblocks = grid_stream;
    bitonic_local_kernel<<<lblocks, 1024, 0, st>>>(B.skey, B.sidx, np2, 2, np2 < kBitonicChunk ? np2 : kBitonicChunk);
    FPL_LAUNCH_CHECK(ctx);
    for (long long k = 2 * kBitonicChunk; k <= np2; k <<= 1) {
        for (long long j = k >> 1; j >= kBitonicChunk; j >>= 1) {
            bitonic_step_kernel<<<sblocks, 256, 0, st>>>(B.skey, B.sidx, np2, j, k);
            FPL_LAUNCH_CHECK(ctx);
        }
        bitonic_local_kernel<<<lblocks, 1024, 0, st>>>(B.skey, B.sidx, np2, k, k);
        FPL_LAUNCH_CHECK(ctx);
    }
    finish_rows_kernel<<<1, 1024, 0, st>>>(B.skey, B.sidx, (long long)n_det, d, p->buffer_xyz[0],
                                           p->buffer_xyz[1], p->buffer_xyz[2], p->offset_xyz[0],
                                           p->offset_xyz[1], p->offset_xyz[2], d_dets, capacity,
                                           B.n_out, &B.cnt->overflow, d_tout);
    FPL_LAUNCH_CHECK(ctx);
    unsigned long long *h_n = (unsigned long long *)((char *)ctx->h_pinned + 1024);
    FPL_CUDA_CHECK(cudaMemcpyAsync(h_n, B.n_out, 8, cudaMemcpyDeviceToHost, st));
    FPL_CUDA_CHECK(cudaStreamSynchronize(st));
    *n_rows = *h_n;
    return FPL_OK;
}

// classic path: threshold known; candidates = smooth > threshold, compacted by a dense pass
static int detect_impl(fpl_ctx *ctx, const float *d_smooth, int64_t Z, int64_t Y, int64_t X,
                       const fpl_v2o_params *p, double threshold, long long cand_cap, double *d_dets,
                       int64_t capacity, int64_t *h_count, int64_t *h_stats, cudaStream_t st) {
    const long long n = Z * Y * X;
    const int r = p->obj_min_dist;
    Dims d{Z, Y, X};
    long long det_cap = capacity > 0 ? capacity : 1;
    DetectBuffers B;
    FPL_TRY(take_detect_buffers(ctx, Z, Y, X, cand_cap, det_cap, B, st));

    const int grid_stream = ctx->sm_count * 8;
    fpl::ProfScope prof(ctx, st, fpl::PROF_NMS, 4.0 * (double)n);
    compact_candidates_kernel<<<grid_stream, 256, 0, st>>>(d_smooth, n, threshold, B.a_idx, B.a_val,
                                                          cand_cap, B.cnt);
    FPL_LAUNCH_CHECK(ctx);
    FPL_REQUIRE(r <= 27 + 4, "voxel2obj: radius outside the brick-grid limits");
    brick_max_kernel<<<B.gz * B.gy, 256, B.gx * sizeof(float), st>>>(d_smooth, d, B.gy, B.gx, B.grid);
    FPL_LAUNCH_CHECK(ctx);
    Counters *h_cnt = (Counters *)ctx->h_pinned;
    FPL_CUDA_CHECK(cudaMemcpyAsync(h_cnt, B.cnt, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    FPL_CUDA_CHECK(cudaStreamSynchronize(st));
    if (h_cnt->overflow) {
        fpl::set_error("voxel2obj: candidate list overflow (%llu > %lld)", h_cnt->n_cand, cand_cap);
        return FPL_EOVERFLOW;
    }
    const unsigned long long n_candidates = h_cnt->n_cand;
    long long rounds = 0;
    FPL_TRY(run_rounds(ctx, B, d_smooth, d, r, n_candidates, &rounds, h_cnt, capacity, st));
    const unsigned long long n_det = h_cnt->n_det * (rounds > 0 ? 1 : 0);
    unsigned long long n_rows = 0;
    FPL_TRY(finish_detections(ctx, B, n_det, d, p, d_dets, capacity, nullptr, &n_rows, st));
    if (h_count) *h_count = (int64_t)n_rows;
    if (h_stats) {
        h_stats[0] = (int64_t)n_candidates;
        h_stats[1] = rounds;
        h_stats[2] = (int64_t)h_cnt->ball_checks;
        h_stats[3] = (int64_t)n_det;
        h_stats[4] = h_stats[5] = h_stats[6] = h_stats[7] = 0;
    }
    return FPL_OK;
}

// fused path of fpl_voxel2obj (see the comment above select_copy_kernel).  Returns FPL_OK with *done = false
// when the map does not suit it (candidate superset too large: massive ties / saturated maps); the caller
// then runs the classic path on the smoothed map.
struct FastHost {                  // pinned scratch layout
    Counters cnt;
    unsigned long long n_above[2], n_bin[2], nan_count;
    ThreshOut tout;
};

__global__ void fast_collect_kernel(const SelectState *st, int ns, const Counters *cnt, FastHost *out) {
    if (threadIdx.x || blockIdx.x) return;
    out->cnt = *cnt;
    for (int i = 0; i < 2; ++i) { out->n_above[i] = st[i < ns ? i : 0].n_above; out->n_bin[i] = st[i < ns ? i : 0].n_bin; }
    out->nan_count = st[0].nan_count;
}

static int voxel2obj_fast(fpl_ctx *ctx, const float *d_smooth, int64_t Z, int64_t Y, int64_t X,
                          const fpl_v2o_params *p, SelectState *d_states, bool level1_done, ThreshOut *d_tout,
                          long long list_cap, double *d_dets, int64_t capacity, int64_t *h_count,
                          double *h_threshold, int64_t *h_stats, cudaStream_t st, bool *done) {
    *done = false;
    const long long n = Z * Y * X;
    const int r = p->obj_min_dist;
    Dims d{Z, Y, X};
    const int ns = p->rank_hi != p->rank_lo ? 2 : 1;
    const int grid_stream = ctx->sm_count * 8;
    long long det_cap = capacity > 0 ? capacity : 1;
    DetectBuffers B;
    FPL_TRY(take_detect_buffers(ctx, Z, Y, X, list_cap, det_cap, B, st));
    FPL_REQUIRE(r <= 27 + 4, "voxel2obj: radius outside the brick-grid limits");
    FastHost *h = (FastHost *)ctx->h_pinned;
    FastHost *d_host = (FastHost *)ctx->arena.take(sizeof(FastHost));
    FPL_REQUIRE(d_host != nullptr && sizeof(FastHost) <= 1024, "voxel2obj: workspace sizing error");
    long long rounds = 0;
    {
        fpl::ProfScope prof(ctx, st, fpl::PROF_SELECT, 4.0 * (double)n);
        if (!level1_done) {                       // Gaussian ran through the generic kernel / sigma == 0
            hist1_kernel<<<ctx->sm_count * 4, 512, 0, st>>>(d_smooth, n, &d_states[0]);
            FPL_LAUNCH_CHECK(ctx);
        }
        if (ns > 1) { select_copy_kernel<<<1, 256, 0, st>>>(&d_states[0], &d_states[1]); FPL_LAUNCH_CHECK(ctx); }
        for (int i = 0; i < ns; ++i) { select_scan_kernel<<<1, 1024, 0, st>>>(&d_states[i], 21, 2048); FPL_LAUNCH_CHECK(ctx); }
        const size_t sm1 = (size_t)(((B.gx + 31) & ~31) + ns * 2048) * 4;
        int g1 = ctx->sm_count * 2; if (g1 > B.gz * B.gy) g1 = B.gz * B.gy;
        dense_pass1_kernel<<<g1, 256, sm1, st>>>(d_smooth, d, B.gz, B.gy, B.gx, B.grid, d_states, ns, B.w_idx, B.w_val,
                                                B.list_cap, B.cnt);
        FPL_LAUNCH_CHECK(ctx);
        for (int i = 0; i < ns; ++i) { select_scan_kernel<<<1, 1024, 0, st>>>(&d_states[i], 10, 2048); FPL_LAUNCH_CHECK(ctx); }
        fast_collect_kernel<<<1, 32, 0, st>>>(d_states, ns, B.cnt, d_host);
        FPL_LAUNCH_CHECK(ctx);
        FPL_CUDA_CHECK(cudaMemcpyAsync(h, d_host, sizeof(FastHost), cudaMemcpyDeviceToHost, st));
        FPL_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    // every voxel at or above the level-2 cut-off (class of rank_lo) is a potential list entry
    const unsigned long long superset = h->n_above[0] + h->n_bin[0];
    if (h->nan_count == 0 && (h->cnt.overflow || superset > (unsigned long long)list_cap)) return FPL_OK;   // -> classic path
    if (h->nan_count) {
        // np.percentile of a map with NaNs is NaN and `pred > nan` selects nothing (fplobjdetect.py:183-185)
        if (h_count) *h_count = 0;
        if (h_threshold) *h_threshold = nan("");
        if (h_stats) for (int i = 0; i < 8; ++i) h_stats[i] = 0;
        *done = true;
        return FPL_OK;
    }
    unsigned long long n_first = 0;
    {
        fpl::ProfScope prof(ctx, st, fpl::PROF_NMS, 4.0 * (double)n);
        // round 1: the local maxima are the worklist
        ++rounds;
        nms_ballcheck_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(d_smooth, B.sup, d, r, B.grid, B.gz, B.gy, B.gx, B.w_idx, B.w_val,
                                                               B.det_idx, B.det_val, B.sel_idx, B.det_cap, B.cnt, &d_states[0]);
        FPL_LAUNCH_CHECK(ctx);
        nms_suppress_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(B.sup, d, r, B.sel_idx, B.cnt);
        FPL_LAUNCH_CHECK(ctx);
        fast_reset_kernel<<<1, 32, 0, st>>>(B.cnt);
        FPL_LAUNCH_CHECK(ctx);
        dense_pass2_kernel<<<grid_stream, 256, 0, st>>>(d_smooth, n, B.sup, d_states, ns, B.a_idx, B.a_val, B.list_cap, B.cnt);
        FPL_LAUNCH_CHECK(ctx);
        for (int i = 0; i < ns; ++i) { select_scan_kernel<<<1, 1024, 0, st>>>(&d_states[i], 0, 1024); FPL_LAUNCH_CHECK(ctx); }
        threshold_kernel<<<1, 32, 0, st>>>(&d_states[0], &d_states[ns - 1], p->gamma, p->thd, d_tout);
        FPL_LAUNCH_CHECK(ctx);
        fast_collect_kernel<<<1, 32, 0, st>>>(d_states, ns, B.cnt, d_host);
        FPL_LAUNCH_CHECK(ctx);
        FPL_CUDA_CHECK(cudaMemcpyAsync(h, d_host, sizeof(FastHost), cudaMemcpyDeviceToHost, st));
        FPL_CUDA_CHECK(cudaMemcpyAsync(&h->tout, d_tout, sizeof(ThreshOut), cudaMemcpyDeviceToHost, st));
        FPL_CUDA_CHECK(cudaStreamSynchronize(st));
        if (h->cnt.overflow) {
            fpl::set_error("voxel2obj: detection / candidate capacity too small (detections %llu, capacity %lld)",
                           h->cnt.n_det, (long long)capacity);
            return FPL_EOVERFLOW;
        }
        n_first = h->cnt.n_det;
        const double threshold = h->tout.thresh;
        const unsigned long long n_gt = h->n_above[ns - 1];       // values above the order statistic (stats only)
        const unsigned long long ball_first = h->cnt.ball_checks;
        if (h_threshold) *h_threshold = threshold;
        Counters *h_cnt = &h->cnt;
        Counters last = *h_cnt;
        if (h_cnt->n_cand > 0) {
            FPL_TRY(run_rounds(ctx, B, d_smooth, d, r, h_cnt->n_cand, &rounds, h_cnt, capacity, st));
            last = *h_cnt;
        }
        const unsigned long long n_det = last.n_det;
        unsigned long long n_rows = 0;
        FPL_TRY(finish_detections(ctx, B, n_det, d, p, d_dets, capacity, d_tout, &n_rows, st));
        if (h_count) *h_count = (int64_t)n_rows;
        if (h_stats) {
            h_stats[0] = (int64_t)n_gt;
            h_stats[1] = rounds;
            h_stats[2] = (int64_t)(last.ball_checks > ball_first ? last.ball_checks : ball_first);
            h_stats[3] = (int64_t)n_det;
            h_stats[4] = (int64_t)superset; h_stats[5] = (int64_t)n_first; h_stats[6] = 1; h_stats[7] = 0;
        }
    }
    *done = true;
    return FPL_OK;
}

#include "detect_approx.cuh"

// ------------------------------------------------------------------------------------------------
// slab sessions: building blocks of the exact multi-GPU voxel2obj (semantics S2).  A rank holds the
// smoothed map of the planes it owns plus an r-wide halo; it decides only for owned voxels, and after every
// round the points selected by ALL ranks are applied to its suppression map -- so its validity map equals the
// single-GPU one on its extended slab at the start of every round, and the union of the selections is the
// single-GPU result.  Collectives (halo exchange, histogram all-reduce, per-round all-gather of the selected
// points) are done by the host program (flypylib_b200/multi_gpu.py) between these calls.
// ------------------------------------------------------------------------------------------------
struct SlabSession {
    fpl_ctx *ctx;
    const float *smooth;            // extended slab (Ze,Y,X), caller-owned
    Dims d; int r;
    unsigned long long own_lo, own_hi;
    DetectBuffers B;
    void *mem; size_t mem_cap;      // one device block behind all buffers (recycled through fpl_ctx::slab_cache)
    unsigned long long *a_idx, *b_idx; float *a_val, *b_val;
    unsigned long long remaining;   // list A entries (from round 2 on)
    long long rounds;
    ApproxState *state;             // cut-off of the dense pass / the compaction (exact: smallest float above the threshold)
    bool first;                     // round 1 works on the local-maximum worklist of the dense pass
    unsigned long long n_first;     // its size
};

}  // namespace v2o
}  // namespace fpl

using namespace fpl::v2o;

extern "C" {

int fpl_v2o_hist_level(fpl_ctx *ctx, const float *d_v, int64_t n, uint32_t prefix, uint32_t prefix_mask, int shift,
                       int bins, uint64_t *d_hist, uint64_t *d_nan, void *stream) {
    FPL_REQUIRE(ctx && d_v && d_hist && n >= 0 && bins > 0 && bins <= 2048, "fpl_v2o_hist_level: bad argument");
    FPL_CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (ctx->arena.cap < sizeof(SelectState) + 4096) FPL_CUDA_CHECK(cudaStreamSynchronize(st));      // (the arena is about to be re-allocated)
    FPL_TRY(ctx->arena.reserve(sizeof(SelectState) + 4096));
    ctx->arena.reset();
    SelectState *s = (SelectState *)ctx->arena.take(sizeof(SelectState));
    select_init_kernel<<<1, 256, 0, st>>>(s, 0ULL, 0ULL);
    FPL_LAUNCH_CHECK(ctx);
    select_set_prefix_kernel<<<1, 32, 0, st>>>(s, prefix, prefix_mask);      // the class to histogram
    FPL_LAUNCH_CHECK(ctx);
    if (n > 0) {
        select_hist_kernel<<<ctx->sm_count * 4, 512, 0, st>>>(d_v, n, s, shift, bins, d_nan != nullptr);
        FPL_LAUNCH_CHECK(ctx);
    }
    hist_fold_kernel<<<8, 256, 0, st>>>(s, (unsigned long long *)d_hist);      // sum of the partial histograms
    FPL_LAUNCH_CHECK(ctx);
    if (d_nan) FPL_CUDA_CHECK(cudaMemcpyAsync(d_nan, &s->nan_count, sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    return FPL_OK;                  // stream-ordered: the caller's read of d_hist synchronises
}

int fpl_v2o_slab_begin(fpl_ctx *ctx, const float *d_smooth_ext, int64_t Ze, int64_t Y, int64_t X,
                       const fpl_v2o_params *p, double threshold, int64_t own_lo, int64_t own_hi, int64_t list_cap,
                       int64_t det_cap, void **session, int64_t *h_n_candidates, void *stream) {
    FPL_REQUIRE(ctx && d_smooth_ext && session && own_lo >= 0 && own_hi >= own_lo && own_hi <= Ze && list_cap > 0 &&
                det_cap > 0, "fpl_v2o_slab_begin: bad argument");
    FPL_TRY(check_params(p, Ze, Y, X));
    FPL_REQUIRE(p->obj_min_dist <= 27 + 4, "voxel2obj: radius outside the brick-grid limits");
    FPL_CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    SlabSession *S = new SlabSession();
    S->ctx = ctx; S->smooth = d_smooth_ext; S->d = Dims{Ze, Y, X}; S->r = p->obj_min_dist;
    S->own_lo = (unsigned long long)own_lo * Y * X; S->own_hi = (unsigned long long)own_hi * Y * X;
    S->rounds = 0; S->mem = nullptr;
    // private arena for the session (several sessions may be alive on one device: single-GPU emulation in tests)
    fpl::Arena saved = ctx->arena;
    ctx->arena = fpl::Arena();
    const size_t want = detect_workspace_bytes(Ze, Y, X, list_cap, det_cap) + sizeof(ApproxState) + 256;
    for (size_t i = 0; i < ctx->slab_cache.size(); ++i)            // blocks of finished sessions are kept for the next one
        if (ctx->slab_cache[i].cap >= want) {
            ctx->arena.base = (char *)ctx->slab_cache[i].p; ctx->arena.cap = ctx->slab_cache[i].cap;
            ctx->slab_cache.erase(ctx->slab_cache.begin() + i);
            break;
        }
    int rc = ctx->arena.reserve(want);
    if (rc == FPL_OK) { ctx->arena.reset(); rc = take_detect_buffers(ctx, Ze, Y, X, list_cap, det_cap, S->B, st); }
    S->mem = ctx->arena.base; S->mem_cap = ctx->arena.cap;
    ctx->arena = saved;
    if (rc != FPL_OK) { if (S->mem) cudaFree(S->mem); delete S; return rc; }
    // One dense pass instead of "compact every candidate, then test each against its 26 neighbours": brick maxima + the
    // owned voxels above the threshold that no 26-neighbour exceeds = the worklist of round 1 (nothing is suppressed
    // yet, so "valid and no better valid neighbour" is "local maximum"; ties are left to the exact ball check).
    // The candidate list for the later rounds is gathered after the first suppression, brick by brick.
    S->state = (ApproxState *)((char *)S->mem + detect_workspace_bytes(Ze, Y, X, list_cap, det_cap));
    ApproxState hs;
    memset(&hs, 0, sizeof(hs));
    {   // candidates are  (double)v > threshold  and  v > 0  (fplobjdetect.py:184, :205): as a float cut-off
        float c = (float)threshold;
        if (threshold != threshold) c = INFINITY;
        else if (!((double)c > threshold)) c = nextafterf(c, INFINITY);
        while ((double)nextafterf(c, -INFINITY) > threshold) c = nextafterf(c, -INFINITY);
        if (!(c > 0.f)) c = 1.401298464324817e-45f;
        hs.cutA = c; hs.Lb = INFINITY; hs.Hb = INFINITY;
    }
    FPL_CUDA_CHECK(cudaMemcpyAsync(S->state, &hs, sizeof(hs), cudaMemcpyHostToDevice, st));
    FPL_CUDA_CHECK(cudaMemsetAsync(S->B.grid, 0, brick_grid_bytes(Ze, Y, X), st));
    {
        const long long items = Ze * ((Y + kP1Strip - 1) / kP1Strip) * ((X + kP1Cols - 1) / kP1Cols);
        long long g1 = (items + 3) / 4;
        if (g1 > (long long)ctx->sm_count * 64) g1 = (long long)ctx->sm_count * 64;
        if ((X % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_smooth_ext) & 15) == 0))
            approx_pass1_kernel<true><<<(unsigned)g1, 128, 0, st>>>(d_smooth_ext, S->d, S->B.gy, S->B.gx, S->B.grid, S->state, nullptr, nullptr, 0,
                                                                    S->B.w_idx, S->B.w_val, S->B.list_cap, S->B.cnt, (int)own_lo, (int)own_hi);
        else
            approx_pass1_kernel<false><<<(unsigned)g1, 128, 0, st>>>(d_smooth_ext, S->d, S->B.gy, S->B.gx, S->B.grid, S->state, nullptr, nullptr, 0,
                                                                     S->B.w_idx, S->B.w_val, S->B.list_cap, S->B.cnt, (int)own_lo, (int)own_hi);
        FPL_LAUNCH_CHECK(ctx);
    }
    // no host round trip here: the size of the worklist (= "alive" of round 1) and a possible overflow travel in the
    // header of the first exchange block (fpl_v2o_slab_round_pack) or are read by fpl_v2o_slab_round
    S->first = true;
    S->n_first = ~0ULL;             // unknown on the host until the first round
    S->remaining = 1;
    S->a_idx = S->B.a_idx; S->b_idx = S->B.b_idx; S->a_val = S->B.a_val; S->b_val = S->B.b_val;
    if (h_n_candidates) *h_n_candidates = -1;
    *session = S;
    return FPL_OK;
}

// one round, decision half: worklist of the owned valid candidates -> ball check -> newly selected points.
// d_sel_zyx receives their (z,y,x) in slab coordinates; *h_alive_owned = valid owned candidates at round start.
int fpl_v2o_slab_round(void *session, int64_t *d_sel_zyx, int64_t sel_cap, int64_t *h_n_sel, int64_t *h_alive_owned,
                       void *stream) {
    SlabSession *S = (SlabSession *)session;
    FPL_REQUIRE(S && d_sel_zyx && h_n_sel && h_alive_owned, "fpl_v2o_slab_round: NULL argument");
    fpl_ctx *ctx = S->ctx;
    FPL_CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    Counters *h_cnt = (Counters *)ctx->h_pinned;
    *h_n_sel = 0; *h_alive_owned = 0;
    const bool first = S->first;
    if (!first && S->rounds == 1) {
        // entering round 2: every rank's round-1 balls are in the suppression map now -> gather the candidates that
        // are still valid (bricks whose maximum reaches the cut-off only), owned planes and halo alike
        fast_reset_kernel<<<1, 32, 0, st>>>(S->B.cnt);
        FPL_LAUNCH_CHECK(ctx);
        approx_compact_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(S->smooth, S->d, S->B.grid, S->B.gz, S->B.gy, S->B.gx, S->B.sup, S->state,
                                                                  S->a_idx, S->a_val, S->B.list_cap, S->B.cnt);
        FPL_LAUNCH_CHECK(ctx);
        FPL_CUDA_CHECK(cudaMemcpyAsync(h_cnt, S->B.cnt, sizeof(Counters), cudaMemcpyDeviceToHost, st));
        FPL_CUDA_CHECK(cudaStreamSynchronize(st));
        if (h_cnt->overflow) { fpl::set_error("voxel2obj slab: candidate list overflow"); return FPL_EOVERFLOW; }
        S->remaining = h_cnt->n_cand;
    }
    if (S->remaining == 0 && !first) { ++S->rounds; return FPL_OK; }
    ++S->rounds;
    if (!first) {
        long long fblocks = (long long)((S->remaining + 255) / 256);
        if (fblocks > ctx->sm_count * 8) fblocks = ctx->sm_count * 8;
        nms_filter_kernel<<<(unsigned)fblocks, 256, 0, st>>>(S->smooth, S->B.sup, S->d, S->a_idx, S->a_val, S->b_idx, S->b_val,
                                                            S->B.w_idx, S->B.w_val, S->B.list_cap, S->B.cnt, S->own_lo, S->own_hi);
        FPL_LAUNCH_CHECK(ctx);
    }
    nms_ballcheck_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(S->smooth, S->B.sup, S->d, S->r, S->B.grid, S->B.gz, S->B.gy, S->B.gx,
                                                           S->B.w_idx, S->B.w_val, S->B.det_idx, S->B.det_val, S->B.sel_idx,
                                                           S->B.det_cap, S->B.cnt, nullptr);
    FPL_LAUNCH_CHECK(ctx);
    FPL_CUDA_CHECK(cudaMemcpyAsync(h_cnt, S->B.cnt, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    FPL_CUDA_CHECK(cudaStreamSynchronize(st));
    if (h_cnt->overflow) { fpl::set_error("voxel2obj slab: detection capacity too small"); return FPL_EOVERFLOW; }
    FPL_REQUIRE((int64_t)h_cnt->n_sel_round <= sel_cap, "fpl_v2o_slab_round: selection buffer too small");
    if (h_cnt->n_sel_round) {
        idx_to_zyx_kernel<<<64, 256, 0, st>>>(S->B.sel_idx, (long long)h_cnt->n_sel_round, S->d, (long long *)d_sel_zyx);
        FPL_LAUNCH_CHECK(ctx);
    }
    *h_n_sel = (int64_t)h_cnt->n_sel_round;
    // round 1: "alive" = owned local maxima above the threshold (zero on every rank <=> there is no candidate at all)
    *h_alive_owned = first ? (int64_t)h_cnt->n_work : (int64_t)h_cnt->n_alive_owned;
    round_reset_kernel<<<1, 32, 0, st>>>(S->B.cnt);
    FPL_LAUNCH_CHECK(ctx);
    if (first) { S->first = false; return FPL_OK; }
    S->remaining = h_cnt->n_next;
    unsigned long long *ti = S->a_idx; S->a_idx = S->b_idx; S->b_idx = ti;
    float *tv = S->a_val; S->a_val = S->b_val; S->b_val = tv;
    return FPL_OK;
}

// one round, update half: suppress the balls of the points selected by all ranks (slab coordinates, may lie outside)
int fpl_v2o_slab_suppress(void *session, const int64_t *d_zyx, int64_t n_pts, void *stream) {
    SlabSession *S = (SlabSession *)session;
    FPL_REQUIRE(S && (d_zyx || n_pts == 0), "fpl_v2o_slab_suppress: NULL argument");
    if (n_pts <= 0) return FPL_OK;
    FPL_CUDA_CHECK(cudaSetDevice(S->ctx->device));
    nms_suppress_zyx_kernel<<<S->ctx->sm_count * 4, 256, 0, (cudaStream_t)stream>>>(S->B.sup, S->d, S->r, (const long long *)d_zyx, n_pts);
    FPL_LAUNCH_CHECK(S->ctx);
    return FPL_OK;
}

// One round without a host round trip: decision half -> exchange block (header + selected points, z global).  The
// caller all-gathers the blocks of all ranks (one collective) and hands them to fpl_v2o_slab_apply_blocks.
int fpl_v2o_slab_round_pack(void *session, int64_t *d_block, int64_t sel_cap, int64_t z_offset, void *stream) {
    SlabSession *S = (SlabSession *)session;
    FPL_REQUIRE(S && d_block && sel_cap > 0, "fpl_v2o_slab_round_pack: bad argument");
    fpl_ctx *ctx = S->ctx;
    FPL_CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const bool first = S->first;
    const int grid_stream = ctx->sm_count * 8;
    if (!first && S->rounds == 1) {         // entering round 2: gather the candidates that survived every rank's round 1
        fast_reset_kernel<<<1, 32, 0, st>>>(S->B.cnt);
        FPL_LAUNCH_CHECK(ctx);
        approx_compact_kernel<<<grid_stream, 256, 0, st>>>(S->smooth, S->d, S->B.grid, S->B.gz, S->B.gy, S->B.gx, S->B.sup, S->state,
                                                       S->a_idx, S->a_val, S->B.list_cap, S->B.cnt);
        FPL_LAUNCH_CHECK(ctx);
    }
    ++S->rounds;
    if (!first) {
        nms_filter_kernel<<<grid_stream, 256, 0, st>>>(S->smooth, S->B.sup, S->d, S->a_idx, S->a_val, S->b_idx, S->b_val,
                                                     S->B.w_idx, S->B.w_val, S->B.list_cap, S->B.cnt, S->own_lo, S->own_hi);
        FPL_LAUNCH_CHECK(ctx);
    }
    nms_ballcheck_kernel<<<grid_stream, 256, 0, st>>>(S->smooth, S->B.sup, S->d, S->r, S->B.grid, S->B.gz, S->B.gy, S->B.gx,
                                                     S->B.w_idx, S->B.w_val, S->B.det_idx, S->B.det_val, S->B.sel_idx,
                                                     S->B.det_cap, S->B.cnt, nullptr);
    FPL_LAUNCH_CHECK(ctx);
    slab_pack_kernel<<<64, 256, 0, st>>>(S->B.sel_idx, S->B.cnt, S->d, (long long)z_offset, (long long *)d_block, (long long)sel_cap,
                                         first ? 1 : 0, S->n_first);
    FPL_LAUNCH_CHECK(ctx);
    round_reset_kernel<<<1, 32, 0, st>>>(S->B.cnt);
    FPL_LAUNCH_CHECK(ctx);
    if (first) S->first = false;
    else {
        unsigned long long *ti = S->a_idx; S->a_idx = S->b_idx; S->b_idx = ti;
        float *tv = S->a_val; S->a_val = S->b_val; S->b_val = tv;
    }
    return FPL_OK;
}

// Update half for the gathered blocks of all ranks (n_blocks x (1 + sel_cap) x 3 int64, z global; z_offset = global z of
// this slab's plane 0).
int fpl_v2o_slab_apply_blocks(void *session, const int64_t *d_blocks, int32_t n_blocks, int64_t sel_cap, int64_t z_offset,
                              void *stream) {
    SlabSession *S = (SlabSession *)session;
    FPL_REQUIRE(S && d_blocks && n_blocks > 0 && sel_cap > 0, "fpl_v2o_slab_apply_blocks: bad argument");
    FPL_CUDA_CHECK(cudaSetDevice(S->ctx->device));
    nms_suppress_blocks_kernel<<<S->ctx->sm_count * 4, 256, 0, (cudaStream_t)stream>>>(S->B.sup, S->d, S->r, (const long long *)d_blocks,
                                                                                       n_blocks, (long long)sel_cap, (long long)z_offset);
    FPL_LAUNCH_CHECK(S->ctx);
    return FPL_OK;
}

// detections of the owned planes: rows (z, y, x, conf) in slab coordinates, unordered
int fpl_v2o_slab_end(void *session, double *d_rows, int64_t capacity, int64_t *h_count, int64_t *h_rounds, void *stream) {
    SlabSession *S = (SlabSession *)session;
    FPL_REQUIRE(S && h_count, "fpl_v2o_slab_end: NULL argument");
    fpl_ctx *ctx = S->ctx;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = FPL_OK;
    Counters *h_cnt = (Counters *)ctx->h_pinned;
    if (cudaMemcpyAsync(h_cnt, S->B.cnt, sizeof(Counters), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) rc = FPL_ECUDA;
    const long long n_det = rc == FPL_OK ? (long long)h_cnt->n_det : 0;
    if (rc == FPL_OK && n_det > capacity) { fpl::set_error("fpl_v2o_slab_end: row buffer too small"); rc = FPL_EOVERFLOW; }
    if (rc == FPL_OK && n_det > 0) {
        det_rows_kernel<<<64, 256, 0, st>>>(S->B.det_idx, S->B.det_val, n_det, S->d, d_rows);
        if (cudaStreamSynchronize(st) != cudaSuccess) rc = FPL_ECUDA;
    }
    *h_count = n_det;
    if (h_rounds) *h_rounds = S->rounds;
    if (rc == FPL_OK && ctx->slab_cache.size() < 4) ctx->slab_cache.push_back(fpl::PoolBuf{S->mem, S->mem_cap, false});
    else cudaFree(S->mem);
    delete S;
    return rc;
}

}  // extern "C"

namespace fpl {
namespace v2o {
}  // namespace v2o
}  // namespace fpl

using namespace fpl::v2o;

static int g_force_generic_gauss = 0;

extern "C" {

// test hook: route every Gaussian pass through the generic one-thread-per-output kernel
int fpl_debug_force_generic_gauss(int on) { g_force_generic_gauss = on; return FPL_OK; }
// test hook: margin of the certified-FMA Gaussian chain in double ulps (default 64; 0 = exact chain only;
// 1 << 28 = every output fails the certificate and is recomputed exactly)
int fpl_debug_gauss_cert(int margin) { fpl::v2o::g_gauss_cert = (unsigned)margin; return FPL_OK; }

int fpl_v2o_smooth(fpl_ctx *ctx, const float *d_pred, int64_t Z, int64_t Y, int64_t X,
                   const fpl_v2o_params *p, float *d_smooth, void *stream) {
    FPL_REQUIRE(ctx && d_pred && d_smooth, "fpl_v2o_smooth: NULL argument");
    FPL_TRY(check_params(p, Z, Y, X));
    FPL_REQUIRE(d_pred != d_smooth, "fpl_v2o_smooth: d_smooth may not alias d_pred");
    FPL_CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)Z * Y * X;
    if (ctx->arena.cap < n * sizeof(float) + 4096) FPL_CUDA_CHECK(cudaStreamSynchronize(st));   // the arena is about to be re-allocated
    FPL_TRY(ctx->arena.reserve(n * sizeof(float) + 4096));
    ctx->arena.reset();
    float *tmp = (float *)ctx->arena.take(n * sizeof(float));
    return smooth_impl(ctx, d_pred, Z, Y, X, p, d_smooth, tmp, st, g_force_generic_gauss != 0);
}

int fpl_v2o_threshold(fpl_ctx *ctx, const float *d_smooth, int64_t Z, int64_t Y, int64_t X,
                      const fpl_v2o_params *p, double *h_out, void *stream) {
    FPL_REQUIRE(ctx && d_smooth && h_out, "fpl_v2o_threshold: NULL argument");
    FPL_TRY(check_params(p, Z, Y, X));
    FPL_CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    FPL_CUDA_CHECK(cudaStreamSynchronize(st));
    FPL_TRY(ctx->arena.reserve(2 * sizeof(SelectState) + sizeof(ThreshOut) + 4096));
    ctx->arena.reset();
    SelectState *states = (SelectState *)ctx->arena.take(2 * sizeof(SelectState));
    ThreshOut *tout = (ThreshOut *)ctx->arena.take(sizeof(ThreshOut));
    FPL_TRY(threshold_impl(ctx, d_smooth, Z, Y, X, p, tout, states, st));
    ThreshOut *h = (ThreshOut *)ctx->h_pinned;
    FPL_CUDA_CHECK(cudaMemcpyAsync(h, tout, sizeof(ThreshOut), cudaMemcpyDeviceToHost, st));
    FPL_CUDA_CHECK(cudaStreamSynchronize(st));
    h_out[0] = h->thresh; h_out[1] = (double)h->v_lo; h_out[2] = (double)h->v_hi;
    h_out[3] = (double)h->nan_count;
    return FPL_OK;
}

static long long candidate_capacity(const fpl_v2o_params *p, int64_t Z, int64_t Y, int64_t X) {
    const int r = p->obj_min_dist;
    const long long n_pad = (long long)(Z + 2 * r) * (Y + 2 * r) * (X + 2 * r);
    // values strictly above the threshold are at most the elements above order statistic rank_lo
    long long cap = n_pad - p->rank_lo;
    const long long n = Z * Y * X;
    if (cap > n) cap = n;
    if (cap < 1) cap = 1;
    return cap;
}

int fpl_v2o_detect(fpl_ctx *ctx, const float *d_smooth, int64_t Z, int64_t Y, int64_t X,
                   const fpl_v2o_params *p, double threshold, double *d_dets, int64_t capacity,
                   int64_t *h_count, int64_t *h_stats, void *stream) {
    FPL_REQUIRE(ctx && d_smooth && (d_dets || capacity == 0), "fpl_v2o_detect: NULL argument");
    FPL_TRY(check_params(p, Z, Y, X));
    FPL_CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    FPL_CUDA_CHECK(cudaStreamSynchronize(st));
    // without a percentile rank the only safe bound is every voxel
    long long cand_cap = (p->rank_lo > 0) ? candidate_capacity(p, Z, Y, X) : (long long)(Z * Y * X);
    FPL_TRY(ctx->arena.reserve(detect_workspace_bytes(Z, Y, X, cand_cap, capacity > 0 ? capacity : 1)));
    ctx->arena.reset();
    return detect_impl(ctx, d_smooth, Z, Y, X, p, threshold, cand_cap, d_dets, capacity, h_count,
                       h_stats, st);
}

// test / diagnosis hook: why the last fpl_voxel2obj on ctx left the two-tier path (codes: detect_approx.cuh; 0 = it ran,
// -1 = skipped by the adaptive back-off); reset_backoff != 0 also clears the back-off state
int fpl_debug_v2o_decline_reason(fpl_ctx *ctx, int reset_backoff) {
    if (!ctx) return -2;
    if (reset_backoff) { ctx->v2o_skip = 0; ctx->v2o_fail_streak = 0; }
    return ctx->v2o_decline;
}

// test / experiment hook: 0 = the sigma-5 Gaussian uses the generic (uniform-register taps) instantiation too
int fpl_debug_gauss_imm(int on) { fpl::v2o::g_gauss_imm = on; return FPL_OK; }

int fpl_debug_v2o_decline_info(fpl_ctx *ctx, long long *out8) {
    if (!ctx || !out8) return -2;
    for (int i = 0; i < 8; ++i) out8[i] = ctx->v2o_info[i];
    return ctx->v2o_decline;
}

static int g_v2o_classic = 0;
// test hook: 0 = default (two-tier path when the map qualifies, else the fused exact path), 1 = classic exact path
// (five dense passes), 2 = fused exact path only (never the two-tier path)
int fpl_debug_v2o_classic(int mode) { g_v2o_classic = mode; return FPL_OK; }

int fpl_voxel2obj(fpl_ctx *ctx, const float *d_pred, int64_t Z, int64_t Y, int64_t X,
                  const fpl_v2o_params *p, double *d_dets, int64_t capacity, int64_t *h_count,
                  double *h_threshold, int64_t *h_stats, void *stream) {
    FPL_REQUIRE(ctx && d_pred && (d_dets || capacity == 0), "fpl_voxel2obj: NULL argument");
    FPL_TRY(check_params(p, Z, Y, X));
    FPL_CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    FPL_CUDA_CHECK(cudaStreamSynchronize(st));
    const size_t n = (size_t)Z * Y * X;
    const int r = p->obj_min_dist;
    const long long cand_cap = candidate_capacity(p, Z, Y, X);
    // fused path: lists hold a superset of the candidates (everything from the lower edge of the percentile's
    // 22-bit radix class upwards) -> some slack on top of the exact candidate bound
    long long list_cap = cand_cap + cand_cap / 8 + 65536;
    if (list_cap > (long long)n) list_cap = (long long)n;
    const bool try_fast = g_v2o_classic != 1 && r > 0;
    const bool try_approx = g_v2o_classic == 0 && r > 0;
    size_t need = 2 * (n * sizeof(float) + 512) + 2 * sizeof(SelectState) + sizeof(ThreshOut) + 4096 + 2048 +
                  detect_workspace_bytes(Z, Y, X, try_fast ? list_cap : cand_cap, capacity > 0 ? capacity : 1);
    if (try_approx) need += (size_t)n / 16 + ((size_t)1 << 24) + ((size_t)128 << 20);     // lattice sample, band / narrow / ambiguity lists, exact-value scratch
    FPL_TRY(ctx->arena.reserve(need));
    ctx->arena.reset();
    // the back-off belongs to one kind of map: a call on a volume of another shape / with other parameters starts afresh
    if (ctx->v2o_key[0] != Z || ctx->v2o_key[1] != Y || ctx->v2o_key[2] != X || ctx->v2o_key[3] != r || ctx->v2o_key[4] != p->lw) {
        ctx->v2o_key[0] = Z; ctx->v2o_key[1] = Y; ctx->v2o_key[2] = X; ctx->v2o_key[3] = r; ctx->v2o_key[4] = p->lw;
        ctx->v2o_skip = 0; ctx->v2o_fail_streak = 0;
    }
    if (try_approx && ctx->v2o_skip > 0) {
        // adaptive: the last call(s) on this context did not qualify (a kind of map whose certificates fail, e.g. a value
        // distribution so tight that too many voxels sit within the bound of the percentile) -- do not pay for the
        // attempt every time; retry after 16, 32, 64, 64, .. calls
        --ctx->v2o_skip;
        ctx->v2o_decline = -1;
    } else if (try_approx) {
        // two-tier path: fp32 smoothing with a proven bound, exact values only where a decision needs them
        SelectState *st2 = (SelectState *)ctx->arena.take(2 * sizeof(SelectState));
        ThreshOut *tout2 = (ThreshOut *)ctx->arena.take(sizeof(ThreshOut));
        FPL_REQUIRE(st2 && tout2, "voxel2obj: workspace sizing error");
        bool done = false;
        FPL_TRY(voxel2obj_approx(ctx, d_pred, Z, Y, X, p, st2, tout2, list_cap, d_dets, capacity, h_count, h_threshold,
                                 h_stats, st, &done));
        if (done) { ctx->v2o_fail_streak = 0; return FPL_OK; }
        FPL_CUDA_CHECK(cudaStreamSynchronize(st));
        ctx->arena.reset();                 // did not qualify / certificate failed: exact path from scratch
        if (ctx->v2o_decline > 5) {         // a data-dependent decline cost a Gaussian pass: back off
            if (ctx->v2o_fail_streak < 3) ++ctx->v2o_fail_streak;
            ctx->v2o_skip = 8 << ctx->v2o_fail_streak;          // 16, 32, 64, 64, .. calls go straight to the exact path
        }
    }
    float *smooth = (float *)ctx->arena.take(n * sizeof(float));
    float *tmp = (float *)ctx->arena.take(n * sizeof(float));
    SelectState *states = (SelectState *)ctx->arena.take(2 * sizeof(SelectState));
    ThreshOut *tout = (ThreshOut *)ctx->arena.take(sizeof(ThreshOut));
    if (try_fast) {
        const unsigned long long n_pad = (unsigned long long)(Z + 2 * r) * (Y + 2 * r) * (X + 2 * r);
        FPL_REQUIRE(p->rank_lo >= 0 && (unsigned long long)p->rank_lo < n_pad && p->rank_hi >= 0 &&
                    (unsigned long long)p->rank_hi < n_pad, "voxel2obj: percentile ranks out of range");
        select_init_kernel<<<1, 256, 0, st>>>(&states[0], (unsigned long long)p->rank_lo, n_pad - n);
        FPL_LAUNCH_CHECK(ctx);
        select_init_kernel<<<1, 256, 0, st>>>(&states[1], (unsigned long long)p->rank_hi, n_pad - n);
        FPL_LAUNCH_CHECK(ctx);
        FPL_TRY(smooth_impl(ctx, d_pred, Z, Y, X, p, smooth, tmp, st, g_force_generic_gauss != 0));
        const size_t mark = ctx->arena.used;
        bool done = false;
        FPL_TRY(voxel2obj_fast(ctx, smooth, Z, Y, X, p, states, false, tout, list_cap, d_dets, capacity, h_count,
                               h_threshold, h_stats, st, &done));
        if (done) return FPL_OK;
        ctx->arena.used = mark;             // superset too large for the lists: classic path on the smoothed map
    } else {
        FPL_TRY(smooth_impl(ctx, d_pred, Z, Y, X, p, smooth, tmp, st, g_force_generic_gauss != 0));
    }
    FPL_TRY(threshold_impl(ctx, smooth, Z, Y, X, p, tout, states, st));
    ThreshOut *h = (ThreshOut *)((char *)ctx->h_pinned + 2048);
    FPL_CUDA_CHECK(cudaMemcpyAsync(h, tout, sizeof(ThreshOut), cudaMemcpyDeviceToHost, st));
    FPL_CUDA_CHECK(cudaStreamSynchronize(st));
    const double threshold = h->thresh;
    if (h_threshold) *h_threshold = threshold;
    return detect_impl(ctx, smooth, Z, Y, X, p, threshold, cand_cap, d_dets, capacity, h_count, h_stats, st);
}

}  // extern "C"
