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

namespace fpl {
namespace v2o {

constexpr int kMaxLw = 32;                 // sigma <= 16 (truncate 2.0)
struct Taps { double w[2 * kMaxLw + 1]; }; // correlate-order taps, centre at w[lw]

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
constexpr int kRun = 8;
constexpr int kRunsPerThread = 4;
constexpr int kTN = kGroups * kRunsPerThread * kRun;   // 64 outputs along the axis per block

template <int LW>
__device__ __forceinline__ void run_from_window(const double (&x)[kRun + 2 * LW], const Taps &taps,
                                                float (&res)[kRun]) {
#pragma unroll
    for (int q = 0; q < kRun; ++q) {
        const int c = q + LW;
        double tmp = __dmul_rn(x[c], taps.w[LW]);
#pragma unroll
        for (int j = -LW; j < 0; ++j)
            tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(x[c + j], x[c - j]), taps.w[LW + j]));
        res[q] = (float)tmp;
    }
}

// filtered axis has stride `inner` (> 1 in practice: z and y passes); lines are contiguous in memory
template <int LW>
__global__ void __launch_bounds__(kLines * kGroups)
gauss_strided_kernel(const float *__restrict__ in, float *__restrict__ out, long long n,
                     long long inner, int r, Taps taps) {
    __shared__ float tile[kTN + 2 * LW][kLines];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const unsigned nchunks = (unsigned)((n + kTN - 1) / kTN);
    const long long n0 = (long long)(blockIdx.x % nchunks) * kTN;      // n-chunk fastest: blocks that
    const long long i = (long long)(blockIdx.x / nchunks) * kLines + tx; // share halo rows are co-resident
    const long long o = blockIdx.y;
    const float *src = in + o * n * inner;
    float *dst = out + o * n * inner;
    const bool live = i < inner;
    {
        // stage the (kTN + 2 LW) x 64 source tile: loads are issued in batches of kBatch rows before any is
        // stored, so that kBatch global loads per thread are in flight (the pass is latency bound otherwise)
        constexpr int kRowsT = (kTN + 2 * LW + kGroups - 1) / kGroups;
        constexpr int kBatch = 14;
        // no reflection can occur when the tile (with halo) stays inside the zero-padded line
        const bool simple = (n0 - LW + r >= 0) && (n0 + kTN + LW - 1 < n + r);
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
        run_from_window<LW>(x, taps, res);
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
                    long long n, int r, Taps taps) {
    constexpr int kPitch = kLines + 1;
    __shared__ float tile[(kTN + 2 * LW) * kPitch];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * kLines + tx;
    const unsigned nchunks = (unsigned)((n + kTN - 1) / kTN);
    const long long n0 = (long long)(blockIdx.x % nchunks) * kTN;
    const long long row0 = (long long)(blockIdx.x / nchunks) * kLines;
    constexpr int kW = kTN + 2 * LW;
    {
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
        run_from_window<LW>(x, taps, res[run]);
    }
    __syncthreads();
#pragma unroll
    for (int run = 0; run < kRunsPerThread; ++run) {
        const int base = (ty * kRunsPerThread + run) * kRun;
#pragma unroll
        for (int q = 0; q < kRun; ++q) tile[(base + q) * kPitch + tx] = res[run][q];
    }
    __syncthreads();
    for (int idx = tid; idx < kLines * kTN; idx += kLines * kGroups) {
        int rr = idx / kTN, xx = idx - rr * kTN;
        if (row0 + rr < rows && n0 + xx < n) out[(row0 + rr) * n + n0 + xx] = tile[xx * kPitch + rr];
    }
}

template <int LW>
static int launch_pass_fast(fpl_ctx *ctx, const float *in, float *out, long long outer, long long n,
                            long long inner, int r, const Taps &taps, cudaStream_t st) {
    dim3 block(kLines, kGroups);
    const long long nchunks = (n + kTN - 1) / kTN;
    if (inner == 1) {
        const long long rows = outer;
        const long long gx = nchunks * ((rows + kLines - 1) / kLines);
        FPL_REQUIRE(gx < 2147483647LL, "gauss pass: volume too large for one launch");
        gauss_contig_kernel<LW><<<dim3((unsigned)gx), block, 0, st>>>(in, out, rows, n, r, taps);
        FPL_LAUNCH_CHECK(ctx);
        return FPL_OK;
    }
    const long long gx = nchunks * ((inner + kLines - 1) / kLines);
    FPL_REQUIRE(gx < 2147483647LL && outer <= 65535, "gauss pass: volume too large for one launch");
    gauss_strided_kernel<LW><<<dim3((unsigned)gx, (unsigned)outer), block, 0, st>>>(in, out, n, inner, r, taps);
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
    unsigned long long hist[2048];
};

constexpr int kHistBits0 = 11, kHistBits1 = 11, kHistBits2 = 10;

__global__ void select_init_kernel(SelectState *s, unsigned long long rank,
                                   unsigned long long extra_zeros) {
    int t = threadIdx.x;
    for (int b = t; b < 2048; b += blockDim.x) s->hist[b] = 0;
    if (t == 0) { s->rank = rank; s->prefix = 0; s->prefix_mask = 0; s->extra_zeros = extra_zeros; s->nan_count = 0; }
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
        if (h[b]) atomicAdd(&s->hist[b], (unsigned long long)h[b]);
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
        unsigned long long c = b < bins ? s->hist[b] : 0ULL;
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
    __syncthreads();
    for (int b = t; b < bins; b += 1024) s->hist[b] = 0;
    if (t == 0) {
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
    unsigned long long pad;
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
                  long long w_capacity, Counters *cnt) {
    const unsigned long long nA = cnt->n_cand;
    const unsigned lane = threadIdx.x & 31;
    for (unsigned long long i0 = (unsigned long long)blockIdx.x * blockDim.x; i0 < nA;
         i0 += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long i = i0 + threadIdx.x;
        bool alive = false, is_work = false;
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
            is_work = !better;
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
                     long long det_capacity, Counters *cnt) {
    __shared__ int s_list[1024];
    __shared__ int s_n, found;
    const unsigned long long nW = cnt->n_work;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int r2 = r * r;
    for (unsigned long long w = blockIdx.x; w < nW; w += gridDim.x) {
        if (threadIdx.x == 0) { found = 0; s_n = 0; }
        __syncthreads();
        const unsigned long long idx = w_idx[w];
        const float val = w_val[w];
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
            for (unsigned long long wd = q0 >> 5; wd <= (q1 >> 5); ++wd) {
                unsigned long long lo = wd << 5;
                unsigned b0 = q0 > lo ? (unsigned)(q0 - lo) : 0u;
                unsigned b1 = q1 < lo + 31 ? (unsigned)(q1 - lo) : 31u;
                unsigned mask = (b1 == 31u ? 0xffffffffu : ((1u << (b1 + 1)) - 1u)) & ~((1u << b0) - 1u);
                atomicOr(&sup[wd], mask);
            }
        }
    }
}

__global__ void round_reset_kernel(Counters *cnt) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        cnt->n_cand = cnt->n_next; cnt->n_next = 0; cnt->n_work = 0; cnt->n_sel_round = 0;
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

// fplobjdetect.py:233-253: columns (x,y,z,conf) float64; coordinates are already "un-padded"
// (interior indices); keep rows with b <= coord < size-b (buffer in x,y,z order); add the offset.
// Order-preserving compaction by one block.
__global__ void __launch_bounds__(1024)
finish_rows_kernel(const unsigned *skey, const unsigned long long *sidx, long long n, Dims d,
                   int bx, int by, int bz, double ox, double oy, double oz, double *rows,
                   long long capacity, unsigned long long *n_out, unsigned long long *overflow) {
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
            c = (double)key2f(~skey[i]);
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

struct DetectBuffers {
    unsigned *sup; size_t sup_words;
    unsigned long long *a_idx, *b_idx, *w_idx, *det_idx, *sel_idx, *sidx;
    float *a_val, *b_val, *w_val, *det_val;
    unsigned *skey;
    Counters *cnt;
    unsigned long long *n_out;
    float *grid;                    // brick maxima
    long long cand_cap, work_cap, det_cap, sort_cap;
};

static long long next_pow2(long long v) { long long p = 1; while (p < v) p <<= 1; return p; }

static size_t detect_workspace_bytes(long long n, long long cand_cap, long long det_cap) {
    size_t b = 0;
    auto add = [&](size_t x) { b += (x + 255) & ~size_t(255); };
    add(((size_t)n + 31) / 32 * 4);
    add(cand_cap * 8); add(cand_cap * 8); add(cand_cap * 4); add(cand_cap * 4);   // A, B
    add(cand_cap * 8); add(cand_cap * 4);                                          // worklist
    add(det_cap * 8); add(det_cap * 4); add(det_cap * 8);                          // det, sel
    long long sp = next_pow2(det_cap);
    add(sp * 4); add(sp * 8);
    add(sizeof(Counters)); add(64);
    add(((size_t)n / 64 + 4 * (size_t)cbrt((double)n) * (size_t)cbrt((double)n) + 4096) * 4);   // brick grid (generous)
    return b + 4096;
}

static int detect_impl(fpl_ctx *ctx, const float *d_smooth, int64_t Z, int64_t Y, int64_t X,
                       const fpl_v2o_params *p, double threshold, long long cand_cap, double *d_dets,
                       int64_t capacity, int64_t *h_count, int64_t *h_stats, cudaStream_t st) {
    const long long n = Z * Y * X;
    const int r = p->obj_min_dist;
    Dims d{Z, Y, X};
    long long det_cap = capacity > 0 ? capacity : 1;
    // worklist capacity: every candidate could in principle be a local maximum
    DetectBuffers B;
    fpl::Arena &A = ctx->arena;
    B.sup_words = ((size_t)n + 31) / 32;
    B.sup = (unsigned *)A.take(B.sup_words * 4);
    B.a_idx = (unsigned long long *)A.take(cand_cap * 8);
    B.b_idx = (unsigned long long *)A.take(cand_cap * 8);
    B.a_val = (float *)A.take(cand_cap * 4);
    B.b_val = (float *)A.take(cand_cap * 4);
    B.w_idx = (unsigned long long *)A.take(cand_cap * 8);
    B.w_val = (float *)A.take(cand_cap * 4);
    B.det_idx = (unsigned long long *)A.take(det_cap * 8);
    B.det_val = (float *)A.take(det_cap * 4);
    B.sel_idx = (unsigned long long *)A.take(det_cap * 8);
    long long sp = next_pow2(det_cap);
    B.skey = (unsigned *)A.take(sp * 4);
    B.sidx = (unsigned long long *)A.take(sp * 8);
    B.cnt = (Counters *)A.take(sizeof(Counters));
    B.n_out = (unsigned long long *)A.take(64);
    const int gz = (int)((Z + kBrick - 1) / kBrick), gy = (int)((Y + kBrick - 1) / kBrick), gx = (int)((X + kBrick - 1) / kBrick);
    B.grid = (float *)A.take((size_t)gz * gy * gx * 4);
    if (!B.grid) { fpl::set_error("voxel2obj: workspace too small for the brick grid"); return FPL_ENOMEM; }
    if (!B.sup || !B.a_idx || !B.b_idx || !B.a_val || !B.b_val || !B.w_idx || !B.w_val || !B.det_idx ||
        !B.det_val || !B.sel_idx || !B.skey || !B.sidx || !B.cnt || !B.n_out) {
        fpl::set_error("voxel2obj: internal workspace sizing error");
        return FPL_ENOMEM;
    }
    FPL_CUDA_CHECK(cudaMemsetAsync(B.sup, 0, B.sup_words * 4, st));
    FPL_CUDA_CHECK(cudaMemsetAsync(B.cnt, 0, sizeof(Counters), st));
    FPL_CUDA_CHECK(cudaMemsetAsync(B.n_out, 0, 64, st));

    const int grid_stream = ctx->sm_count * 8;
    fpl::ProfScope prof(ctx, st, fpl::PROF_NMS, 4.0 * (double)n);
    compact_candidates_kernel<<<grid_stream, 256, 0, st>>>(d_smooth, n, threshold, B.a_idx, B.a_val,
                                                          cand_cap, B.cnt);
    FPL_LAUNCH_CHECK(ctx);

    FPL_REQUIRE(gz < 2048 && gy < 1024 && gx < 1024 && r <= 27 + 4, "voxel2obj: volume / radius outside the brick-grid limits");
    brick_max_kernel<<<gz * gy, 256, gx * sizeof(float), st>>>(d_smooth, d, gy, gx, B.grid);
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
    unsigned long long remaining = n_candidates;
    unsigned long long *a_idx = B.a_idx, *b_idx = B.b_idx;
    float *a_val = B.a_val, *b_val = B.b_val;
    while (remaining > 0) {
        ++rounds;
        long long fblocks = (long long)((remaining + 255) / 256);
        if (fblocks > grid_stream) fblocks = grid_stream;
        nms_filter_kernel<<<(unsigned)fblocks, 256, 0, st>>>(d_smooth, B.sup, d, a_idx, a_val, b_idx,
                                                            b_val, B.w_idx, B.w_val, cand_cap, B.cnt);
        FPL_LAUNCH_CHECK(ctx);
        nms_ballcheck_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(d_smooth, B.sup, d, r, B.grid, gz, gy, gx, B.w_idx, B.w_val,
                                                               B.det_idx, B.det_val, B.sel_idx, det_cap,
                                                               B.cnt);
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
        if (h_cnt->n_next == h_cnt->n_sel_round) {
            // everything left was selected this round
            remaining = 0;
        }
    }
    const unsigned long long n_det = h_cnt->n_det * (rounds > 0 ? 1 : 0);
    unsigned long long n_rows = 0;
    if (n_det > 0) {
        long long np2 = next_pow2((long long)n_det);
        int sblocks = (int)((np2 + 255) / 256); if (sblocks > grid_stream) sblocks = grid_stream;
        sort_prepare_kernel<<<sblocks, 256, 0, st>>>(B.det_idx, B.det_val, (long long)n_det, np2, B.skey, B.sidx);
        FPL_LAUNCH_CHECK(ctx);
        for (long long k = 2; k <= np2; k <<= 1)
            for (long long j = k >> 1; j > 0; j >>= 1) {
                bitonic_step_kernel<<<sblocks, 256, 0, st>>>(B.skey, B.sidx, np2, j, k);
                FPL_LAUNCH_CHECK(ctx);
            }
        finish_rows_kernel<<<1, 1024, 0, st>>>(B.skey, B.sidx, (long long)n_det, d, p->buffer_xyz[0],
                                               p->buffer_xyz[1], p->buffer_xyz[2], p->offset_xyz[0],
                                               p->offset_xyz[1], p->offset_xyz[2], d_dets, capacity,
                                               B.n_out, &B.cnt->overflow);
        FPL_LAUNCH_CHECK(ctx);
        unsigned long long *h_n = (unsigned long long *)((char *)ctx->h_pinned + 1024);
        FPL_CUDA_CHECK(cudaMemcpyAsync(h_n, B.n_out, 8, cudaMemcpyDeviceToHost, st));
        FPL_CUDA_CHECK(cudaStreamSynchronize(st));
        n_rows = *h_n;
    }
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

}  // namespace v2o
}  // namespace fpl

using namespace fpl::v2o;

static int g_force_generic_gauss = 0;

extern "C" {

// test hook: route every Gaussian pass through the generic one-thread-per-output kernel
int fpl_debug_force_generic_gauss(int on) { g_force_generic_gauss = on; return FPL_OK; }

int fpl_v2o_smooth(fpl_ctx *ctx, const float *d_pred, int64_t Z, int64_t Y, int64_t X,
                   const fpl_v2o_params *p, float *d_smooth, void *stream) {
    FPL_REQUIRE(ctx && d_pred && d_smooth, "fpl_v2o_smooth: NULL argument");
    FPL_TRY(check_params(p, Z, Y, X));
    FPL_REQUIRE(d_pred != d_smooth, "fpl_v2o_smooth: d_smooth may not alias d_pred");
    FPL_CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)Z * Y * X;
    FPL_CUDA_CHECK(cudaStreamSynchronize(st));   // arena may be re-allocated
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
    FPL_TRY(ctx->arena.reserve(detect_workspace_bytes(Z * Y * X, cand_cap, capacity > 0 ? capacity : 1)));
    ctx->arena.reset();
    return detect_impl(ctx, d_smooth, Z, Y, X, p, threshold, cand_cap, d_dets, capacity, h_count,
                       h_stats, st);
}

int fpl_voxel2obj(fpl_ctx *ctx, const float *d_pred, int64_t Z, int64_t Y, int64_t X,
                  const fpl_v2o_params *p, double *d_dets, int64_t capacity, int64_t *h_count,
                  double *h_threshold, int64_t *h_stats, void *stream) {
    FPL_REQUIRE(ctx && d_pred && (d_dets || capacity == 0), "fpl_voxel2obj: NULL argument");
    FPL_TRY(check_params(p, Z, Y, X));
    FPL_CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    FPL_CUDA_CHECK(cudaStreamSynchronize(st));
    const size_t n = (size_t)Z * Y * X;
    const long long cand_cap = candidate_capacity(p, Z, Y, X);
    size_t need = 2 * (n * sizeof(float) + 512) + 2 * sizeof(SelectState) + sizeof(ThreshOut) + 4096 +
                  detect_workspace_bytes((long long)n, cand_cap, capacity > 0 ? capacity : 1);
    FPL_TRY(ctx->arena.reserve(need));
    ctx->arena.reset();
    float *smooth = (float *)ctx->arena.take(n * sizeof(float));
    float *tmp = (float *)ctx->arena.take(n * sizeof(float));
    SelectState *states = (SelectState *)ctx->arena.take(2 * sizeof(SelectState));
    ThreshOut *tout = (ThreshOut *)ctx->arena.take(sizeof(ThreshOut));
    FPL_TRY(smooth_impl(ctx, d_pred, Z, Y, X, p, smooth, tmp, st, g_force_generic_gauss != 0));
    FPL_TRY(threshold_impl(ctx, smooth, Z, Y, X, p, tout, states, st));
    ThreshOut *h = (ThreshOut *)((char *)ctx->h_pinned + 2048);
    FPL_CUDA_CHECK(cudaMemcpyAsync(h, tout, sizeof(ThreshOut), cudaMemcpyDeviceToHost, st));
    FPL_CUDA_CHECK(cudaStreamSynchronize(st));
    const double threshold = h->thresh;
    if (h_threshold) *h_threshold = threshold;
    return detect_impl(ctx, smooth, Z, Y, X, p, threshold, cand_cap, d_dets, capacity, h_count, h_stats, st);
}

}  // extern "C"
