// detect_approx.cuh -- two-tier ("certified approximate") form of voxel2obj.  Included by detect.cu inside
// namespace fpl::v2o, after the exact kernels (it reuses their lists, counters, select and sort kernels).
//
// Why: the exact smoothing (SciPy's double chain, three passes, float32 store per pass) needs 93 FP64 operations per
// voxel and bounds the detection path at the FP64 pipe, 4x above its HBM roofline (12 B per voxel: read the map,
// write the smoothed map, read it once).  But voxel2obj only ever LOOKS at exact smoothed values in three places:
// the two order statistics behind the percentile threshold, the confidences of the detections, and comparisons
// between near-equal voxels inside one suppression ball.  Everything else only needs "certainly above / below".
//
// Tier 1: A = fp32 separable Gaussian of the whole map (one fused kernel, one read + one write), with a proven bound
//         |A - S| <= eps * S on every voxel (S = the reference's float32 result).  Requirements, checked at run time
//         (otherwise the exact path runs): inputs finite and >= 0, taps >= 0, lw <= min(r, 10).
//         Bound: per pass A_out = fl(sum fl(w_j) x_j), <= 21 non-negative terms, FMA accumulation -> relative error
//         <= 22u of the pass (u = 2^-24); relative perturbations of the inputs pass through a non-negative filter
//         unamplified, so three passes give < 67u against real arithmetic.  The reference chain rounds once per pass
//         to float32 (double accumulation adds < 1e-14): < 3.1u.  Together |A - S| < 71u * S = 4.3e-6 * S < eps := 2^-17.
//         (Underflow adds < 1e-43 absolute; only values >= 1e-30 are ever compared, see kApxMinCut.)
// Tier 2: exact values (exact_point: the SciPy chain, z -> y -> x, on the 21^3 neighbourhood of ONE voxel) only for
//           * the voxels whose A lies within 3 eps of the two order statistics of the percentile,
//           * the selected points (their confidences, and the final  S > threshold  test),
//           * ball comparisons that A cannot decide (values within 4 eps of each other).
// The greedy selection is the lexicographically-first maximal independent set under the order (S desc, index asc);
// it depends on S only through comparisons inside balls.  Every comparison is either certified by the bound
// (A(q) > A(p)(1+4eps) => S(q) > S(p), likewise below) or resolved with exact values, so every decision -- and hence
// the detection list, its order and its confidences -- is the reference's, bit for bit.
//
// Percentile without radix passes over the map: the Gaussian kernel also writes a 1/64 lattice sample of A.  Order
// statistics of the sample give a band [Lb, Hb] that holds the percentile's two order statistics with overwhelming
// probability; the one dense pass counts the voxels below the band and lists the voxels inside it.  The exact rank
// arithmetic on (count, list) CERTIFIES the band (a miss falls back to the exact path; it cannot produce a wrong answer).

constexpr float kApxEps = 1.0f / 131072.0f;                 // 2^-17
constexpr float kApxUp = 1.0f + 4.0f * kApxEps;             // A(q) > A(p) * kApxUp  =>  S(q) > S(p)
constexpr float kApxDn = 1.0f - 4.0f * kApxEps;             // A(q) < A(p) * kApxDn  =>  S(q) < S(p)
constexpr float kApxMinCut = 1e-30f;                        // below this the relative bound is not claimed
constexpr unsigned kApxBadBits = 0x7e967699u;               // raw bits of 1e38f: inputs at or above (incl. inf/NaN, and
                                                            // every negative value: sign bit) disqualify the map
constexpr int kApxMaxLw = 10;
constexpr int kApxSampleStep = 4;                           // lattice sample: every 4th plane, 1 of 4 rows, every 4th x
constexpr int kApxAmbCap = 1 << 16;                         // ambiguous ball checks per round (more = tie-heavy map)
constexpr int kApxMarginCap = 192;                          // near-equal voxels in one ball

__constant__ float c_gw32[2 * kApxMaxLw + 1];               // float32 taps of the current call (correlate order)

struct ApproxState {                                        // device
    float Lb, Hb, cutA;                                     // band edges and NMS cut-off, in A space
    float a_lo, a_hi;                                       // A values at the two ranks inside the band
    float n_lo, n_hi;                                       // narrow band (exact recompute) edges
    unsigned bad_bits;                                      // max of the raw input bits
    unsigned long long n_below;                             // owned voxels with A < Lb (= owned voxels - n_ge)
    unsigned long long n_ge;                                // owned voxels with A >= Lb, counted by the dense pass
    unsigned long long n_band;                              // band list entries
    unsigned long long n_below_narrow;                      // band entries below the narrow band
    unsigned long long n_narrow;                            // narrow list entries
    unsigned long long n_amb;                               // ambiguous ball checks of the current round
    unsigned long long overflow;                            // any list overflow -> exact path
};

// ------------------------------------------------------------------------------------------------------------------
// exact smoothed value of ONE interior voxel, straight from the probability map: the SciPy chain on its
// (2lw+1)^3 neighbourhood.  Block-cooperative (all threads call it with the same voxel; `sm` holds
// (2lw+1)^2 + (2lw+1) floats).  Valid for lw <= r (no reflection reaches the interior): values outside the volume
// are the pad zeros.  Returns the same value on every thread.
// ------------------------------------------------------------------------------------------------------------------
__device__ float exact_point(const float *__restrict__ pred, const Dims &d, int lw, const Taps &taps, long long z,
                             long long y, long long x, float *sm) {
    const int W = 2 * lw + 1;
    float *T1 = sm, *T2 = sm + W * W;
    const long long plane = d.Y * d.X;
    for (int c = threadIdx.x; c < W * W; c += blockDim.x) {
        const int dy = c / W - lw, dx = c % W - lw;
        const long long yy = y + dy, xx = x + dx;
        float res = 0.f;
        if (yy >= 0 && yy < d.Y && xx >= 0 && xx < d.X) {
            const float *col = pred + yy * d.X + xx;
            // all loads of the column first (independent: one memory latency), then the dependent FP64 chain
            auto at = [&](long long zz) -> float { return (zz >= 0 && zz < d.Z) ? __ldg(col + zz * plane) : 0.f; };
            float lo_v[kApxMaxLw], hi_v[kApxMaxLw];
#pragma unroll
            for (int k = 0; k < kApxMaxLw; ++k) {
                lo_v[k] = k < lw ? at(z - lw + k) : 0.f;
                hi_v[k] = k < lw ? at(z + lw - k) : 0.f;
            }
            double tmp = __dmul_rn((double)at(z), taps.w[lw]);
#pragma unroll
            for (int k = 0; k < kApxMaxLw; ++k)
                if (k < lw) tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn((double)lo_v[k], (double)hi_v[k]), taps.w[k]));
            res = (float)tmp;
        }
        T1[c] = res;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < W; c += blockDim.x) {
        auto at = [&](int dy) -> double { return (double)T1[(dy + lw) * W + c]; };
        double tmp = __dmul_rn(at(0), taps.w[lw]);
        for (int j = -lw; j < 0; ++j) tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(at(j), at(-j)), taps.w[lw + j]));
        T2[c] = (float)tmp;
    }
    __syncthreads();
    double tmp = __dmul_rn((double)T2[lw], taps.w[lw]);
    for (int j = -lw; j < 0; ++j)
        tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn((double)T2[lw + j], (double)T2[lw - j]), taps.w[lw + j]));
    __syncthreads();                     // the scratch may be overwritten by the next call
    return (float)tmp;
}

// Exact values of a LIST of voxels (the percentile's narrow band, the detections): the same chain as exact_point, but
// as three flat kernels -- one thread per (voxel, column) for the z pass, per (voxel, x offset) for the y pass, per voxel
// for the x pass -- so that the per-voxel dependency chain is hidden by parallelism instead of paid per block.
// t1: n * (2lw+1)^2 floats, t2: n * (2lw+1) floats of scratch.
constexpr int kExactBatch = 1 << 16;
__global__ void __launch_bounds__(256)
exact_z_kernel(const float *__restrict__ pred, Dims d, int lw, Taps taps, const unsigned long long *__restrict__ idx,
               long long n, float *__restrict__ t1) {
    const int W = 2 * lw + 1, W2 = W * W;
    const long long plane = d.Y * d.X, total = n * W2;
    for (long long gi = blockIdx.x * (long long)blockDim.x + threadIdx.x; gi < total; gi += (long long)gridDim.x * blockDim.x) {
        const long long i = gi / W2;
        const int c = (int)(gi - i * W2);
        const unsigned long long q = idx[i];
        const long long x = (long long)(q % (unsigned long long)d.X) + (c % W - lw);
        const long long y = (long long)((q / (unsigned long long)d.X) % (unsigned long long)d.Y) + (c / W - lw);
        const long long z = (long long)(q / ((unsigned long long)d.X * d.Y));
        float res = 0.f;
        if (y >= 0 && y < d.Y && x >= 0 && x < d.X) {
            const float *col = pred + y * d.X + x;
            auto at = [&](long long zz) -> float { return (zz >= 0 && zz < d.Z) ? __ldg(col + zz * plane) : 0.f; };
            float lo_v[kApxMaxLw], hi_v[kApxMaxLw];
#pragma unroll
            for (int k = 0; k < kApxMaxLw; ++k) {
                lo_v[k] = k < lw ? at(z - lw + k) : 0.f;
                hi_v[k] = k < lw ? at(z + lw - k) : 0.f;
            }
            double tmp = __dmul_rn((double)at(z), taps.w[lw]);
#pragma unroll
            for (int k = 0; k < kApxMaxLw; ++k)
                if (k < lw) tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn((double)lo_v[k], (double)hi_v[k]), taps.w[k]));
            res = (float)tmp;
        }
        t1[gi] = res;
    }
}
// in: n x rows x W (filter along rows, i.e. stride W) -> out: n x W; rows == W
__global__ void __launch_bounds__(256)
exact_y_kernel(const float *__restrict__ t1, int lw, Taps taps, long long n, float *__restrict__ t2) {
    const int W = 2 * lw + 1;
    const long long total = n * W;
    for (long long gi = blockIdx.x * (long long)blockDim.x + threadIdx.x; gi < total; gi += (long long)gridDim.x * blockDim.x) {
        const long long i = gi / W;
        const int dx = (int)(gi - i * W);
        const float *src = t1 + i * W * W + dx;
        double tmp = __dmul_rn((double)src[lw * W], taps.w[lw]);
        for (int k = 0; k < lw; ++k)
            tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn((double)src[k * W], (double)src[(2 * lw - k) * W]), taps.w[k]));
        t2[gi] = (float)tmp;
    }
}
__global__ void __launch_bounds__(256)
exact_x_kernel(const float *__restrict__ t2, int lw, Taps taps, long long n, float *__restrict__ out) {
    const int W = 2 * lw + 1;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float *src = t2 + i * W;
        double tmp = __dmul_rn((double)src[lw], taps.w[lw]);
        for (int k = 0; k < lw; ++k)
            tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn((double)src[k], (double)src[2 * lw - k]), taps.w[k]));
        out[i] = (float)tmp;
    }
}

static int exact_list(fpl_ctx *ctx, const float *d_pred, Dims d, int lw, const Taps &taps, const unsigned long long *idx,
                      long long n, float *out, float *t1, float *t2, cudaStream_t st) {
    const int W = 2 * lw + 1;
    for (long long b0 = 0; b0 < n; b0 += kExactBatch) {
        const long long nb = n - b0 < kExactBatch ? n - b0 : kExactBatch;
        long long blocks = (nb * W * W + 255) / 256;
        if (blocks > (long long)ctx->sm_count * 64) blocks = (long long)ctx->sm_count * 64;
        exact_z_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_pred, d, lw, taps, idx + b0, nb, t1);
        FPL_LAUNCH_CHECK(ctx);
        long long b2 = (nb * W + 255) / 256; if (b2 > (long long)ctx->sm_count * 32) b2 = (long long)ctx->sm_count * 32;
        exact_y_kernel<<<(unsigned)b2, 256, 0, st>>>(t1, lw, taps, nb, t2);
        FPL_LAUNCH_CHECK(ctx);
        long long b3 = (nb + 255) / 256;
        exact_x_kernel<<<(unsigned)b3, 256, 0, st>>>(t2, lw, taps, nb, out + b0);
        FPL_LAUNCH_CHECK(ctx);
    }
    return FPL_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Tier 1: fused fp32 separable Gaussian, z-streamed.  A block owns a 32 x 32 (y,x) tile and a chunk of z; per input
// plane it stages the (32+2lw) x (32+2*lwa) patch in shared memory (cp.async, zero fill outside the volume = the pad
// zeros), runs the x pass (8 outputs per thread from a register window), the y pass (4 outputs per thread) and feeds
// the 4 results into 2lw+1 running z accumulators per column that live in registers (scatter form: the shift is free
// because every FMA writes the neighbouring accumulator).  One global read, one global write per voxel.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kGT = 32;

template <int LW>
struct GaussCfg {
    static constexpr int W = 2 * LW + 1;
    static constexpr int PH = kGT + 2 * LW;                     // patch rows
    static constexpr int LWA = (LW + 3) & ~3;                   // x halo rounded up to whole 16-byte chunks
    static constexpr int PWV = kGT + 2 * LWA;                   // patch columns (x origin = tile x0 - LWA)
    static constexpr int XOFF = LWA - LW;                       // patch column of tap -LW of tile column 0
    static constexpr int WINV = (XOFF + 8 + 2 * LW + 3) & ~3;   // x-pass window of one 8-output run, whole float4s
    static constexpr int PITCH0 = PWV > 24 + WINV ? PWV : 24 + WINV;
    static constexpr int PITCH = ((PITCH0 / 4) % 2 == 1) ? PITCH0 : PITCH0 + 4;   // pitch / 4 odd: conflict-free LDS.128
    static constexpr int XS_PITCH = kGT;
    static constexpr size_t SMEM = sizeof(float) * (size_t)(2 * PH * PITCH + 2 * PH * XS_PITCH);
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc, bool valid) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc, bool valid) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// packed fp32: one FFMA2 issues two fused multiply-adds (sm_100 fma.rn.f32x2) -- the kernel is issue-bound
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// float32 taps of sigma = 5, truncate 2.0 (lw = 10) -- the reference's smoothing_sigma everywhere (scripts/*.py,
// fplobjdetect.py:844) -- as literals: the IMM instantiation multiplies by immediates (FFMA with an immediate operand
// issues at twice the rate of the register form, /opt/skills/guides/B300_MICROARCH.md "Pipe rates").  The host selects
// it only when the call's float32 taps equal this table bit for bit.
__device__ constexpr float kTapsSigma5[21] = {1.1194727e-02f, 1.6369877e-02f, 2.299882e-02f, 3.1045157e-02f, 4.02634e-02f, 5.0171286e-02f, 6.0065933e-02f, 6.909227e-02f, 7.6358765e-02f, 8.108053e-02f, 8.271846e-02f, 8.108053e-02f, 7.6358765e-02f, 6.909227e-02f, 6.0065933e-02f, 5.0171286e-02f, 4.02634e-02f, 3.1045157e-02f, 2.299882e-02f, 1.6369877e-02f, 1.1194727e-02f};

template <int LW, bool IMM>
__global__ void __launch_bounds__(256)
gauss32_kernel(const float *__restrict__ in, float *__restrict__ out, float *__restrict__ sample, Dims d, int zc_len,
               int sy, int sx, int vec_ok, ApproxState *state) {
    using C = GaussCfg<LW>;
    extern __shared__ __align__(16) float g_sm[];
    float *patch = g_sm;                                   // [2][PH][PITCH]
    float *xs = g_sm + 2 * C::PH * C::PITCH;               // [2][PH][32]
    const int t = threadIdx.x;
    const int X = (int)d.X, Y = (int)d.Y, Z = (int)d.Z;    // (the host checks that the extents fit 32 bits)
    const int tx0 = blockIdx.x * kGT, ty0 = blockIdx.y * kGT;
    const int zc0 = blockIdx.z * zc_len;
    const int zc1 = zc0 + zc_len < Z ? zc0 + zc_len : Z;
    const int z_begin = zc0 - LW, z_end = zc1 + LW;
    const long long plane = d.Y * d.X;

    // ---- staging plan of this thread: up to kLd chunks (16 B, vector layout) or elements (4 B) of the patch; the
    // (row, column) part of every source address is fixed, only the plane advances
    constexpr int kChunks = C::PH * (C::PWV / 4), kElems = C::PH * C::PWV;
    constexpr int kLdV = (kChunks + 255) / 256, kLdS = (kElems + 255) / 256;
    const float *src_v[kLdV]; int dst_v[kLdV]; bool ok_v[kLdV];
#pragma unroll
    for (int i = 0; i < kLdV; ++i) {
        const int c = t + 256 * i;
        const int row = c / (C::PWV / 4), c4 = c - row * (C::PWV / 4);
        const int gy = ty0 - LW + row, gx = tx0 - C::LWA + 4 * c4;
        ok_v[i] = c < kChunks && gy >= 0 && gy < Y && gx >= 0 && gx < X;       // X % 4 == 0: a chunk is wholly in or out
        src_v[i] = in + (ok_v[i] ? (long long)gy * X + gx : 0LL);
        dst_v[i] = c < kChunks ? row * C::PITCH + 4 * c4 : -1;
    }
    auto issue = [&](int gz, int buf) {                    // stage the patch of plane gz (zeros outside the volume)
        float *dst = patch + buf * C::PH * C::PITCH;
        const bool zok = gz >= 0 && gz < Z;
        const long long zoff = zok ? (long long)gz * plane : 0LL;
        if (vec_ok) {
#pragma unroll
            for (int i = 0; i < kLdV; ++i)
                if (dst_v[i] >= 0) cp_async16(dst + dst_v[i], src_v[i] + zoff, zok && ok_v[i]);
        } else {
#pragma unroll 2
            for (int i = 0; i < kLdS; ++i) {
                const int e = t + 256 * i;
                if (e < kElems) {
                    const int row = e / C::PWV, col = e - row * C::PWV;
                    const int gy = ty0 - LW + row, gx = tx0 - C::LWA + col;
                    const bool ok = zok && gy >= 0 && gy < Y && gx >= 0 && gx < X;
                    cp_async4(dst + row * C::PITCH + col, ok ? (const void *)(in + zoff + (long long)gy * X + gx) : (const void *)in, ok);
                }
            }
        }
        cp_async_commit();
    };

    // x-pass role: row = t / 4 (patch row), run = t % 4 (8 consecutive tile columns)
    const int xrow = t >> 2, xrun = t & 3;
    const bool x_active = xrow < C::PH;
    const bool x_checks = xrow >= LW && xrow < LW + kGT;    // rows of the tile itself: every input voxel is checked once
    // y/z-pass role: columns 2*cx2, 2*cx2+1 (one float2), rows cy0, cy0+1
    const int cx2 = t & 15, cy0 = (t >> 4) * 2;
    float2 acc[2][C::W];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int k = 0; k < C::W; ++k) acc[c][k] = make_float2(0.f, 0.f);
    float2 w2[C::W];                                        // taps duplicated into both halves
#pragma unroll
    for (int j = 0; j < C::W; ++j) w2[j] = make_float2(c_gw32[j], c_gw32[j]);
    static_assert(!IMM || LW == 10, "immediate taps exist for lw = 10 (sigma 5) only");
    // c + w[j] * a on both halves: packed FFMA2 with the tap in a uniform register, or two FFMA with the tap as an
    // immediate (j is a compile-time constant after unrolling, so kTapsSigma5[j] folds into the instruction)
#define FPL_MAC2(j, av, cv) (IMM ? make_float2(fmaf(kTapsSigma5[(j) < 21 ? (j) : 0], (av).x, (cv).x), \
                                               fmaf(kTapsSigma5[(j) < 21 ? (j) : 0], (av).y, (cv).y)) \
                                 : ffma2(w2[j], av, cv))
    unsigned bad = 0;
    // output / sample addressing of this thread (fixed in y and x)
    const int gx = tx0 + 2 * cx2, gy = ty0 + cy0;
    const bool st_pair = (X % 2 == 0) && gx + 1 < X;       // 8-byte stores need even row pitch
    const bool row_ok0 = gy < Y && gx < X, row_ok1 = gy + 1 < Y && gx < X;
    float *optr = out + (long long)zc0 * plane + (long long)gy * X + gx;

    // x pass of one staged plane: patch[pbuf] -> xs[xbuf]
    auto x_pass = [&](int pbuf, int xbuf) {
        if (x_active) {
            const float *prow = patch + pbuf * C::PH * C::PITCH + xrow * C::PITCH + 8 * xrun;
            float win[C::WINV + 1];
#pragma unroll
            for (int k = 0; k < C::WINV / 4; ++k) {
                const float4 q = *reinterpret_cast<const float4 *>(prow + 4 * k);
                win[4 * k] = q.x; win[4 * k + 1] = q.y; win[4 * k + 2] = q.z; win[4 * k + 3] = q.w;
            }
            win[C::WINV] = 0.f;
            if (x_checks) {
#pragma unroll
                for (int q = 0; q < 8; ++q) bad = max(bad, __float_as_uint(win[C::LWA + q]));
            }
            float2 o2[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                float2 a = make_float2(0.f, 0.f);
#pragma unroll
                for (int j = 0; j < C::W; ++j)
                    a = FPL_MAC2(j, make_float2(win[C::XOFF + 2 * m + j], win[C::XOFF + 2 * m + j + 1]), a);
                o2[m] = a;
            }
            float4 *xo = reinterpret_cast<float4 *>(xs + xbuf * C::PH * C::XS_PITCH + xrow * C::XS_PITCH + 8 * xrun);
            xo[0] = make_float4(o2[0].x, o2[0].y, o2[1].x, o2[1].y);
            xo[1] = make_float4(o2[2].x, o2[2].y, o2[3].x, o2[3].y);
        }
    };
    // Software pipeline with ONE block barrier per plane: in phase zi the y/z pass of plane zi (from xs[i&1]) and the x
    // pass of plane zi+1 (patch[(i+1)&1] -> xs[(i+1)&1]) run back to back, while cp.async stages plane zi+2.
    issue(z_begin, 0);
    if (z_begin + 1 < z_end) issue(z_begin + 1, 1);
    if (z_begin + 1 < z_end) asm volatile("cp.async.wait_group 1;\n" ::: "memory"); else cp_async_wait_all();
    __syncthreads();
    x_pass(0, 0);
    for (int zi = z_begin; zi < z_end; ++zi) {
        const int i = zi - z_begin;
        cp_async_wait_all();
        __syncthreads();                                   // xs[i&1] complete; patch[(i+1)&1] staged; patch[i&1] free
        if (zi + 2 < z_end) issue(zi + 2, i & 1);
        {
            const float *xsb = xs + (i & 1) * C::PH * C::XS_PITCH;
            float2 win[2 + 2 * LW];
#pragma unroll
            for (int k = 0; k < 2 + 2 * LW; ++k) win[k] = *reinterpret_cast<const float2 *>(xsb + (cy0 + k) * C::XS_PITCH + 2 * cx2);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                // two independent partial sums (even / odd taps): twice the ILP of one 21-long dependent chain
                float2 pe = make_float2(0.f, 0.f), po = make_float2(0.f, 0.f);
#pragma unroll
                for (int j = 0; j < C::W; j += 2) pe = FPL_MAC2(j, win[c + j], pe);
#pragma unroll
                for (int j = 1; j < C::W; j += 2) po = FPL_MAC2(j, win[c + j], po);
                const float2 p = __fadd2_rn(pe, po);
                // running z accumulators: after this plane acc[c][k] belongs to output plane zi - LW + k
#pragma unroll
                for (int k = 0; k < C::W - 1; ++k) acc[c][k] = FPL_MAC2(C::W - 1 - k, p, acc[c][k + 1]);
                acc[c][C::W - 1] = IMM ? make_float2(kTapsSigma5[0] * p.x, kTapsSigma5[0] * p.y) : make_float2(w2[0].x * p.x, w2[0].x * p.y);
            }
        }
        const int o = zi - LW;
        if (o >= zc0 && o < zc1) {
            if (st_pair) {
                if (row_ok0) *reinterpret_cast<float2 *>(optr) = acc[0][0];
                if (row_ok1) *reinterpret_cast<float2 *>(optr + X) = acc[1][0];
            } else {
                if (row_ok0) { optr[0] = acc[0][0].x; if (gx + 1 < X) optr[1] = acc[0][0].y; }
                if (row_ok1) { optr[X] = acc[1][0].x; if (gx + 1 < X) optr[X + 1] = acc[1][0].y; }
            }
            optr += plane;
            if ((o & (kApxSampleStep - 1)) == 0 && (cx2 & 1) == 0 && row_ok0) {
                // lattice sample: plane o, row 4*jy + phase(o) (clamped to the last row), column gx (a multiple of 4)
                int want = (gy & ~3) + ((o / kApxSampleStep) & 3);
                if (want > Y - 1) want = Y - 1;
                if (want == gy || want == gy + 1)
                    sample[((long long)(o / kApxSampleStep) * sy + (gy >> 2)) * sx + (gx >> 2)] = want == gy ? acc[0][0].x : acc[1][0].x;
            }
        }
        if (zi + 1 < z_end) x_pass((i + 1) & 1, (i + 1) & 1);
    }
    // negative / non-finite / huge inputs disqualify the map (the bound needs non-negative finite terms)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) bad = max(bad, __shfl_xor_sync(0xffffffffu, bad, off));
    if ((t & 31) == 0 && bad >= kApxBadBits) atomicMax(&state->bad_bits, bad);
}
#undef FPL_MAC2


// band [Lb, Hb] from the two sample order statistics; NMS cut-off
__global__ void approx_band_kernel(const SelectState *lo, const SelectState *hi, double thd, ApproxState *s) {
    if (threadIdx.x || blockIdx.x) return;
    const float L = key2f(lo->prefix), H = key2f(hi->prefix);
    s->Lb = L * (1.0f - 6.0f * kApxEps);
    s->Hb = H * (1.0f + 6.0f * kApxEps);
    // every voxel with S > threshold >= S_(rank_lo) has A > cutA (threshold >= a_lo / (1+eps), a_lo >= Lb / (1 - 6 eps))
    float cut = s->Lb * (1.0f - 4.0f * kApxEps);
    if (thd > 0.0 && thd < 3e38) {
        const float tc = (float)thd * (1.0f - 4.0f * kApxEps);
        if (tc > cut) cut = tc;
    }
    s->cutA = cut;
}

// how many sample values lie inside [Lb, Hb]: x 64 = the expected size of the band list.  A map whose values crowd
// around the percentile within the bound (tight distributions) is handed to the exact path before the dense pass.
__global__ void __launch_bounds__(256)
approx_sample_count_kernel(const float *__restrict__ sample, long long n_s, const ApproxState *S, unsigned long long *count) {
    const float Lb = S->Lb, Hb = S->Hb;
    unsigned c = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_s; i += (long long)gridDim.x * blockDim.x) {
        const float v = __ldg(sample + i);
        c += (v >= Lb && v <= Hb) ? 1u : 0u;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, (unsigned long long)c);
}

// ------------------------------------------------------------------------------------------------------------------
// the ONE dense pass over A: count of voxels below the band, list of the voxels inside it, 8^3 brick maxima, and the
// worklist of the first NMS round -- voxels >= cutA that no 26-neighbour certainly beats.  Warps are independent (no
// shared memory, no barriers).  A warp takes (plane z, strip of 32 rows) items and walks down the rows with a rolling
// 3-row window; a lane owns 4 consecutive x (one 16-byte load per row), lanes 0 and 31 only feed their neighbours'
// x +- 1 tests, so a warp produces 120 columns per pass and needs no edge loads.  The rows of the next group of 4 are
// in flight while the current group is processed.  Per voxel: separable 3x3 maximum from registers + shuffles; the few
// survivors fetch their 18 neighbours of the planes above / below through L1/L2.  Brick maxima are combined with
// integer atomicMax on the zero-initialised grid (the values are non-negative floats).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kP1Strip = 32, kP1Cols = 120, kP1Group = 4;
__device__ __align__(16) unsigned g_nan_row[4] = {0x7fc00000u, 0x7fc00000u, 0x7fc00000u, 0x7fc00000u};   // what lanes outside the volume load
template <bool VEC>
__global__ void __launch_bounds__(128, 5)
approx_pass1_kernel(const float *__restrict__ A, Dims d, int gy, int gx, float *__restrict__ g,
                    ApproxState *S, unsigned long long *band_idx, float *band_val, long long band_cap,
                    unsigned long long *w_idx, float *w_val, long long w_cap, Counters *cnt, int own_z0, int own_z1) {
    // planes [own_z0, own_z1) are owned (z-slab ranks: the rest is halo -- it feeds the neighbourhood tests and the brick
    // maxima, but is neither counted, listed nor put on the worklist)
    const float Lb = S->Lb, Hb = S->Hb, cutA = S->cutA;
    const float lowA = fminf(Lb, cutA);                     // (a thd above the percentile lifts cutA over the band)
    const int lane = threadIdx.x & 31;
    const int X = (int)d.X, Y = (int)d.Y, Z = (int)d.Z;
    const long long plane = d.Y * d.X;
    const int strips = (Y + kP1Strip - 1) / kP1Strip, xchunks = (X + kP1Cols - 1) / kP1Cols;
    const long long n_items = (long long)Z * strips * xchunks;
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long n_warps = (long long)gridDim.x * (blockDim.x >> 5);
    const bool lane_owns = lane >= 1 && lane <= 30;
    unsigned long long n_ge = 0;                            // owned voxels with A >= Lb (the host turns it into "below")
    for (long long it = warp0; it < n_items; it += n_warps) {
        // x chunk fastest, then strip, then plane: concurrently running warps share halo rows / planes in L1/L2
        const int xcI = (int)(it % xchunks);
        const int sI = (int)((it / xchunks) % strips);
        const int z = (int)(it / ((long long)xchunks * strips));
        const int x = xcI * kP1Cols - 4 + 4 * lane, y0 = sI * kP1Strip;
        const int y_end = min(y0 + kP1Strip, Y);            // rows [y0, y_end) are produced
        const float *pz = A + (long long)z * plane;
        const bool owned = z >= own_z0 && z < own_z1;       // warp-uniform
        const bool col_ok = x >= 0 && x < X;
        const bool owns = owned && lane_owns;
        // values outside the volume are NaN: every comparison with them is false (not counted, not a candidate, never
        // "better") and fmaxf / FMNMX ignore them -- no per-voxel range tests in the hot loop.  Lanes whose columns lie
        // outside read a NaN row (stride 0) instead of the map: the load itself needs no predicate either.
        const float kOut = __int_as_float(0x7fc00000);
        const float *lane_base = col_ok ? pz + x : reinterpret_cast<const float *>(g_nan_row);
        const long long lane_stride = col_ok ? (long long)X : 0LL;
        auto load_row = [&](int yy, float (&o)[4]) {
            if (yy >= 0 && yy < Y) {                            // warp-uniform
                const float *rp = lane_base + (long long)yy * lane_stride;
                if (VEC) { const float4 q = __ldg(reinterpret_cast<const float4 *>(rp)); o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w; }
                else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) o[e] = (!col_ok || x + e < X) ? __ldg(rp + e) : kOut;
                }
            } else { o[0] = o[1] = o[2] = o[3] = kOut; }
        };
        float pv[4] = {kOut, kOut, kOut, kOut};             // values of the previous row
        float m3pp[4] = {kOut, kOut, kOut, kOut}, m3p[4] = {kOut, kOut, kOut, kOut};
        float bm = 0.f;                                     // running brick maximum (8 rows)
        unsigned surv = 0, bandm = 0;                       // bit 4 k + e of the current group: output row yb + k - 1, column x + e
        // one group = 4 consecutive rows; row yy - 1 is decided when row yy arrives.  Uniform control flow only (the
        // shuffles must not sit behind divergent branches): rows past the strip are loaded as -inf / masked out.
        auto process_group = [&](const float (&c)[kP1Group][4], int yb) {
#pragma unroll
            for (int k = 0; k < kP1Group; ++k) {
                const float lf = __shfl_up_sync(0xffffffffu, c[k][3], 1), rt = __shfl_down_sync(0xffffffffu, c[k][0], 1);
                float m3c[4];
                m3c[0] = fmaxf(fmaxf(lf, c[k][0]), c[k][1]);
                m3c[1] = fmaxf(fmaxf(c[k][0], c[k][1]), c[k][2]);
                m3c[2] = fmaxf(fmaxf(c[k][1], c[k][2]), c[k][3]);
                m3c[3] = fmaxf(fmaxf(c[k][2], c[k][3]), rt);
                const int yo = yb + k - 1;                  // output row
                const bool row_out = yo >= y0 && yo < y_end;        // warp-uniform
                if (row_out) {
                    // common path, per row of 4 voxels: row maximum (brick maximum, "any candidate?") and the below-band count
                    const float rmax = fmaxf(fmaxf(pv[0], pv[1]), fmaxf(pv[2], pv[3]));
                    bm = fmaxf(bm, rmax);
                    if (owns && rmax >= lowA) {                     // rare: some voxel of the row reaches the band / the cut-off
                        {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float f = pv[e];
                                n_ge += f >= Lb ? 1u : 0u;
                                const float m9 = fmaxf(fmaxf(m3pp[e], m3p[e]), m3c[e]);
                                // band member / nobody in the plane certainly beats it (near-equal neighbours are left to
                                // the ball check): both are rare and handled after the group, out of the unrolled code
                                if (f >= Lb && f <= Hb) bandm |= 1u << (4 * k + e);
                                if (f >= cutA && !(m9 > f * kApxUp)) surv |= 1u << (4 * k + e);
                            }
                        }
                    }
                }
                // brick maximum: flushed every 8 rows (strips start on brick boundaries); lanes (odd, odd+1) share a brick
                if (row_out && ((yo & (kBrick - 1)) == kBrick - 1 || yo == y_end - 1)) {        // warp-uniform
                    const float other = __shfl_down_sync(0xffffffffu, bm, 1);
                    if ((lane & 1) && lane <= 29 && col_ok && fmaxf(bm, other) > 0.f)
                        atomicMax(reinterpret_cast<int *>(g + ((size_t)(z / kBrick) * gy + yo / kBrick) * gx + x / kBrick),
                                  __float_as_int(fmaxf(bm, other)));
                    bm = 0.f;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) { pv[e] = c[k][e]; m3pp[e] = m3p[e]; m3p[e] = m3c[e]; }
            }
        };
        auto rare = [&](int yb) {
#pragma unroll 1
            while (bandm) {
                const int b = __ffs((int)bandm) - 1;
                bandm &= bandm - 1u;
                const unsigned long long idx = (unsigned long long)((long long)z * plane + (long long)(yb + (b >> 2) - 1) * X + x + (b & 3));
                const unsigned long long pos = atomicAdd(&S->n_band, 1ULL);
                if ((long long)pos < band_cap) { band_idx[pos] = idx; band_val[pos] = __ldg(A + idx); }
                else atomicAdd(&S->overflow, 1ULL);
            }
#pragma unroll 1
            while (surv) {
                const int b = __ffs((int)surv) - 1;
                surv &= surv - 1u;
                const int yo = yb + (b >> 2) - 1, xx = x + (b & 3);
                const unsigned long long idx = (unsigned long long)((long long)z * plane + (long long)yo * X + xx);
                const float val = __ldg(A + idx);
                const float hi = val * kApxUp;
                bool ok = true;                             // the 18 neighbours of the planes above / below
#pragma unroll
                for (int dz = -1; dz <= 1; dz += 2) {
                    const int zz = z + dz;
                    if (zz < 0 || zz >= Z) continue;
#pragma unroll
                    for (int dy = -1; dy <= 1; ++dy) {
                        const int y2 = yo + dy;
                        if (y2 < 0 || y2 >= Y) continue;
                        const float *rp = A + (long long)zz * plane + (long long)y2 * X;
#pragma unroll
                        for (int dx = -1; dx <= 1; ++dx) {
                            const int x2 = xx + dx;
                            if (x2 >= 0 && x2 < X && __ldg(rp + x2) > hi) ok = false;
                        }
                    }
                }
                if (ok) {
                    const unsigned long long pos = atomicAdd(&cnt->n_work, 1ULL);
                    if ((long long)pos < w_cap) { w_idx[pos] = idx; w_val[pos] = val; }
                    else atomicAdd(&cnt->overflow, 1ULL);
                }
            }
        };
        // rows y0-1 .. y_end are consumed in groups of 4; the loads of the next group are in flight while the current
        // one is processed (one copy of the processing code: it must stay inside the instruction cache)
        float cur[kP1Group][4], nxt[kP1Group][4];
#pragma unroll
        for (int k = 0; k < kP1Group; ++k) load_row(y0 - 1 + k, nxt[k]);
#pragma unroll 1
        for (int yb = y0 - 1; yb <= y_end; yb += kP1Group) {
#pragma unroll
            for (int k = 0; k < kP1Group; ++k)
#pragma unroll
                for (int e = 0; e < 4; ++e) cur[k][e] = nxt[k][e];
#pragma unroll
            for (int k = 0; k < kP1Group; ++k) load_row(yb + kP1Group + k <= y_end ? yb + kP1Group + k : -1, nxt[k]);
            process_group(cur, yb);
            rare(yb);
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) n_ge += __shfl_xor_sync(0xffffffffu, n_ge, off);
    if (lane == 0 && n_ge) atomicAdd(&S->n_ge, n_ge);
}

// a_lo / a_hi (A at the two ranks, found by the radix select on the band list) -> narrow band edges
__global__ void approx_narrow_setup_kernel(const SelectState *lo, const SelectState *hi, ApproxState *s) {
    if (threadIdx.x || blockIdx.x) return;
    s->a_lo = key2f(lo->prefix); s->a_hi = key2f(hi->prefix);
    s->n_lo = s->a_lo * (1.0f - 3.0f * kApxEps);
    s->n_hi = s->a_hi * (1.0f + 3.0f * kApxEps);
}

// band entries inside the narrow band -> list (exact values follow); entries below it are only counted
__global__ void approx_narrow_kernel(const unsigned long long *__restrict__ band_idx, const float *__restrict__ band_val,
                                     ApproxState *s, unsigned long long *n_idx, long long n_cap) {
    const unsigned long long n = s->n_band;
    const float lo = s->n_lo, hi = s->n_hi;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float a = band_val[i];
        if (a < lo) atomicAdd(&s->n_below_narrow, 1ULL);
        else if (a <= hi) {
            const unsigned long long pos = atomicAdd(&s->n_narrow, 1ULL);
            if ((long long)pos < n_cap) n_idx[pos] = band_idx[i];
            else atomicAdd(&s->overflow, 1ULL);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// NMS rounds on A with certified comparisons
// ------------------------------------------------------------------------------------------------------------------
// list A -> list B (still valid) and worklist (valid, owned, and no valid 26-neighbour certainly better)
__global__ void __launch_bounds__(256)
approx_filter_kernel(const float *__restrict__ v, const unsigned *__restrict__ sup, Dims d,
                     const unsigned long long *__restrict__ a_idx, const float *__restrict__ a_val,
                     unsigned long long *b_idx, float *b_val, unsigned long long *w_idx, float *w_val,
                     long long w_capacity, Counters *cnt) {
    const unsigned long long nA = cnt->n_cand;
    const unsigned lane = threadIdx.x & 31;
    for (unsigned long long i0 = (unsigned long long)blockIdx.x * blockDim.x; i0 < nA;
         i0 += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long i = i0 + threadIdx.x;
        bool alive = false, is_work = false;
        unsigned long long idx = 0; float val = 0.f;
        if (i < nA) { idx = a_idx[i]; val = a_val[i]; alive = !is_suppressed(sup, idx); }
        if (alive) {
            long long x, y, z;
            decode_idx(idx, d, z, y, x);
            const float hi = val * kApxUp;
            bool better = false;
            for (int dz = -1; dz <= 1 && !better; ++dz) {
                const long long zz = z + dz; if (zz < 0 || zz >= d.Z) continue;
                for (int dy = -1; dy <= 1 && !better; ++dy) {
                    const long long yy = y + dy; if (yy < 0 || yy >= d.Y) continue;
                    for (int dx = -1; dx <= 1; ++dx) {
                        const long long xx = x + dx; if (xx < 0 || xx >= d.X) continue;
                        if (!(dz | dy | dx)) continue;
                        const unsigned long long q = ((unsigned long long)zz * d.Y + yy) * d.X + xx;
                        if (__ldg(v + q) > hi && !is_suppressed(sup, q)) { better = true; break; }
                    }
                }
            }
            is_work = !better;
        }
        const unsigned m_alive = __ballot_sync(0xffffffffu, alive);
        const unsigned m_work = __ballot_sync(0xffffffffu, is_work);
        unsigned long long base_b = 0, base_w = 0;
        if (lane == 0) {
            if (m_alive) base_b = atomicAdd(&cnt->n_next, (unsigned long long)__popc(m_alive));
            if (m_work) base_w = atomicAdd(&cnt->n_work, (unsigned long long)__popc(m_work));
        }
        base_b = __shfl_sync(0xffffffffu, base_b, 0);
        base_w = __shfl_sync(0xffffffffu, base_w, 0);
        const unsigned below = (1u << lane) - 1u;
        if (alive) { const unsigned long long p = base_b + __popc(m_alive & below); b_idx[p] = idx; b_val[p] = val; }
        if (is_work) {
            const unsigned long long p = base_w + __popc(m_work & below);
            if ((long long)p < w_capacity) { w_idx[p] = idx; w_val[p] = val; }
            else atomicAdd(&cnt->overflow, 1ULL);
        }
    }
}

// ball scan shared by the check and the resolve kernels: lists the bricks of the ball of (z,y,x) whose maximum
// reaches `floor_val` into s_list (block-cooperative; returns the count, capped at 1024)
__device__ __forceinline__ int approx_list_bricks(const float *__restrict__ g, int gy, int gx, const Dims &d, int r, int z,
                                                  int y, int x, float floor_val, int *s_list, int *s_n) {
    const int r2 = r * r;
    if (threadIdx.x == 0) *s_n = 0;
    __syncthreads();
    const int bz0 = max(z - r, 0) / kBrick, bz1 = (int)(min((long long)z + r, d.Z - 1) / kBrick);
    const int by0 = max(y - r, 0) / kBrick, by1 = (int)(min((long long)y + r, d.Y - 1) / kBrick);
    const int bx0 = max(x - r, 0) / kBrick, bx1 = (int)(min((long long)x + r, d.X - 1) / kBrick);
    const int nbz = bz1 - bz0 + 1, nby = by1 - by0 + 1, nbx = bx1 - bx0 + 1;
    for (int i = threadIdx.x; i < nbz * nby * nbx; i += blockDim.x) {
        const int bx = bx0 + i % nbx, by = by0 + (i / nbx) % nby, bz = bz0 + i / (nbx * nby);
        const float gm = __ldg(g + ((size_t)bz * gy + by) * gx + bx);
        if (!(gm >= floor_val)) continue;
        const int cz = min(max(z, bz * kBrick), bz * kBrick + kBrick - 1);
        const int cy = min(max(y, by * kBrick), by * kBrick + kBrick - 1);
        const int cx = min(max(x, bx * kBrick), bx * kBrick + kBrick - 1);
        if ((cz - z) * (cz - z) + (cy - y) * (cy - y) + (cx - x) * (cx - x) > r2) continue;
        const int slot = atomicAdd(s_n, 1);
        if (slot < 1024) s_list[slot] = (bz << 20) | (by << 10) | bx;
    }
    __syncthreads();
    return min(*s_n, 1024);
}

// one WARP per worklist entry (no block barriers: entries are independent).  Outcome per entry p: some valid voxel of
// the ball certainly beats it -> nothing; every other valid voxel of the ball is certainly worse -> selected;
// otherwise -> ambiguous list (exact resolve).  A voxel that matters has A >= A(p)(1 - 4 eps), so only bricks whose
// maximum reaches that are scanned (the brick grid stays in L2); for an isolated peak that is its own blob.
// EXACT: the same scan on an exact smoothed map with the reference's order (larger value, or equal value and lower flat
// index) -- no margin, no ambiguous outcome; used by the exact rounds (single GPU and slab sessions).
constexpr int kBcListCap = 768;                             // >= 9^3 bricks can touch a ball of radius <= 31
template <bool EXACT>
__global__ void __launch_bounds__(256)
approx_ballcheck_kernel(const float *__restrict__ v, const unsigned *__restrict__ sup, Dims d, int r,
                        const float *__restrict__ g, int gy, int gx, const unsigned long long *__restrict__ w_idx,
                        const float *__restrict__ w_val, unsigned long long *det_idx, float *det_val,
                        unsigned long long *sel_idx, long long det_capacity, Counters *cnt, ApproxState *S,
                        unsigned long long *amb_idx, unsigned long long own_lo, unsigned long long own_hi) {
    __shared__ int s_list_all[8][kBcListCap];
    const unsigned long long nW = cnt->n_work;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int *s_list = s_list_all[warp];
    const int r2 = r * r;
    const float cutA = S ? S->cutA : -INFINITY;
    const unsigned long long w0 = (unsigned long long)blockIdx.x * 8 + warp, wstride = (unsigned long long)gridDim.x * 8;
    for (unsigned long long w = w0; w < nW; w += wstride) {
        const unsigned long long idx = w_idx[w];
        const float val = w_val[w];
        if (!(val >= cutA) || idx < own_lo || idx >= own_hi) continue;      // warp-uniform
        const int x = (int)(idx % (unsigned long long)d.X);
        const int y = (int)((idx / (unsigned long long)d.X) % (unsigned long long)d.Y);
        const int z = (int)(idx / ((unsigned long long)d.X * d.Y));
        const float lo = EXACT ? val : val * kApxDn, hi = EXACT ? val : val * kApxUp;
        // bricks of the ball whose maximum reaches lo
        const int bz0 = max(z - r, 0) / kBrick, bz1 = (int)(min((long long)z + r, d.Z - 1) / kBrick);
        const int by0 = max(y - r, 0) / kBrick, by1 = (int)(min((long long)y + r, d.Y - 1) / kBrick);
        const int bx0 = max(x - r, 0) / kBrick, bx1 = (int)(min((long long)x + r, d.X - 1) / kBrick);
        const int nby = by1 - by0 + 1, nbx = bx1 - bx0 + 1, nb = (bz1 - bz0 + 1) * nby * nbx;
        int nlist = 0;
        for (int i0 = 0; i0 < nb; i0 += 32) {
            const int i = i0 + lane;
            bool take = false;
            int code = 0;
            if (i < nb) {
                const int bx = bx0 + i % nbx, by = by0 + (i / nbx) % nby, bz = bz0 + i / (nbx * nby);
                const float gm = __ldg(g + ((size_t)bz * gy + by) * gx + bx);
                if (gm >= lo) {
                    const int cz = min(max(z, bz * kBrick), bz * kBrick + kBrick - 1);
                    const int cy = min(max(y, by * kBrick), by * kBrick + kBrick - 1);
                    const int cx = min(max(x, bx * kBrick), bx * kBrick + kBrick - 1);
                    take = (cz - z) * (cz - z) + (cy - y) * (cy - y) + (cx - x) * (cx - x) <= r2;
                    code = (bz << 20) | (by << 10) | bx;
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, take);
            if (take) { const int slot = nlist + __popc(m & ((1u << lane) - 1u)); if (slot < kBcListCap) s_list[slot] = code; }
            nlist += __popc(m);
        }
        __syncwarp();
        if (nlist > kBcListCap) { if (lane == 0) atomicAdd(S ? &S->overflow : &cnt->overflow, 1ULL); continue; }
        bool found = false, amb = false;
        for (int li = 0; li < nlist && !found; ++li) {
            const int code = s_list[li];
            const int bz = code >> 20, by = (code >> 10) & 1023, bx = code & 1023;
            bool hit = false, near = false;
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
                    if (vq >= lo && q != idx && !is_suppressed(sup, q)) {
                        if (EXACT) { if (vq > val || q < idx) hit = true; }         // equal value: the lower flat index is better
                        else if (vq > hi) hit = true;
                        else near = true;
                    }
                }
            }
            found = __any_sync(0xffffffffu, hit);
            amb = amb || __any_sync(0xffffffffu, near);
        }
        __syncwarp();
        if (lane == 0 && !found) {
            if (!EXACT && amb) {
                const unsigned long long a = atomicAdd(&S->n_amb, 1ULL);
                if (a < (unsigned long long)kApxAmbCap) amb_idx[a] = idx;
                else atomicAdd(&S->overflow, 1ULL);
            } else {
                const unsigned long long p = atomicAdd(&cnt->n_det, 1ULL);
                const unsigned long long s = atomicAdd(&cnt->n_sel_round, 1ULL);
                if ((long long)p < det_capacity) { det_idx[p] = idx; det_val[p] = val; sel_idx[s] = idx; }
                else atomicAdd(&cnt->overflow, 1ULL);
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&cnt->ball_checks, nW);
}

// one block per ambiguous entry p (no valid voxel of its ball certainly beats it, some are within the margin): exact
// values of p and of those voxels decide, with the reference's tie rule (lower flat index wins).
__global__ void __launch_bounds__(256)
approx_resolve_kernel(const float *__restrict__ pred, const float *__restrict__ v, const unsigned *__restrict__ sup,
                      Dims d, int r, int lw, Taps taps, const float *__restrict__ g, int gy, int gx,
                      const unsigned long long *__restrict__ amb_idx, unsigned long long *det_idx, float *det_val,
                      unsigned long long *sel_idx, long long det_capacity, Counters *cnt, ApproxState *S) {
    extern __shared__ float ex_sm[];                        // exact_point scratch
    __shared__ int s_list[1024];
    __shared__ unsigned long long s_q[kApxMarginCap];
    __shared__ int s_n, s_nq;
    const unsigned long long nA = S->n_amb < (unsigned long long)kApxAmbCap ? S->n_amb : (unsigned long long)kApxAmbCap;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int r2 = r * r;
    for (unsigned long long a = blockIdx.x; a < nA; a += gridDim.x) {
        const unsigned long long idx = amb_idx[a];
        const float val = __ldg(v + idx);
        const int x = (int)(idx % (unsigned long long)d.X);
        const int y = (int)((idx / (unsigned long long)d.X) % (unsigned long long)d.Y);
        const int z = (int)(idx / ((unsigned long long)d.X * d.Y));
        const float lo = val * kApxDn;
        if (threadIdx.x == 0) s_nq = 0;
        const int nlist = approx_list_bricks(g, gy, gx, d, r, z, y, x, lo, s_list, &s_n);
        for (int li = warp; li < nlist; li += nwarps) {
            const int code = s_list[li];
            const int bz = code >> 20, by = (code >> 10) & 1023, bx = code & 1023;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int row = lane + 32 * h;
                const long long zz = (long long)bz * kBrick + row / kBrick, yy = (long long)by * kBrick + row % kBrick;
                if (zz >= d.Z || yy >= d.Y) continue;
                const int dzy = (int)((zz - z) * (zz - z) + (yy - y) * (yy - y));
                if (dzy > r2) continue;
                const unsigned long long rowbase = ((unsigned long long)zz * d.Y + yy) * d.X;
                for (int e = 0; e < kBrick; ++e) {
                    const long long xx = (long long)bx * kBrick + e;
                    if (xx >= d.X) break;
                    const int ddx = (int)(xx - x);
                    if (dzy + ddx * ddx > r2) continue;
                    const unsigned long long q = rowbase + xx;
                    if (__ldg(v + q) >= lo && q != idx && !is_suppressed(sup, q)) {
                        const int slot = atomicAdd(&s_nq, 1);
                        if (slot < kApxMarginCap) s_q[slot] = q;
                    }
                }
            }
        }
        __syncthreads();
        const int nq = s_nq;
        if (nq > kApxMarginCap) {                           // plateau: not this path's business
            if (threadIdx.x == 0) atomicAdd(&S->overflow, 1ULL);
            __syncthreads();
            continue;
        }
        const float sp = exact_point(pred, d, lw, taps, z, y, x, ex_sm);
        bool lose = false;
        for (int k = 0; k < nq; ++k) {
            const unsigned long long q = s_q[k];
            const long long qx = (long long)(q % (unsigned long long)d.X);
            const long long qy = (long long)((q / (unsigned long long)d.X) % (unsigned long long)d.Y);
            const long long qz = (long long)(q / ((unsigned long long)d.X * d.Y));
            const float sq = exact_point(pred, d, lw, taps, qz, qy, qx, ex_sm);
            if (sq > sp || (sq == sp && q < idx)) lose = true;
        }
        if (threadIdx.x == 0 && !lose) {
            const unsigned long long p = atomicAdd(&cnt->n_det, 1ULL);
            const unsigned long long s = atomicAdd(&cnt->n_sel_round, 1ULL);
            if ((long long)p < det_capacity) { det_idx[p] = idx; det_val[p] = val; sel_idx[s] = idx; }
            else atomicAdd(&cnt->overflow, 1ULL);
        }
        __syncthreads();
    }
}

// survivors for the later rounds: valid voxels >= cutA, gathered brick by brick -- only bricks whose maximum reaches
// the cut-off are read (a warp per brick, 64 rows of 32 bytes)
__global__ void __launch_bounds__(256)
approx_compact_kernel(const float *__restrict__ v, Dims d, const float *__restrict__ g, int gz, int gy, int gx,
                      const unsigned *__restrict__ sup, const ApproxState *S, unsigned long long *cand_idx,
                      float *cand_val, long long capacity, Counters *cnt) {
    const float cutA = S->cutA;
    const int lane = threadIdx.x & 31;
    const long long n_bricks = (long long)gz * gy * gx;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    const bool vec_ok = (d.X % 4 == 0) && ((reinterpret_cast<uintptr_t>(v) & 15) == 0);
    for (long long b0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; b0 < n_bricks; b0 += warps * 32) {
        // a warp looks at 32 consecutive bricks of the grid at once and then visits those that qualify
        const long long bme = b0 + lane;
        const bool q = bme < n_bricks && __ldg(g + bme) >= cutA;
        unsigned todo = __ballot_sync(0xffffffffu, q);
        while (todo) {
            const int bi = __ffs((int)todo) - 1;
            todo &= todo - 1u;
            const long long b = b0 + bi;
            const int bx = (int)(b % gx), by = (int)((b / gx) % gy), bz = (int)(b / ((long long)gx * gy));
            float f[2][8];
            unsigned mask = 0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int row = lane + 32 * h;
                const long long zz = (long long)bz * kBrick + row / kBrick, yy = (long long)by * kBrick + row % kBrick;
                const long long x0 = (long long)bx * kBrick;
                const bool rok = zz < d.Z && yy < d.Y;
                const unsigned long long rowbase = ((unsigned long long)zz * d.Y + yy) * d.X + x0;
                if (rok && vec_ok && x0 + 8 <= d.X) {
                    const float4 a4 = __ldg(reinterpret_cast<const float4 *>(v + rowbase));
                    const float4 b4 = __ldg(reinterpret_cast<const float4 *>(v + rowbase) + 1);
                    f[h][0] = a4.x; f[h][1] = a4.y; f[h][2] = a4.z; f[h][3] = a4.w;
                    f[h][4] = b4.x; f[h][5] = b4.y; f[h][6] = b4.z; f[h][7] = b4.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[h][e] = (rok && x0 + e < d.X) ? __ldg(v + rowbase + e) : -INFINITY;
                }
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (f[h][e] >= cutA && !is_suppressed(sup, rowbase + e)) mask |= 1u << (h * 8 + e);
            }
            const unsigned cnt_me = __popc(mask);
            unsigned pre = cnt_me;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, pre, o);
                if (lane >= o) pre += t;
            }
            const unsigned total = __shfl_sync(0xffffffffu, pre, 31);
            unsigned long long wbase = 0;
            if (total) {
                if (lane == 0) wbase = atomicAdd(&cnt->n_cand, (unsigned long long)total);
                wbase = __shfl_sync(0xffffffffu, wbase, 0);
            }
            unsigned long long pos = wbase + (pre - cnt_me);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int row = lane + 32 * h;
                const long long zz = (long long)bz * kBrick + row / kBrick, yy = (long long)by * kBrick + row % kBrick;
                const unsigned long long rowbase = ((unsigned long long)zz * d.Y + yy) * d.X + (unsigned long long)bx * kBrick;
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (mask & (1u << (h * 8 + e))) {
                        if ((long long)pos < capacity) { cand_idx[pos] = rowbase + e; cand_val[pos] = f[h][e]; }
                        else atomicAdd(&cnt->overflow, 1ULL);
                        ++pos;
                    }
            }
        }
    }
}

// Suppression balls, brick-guided: the validity bit of a voxel is only ever consulted for voxels with A >= cutA (1 - 4 eps)
// (filter, ball check, resolve, compaction), and those live in bricks whose maximum reaches gcut = cutA (1 - 4 eps)^2.
// A block first marks which of the <= 9^3 bricks around its point qualify (one bit each, shared memory); rows of the
// ball whose 64-bit words only cover other bricks are skipped -- on sparse maps that is most of the 4 800 atomics a
// ball of radius 27 would otherwise issue.
__global__ void __launch_bounds__(256)
approx_suppress_kernel(unsigned *sup, Dims d, int r, const float *__restrict__ g, int gy, int gx, const ApproxState *S,
                       const unsigned long long *__restrict__ sel_idx, const Counters *cnt) {
    __shared__ unsigned s_rowbits[10 * 10];                 // [brick z][brick y] -> bit per brick x (relative)
    const unsigned long long nS = cnt->n_sel_round;
    const int side = 2 * r + 1;
    const float gcut = S->cutA * kApxDn * kApxDn;
    for (unsigned long long s = blockIdx.x; s < nS; s += gridDim.x) {
        const unsigned long long idx = sel_idx[s];
        const long long x = (long long)(idx % (unsigned long long)d.X);
        const long long y = (long long)((idx / (unsigned long long)d.X) % (unsigned long long)d.Y);
        const long long z = (long long)(idx / ((unsigned long long)d.X * d.Y));
        const int bz0 = (int)(max(z - r, 0LL) / kBrick), by0 = (int)(max(y - r, 0LL) / kBrick), bx0 = (int)(max(x - r, 0LL) / kBrick);
        const int bz1 = (int)(min(z + r, d.Z - 1) / kBrick), by1 = (int)(min(y + r, d.Y - 1) / kBrick), bx1 = (int)(min(x + r, d.X - 1) / kBrick);
        const int nbz = bz1 - bz0 + 1, nby = by1 - by0 + 1, nbx = bx1 - bx0 + 1;      // <= 9 each for r <= 31
        __syncthreads();
        for (int i = threadIdx.x; i < nbz * nby; i += blockDim.x) {
            const int bz = bz0 + i / nby, by = by0 + i % nby;
            unsigned bits = 0;
            for (int k = 0; k < nbx; ++k)
                if (__ldg(g + ((size_t)bz * gy + by) * gx + bx0 + k) >= gcut) bits |= 1u << k;
            s_rowbits[(i / nby) * 10 + (i % nby)] = bits;
        }
        __syncthreads();
        for (int row = threadIdx.x; row < side * side; row += blockDim.x) {
            const int dz = row / side - r, dy = row % side - r;
            const int rem = r * r - dz * dz - dy * dy;
            if (rem < 0) continue;
            const long long zz = z + dz, yy = y + dy;
            if (zz < 0 || zz >= d.Z || yy < 0 || yy >= d.Y) continue;
            const unsigned bits = s_rowbits[((int)(zz / kBrick) - bz0) * 10 + ((int)(yy / kBrick) - by0)];
            if (!bits) continue;
            const int hw = isqrt_floor(rem);
            const long long x0 = x - hw < 0 ? 0 : x - hw;
            const long long x1 = x + hw >= d.X ? d.X - 1 : x + hw;
            const unsigned long long rowbase = ((unsigned long long)zz * d.Y + yy) * d.X;
            const unsigned long long q0 = rowbase + x0, q1 = rowbase + x1;
            unsigned long long *sup64 = reinterpret_cast<unsigned long long *>(sup);
            for (unsigned long long wd = q0 >> 6; wd <= (q1 >> 6); ++wd) {
                const unsigned long long lo = wd << 6;
                const unsigned b0 = q0 > lo ? (unsigned)(q0 - lo) : 0u;
                const unsigned b1 = q1 < lo + 63 ? (unsigned)(q1 - lo) : 63u;
                // bricks (in x) this word segment covers
                const int kx0 = (int)((lo + b0 - rowbase) / kBrick) - bx0, kx1 = (int)((lo + b1 - rowbase) / kBrick) - bx0;
                const unsigned seg = ((kx1 >= 31 ? 0xffffffffu : ((1u << (kx1 + 1)) - 1u)) & ~((1u << kx0) - 1u));
                if (!(bits & seg)) continue;
                const unsigned long long mask = (b1 == 63u ? ~0ULL : ((1ULL << (b1 + 1)) - 1ULL)) & ~((1ULL << b0) - 1ULL);
                atomicOr(&sup64[wd], mask);
            }
        }
    }
}

__global__ void approx_round_reset_kernel(Counters *cnt, ApproxState *S) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        cnt->n_cand = cnt->n_next; cnt->n_next = 0; cnt->n_work = 0; cnt->n_sel_round = 0; cnt->n_alive_owned = 0;
        S->n_amb = 0;
    }
}
__global__ void approx_first_reset_kernel(Counters *cnt, ApproxState *S) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        cnt->n_cand = 0; cnt->n_next = 0; cnt->n_work = 0; cnt->n_sel_round = 0; S->n_amb = 0;
    }
}
__global__ void approx_set_thresh_kernel(ThreshOut *t, double thresh, float v_lo, float v_hi) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { t->thresh = thresh; t->v_lo = v_lo; t->v_hi = v_hi; t->nan_count = 0; }
}

struct ApproxHost {                 // pinned scratch layout
    Counters cnt;
    ApproxState st;
};
__global__ void approx_collect_kernel(const Counters *cnt, const ApproxState *S, ApproxHost *out) {
    if (threadIdx.x || blockIdx.x) return;
    out->cnt = *cnt; out->st = *S;
}

template <int LW, bool IMM = false>
static int launch_gauss32(fpl_ctx *ctx, const float *in, float *out, float *sample, Dims d, int sy, int sx,
                          ApproxState *state, cudaStream_t st) {
    using C = GaussCfg<LW>;
    static bool attr_done = false;
    if (!attr_done) {
        FPL_CUDA_CHECK(cudaFuncSetAttribute(gauss32_kernel<LW, IMM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        attr_done = true;
    }
    const long long txn = (d.X + kGT - 1) / kGT, tyn = (d.Y + kGT - 1) / kGT;
    // z chunks: enough blocks for a few waves, but chunks long enough that the 2 lw warm-up planes stay cheap
    long long want = (8LL * ctx->sm_count + txn * tyn - 1) / (txn * tyn);
    long long zc = (d.Z + want - 1) / want;
    if (zc < 48) zc = 48;
    if (zc > 256) zc = 256;
    if (zc > d.Z) zc = d.Z;
    zc = (zc + kApxSampleStep - 1) / kApxSampleStep * kApxSampleStep;
    const long long nzc = (d.Z + zc - 1) / zc;
    FPL_REQUIRE(tyn <= 65535 && nzc <= 65535 && txn < 2147483647LL && d.X < (1LL << 30) && d.Y < (1LL << 30) && d.Z < (1LL << 30),
                "gauss32: volume too large for one launch");
    const int vec_ok = (d.X % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
    gauss32_kernel<LW, IMM><<<dim3((unsigned)txn, (unsigned)tyn, (unsigned)nzc), 256, C::SMEM, st>>>(in, out, sample, d, (int)zc, sy, sx,
                                                                                         vec_ok, state);
    FPL_LAUNCH_CHECK(ctx);
    return FPL_OK;
}

static bool approx_lw_supported(int lw) { return lw == 2 || lw == 3 || lw == 4 || lw == 8 || lw == 10; }

static int g_gauss_imm = 1;     // test / experiment hook (fpl_debug_gauss_imm): 0 = never use the immediate-tap instantiation

static int launch_gauss32_any(fpl_ctx *ctx, int lw, const float *w32, const float *in, float *out, float *sample, Dims d,
                              int sy, int sx, ApproxState *state, cudaStream_t st) {
    static const int env_imm = getenv("FPL_GAUSS_IMM") ? atoi(getenv("FPL_GAUSS_IMM")) : -1;     // A/B measurements
    if (lw == 10 && (env_imm >= 0 ? env_imm : g_gauss_imm)) {     // the reference's sigma = 5: taps as immediates, when they are exactly those
        bool same = true;
        for (int i = 0; i < 21; ++i) same = same && memcmp(&w32[i], &kTapsSigma5[i], 4) == 0;
        if (same) return launch_gauss32<10, true>(ctx, in, out, sample, d, sy, sx, state, st);
    }
    switch (lw) {
        case 2:  return launch_gauss32<2>(ctx, in, out, sample, d, sy, sx, state, st);
        case 3:  return launch_gauss32<3>(ctx, in, out, sample, d, sy, sx, state, st);
        case 4:  return launch_gauss32<4>(ctx, in, out, sample, d, sy, sx, state, st);
        case 8:  return launch_gauss32<8>(ctx, in, out, sample, d, sy, sx, state, st);
        case 10: return launch_gauss32<10>(ctx, in, out, sample, d, sy, sx, state, st);
        default: break;
    }
    fpl::set_error("gauss32: unsupported half width %d", lw);
    return FPL_EINVAL;
}

// NumPy _lerp in float32 + np.maximum(., thd), on the host (same operations as threshold_kernel)
static double approx_threshold_value(float a, float b, float gamma, double thd) {
    volatile float dlt = b - a;
    volatile float prod = dlt * gamma;
    volatile float res = a + prod;
    if (gamma >= 0.5f) {
        volatile float omg = 1.0f - gamma;
        volatile float prod2 = dlt * omg;
        res = b - prod2;
    }
    const double p = (double)res;
    if (p != p || thd != thd) return nan("");
    return p > thd ? p : thd;
}

static size_t approx_workspace_bytes(int64_t Z, int64_t Y, int64_t X, long long band_cap, long long narrow_cap) {
    const size_t n = (size_t)Z * Y * X;
    const size_t ns = (size_t)((Z + kApxSampleStep - 1) / kApxSampleStep) * ((Y + 3) / 4) * ((X + 3) / 4);
    return n * 4 + ns * 4 + (size_t)band_cap * 12 + (size_t)narrow_cap * 12 + (size_t)kApxAmbCap * 8 +
           sizeof(ApproxState) + sizeof(ApproxHost) + (size_t)kExactBatch * (21 * 21 + 21) * 4 + 16 * 256 + 4096;
}

// The two-tier path of fpl_voxel2obj.  *done = false (with FPL_OK) when the map or the parameters do not qualify, or
// when a certificate fails: the caller then runs the exact path.  Never returns a result it has not certified.
// why the two-tier path handed the call to the exact path (fpl_debug_v2o_decline_reason; 0 = it did not):
// 1 parameters / shape, 2 taps, 3 percentile at or near zero, 4 rank beyond the interior, 5 sample band touches an end,
// 6 inputs negative / non-finite / huge or cut-off too small or list overflow, 16 the sample predicts a band list
// overflow, 17 .. too many exact recomputations (values crowd around the percentile within the bound), 7 band certificate, 8 narrow list
// overflow, 9 narrow band leaves the listed band, 10 narrow rank certificate, 11 NaN threshold, 12.. overflow in a round
static int decline(fpl_ctx *ctx, int code) {
    ctx->v2o_decline = code;
    const ApproxHost *h = (const ApproxHost *)ctx->h_pinned;     // last snapshot (diagnosis only)
    ctx->v2o_info[0] = (long long)h->st.n_below; ctx->v2o_info[1] = (long long)h->st.n_band;
    ctx->v2o_info[2] = (long long)h->st.n_narrow; ctx->v2o_info[3] = (long long)h->st.n_below_narrow;
    ctx->v2o_info[4] = (long long)h->st.overflow; ctx->v2o_info[5] = (long long)h->cnt.overflow;
    ctx->v2o_info[6] = (long long)h->st.bad_bits; ctx->v2o_info[7] = (long long)(h->st.cutA * 1e9f);
    return FPL_OK;
}

static int voxel2obj_approx(fpl_ctx *ctx, const float *d_pred, int64_t Z, int64_t Y, int64_t X, const fpl_v2o_params *p,
                            SelectState *d_states, ThreshOut *d_tout, long long list_cap, double *d_dets,
                            int64_t capacity, int64_t *h_count, double *h_threshold, int64_t *h_stats,
                            cudaStream_t st, bool *done) {
    *done = false;
    ctx->v2o_decline = 0;
    const long long n = Z * Y * X;
    const int r = p->obj_min_dist, lw = p->lw;
    if (r <= 0 || r > 27 + 4 || lw < 0 || lw > r || !approx_lw_supported(lw) || n < 32768 || p->thd != p->thd) return decline(ctx, 1);
    Taps taps;
    memset(&taps, 0, sizeof(taps));
    float w32[2 * kApxMaxLw + 1];
    for (int i = 0; i < 2 * lw + 1; ++i) {
        taps.w[i] = p->h_weights[i];
        if (!(taps.w[i] >= 0.0) || !(taps.w[i] <= 1.0)) return decline(ctx, 2);
        w32[i] = (float)taps.w[i];
    }
    const unsigned long long n_pad = (unsigned long long)(Z + 2 * r) * (Y + 2 * r) * (X + 2 * r);
    const unsigned long long extra = n_pad - (unsigned long long)n;
    const int ns = p->rank_hi != p->rank_lo ? 2 : 1;
    // ranks among the interior values (the border zeros are the smallest values of a non-negative map)
    if ((unsigned long long)p->rank_lo < extra + 16) return decline(ctx, 3);         // percentile at / near zero: exact path
    const unsigned long long k_lo = (unsigned long long)p->rank_lo - extra, k_hi = (unsigned long long)p->rank_hi - extra;
    if (k_hi >= (unsigned long long)n) return decline(ctx, 4);
    Dims d{Z, Y, X};
    const int sy = (int)((Y + 3) / 4), sx = (int)((X + 3) / 4);
    const long long n_s = ((Z + kApxSampleStep - 1) / kApxSampleStep) * (long long)sy * sx;
    // sample ranks: +- 6 sigma of the binomial rank fluctuation (+ slack); a miss is caught by the certificate
    const double frac_lo = (double)k_lo / (double)n, frac_hi = (double)k_hi / (double)n;
    const double sig = sqrt((double)n_s * frac_lo * (1.0 - frac_lo));
    long long r_lo = (long long)floor(frac_lo * (double)(n_s - 1) - 6.0 * sig - 4.0);
    long long r_hi = (long long)ceil(frac_hi * (double)(n_s - 1) + 6.0 * sig + 4.0);
    if (r_lo < 1 || r_hi > n_s - 2) return decline(ctx, 5);                           // band would touch an end of the sample
    long long band_cap = (long long)(4.0 * (double)(r_hi - r_lo + 1) / (double)n_s * (double)n) + 65536;
    if (band_cap > n) band_cap = n;
    const long long narrow_cap = 1 << 18;
    long long det_cap = capacity > 0 ? capacity : 1;

    float *A = (float *)ctx->arena.take((size_t)n * 4);
    float *sample = (float *)ctx->arena.take((size_t)n_s * 4);
    unsigned long long *band_idx = (unsigned long long *)ctx->arena.take((size_t)band_cap * 8);
    float *band_val = (float *)ctx->arena.take((size_t)band_cap * 4);
    unsigned long long *nar_idx = (unsigned long long *)ctx->arena.take((size_t)narrow_cap * 8);
    float *nar_val = (float *)ctx->arena.take((size_t)narrow_cap * 4);
    unsigned long long *amb_idx = (unsigned long long *)ctx->arena.take((size_t)kApxAmbCap * 8);
    ApproxState *S = (ApproxState *)ctx->arena.take(sizeof(ApproxState));
    ApproxHost *d_host = (ApproxHost *)ctx->arena.take(sizeof(ApproxHost));
    const int W = 2 * lw + 1;
    float *ex_t1 = (float *)ctx->arena.take((size_t)kExactBatch * W * W * 4);
    float *ex_t2 = (float *)ctx->arena.take((size_t)kExactBatch * W * 4);
    FPL_REQUIRE(A && sample && band_idx && band_val && nar_idx && nar_val && amb_idx && S && d_host && ex_t1 && ex_t2 &&
                sizeof(ApproxHost) <= 1024, "voxel2obj: workspace sizing error (two-tier path)");
    DetectBuffers B;
    FPL_TRY(take_detect_buffers(ctx, Z, Y, X, list_cap, det_cap, B, st));
    ApproxHost *h = (ApproxHost *)ctx->h_pinned;
    const int grid_stream = ctx->sm_count * 8;
    const size_t ex_smem = sizeof(float) * (size_t)((2 * lw + 1) * (2 * lw + 1) + (2 * lw + 1));
    auto collect = [&]() -> int {
        approx_collect_kernel<<<1, 32, 0, st>>>(B.cnt, S, d_host);
        FPL_LAUNCH_CHECK(ctx);
        FPL_CUDA_CHECK(cudaMemcpyAsync(h, d_host, sizeof(ApproxHost), cudaMemcpyDeviceToHost, st));
        FPL_CUDA_CHECK(cudaStreamSynchronize(st));
        return FPL_OK;
    };

    // ---- tier 1: A, sample, band -----------------------------------------------------------------------------
    {
        fpl::ProfScope prof(ctx, st, fpl::PROF_GAUSS, 8.0 * (double)n);
        FPL_CUDA_CHECK(cudaMemsetAsync(S, 0, sizeof(ApproxState), st));
        FPL_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_gw32, w32, sizeof(float) * (2 * lw + 1), 0, cudaMemcpyHostToDevice, st));
        FPL_TRY(launch_gauss32_any(ctx, lw, w32, d_pred, A, sample, d, sy, sx, S, st));
    }
    {
        fpl::ProfScope prof(ctx, st, fpl::PROF_SELECT, 4.0 * (double)n);
        FPL_TRY(select_rank(ctx, sample, n_s, 0ULL, (unsigned long long)r_lo, &d_states[0], st));
        FPL_TRY(select_rank(ctx, sample, n_s, 0ULL, (unsigned long long)r_hi, &d_states[1], st));
        approx_band_kernel<<<1, 32, 0, st>>>(&d_states[0], &d_states[1], p->thd, S);
        FPL_LAUNCH_CHECK(ctx);
        approx_sample_count_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(sample, n_s, S, &S->n_narrow);     // (n_narrow is unused until later)
        FPL_LAUNCH_CHECK(ctx);
        FPL_TRY(collect());
        if (h->st.bad_bits >= kApxBadBits || !(h->st.cutA >= kApxMinCut)) return decline(ctx, 6);
        if ((double)h->st.n_narrow * 64.0 * 1.5 > (double)band_cap) return decline(ctx, 16);     // band list would overflow
        {   // the voxels within 3 eps of the percentile are recomputed exactly (~25 ns each) and tight value distributions
            // also mean many undecidable ball comparisons: predict the size of that narrow band from the sample (its share
            // of the listed band = ratio of the widths) and leave maps that crowd around the percentile to the exact path
            const double band_w = (double)h->st.Hb - (double)h->st.Lb;
            const double narrow_w = 6.0 * (double)kApxEps * 0.5 * ((double)h->st.Hb + (double)h->st.Lb);
            const double pred = (double)h->st.n_narrow * 64.0 * (band_w > 0.0 ? (narrow_w < band_w ? narrow_w / band_w : 1.0) : 1.0);
            const double budget = (double)n / 16384.0 > 8192.0 ? (double)n / 16384.0 : 8192.0;
            if (pred > budget) return decline(ctx, 17);
        }
        FPL_CUDA_CHECK(cudaMemsetAsync(&S->n_narrow, 0, sizeof(unsigned long long), st));
        FPL_CUDA_CHECK(cudaMemsetAsync(B.grid, 0, brick_grid_bytes(Z, Y, X), st));
        const long long items = Z * ((Y + kP1Strip - 1) / kP1Strip) * ((X + kP1Cols - 1) / kP1Cols);
        long long g1 = (items + 3) / 4;
        if (g1 > (long long)ctx->sm_count * 64) g1 = (long long)ctx->sm_count * 64;
        if ((X % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0))
            approx_pass1_kernel<true><<<(unsigned)g1, 128, 0, st>>>(A, d, B.gy, B.gx, B.grid, S, band_idx, band_val, band_cap, B.w_idx,
                                                                    B.w_val, B.list_cap, B.cnt, 0, (int)Z);
        else
            approx_pass1_kernel<false><<<(unsigned)g1, 128, 0, st>>>(A, d, B.gy, B.gx, B.grid, S, band_idx, band_val, band_cap, B.w_idx,
                                                                     B.w_val, B.list_cap, B.cnt, 0, (int)Z);
        FPL_LAUNCH_CHECK(ctx);
        FPL_TRY(collect());
        h->st.n_below = (unsigned long long)n - h->st.n_ge;        // every interior voxel is either below the band or counted
    }
    if (h->st.bad_bits >= kApxBadBits || !(h->st.cutA >= kApxMinCut) || h->st.overflow || h->cnt.overflow) return decline(ctx, 6);
    // certificate of the band: both ranks fall inside the listed voxels
    if (k_lo < h->st.n_below || k_hi - h->st.n_below >= h->st.n_band) return decline(ctx, 7);
    const unsigned long long j_lo = k_lo - h->st.n_below, j_hi = k_hi - h->st.n_below;
    float s_lo = 0.f, s_hi = 0.f;
    {
        fpl::ProfScope prof(ctx, st, fpl::PROF_SELECT, 0.0);
        FPL_TRY(select_rank(ctx, band_val, (long long)h->st.n_band, 0ULL, j_lo, &d_states[0], st));
        SelectState *hi_state = &d_states[0];
        if (ns > 1) {
            hi_state = &d_states[1];
            FPL_TRY(select_rank(ctx, band_val, (long long)h->st.n_band, 0ULL, j_hi, hi_state, st));
        }
        approx_narrow_setup_kernel<<<1, 32, 0, st>>>(&d_states[0], hi_state, S);
        FPL_LAUNCH_CHECK(ctx);
        int nb = (int)((h->st.n_band + 255) / 256); if (nb > grid_stream) nb = grid_stream; if (nb < 1) nb = 1;
        approx_narrow_kernel<<<nb, 256, 0, st>>>(band_idx, band_val, S, nar_idx, narrow_cap);
        FPL_LAUNCH_CHECK(ctx);
        FPL_TRY(collect());
        if (h->st.overflow || h->st.n_narrow > (unsigned long long)narrow_cap) return decline(ctx, 8);
        FPL_TRY(exact_list(ctx, d_pred, d, lw, taps, nar_idx, (long long)h->st.n_narrow, nar_val, ex_t1, ex_t2, st));
        // the narrow band must lie inside the listed band (else voxels outside the list could belong to it)
        if (!(h->st.n_lo >= h->st.Lb) || !(h->st.n_hi <= h->st.Hb)) return decline(ctx, 9);
        const unsigned long long below = h->st.n_below_narrow;
        if (j_lo < below || j_hi - below >= h->st.n_narrow) return decline(ctx, 10);
        std::vector<float> vals((size_t)h->st.n_narrow);
        FPL_CUDA_CHECK(cudaMemcpyAsync(vals.data(), nar_val, sizeof(float) * vals.size(), cudaMemcpyDeviceToHost, st));
        FPL_CUDA_CHECK(cudaStreamSynchronize(st));
        std::sort(vals.begin(), vals.end());
        s_lo = vals[(size_t)(j_lo - below)];
        s_hi = vals[(size_t)(j_hi - below)];
    }
    const double threshold = approx_threshold_value(s_lo, s_hi, p->gamma, p->thd);
    if (threshold != threshold) return decline(ctx, 11);
    if (h_threshold) *h_threshold = threshold;
    approx_set_thresh_kernel<<<1, 32, 0, st>>>(d_tout, threshold, s_lo, s_hi);
    FPL_LAUNCH_CHECK(ctx);

    // ---- NMS rounds with certified comparisons -----------------------------------------------------------------
    long long rounds = 0;
    unsigned long long n_first = 0, n_amb_total = 0;
    {
        fpl::ProfScope prof(ctx, st, fpl::PROF_NMS, 4.0 * (double)n);
        unsigned long long *a_idx = B.a_idx, *b_idx = B.b_idx;
        float *a_val = B.a_val, *b_val = B.b_val;
        unsigned long long remaining = 1;           // round 1 works on the worklist of the dense pass
        bool first = true;
        while (remaining > 0) {
            ++rounds;
            if (!first) {
                long long fblocks = (long long)((remaining + 255) / 256);
                if (fblocks > grid_stream) fblocks = grid_stream;
                approx_filter_kernel<<<(unsigned)fblocks, 256, 0, st>>>(A, B.sup, d, a_idx, a_val, b_idx, b_val, B.w_idx, B.w_val,
                                                                        B.list_cap, B.cnt);
                FPL_LAUNCH_CHECK(ctx);
            }
            approx_ballcheck_kernel<false><<<ctx->sm_count * 8, 256, 0, st>>>(A, B.sup, d, r, B.grid, B.gy, B.gx, B.w_idx, B.w_val, B.det_idx,
                                                                       B.det_val, B.sel_idx, B.det_cap, B.cnt, S, amb_idx, 0ULL, ~0ULL);
            FPL_LAUNCH_CHECK(ctx);
            approx_resolve_kernel<<<ctx->sm_count * 2, 256, ex_smem, st>>>(d_pred, A, B.sup, d, r, lw, taps, B.grid, B.gy, B.gx, amb_idx,
                                                                           B.det_idx, B.det_val, B.sel_idx, B.det_cap, B.cnt, S);
            FPL_LAUNCH_CHECK(ctx);
            approx_suppress_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(B.sup, d, r, B.grid, B.gy, B.gx, S, B.sel_idx, B.cnt);
            FPL_LAUNCH_CHECK(ctx);
            if (first) {
                approx_collect_kernel<<<1, 32, 0, st>>>(B.cnt, S, d_host);
                FPL_LAUNCH_CHECK(ctx);
                approx_first_reset_kernel<<<1, 32, 0, st>>>(B.cnt, S);
                FPL_LAUNCH_CHECK(ctx);
                int gc = ctx->sm_count * 8;
                approx_compact_kernel<<<gc, 256, 0, st>>>(A, d, B.grid, B.gz, B.gy, B.gx, B.sup, S, a_idx, a_val, B.list_cap, B.cnt);
                FPL_LAUNCH_CHECK(ctx);
                FPL_CUDA_CHECK(cudaMemcpyAsync(h, d_host, sizeof(ApproxHost), cudaMemcpyDeviceToHost, st));
                FPL_CUDA_CHECK(cudaMemcpyAsync(&h->cnt.n_cand, &B.cnt->n_cand, 8, cudaMemcpyDeviceToHost, st));
                FPL_CUDA_CHECK(cudaMemcpyAsync(&h->cnt.overflow, &B.cnt->overflow, 8, cudaMemcpyDeviceToHost, st));
                FPL_CUDA_CHECK(cudaStreamSynchronize(st));
                if (h->st.overflow) return decline(ctx, 12);
                if (h->cnt.overflow) {
                    if (h->cnt.n_det > (unsigned long long)det_cap) {
                        fpl::set_error("voxel2obj: detection capacity %lld too small (need > %llu)", (long long)capacity, h->cnt.n_det);
                        return FPL_EOVERFLOW;
                    }
                    return decline(ctx, 13);
                }
                n_first = h->cnt.n_det;
                n_amb_total += h->st.n_amb;
                remaining = h->cnt.n_cand;
                first = false;
                continue;
            }
            approx_collect_kernel<<<1, 32, 0, st>>>(B.cnt, S, d_host);
            FPL_LAUNCH_CHECK(ctx);
            approx_round_reset_kernel<<<1, 32, 0, st>>>(B.cnt, S);
            FPL_LAUNCH_CHECK(ctx);
            FPL_CUDA_CHECK(cudaMemcpyAsync(h, d_host, sizeof(ApproxHost), cudaMemcpyDeviceToHost, st));
            FPL_CUDA_CHECK(cudaStreamSynchronize(st));
            if (h->st.overflow) return decline(ctx, 14);
            if (h->cnt.overflow) {
                if (h->cnt.n_det > (unsigned long long)det_cap) {
                    fpl::set_error("voxel2obj: detection capacity %lld too small (need > %llu)", (long long)capacity, h->cnt.n_det);
                    return FPL_EOVERFLOW;
                }
                return decline(ctx, 15);
            }
            if (h->cnt.n_sel_round == 0 && h->cnt.n_next > 0) {
                fpl::set_error("voxel2obj: NMS round made no progress (internal error, two-tier path)");
                return FPL_ECUDA;
            }
            n_amb_total += h->st.n_amb;
            remaining = h->cnt.n_next;
            unsigned long long *ti = a_idx; a_idx = b_idx; b_idx = ti;
            float *tv = a_val; a_val = b_val; b_val = tv;
            if (h->cnt.n_next == h->cnt.n_sel_round) remaining = 0;
        }
        // tier 2 for the detections: exact confidences; finish_rows drops the selected points that are not above the threshold
        const unsigned long long n_det = h->cnt.n_det;
        if (n_det > 0) FPL_TRY(exact_list(ctx, d_pred, d, lw, taps, B.det_idx, (long long)n_det, B.det_val, ex_t1, ex_t2, st));
        unsigned long long n_rows = 0;
        FPL_TRY(finish_detections(ctx, B, n_det, d, p, d_dets, capacity, d_tout, &n_rows, st));
        if (h_count) *h_count = (int64_t)n_rows;
        if (h_stats) {
            h_stats[0] = (int64_t)h->st.n_band;
            h_stats[1] = rounds;
            h_stats[2] = (int64_t)h->cnt.ball_checks;
            h_stats[3] = (int64_t)n_det;
            h_stats[4] = (int64_t)h->st.n_narrow; h_stats[5] = (int64_t)n_first; h_stats[6] = 2; h_stats[7] = (int64_t)n_amb_total;
        }
    }
    *done = true;
    return FPL_OK;
}
