// detect_seg.cu -- segmentation-aware greedy suppression: the seg / seg_dilate / seg_force branch of voxel2obj
// (flypylib/fplobjdetect.py:161-165, 192-195, 213-224).
//
// With a segmentation the suppression set of a selected point q is no longer a ball: it is
//     ball_r(q)  AND  dilate_D( seg == seg[q] inside the (2r+1)^3 cube of q )      [OR  ball_f(q) when seg_force = f]
// (binary_dilation with SciPy's default 6-connected structure, D = seg_dilate iterations, border value 0 at the faces
// of the cube; voxels outside the volume carry label 0 = the reference's zero padding).  The relation is not
// symmetric, so the rounds formulation of the plain path does not apply; the loop runs as the reference writes it,
// sequentially over the candidates in (value desc, index asc) order -- but entirely on the device, in ONE persistent
// CTA: per selected point the cube mask is built as 64-bit row masks in shared memory (r <= 31), dilated in place
// (bit shifts along x, neighbouring rows / planes along y / z), intersected with the ball rows and OR-ed into a 1 bit
// per voxel suppression map; the next point is the first candidate of the sorted list whose bit is still clear.
#include "common.cuh"

namespace fpl {
namespace v2oseg {

constexpr int kThreads = 1024;

struct Args {
    const float *val;            // candidate values, sorted (value desc, flat index asc)
    const long long *idx;        // flat interior index (z*Y + y)*X + x of every candidate
    long long n_cand;
    const long long *seg;        // (Z,Y,X) labels or nullptr
    unsigned int *supp;          // 1 bit per interior voxel, zero on entry
    long long Z, Y, X;
    int r, dilate, force;        // dilate < 0: no dilation; force <= 0: no forced inner ball
    int bx, by, bz;              // buffer (x,y,z)
    double ox, oy, oz;           // volume offset
    double *rows; long long cap; // output (x, y, z, conf)
    long long *out_count;        // [0] rows written, [1] points selected (before the buffer crop), [2] overflow flag
};

__device__ __forceinline__ bool is_supp(const unsigned int *supp, long long i) {
    return (__ldcg(supp + (i >> 5)) >> (i & 31)) & 1u;
}

__global__ void __launch_bounds__(kThreads, 1)
seg_greedy_kernel(const Args a) {
    extern __shared__ unsigned long long s_rows[];          // 2 x S*S row masks
    __shared__ long long s_sel;
    const int S = 2 * a.r + 1, nrows = S * S;
    unsigned long long *cur = s_rows, *nxt = s_rows + nrows;
    const unsigned long long full = S == 64 ? ~0ull : ((1ull << S) - 1ull);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    long long cursor = 0, n_rows_out = 0, n_sel = 0;
    const long long r2 = (long long)a.r * a.r, f2 = (long long)a.force * a.force;
    for (;;) {
        // ---- next candidate whose voxel is not suppressed: 1024 entries of the sorted list per step
        for (;;) {
            if (tid == 0) s_sel = -1;                            // = ULLONG_MAX for the unsigned minimum below
            __syncthreads();
            const long long i = cursor + tid;
            if (i < a.n_cand && !is_supp(a.supp, a.idx[i])) atomicMin((unsigned long long *)&s_sel, (unsigned long long)i);
            __syncthreads();
            const long long found = s_sel;
            __syncthreads();                                     // everybody has read s_sel before it is reset
            if (found >= 0) { cursor = found; break; }
            cursor += kThreads;
            if (cursor >= a.n_cand) { cursor = -1; break; }
        }
        if (cursor < 0) break;
        const float v = a.val[cursor];
        if (!(v > 0.f)) break;                               // fplobjdetect.py:204-205
        const long long q = a.idx[cursor];
        const long long qz = q / (a.Y * a.X), qy = (q / a.X) % a.Y, qx = q % a.X;
        ++n_sel;
        if (tid == 0) {
            const bool keep = qx >= a.bx && qy >= a.by && qz >= a.bz && qx < a.X - a.bx && qy < a.Y - a.by && qz < a.Z - a.bz;
            if (keep) {
                if (n_rows_out < a.cap) {
                    double *o = a.rows + 4 * n_rows_out;
                    o[0] = (double)qx + a.ox; o[1] = (double)qy + a.oy; o[2] = (double)qz + a.oz; o[3] = (double)v;
                } else {
                    a.out_count[2] = 1;
                }
            }
        }
        {
            const bool keep = qx >= a.bx && qy >= a.by && qz >= a.bz && qx < a.X - a.bx && qy < a.Y - a.by && qz < a.Z - a.bz;
            if (keep && n_rows_out < a.cap) ++n_rows_out;
        }
        // ---- cube mask: seg == seg[q] (voxels outside the volume carry label 0)
        if (a.seg) {
            const long long id = a.seg[q];
            for (int row = warp; row < nrows; row += kThreads / 32) {
                const long long z = qz + row / S - a.r, y = qy + row % S - a.r;
                const bool row_in = z >= 0 && z < a.Z && y >= 0 && y < a.Y;
                unsigned long long m = 0;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int dx = lane + 32 * h;
                    bool hit = false;
                    if (dx < S) {
                        const long long x = qx + dx - a.r;
                        const long long lab = (row_in && x >= 0 && x < a.X) ? __ldg(a.seg + (z * a.Y + y) * a.X + x) : 0;
                        hit = lab == id;
                    }
                    const unsigned int b = __ballot_sync(0xffffffffu, hit);
                    m |= (unsigned long long)b << (32 * h);
                }
                if (lane == 0) cur[row] = m & full;
            }
            __syncthreads();
            for (int itn = 0; itn < a.dilate; ++itn) {
                for (int row = tid; row < nrows; row += kThreads) {
                    const int dz = row / S, dy = row % S;
                    unsigned long long m = cur[row];
                    m |= (m << 1) | (m >> 1);
                    if (dy > 0) m |= cur[row - 1];
                    if (dy < S - 1) m |= cur[row + 1];
                    if (dz > 0) m |= cur[row - S];
                    if (dz < S - 1) m |= cur[row + S];
                    nxt[row] = m & full;
                }
                __syncthreads();
                unsigned long long *t = cur; cur = nxt; nxt = t;
            }
        }
        // ---- suppression rows: (mask AND ball) OR forced inner ball -> global bit map
        for (int row = tid; row < nrows; row += kThreads) {
            const int dz = row / S - a.r, dy = row % S - a.r;
            const long long z = qz + dz, y = qy + dy;
            if (z < 0 || z >= a.Z || y < 0 || y >= a.Y) continue;
            const long long rem = r2 - (long long)dz * dz - (long long)dy * dy;
            unsigned long long bits = 0;
            if (rem >= 0) {
                int w = (int)sqrtf((float)rem);
                while ((long long)(w + 1) * (w + 1) <= rem) ++w;
                while ((long long)w * w > rem) --w;
                unsigned long long ball = (w >= 31 ? ~0ull : ((1ull << (2 * w + 1)) - 1ull)) << (a.r - w);
                bits = a.seg ? (cur[row] & ball) : ball;
            }
            if (a.force > 0) {
                const long long remf = f2 - (long long)dz * dz - (long long)dy * dy;
                if (remf >= 0) {
                    int w = (int)sqrtf((float)remf);
                    while ((long long)(w + 1) * (w + 1) <= remf) ++w;
                    while ((long long)w * w > remf) --w;
                    bits |= ((1ull << (2 * w + 1)) - 1ull) << (a.r - w);
                }
            }
            bits &= full;
            // clip to the volume in x and scatter into 32-bit words
            const long long x0 = qx - a.r;
            for (int part = 0; part < 2 && bits; ++part) {
                unsigned int chunk = (unsigned int)(bits >> (32 * part));
                while (chunk) {
                    const int b = __ffs(chunk) - 1;
                    // run of set bits starting at b
                    unsigned int run = chunk >> b;
                    const int len = run == 0xffffffffu ? 32 - b : __ffs(~run) - 1;
                    chunk = (len + b >= 32) ? 0u : (chunk & ~(((1u << len) - 1u) << b));
                    long long xs = x0 + 32 * part + b, xe = xs + len;          // [xs, xe)
                    if (xs < 0) xs = 0;
                    if (xe > a.X) xe = a.X;
                    if (xs >= xe) continue;
                    long long g0 = (z * a.Y + y) * a.X + xs, g1 = g0 + (xe - xs);     // bit range in the map
                    while (g0 < g1) {
                        const long long wi = g0 >> 5;
                        const int lo = (int)(g0 & 31);
                        const long long wend = (wi + 1) << 5;
                        const int hi = (int)((g1 < wend ? g1 : wend) - (wi << 5));   // exclusive
                        const unsigned int mask = (hi == 32 ? ~0u : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
                        atomicOr(a.supp + wi, mask);
                        g0 = wend;
                    }
                }
            }
        }
        __threadfence();
        __syncthreads();
        ++cursor;
        if (cursor >= a.n_cand) break;
    }
    if (tid == 0) { a.out_count[0] = n_rows_out; a.out_count[1] = n_sel; }
}

}  // namespace v2oseg
}  // namespace fpl

extern "C" {

// Stage C of voxel2obj with a segmentation (not part of the plain hot path).  d_val / d_idx: the candidates
// (smoothed value > threshold) sorted by (value desc, flat interior index asc); d_seg: (Z,Y,X) int64 labels or NULL
// (then the suppression set is the plain ball: the reference's loop with seg=None); d_supp: ceil(Z*Y*X/32) zeroed
// uint32; seg_dilate < 0: no dilation (seg_dilate=None); seg_force <= 0: none.  d_rows: capacity x 4 doubles
// (x, y, z, conf) after un-padding, buffer crop and offset (fplobjdetect.py:233-253); d_count: 3 int64 (rows,
// selected before the crop, overflow flag).  Does not synchronise.
int fpl_v2o_detect_seg(fpl_ctx *ctx, const float *d_val, const int64_t *d_idx, int64_t n_cand, const int64_t *d_seg,
                       uint32_t *d_supp, int64_t Z, int64_t Y, int64_t X, const fpl_v2o_params *p, int32_t seg_dilate,
                       int32_t seg_force, double *d_rows, int64_t capacity, int64_t *d_count, void *stream) {
    FPL_REQUIRE(ctx && d_val && d_idx && d_supp && p && d_rows && d_count, "fpl_v2o_detect_seg: NULL argument");
    FPL_REQUIRE(p->obj_min_dist >= 0 && p->obj_min_dist <= 31,
                "fpl_v2o_detect_seg: obj_min_dist must be <= 31 (64-bit row masks), got %d", p->obj_min_dist);
    FPL_REQUIRE(seg_force <= p->obj_min_dist, "fpl_v2o_detect_seg: seg_force must not exceed obj_min_dist");
    FPL_REQUIRE(Z > 0 && Y > 0 && X > 0 && n_cand >= 0, "fpl_v2o_detect_seg: bad extents");
    FPL_CUDA_CHECK(cudaSetDevice(ctx->device));
    fpl::v2oseg::Args a{};
    a.val = d_val; a.idx = (const long long *)d_idx; a.n_cand = n_cand; a.seg = (const long long *)d_seg; a.supp = d_supp;
    a.Z = Z; a.Y = Y; a.X = X; a.r = p->obj_min_dist; a.dilate = seg_dilate; a.force = seg_force;
    a.bx = p->buffer_xyz[0]; a.by = p->buffer_xyz[1]; a.bz = p->buffer_xyz[2];
    a.ox = p->offset_xyz[0]; a.oy = p->offset_xyz[1]; a.oz = p->offset_xyz[2];
    a.rows = d_rows; a.cap = capacity; a.out_count = (long long *)d_count;
    const int S = 2 * a.r + 1;
    const size_t smem = (size_t)2 * S * S * sizeof(unsigned long long);
    FPL_CUDA_CHECK(cudaFuncSetAttribute(fpl::v2oseg::seg_greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 63 * 63 * 8));
    cudaStream_t st = (cudaStream_t)stream;
    FPL_CUDA_CHECK(cudaMemsetAsync(d_count, 0, 3 * sizeof(int64_t), st));
    fpl::v2oseg::seg_greedy_kernel<<<1, fpl::v2oseg::kThreads, smem, st>>>(a);
    FPL_LAUNCH_CHECK(ctx);
    return FPL_OK;
}

}  // extern "C"
