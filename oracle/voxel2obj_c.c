/* TEST INFRASTRUCTURE ONLY (oracle).  Plain-C restatement of the two expensive stages of
 * flypylib's voxel2obj (reference: flypylib/fplobjdetect.py:158-175 smoothing of the padded
 * map, :184-231 greedy non-max suppression; ball footprint flypylib/fplutils.py:14-22).
 * Never linked into, loaded by, or called from the product library.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC   (see oracle/Makefile).
 * -ffp-contract=off matters: SciPy's NI_Correlate1D (symmetric branch) performs a separate
 * double multiply and double add per tap pair; the installed _nd_image.so has no FMA.
 *
 * Stage 1  fpl_oracle_smooth_padded: np.pad(pred, r) -> scipy gaussian_filter(sigma,
 *          truncate=2.0) [axis 0,1,2; double line buffer; 'reflect'; float32 store per
 *          axis] -> zero the r-wide border on all six faces.
 * Stage 2  fpl_oracle_greedy: candidates = s > thresh; repeatedly emit the best remaining
 *          valid candidate (largest value, lowest flat index on ties), clear the ball
 *          dz^2+dy^2+dx^2 <= r^2 around it in the validity map; stop at value <= 0.
 *          Implemented by visiting candidates in sorted priority order, which selects
 *          exactly the points the reference loop selects, in the same order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline int64_t reflect_idx(int64_t i, int64_t n)
{
    int64_t period = 2 * n;
    int64_t m = i % period;
    if (m < 0) m += period;
    return m >= n ? period - 1 - m : m;
}

/* one axis pass over a (n0,n1,n2) float32 array, in place, along `axis` */
static void pass_axis(float *a, int64_t n0, int64_t n1, int64_t n2, int axis,
                      const double *w, int lw)
{
    int64_t dims[3] = {n0, n1, n2};
    int64_t strides[3] = {n1 * n2, n2, 1};
    int64_t n = dims[axis], st = strides[axis];
    int o1 = (axis + 1) % 3, o2 = (axis + 2) % 3;
    int64_t nlines = dims[o1] * dims[o2];
#pragma omp parallel
    {
        double *line = (double *)malloc(sizeof(double) * (size_t)(n + 2 * lw));
        float *res = (float *)malloc(sizeof(float) * (size_t)n);
#pragma omp for schedule(static)
        for (int64_t l = 0; l < nlines; ++l) {
            int64_t i1 = l / dims[o2], i2 = l % dims[o2];
            float *base = a + i1 * strides[o1] + i2 * strides[o2];
            for (int64_t i = -lw; i < n + lw; ++i)
                line[i + lw] = (double)base[reflect_idx(i, n) * st];
            const double *fw = w + lw; /* centre */
            for (int64_t c = 0; c < n; ++c) {
                const double *x = line + lw + c;
                double tmp = x[0] * fw[0];
                for (int j = -lw; j < 0; ++j)
                    tmp += (x[j] + x[-j]) * fw[j];
                res[c] = (float)tmp;
            }
            for (int64_t c = 0; c < n; ++c)
                base[c * st] = res[c];
        }
        free(line);
        free(res);
    }
}

/* out: (Z+2r, Y+2r, X+2r) float32, caller-zeroed.  lw < 0 means "sigma == 0: no smoothing". */
int fpl_oracle_smooth_padded(const float *pred, int64_t Z, int64_t Y, int64_t X, int r,
                             const double *w, int lw, float *out, int threads)
{
    int64_t PZ = Z + 2 * r, PY = Y + 2 * r, PX = X + 2 * r;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    for (int64_t z = 0; z < Z; ++z)
        for (int64_t y = 0; y < Y; ++y)
            memcpy(out + ((z + r) * PY + (y + r)) * PX + r, pred + (z * Y + y) * X,
                   sizeof(float) * (size_t)X);
    if (lw >= 0) {
        pass_axis(out, PZ, PY, PX, 0, w, lw);
        pass_axis(out, PZ, PY, PX, 1, w, lw);
        pass_axis(out, PZ, PY, PX, 2, w, lw);
    }
    /* zero the r-wide border (r == 0: the reference's x[-0:] = 0 clears everything) */
    for (int64_t z = 0; z < PZ; ++z)
        for (int64_t y = 0; y < PY; ++y) {
            float *row = out + (z * PY + y) * PX;
            int inside = r > 0 && z >= r && z < PZ - r && y >= r && y < PY - r;
            if (!inside) {
                memset(row, 0, sizeof(float) * (size_t)PX);
            } else {
                memset(row, 0, sizeof(float) * (size_t)r);
                memset(row + PX - r, 0, sizeof(float) * (size_t)r);
            }
        }
    return 0;
}

typedef struct { float v; int64_t i; } cand_t;

static int cand_cmp(const void *pa, const void *pb)
{
    const cand_t *a = (const cand_t *)pa, *b = (const cand_t *)pb;
    if (a->v > b->v) return -1;
    if (a->v < b->v) return 1;
    return (a->i > b->i) - (a->i < b->i);
}

/* s: padded smoothed map (PZ,PY,PX).  Returns number of detections (flat padded index,
 * value) in emission order, or -1 when max_out is too small / out of memory. */
int64_t fpl_oracle_greedy(const float *s, int64_t PZ, int64_t PY, int64_t PX, int r,
                          double thresh, int64_t *out_idx, float *out_val, int64_t max_out)
{
    int64_t n = PZ * PY * PX, nc = 0;
    for (int64_t i = 0; i < n; ++i)
        if ((double)s[i] > thresh) ++nc;
    cand_t *c = (cand_t *)malloc(sizeof(cand_t) * (size_t)(nc ? nc : 1));
    uint8_t *valid = (uint8_t *)malloc((size_t)n);
    if (!c || !valid) { free(c); free(valid); return -1; }
    memset(valid, 1, (size_t)n);
    int64_t k = 0;
    for (int64_t i = 0; i < n; ++i)
        if ((double)s[i] > thresh) { c[k].v = s[i]; c[k].i = i; ++k; }
    qsort(c, (size_t)nc, sizeof(cand_t), cand_cmp);
    int64_t nout = 0;
    for (int64_t j = 0; j < nc; ++j) {
        int64_t i = c[j].i;
        if (!valid[i]) continue;
        if (c[j].v <= 0) break;
        if (nout >= max_out) { nout = -1; break; }
        out_idx[nout] = i; out_val[nout] = c[j].v; ++nout;
        int64_t z = i / (PY * PX), y = (i / PX) % PY, x = i % PX;
        for (int dz = -r; dz <= r; ++dz)
            for (int dy = -r; dy <= r; ++dy) {
                int64_t zz = z + dz, yy = y + dy;
                if (zz < 0 || zz >= PZ || yy < 0 || yy >= PY) continue;
                for (int dx = -r; dx <= r; ++dx) {
                    int64_t xx = x + dx;
                    if (xx < 0 || xx >= PX) continue;
                    if (sqrt((double)(dz * dz + dy * dy + dx * dx)) <= (double)r)
                        valid[(zz * PY + yy) * PX + xx] = 0;
                }
            }
    }
    free(c); free(valid);
    return nout;
}
