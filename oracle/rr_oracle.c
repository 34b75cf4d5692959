/*
 * rr_oracle.c -- CPU restatement of river-route's routing hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (river_route_b200/)
 * may import, link or call this file.  Allowed users: tests/, bench.py's
 * cpu_baseline / --impl reference leg, __graft_entry__.smoke().
 *
 * Parity status: PINNED.  Every function below is checked against outputs of
 * the reference's own numba / scipy code run in the build container
 * (oracle/make_golden.py -> tests/golden/ *.npz, tests/test_oracle_golden.py).
 *
 * The reference is pure Python + numba; there is nothing to compile into
 * oracle/_ref.  File:line citations are relative to /root/reference/.
 *
 * All arithmetic is IEEE fp64.  Compile with -ffp-contract=off for the strict
 * variant (the default recipe) -- numba's fastmath=True may contract a*b+c to
 * an FMA, so the reference itself is only defined to ~1 ulp per operation.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__GNUC__)
#define RR_EXPORT __attribute__((visibility("default")))
#else
#define RR_EXPORT
#endif

/* One routing substep over the whole network, in the reference's push form.
 *   rhs must already hold the per-reach "own" terms (c3*q [+ c4_dt*ql ...]).
 *   river_route/routers/_numba_kernels.py:29-39 (muskingum), :70-78 (rapid)
 *   - pass 1 walks columns ascending and pushes c2[row]*q_old[col] downstream
 *   - pass 2 is the unit-lower-triangular forward substitution: q_new[col] is
 *     final when the walk reaches col, then lhs_off*q_new[col] is subtracted
 *     from the downstream row.
 */
static void substep_push(int64_t n, const int32_t *indptr, const int32_t *indices,
                         const double *lhs_off, const double *c2,
                         double *q, double *rhs)
{
    for (int64_t col = 0; col < n; ++col) {
        const double qv = q[col];
        for (int32_t j = indptr[col]; j < indptr[col + 1]; ++j) {
            const int32_t row = indices[j];
            rhs[row] += c2[row] * qv;
        }
    }
    for (int64_t col = 0; col < n; ++col) {
        const double qn = rhs[col];
        q[col] = qn;
        for (int32_t j = indptr[col]; j < indptr[col + 1]; ++j)
            rhs[indices[j]] -= lhs_off[j] * qn;
    }
}

/* Muskingum channel-only loop.  _numba_kernels.py:9-46.
 * out is [num_output_steps][ldo]; q is mutated in place into the final state. */
RR_EXPORT int rr_oracle_muskingum_route(
    int64_t n, const int32_t *indptr, const int32_t *indices, const double *lhs_off,
    const double *c2, const double *c3, double *q,
    double *out, int64_t ldo, int64_t num_output_steps, int64_t num_routing_per_output)
{
    double *rhs = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    double *acc = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    if (!rhs || !acc) { free(rhs); free(acc); return 1; }
    const double inv = 1.0 / (double)num_routing_per_output;      /* :19 */
    for (int64_t o = 0; o < num_output_steps; ++o) {
        memset(acc, 0, sizeof(double) * (size_t)n);               /* :22-23 */
        for (int64_t r = 0; r < num_routing_per_output; ++r) {
            for (int64_t i = 0; i < n; ++i) rhs[i] = c3[i] * q[i]; /* :27-28 */
            substep_push(n, indptr, indices, lhs_off, c2, q, rhs); /* :29-39 */
            for (int64_t i = 0; i < n; ++i) acc[i] += q[i];       /* :41-42 */
        }
        double *row = out + o * ldo;
        for (int64_t i = 0; i < n; ++i) {                         /* :44-46 */
            const double v = acc[i] * inv;
            row[i] = v > 0.0 ? v : 0.0;
        }
    }
    free(rhs); free(acc);
    return 0;
}

/* RapidMuskingum loop.  _numba_kernels.py:49-84.
 * ql is [T][ldq] (time-major), out is [T][ldo]. */
RR_EXPORT int rr_oracle_rapid_route(
    int64_t n, const int32_t *indptr, const int32_t *indices, const double *lhs_off,
    const double *c2, const double *c3, const double *c4_dt, double *q,
    const double *ql, int64_t ldq, double *out, int64_t ldo,
    int64_t T, int64_t num_substeps)
{
    double *rhs = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    double *acc = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    if (!rhs || !acc) { free(rhs); free(acc); return 1; }
    const double inv = 1.0 / (double)num_substeps;                /* :60 */
    for (int64_t t = 0; t < T; ++t) {
        const double *qrow = ql + t * ldq;
        memset(acc, 0, sizeof(double) * (size_t)n);               /* :64-65 */
        for (int64_t s = 0; s < num_substeps; ++s) {
            for (int64_t i = 0; i < n; ++i)                       /* :68-69 */
                rhs[i] = c3[i] * q[i] + c4_dt[i] * qrow[i];
            substep_push(n, indptr, indices, lhs_off, c2, q, rhs); /* :70-78 */
            for (int64_t i = 0; i < n; ++i) acc[i] += q[i];       /* :79-80 */
        }
        double *row = out + t * ldo;
        for (int64_t i = 0; i < n; ++i) {                         /* :82-84 */
            const double v = acc[i] * inv;
            row[i] = v > 0.0 ? v : 0.0;
        }
    }
    free(rhs); free(acc);
    return 0;
}

/* Column-ascending scatter y[row] += data*x[col] (the two SpMVs of unit_route,
 * _numba_kernels.py:126-139). */
static void csc_push(int64_t ncol, const int32_t *indptr, const int32_t *indices,
                     const double *data, const double *x, double *y)
{
    for (int64_t col = 0; col < ncol; ++col) {
        const double v = x[col];
        for (int32_t j = indptr[col]; j < indptr[col + 1]; ++j)
            y[indices[j]] += data[j] * v;
    }
}

/* UnitMuskingum loop.  _numba_kernels.py:88-171.
 * conv is the convolved lateral [T][ldc] over ALL reaches, out is [T][ldo].
 * q_ch / q_full are inner-reach vectors, both mutated. hw_idx/inner_idx int64. */
RR_EXPORT int rr_oracle_unit_route(
    int64_t n_inner, int64_t n_hw,
    const int32_t *lhs_indptr, const int32_t *lhs_indices, const double *lhs_off,
    const int32_t *ai_indptr, const int32_t *ai_indices, const double *ai_data,
    const int32_t *ah_indptr, const int32_t *ah_indices, const double *ah_data,
    const double *c1, const double *c2, const double *c3,
    const int64_t *hw_idx, const int64_t *inner_idx,
    double *q_ch, double *q_full,
    const double *conv, int64_t ldc, double *out, int64_t ldo,
    int64_t T, int64_t num_substeps)
{
    const size_t ni = (size_t)(n_inner > 0 ? n_inner : 1), nh = (size_t)(n_hw > 0 ? n_hw : 1);
    double *buf = (double *)malloc(sizeof(double) * (6 * ni + nh));
    if (!buf) return 1;
    double *rhs = buf, *acc = buf + ni, *ql_in = buf + 2 * ni, *a_in = buf + 3 * ni,
           *a_hw = buf + 4 * ni, *c1aql = buf + 5 * ni, *ql_hw = buf + 6 * ni;
    const double inv = 1.0 / (double)num_substeps;                /* :104 */
    for (int64_t t = 0; t < T; ++t) {
        const double *crow = conv + t * ldc;
        double *orow = out + t * ldo;
        for (int64_t i = 0; i < n_hw; ++i) ql_hw[i] = crow[hw_idx[i]];       /* :116-117 */
        for (int64_t i = 0; i < n_inner; ++i) ql_in[i] = crow[inner_idx[i]]; /* :118-119 */
        for (int64_t i = 0; i < n_hw; ++i) orow[hw_idx[i]] = ql_hw[i];       /* :122-123 */
        memset(a_in, 0, sizeof(double) * (size_t)n_inner);                   /* :126-131 */
        csc_push(n_inner, ai_indptr, ai_indices, ai_data, ql_in, a_in);
        memset(a_hw, 0, sizeof(double) * (size_t)n_inner);                   /* :134-139 */
        csc_push(n_hw, ah_indptr, ah_indices, ah_data, ql_hw, a_hw);
        for (int64_t i = 0; i < n_inner; ++i)                                /* :142-143 */
            c1aql[i] = c1[i] * (a_in[i] + a_hw[i]);
        memset(acc, 0, sizeof(double) * (size_t)n_inner);                    /* :145-146 */
        for (int64_t s = 0; s < num_substeps; ++s) {
            for (int64_t i = 0; i < n_inner; ++i)                            /* :150-151 */
                rhs[i] = c1aql[i] + c2[i] * a_hw[i] + c3[i] * q_ch[i];
            for (int64_t col = 0; col < n_inner; ++col) {                    /* :152-156 */
                const double qv = q_full[col];
                for (int32_t j = lhs_indptr[col]; j < lhs_indptr[col + 1]; ++j) {
                    const int32_t row = lhs_indices[j];
                    rhs[row] += c2[row] * qv;
                }
            }
            for (int64_t col = 0; col < n_inner; ++col) {                    /* :159-162 */
                const double qn = rhs[col];
                q_ch[col] = qn;
                for (int32_t j = lhs_indptr[col]; j < lhs_indptr[col + 1]; ++j)
                    rhs[lhs_indices[j]] -= lhs_off[j] * qn;
            }
            for (int64_t i = 0; i < n_inner; ++i) {                          /* :165-167 */
                q_full[i] = q_ch[i] + ql_in[i];
                acc[i] += q_full[i];
            }
        }
        for (int64_t i = 0; i < n_inner; ++i) {                              /* :169-171 */
            const double v = acc[i] * inv;
            orow[inner_idx[i]] = v > 0.0 ? v : 0.0;
        }
    }
    free(buf);
    return 0;
}

/* Unit-hydrograph convolution with carry-over state, in the exact summation
 * order of UnitHydrograph.convolve_incrementally
 * (river_route/uhkernels/UnitHydrograph.py:64-75): the running state row that
 * is released at time t was seeded by the incoming state and then received the
 * contributions of runoff steps oldest first.
 * UnitHydrograph.convolve (:77-107, scipy fftconvolve) computes the same
 * quantity to ~1e-16 of the column scale (reference test: rtol 1e-12,
 * tests/test_uhkernels.py:52-78).
 *   lateral [T][ldl], kernel [n_ks][ldk], state [n_ks][lds] (in/out),
 *   out [T][ldo].  State update follows :103-105: rows 0..n_ks-2 = the
 *   full-convolution tail at times T..T+n_ks-2, last row = 0. */
RR_EXPORT int rr_oracle_uh_convolve(
    int64_t n, int64_t n_ks, int64_t T,
    const double *lateral, int64_t ldl, const double *kernel, int64_t ldk,
    double *state, int64_t lds, double *out, int64_t ldo)
{
    double *work = (double *)malloc(sizeof(double) * (size_t)(n_ks > 0 ? n_ks : 1));
    if (!work) return 1;
    for (int64_t b = 0; b < n; ++b) {
        for (int64_t k = 0; k < n_ks; ++k) work[k] = state[k * lds + b];
        for (int64_t t = 0; t < T; ++t) {
            const double r = lateral[t * ldl + b];
            for (int64_t k = 0; k < n_ks; ++k) work[k] += kernel[k * ldk + b] * r; /* :71 */
            out[t * ldo + b] = work[0];                                            /* :72 */
            for (int64_t k = 0; k + 1 < n_ks; ++k) work[k] = work[k + 1];          /* :73 */
            work[n_ks - 1] = 0.0;                                                  /* :74 */
        }
        for (int64_t k = 0; k < n_ks; ++k) state[k * lds + b] = work[k];
    }
    free(work);
    return 0;
}

/* Grid-weight transform core: y[t][r] = sum over the stored entries of CSR row
 * r, in stored (ascending column) order, of w * x[t][col]
 * (river_route/runoff.py:292-298; scipy csr_matvecs row-wise axpy order),
 * followed by the in-place tail of runoff.py:309-337:
 *   cumulative -> incremental (backwards in time), optional clip at 0,
 *   NaN -> 0, optional multiply by catchment area.
 * x is [T][ldx] float64 (the reference upcasts f32 grids in the product),
 * y is [T][ldy]. */
RR_EXPORT int rr_oracle_weights_transform(
    int64_t n_rivers, int64_t T,
    const int32_t *indptr, const int32_t *indices, const double *w,
    const double *x, int64_t ldx, double *y, int64_t ldy,
    int cumulative, int force_positive, const double *area /* NULL = depths */)
{
    for (int64_t r = 0; r < n_rivers; ++r) {
        for (int64_t t = 0; t < T; ++t) y[t * ldy + r] = 0.0;
        for (int32_t j = indptr[r]; j < indptr[r + 1]; ++j) {
            const double a = w[j];
            const int64_t c = indices[j];
            for (int64_t t = 0; t < T; ++t) y[t * ldy + r] += a * x[t * ldx + c];
        }
    }
    if (cumulative)                                            /* :310-312 */
        for (int64_t t = T - 1; t > 0; --t)
            for (int64_t r = 0; r < n_rivers; ++r) y[t * ldy + r] -= y[(t - 1) * ldy + r];
    for (int64_t t = 0; t < T; ++t)
        for (int64_t r = 0; r < n_rivers; ++r) {
            double v = y[t * ldy + r];
            if ((force_positive & 1) && v < 0.0) v = 0.0;      /* :313-314 */
            if (!(force_positive & 2) && v != v) v = 0.0;      /* :331-333; bit 1: the caller resamples first (:316-329) */
            if (area) v *= area[r];                            /* :335-336 */
            y[t * ldy + r] = v;
        }
    return 0;
}
