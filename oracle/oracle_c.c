/*
 * C restatement of the pymotion FK / root dual-quaternion path.
 *
 * TEST INFRASTRUCTURE ONLY -- built into oracle/_build/liboracle_c.so by
 * oracle/build_oracle.py; loaded by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py, never by pymotion_b200.
 *
 * Same arithmetic, same order and same precision steps as the reference's
 * NumPy path so it can check full-size batches (1M x 22 and up) that the NumPy
 * oracle would need tens of seconds and tens of GB for:
 *   fk                 /root/reference/pymotion/ops/skeleton.py:16-61
 *                      local R(q^) in float32 (quat.py:411-423, 276-317),
 *                      chain as float64 4x4 products (skeleton.py:44, 55-58)
 *   to_root_dual_quat  ops/skeleton.py:207-244 -- accumulation in float32,
 *                      dual part in float64 (dual_quat.py:32-35)
 *   from_root_dual_quat ops/skeleton.py:173-204 -- float64 in, float64 out
 * Pinned against the tests/golden fixtures by tests/test_oracle_golden.py.
 *
 * Compile with -ffp-contract=off: NumPy never fuses a*b+c, so neither may we.
 * Frames are independent; `#pragma omp parallel for` over frames is the only
 * parallelism (the reference itself is single-threaded).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* quat.py:411-423 then quat.py:276-317, all in float32 */
static void local_rotmat_f32(const float *q, float m[9]) {
    float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    float d = n + 1e-8f;
    float w = q[0] / d, x = q[1] / d, y = q[2] / d, z = q[3] / d;
    float x2 = x + x, y2 = y + y, z2 = z + z;
    float xx = x * x2, yy = y * y2, zz = z * z2;
    float xy = x * y2, xz = x * z2, yz = y * z2;
    float wx = w * x2, wy = w * y2, wz = w * z2;
    m[0] = 1.0f - (yy + zz); m[1] = xy - wz;          m[2] = xz + wy;
    m[3] = xy + wz;          m[4] = 1.0f - (xx + zz); m[5] = yz - wx;
    m[6] = xz - wy;          m[7] = yz + wx;          m[8] = 1.0f - (xx + yy);
}

/* ops/skeleton.py:16-61.  gpos_stride / off_stride are per-frame strides in
 * elements (0 = one shared row, i.e. NumPy broadcasting).  Outputs float64
 * like the reference: pos[F][J][3], rotm[F][J][3][3]. */
int orc_fk_f32(const float *rot, const float *gpos, int64_t gpos_stride, const float *offsets,
               int64_t off_stride, const int64_t *parents, int64_t n_frames, int32_t n_joints,
               double *pos, double *rotm) {
    if (n_joints <= 0) return -1;
    for (int j = 1; j < n_joints; ++j)
        if (parents[j] < 0 || parents[j] >= n_joints) return -2;
#pragma omp parallel
    {
        double *G = (double *)malloc(sizeof(double) * 12 * (size_t)n_joints); /* [J][R(9) | p(3)] */
#pragma omp for schedule(static)
        for (int64_t f = 0; f < n_frames; ++f) {
            const float *q = rot + f * n_joints * 4;
            const float *off = offsets + f * off_stride;
            const float *gp = gpos + f * gpos_stride;
            /* every local transform first (skeleton.py:45-49): a parent index >= the
             * child's own would read the still-local matrix, exactly like the reference */
            for (int j = 0; j < n_joints; ++j) {
                float m[9];
                local_rotmat_f32(q + 4 * j, m);
                double *g = G + 12 * j;
                for (int k = 0; k < 9; ++k) g[k] = (double)m[k];
                const float *t = (j == 0) ? gp : off + 3 * j;
                g[9] = t[0]; g[10] = t[1]; g[11] = t[2];
            }
            for (int j = 1; j < n_joints; ++j) { /* skeleton.py:51-58, index order */
                const double *P = G + 12 * parents[j];
                double *g = G + 12 * j;
                double r[12];
                for (int a = 0; a < 3; ++a) {
                    for (int b = 0; b < 3; ++b)
                        r[3 * a + b] = P[3 * a] * g[b] + P[3 * a + 1] * g[3 + b] + P[3 * a + 2] * g[6 + b];
                    r[9 + a] = P[3 * a] * g[9] + P[3 * a + 1] * g[10] + P[3 * a + 2] * g[11] + P[9 + a];
                }
                memcpy(g, r, sizeof(r));
            }
            double *po = pos + f * n_joints * 3, *ro = rotm + f * n_joints * 9;
            for (int j = 0; j < n_joints; ++j) {
                memcpy(ro + 9 * j, G + 12 * j, 9 * sizeof(double));
                memcpy(po + 3 * j, G + 12 * j + 9, 3 * sizeof(double));
            }
        }
        free(G);
    }
    return 0;
}

/* quat.py:337-361 */
#define QMUL(T, o, a, b)                                                   \
    do {                                                                   \
        T _w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];      \
        T _x = a[0] * b[1] + b[0] * a[1] + a[2] * b[3] - a[3] * b[2];      \
        T _y = a[0] * b[2] + b[0] * a[2] + a[3] * b[1] - a[1] * b[3];      \
        T _z = a[0] * b[3] + b[0] * a[3] + a[1] * b[2] - a[2] * b[1];      \
        o[0] = _w; o[1] = _x; o[2] = _y; o[3] = _z;                        \
    } while (0)

/* quat.py:320-334 with _fast_cross quat.py:653-674 */
#define QROT(T, o, q, v)                                                   \
    do {                                                                   \
        T _tx = (T)2.0 * (q[2] * v[2] - q[3] * v[1]);                      \
        T _ty = (T)2.0 * (q[3] * v[0] - q[1] * v[2]);                      \
        T _tz = (T)2.0 * (q[1] * v[1] - q[2] * v[0]);                      \
        T _ox = v[0] + q[0] * _tx + (q[2] * _tz - q[3] * _ty);             \
        T _oy = v[1] + q[0] * _ty + (q[3] * _tx - q[1] * _tz);             \
        T _oz = v[2] + q[0] * _tz + (q[1] * _ty - q[2] * _tx);             \
        o[0] = _ox; o[1] = _oy; o[2] = _oz;                                \
    } while (0)

/* ops/skeleton.py:207-244 + dual_quat.py:12-36.  float32 accumulation,
 * float64 dual quaternions out: dq[F][J][8]. */
int orc_to_root_dual_quat_f32(const float *rot, const float *gpos, int64_t gpos_stride, const int64_t *parents,
                              const float *offsets, int64_t n_frames, int32_t n_joints, double *dq) {
    if (n_joints <= 0) return -1;
    if (offsets[0] != 0.0f || offsets[1] != 0.0f || offsets[2] != 0.0f) return -3; /* :227 */
    for (int j = 1; j < n_joints; ++j)
        if (parents[j] < 0 || parents[j] >= n_joints) return -2;
#pragma omp parallel
    {
        float *R = (float *)malloc(sizeof(float) * 4 * (size_t)n_joints);
        float *T = (float *)malloc(sizeof(float) * 3 * (size_t)n_joints);
#pragma omp for schedule(static)
        for (int64_t f = 0; f < n_frames; ++f) {
            memcpy(R, rot + f * n_joints * 4, sizeof(float) * 4 * (size_t)n_joints);
            memcpy(T, offsets, sizeof(float) * 3 * (size_t)n_joints);
            memcpy(T, gpos + f * gpos_stride, sizeof(float) * 3);
            for (int j = 1; j < n_joints; ++j) {
                int64_t p = parents[j];
                if (p == 0) continue; /* :236-237 */
                float *rp = R + 4 * p, *tp = T + 3 * p, *rj = R + 4 * j, *tj = T + 3 * j;
                float v[3];
                QROT(float, v, rp, tj);
                tj[0] = v[0] + tp[0]; tj[1] = v[1] + tp[1]; tj[2] = v[2] + tp[2];
                QMUL(float, rj, rp, rj);
            }
            double *o = dq + f * n_joints * 8;
            for (int j = 0; j < n_joints; ++j) {
                double qr[4] = {R[4 * j], R[4 * j + 1], R[4 * j + 2], R[4 * j + 3]};
                double t[4] = {0.0, T[3 * j], T[3 * j + 1], T[3 * j + 2]};
                double qd[4];
                QMUL(double, qd, t, qr);
                for (int k = 0; k < 4; ++k) { o[8 * j + k] = qr[k]; o[8 * j + 4 + k] = 0.5 * qd[k]; }
            }
        }
        free(R); free(T);
    }
    return 0;
}

/* ops/skeleton.py:173-204 + dual_quat.py:62-83, float64 throughout.
 * trans[F][J][3], rots[F][J][4]. */
int orc_from_root_dual_quat_f64(const double *dq, const int64_t *parents, int64_t n_frames, int32_t n_joints,
                                double *trans, double *rots) {
    if (n_joints <= 0) return -1;
    for (int j = 1; j < n_joints; ++j)
        if (parents[j] < 0 || parents[j] >= n_joints) return -2;
#pragma omp parallel for schedule(static)
    for (int64_t f = 0; f < n_frames; ++f) {
        const double *d = dq + f * n_joints * 8;
        double *R = rots + f * n_joints * 4, *T = trans + f * n_joints * 3;
        for (int j = 0; j < n_joints; ++j) {
            const double *qr = d + 8 * j, *qd = d + 8 * j + 4;
            double c[4] = {qr[0], -qr[1], -qr[2], -qr[3]}, m[4];
            QMUL(double, m, qd, c);
            memcpy(R + 4 * j, qr, 4 * sizeof(double));
            T[3 * j] = 2 * m[1]; T[3 * j + 1] = 2 * m[2]; T[3 * j + 2] = 2 * m[3];
        }
        for (int j = n_joints - 1; j >= 1; --j) { /* :194 reversed */
            int64_t p = parents[j];
            if (p == 0) continue;
            double inv[4] = {R[4 * p], -R[4 * p + 1], -R[4 * p + 2], -R[4 * p + 3]};
            double dlt[3] = {T[3 * j] - T[3 * p], T[3 * j + 1] - T[3 * p + 1], T[3 * j + 2] - T[3 * p + 2]};
            double *tj = T + 3 * j, *rj = R + 4 * j;
            QROT(double, tj, inv, dlt);
            QMUL(double, rj, inv, rj);
        }
    }
    return 0;
}
