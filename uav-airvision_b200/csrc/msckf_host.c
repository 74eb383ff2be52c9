/*
 * _msckfhost -- the scalar inner loops of the host MSCKF (uav-airvision_b200/msckf.py) in C:
 *
 *   propagate(...)          IMU batch between two images: per sample the closed-form transition matrix, RK4 state
 *                           prediction, observability-constrained blocks and the 21x21 covariance update
 *                           (reference behaviour: src/msckf.py:251-388); returns the accumulated transition matrix
 *   triangulate(...) / triangulate_world(...)
 *                           Levenberg-Marquardt on inverse depth over all stereo observations of a feature
 *                           (src/feature/feature_position_initializer.py:6-76, feature_observation.py:4-39)
 *   jacobians(...)          per-observation measurement Jacobian blocks with the observability constraint
 *                           (src/msckf.py:443-502)
 *   gate(...)               chi-square gate statistic of every feature of a group (src/msckf.py:605-612)
 *   null_project(...)       projection onto the left null space of the feature Jacobian (src/msckf.py:504-540)
 *
 * All are small fixed-size loops per sample / observation / feature: numpy spends its time on call overhead there
 * (~0.35 ms per IMU sample, ~0.5 ms per triangulated feature), this spends microseconds.  msckf.py keeps the numpy
 * statement of each (tests/test_msckf_host.py compares the two).
 * Plain C, no BLAS; every argument is a C-contiguous float64 buffer.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <math.h>
#include <string.h>

#define N 21

static void skew3(const double* v, double* S) {
    S[0] = 0; S[1] = -v[2]; S[2] = v[1];
    S[3] = v[2]; S[4] = 0; S[5] = -v[0];
    S[6] = -v[1]; S[7] = v[0]; S[8] = 0;
}
static void mat3_mul(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
static void mat3_mul_bt(const double* A, const double* B, double* C) { /* A B^T */
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
}
static void mat3_vec(const double* A, const double* v, double* o) {
    for (int i = 0; i < 3; ++i) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}
static void mat3_tvec(const double* A, const double* v, double* o) { /* A^T v */
    for (int i = 0; i < 3; ++i) o[i] = A[i] * v[0] + A[3 + i] * v[1] + A[6 + i] * v[2];
}
/* utils.py:12-23 */
static void to_rotation(const double* q_in, double* R) {
    const double n = sqrt(q_in[0] * q_in[0] + q_in[1] * q_in[1] + q_in[2] * q_in[2] + q_in[3] * q_in[3]);
    const double q[4] = {q_in[0] / n, q_in[1] / n, q_in[2] / n, q_in[3] / n};
    const double w = q[3], a = 2.0 * w * w - 1.0;
    double S[9];
    skew3(q, S);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[3 * i + j] = (i == j ? a : 0.0) - 2.0 * w * S[3 * i + j] + 2.0 * q[i] * q[j];
}

static void set_block(double* M, int r0, int c0, const double* B, double s) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[(r0 + i) * N + c0 + j] = s * B[3 * i + j];
}

/* msckf.py:340-388 */
static void predict_new_state(double dt, const double* gyro, const double* acc, const double* g, double* q, double* v, double* p) {
    const double gn = sqrt(gyro[0] * gyro[0] + gyro[1] * gyro[1] + gyro[2] * gyro[2]);
    double Om[16] = {0};
    double S[9];
    skew3(gyro, S);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) Om[4 * i + j] = -S[3 * i + j];
        Om[4 * i + 3] = gyro[i];
        Om[12 + i] = -gyro[i];
    }
    double dq[4], dq2[4];
    for (int half = 0; half < 2; ++half) {
        const double f = half ? 0.25 : 0.5;
        double* out = half ? dq2 : dq;
        double c, sOm;                        /* out = (c I + sOm Omega) q   (large rate) or c (I + sOm Omega) q (small) */
        if (gn > 1e-5) {
            c = cos(gn * dt * f);
            sOm = sin(gn * dt * f) / gn;
            for (int i = 0; i < 4; ++i) {
                double acc_ = 0;
                for (int j = 0; j < 4; ++j) acc_ += ((i == j ? c : 0.0) + sOm * Om[4 * i + j]) * q[j];
                out[i] = acc_;
            }
        } else {
            c = cos(gn * dt * f);
            for (int i = 0; i < 4; ++i) {
                double acc_ = 0;
                for (int j = 0; j < 4; ++j) acc_ += (c * ((i == j ? 1.0 : 0.0) + Om[4 * i + j] * dt * f)) * q[j];
                out[i] = acc_;
            }
        }
    }
    double R0[9], Rh[9], Rf[9];
    to_rotation(q, R0);
    to_rotation(dq2, Rh);
    to_rotation(dq, Rf);
    double k1[3], k2[3], k4[3], t[3];
    mat3_tvec(R0, acc, t);
    for (int i = 0; i < 3; ++i) k1[i] = t[i] + g[i];
    mat3_tvec(Rh, acc, t);
    for (int i = 0; i < 3; ++i) k2[i] = t[i] + g[i];           /* k3_v_dot == k2_v_dot */
    mat3_tvec(Rf, acc, t);
    for (int i = 0; i < 3; ++i) k4[i] = t[i] + g[i];
    const double qn = sqrt(dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2] + dq[3] * dq[3]);
    for (int i = 0; i < 3; ++i) {
        const double k1v = v[i] + k1[i] * dt / 2.0, k2v = v[i] + k2[i] * dt / 2, k3v = v[i] + k2[i] * dt;
        p[i] = p[i] + (v[i] + 2 * k1v + 2 * k2v + k3v) * dt / 6.0;
        v[i] = v[i] + (k1[i] + 2 * k2[i] + 2 * k2[i] + k4[i]) * dt / 6.0;
    }
    for (int i = 0; i < 4; ++i) q[i] = dq[i] / qn;
}

/* One IMU sample (msckf.py:274-338).  P: top-left 21x21 block of a matrix with row stride ldp. */
static void process_model(double dt, const double* m_gyro, const double* m_acc, const double* g, const double* noise, double* q,
                          double* p, double* v, const double* bg, const double* ba, double* qn, double* pn, double* vn, double* P,
                          Py_ssize_t ldp, double* phi) {
    double gyro[3], acc[3];
    for (int i = 0; i < 3; ++i) {
        gyro[i] = m_gyro[i] - bg[i];
        acc[i] = m_acc[i] - ba[i];
    }
    double R[9], Rt[9];
    to_rotation(q, R);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Rt[3 * i + j] = R[3 * j + i];
    double A[9], C[9], D[9], S[9], T[9], A2[9], A3[9], CA[9], CA2[9];
    skew3(gyro, S);
    for (int i = 0; i < 9; ++i) A[i] = -S[i] * dt;
    skew3(acc, S);
    mat3_mul(Rt, S, T);
    for (int i = 0; i < 9; ++i) {
        C[i] = -T[i] * dt;
        D[i] = -Rt[i] * dt;
    }
    mat3_mul(A, A, A2);
    mat3_mul(A2, A, A3);
    mat3_mul(C, A, CA);
    mat3_mul(CA, A, CA2);
    memset(phi, 0, sizeof(double) * N * N);
    for (int i = 0; i < N; ++i) phi[i * N + i] = 1.0;
    double B[9], vth[9], pth[9];
    for (int i = 0; i < 9; ++i) {
        const double id = (i % 4 == 0) ? 1.0 : 0.0;
        B[i] = -dt * (id + A[i] / 2.0 + A2[i] / 6.0);
        vth[i] = C[i] + CA[i] / 2.0 + CA2[i] / 6.0;
        pth[i] = dt * (C[i] / 2.0 + CA[i] / 6.0);
    }
    set_block(phi, 0, 3, B, 1.0);
    for (int i = 0; i < 9; ++i) T[i] = -dt * (C[i] / 2.0 + CA[i] / 6.0);
    set_block(phi, 6, 3, T, 1.0);
    set_block(phi, 6, 9, D, 1.0);
    for (int i = 0; i < 9; ++i) T[i] = -dt * dt * C[i] / 6.0;
    set_block(phi, 12, 3, T, 1.0);
    for (int i = 0; i < 3; ++i) phi[(12 + i) * N + 6 + i] = dt;
    set_block(phi, 12, 9, D, dt / 2.0);

    predict_new_state(dt, gyro, acc, g, q, v, p);

    /* observability constraint (msckf.py:311-328) */
    double Rk[9], Rn[9];
    to_rotation(qn, Rk);
    to_rotation(q, Rn);
    mat3_mul_bt(Rn, Rk, T);
    set_block(phi, 0, 0, T, 1.0);
    double u[3], s[3], w[3], d[3], Au[3];
    mat3_vec(Rk, g, u);
    const double uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
    for (int i = 0; i < 3; ++i) s[i] = u[i] / uu;
    for (int i = 0; i < 3; ++i) d[i] = vn[i] - v[i];
    skew3(d, S);
    mat3_vec(S, g, w);
    mat3_vec(vth, u, Au);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) phi[(6 + i) * N + j] = vth[3 * i + j] - (Au[i] - w[i]) * s[j];
    for (int i = 0; i < 3; ++i) d[i] = dt * vn[i] + pn[i] - p[i];
    skew3(d, S);
    mat3_vec(S, g, w);
    mat3_vec(pth, u, Au);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) phi[(12 + i) * N + j] = pth[3 * i + j] - (Au[i] - w[i]) * s[j];

    /* P_II <- Phi (P_II + G Qc G^T dt) Phi^T, symmetrised */
    double M[N * N], X[N * N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) M[i * N + j] = P[i * ldp + j] + (i == j ? noise[i] * dt : 0.0);
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double a = 0;
            for (int k = 0; k < N; ++k) a += phi[i * N + k] * M[k * N + j];
            X[i * N + j] = a;
        }
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double a = 0;
            for (int k = 0; k < N; ++k) a += X[i * N + k] * phi[j * N + k];
            M[i * N + j] = a;
        }
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) P[i * ldp + j] = (M[i * N + j] + M[j * N + i]) / 2.0;
    memcpy(qn, q, 4 * sizeof(double));
    memcpy(pn, p, 3 * sizeof(double));
    memcpy(vn, v, 3 * sizeof(double));
}

static int get_buf(PyObject* o, Py_buffer* b, Py_ssize_t min_doubles, int writable, const char* name) {
    if (PyObject_GetBuffer(o, b, (writable ? PyBUF_WRITABLE : 0) | PyBUF_C_CONTIGUOUS) < 0) return -1;
    if (b->itemsize != 8 || b->len < min_doubles * 8) {
        PyBuffer_Release(b);
        PyErr_Format(PyExc_ValueError, "%s: need a C-contiguous float64 buffer of >= %zd elements", name, min_doubles);
        return -1;
    }
    return 0;
}

/* propagate(q, p, v, bg, ba, q_null, p_null, v_null, P, n, noise, gravity, imu(k x 7), t_state, t_bound, phi_total)
 *   -> (used, processed, t_state)   state arrays and P's 21x21 block updated in place; phi_total (21x21) = product of the
 *   transition matrices of the processed samples (identity when none). */
static PyObject* py_propagate(PyObject* self, PyObject* args) {
    PyObject *oq, *op, *ov, *obg, *oba, *oqn, *opn, *ovn, *oP, *onoise, *og, *oimu, *ophi;
    Py_ssize_t n;
    double t_state, t_bound;
    if (!PyArg_ParseTuple(args, "OOOOOOOOOnOOOddO", &oq, &op, &ov, &obg, &oba, &oqn, &opn, &ovn, &oP, &n, &onoise, &og, &oimu, &t_state,
                          &t_bound, &ophi))
        return NULL;
    Py_buffer b[13];
    PyObject* objs[13] = {oq, op, ov, obg, oba, oqn, opn, ovn, oP, onoise, og, oimu, ophi};
    const Py_ssize_t mins[13] = {4, 3, 3, 3, 3, 4, 3, 3, n * n, N, 3, 0, N * N};
    const int wr[13] = {1, 1, 1, 0, 0, 1, 1, 1, 1, 0, 0, 0, 1};
    static const char* names[13] = {"q", "p", "v", "bg", "ba", "q_null", "p_null", "v_null", "P", "noise", "gravity", "imu", "phi_total"};
    int got = 0;
    for (; got < 13; ++got)
        if (get_buf(objs[got], &b[got], mins[got], wr[got], names[got]) < 0) break;
    if (got < 13 || n < N) {
        for (int i = 0; i < got; ++i) PyBuffer_Release(&b[i]);
        if (got == 13) PyErr_SetString(PyExc_ValueError, "covariance smaller than 21 x 21");
        return NULL;
    }
    double *q = b[0].buf, *p = b[1].buf, *v = b[2].buf, *qn = b[5].buf, *pn = b[6].buf, *vn = b[7].buf, *P = b[8].buf;
    const double *bg = b[3].buf, *ba = b[4].buf, *noise = b[9].buf, *g = b[10].buf, *imu = b[11].buf;
    double* tot = b[12].buf;
    const Py_ssize_t k = b[11].len / (7 * 8);
    Py_ssize_t used = 0, processed = 0;
    double phi[N * N], tmp[N * N];
    memset(tot, 0, sizeof(double) * N * N);
    for (int i = 0; i < N; ++i) tot[i * N + i] = 1.0;
    for (Py_ssize_t i = 0; i < k; ++i) {
        const double* m = imu + 7 * i;
        if (m[0] < t_state) {
            ++used;
            continue;
        }
        if (m[0] > t_bound) break;
        process_model(m[0] - t_state, m + 1, m + 4, g, noise, q, p, v, bg, ba, qn, pn, vn, P, n, phi);
        for (int r = 0; r < N; ++r)
            for (int c = 0; c < N; ++c) {
                double a = 0;
                for (int j = 0; j < N; ++j) a += phi[r * N + j] * tot[j * N + c];
                tmp[r * N + c] = a;
            }
        memcpy(tot, tmp, sizeof tmp);
        ++used;
        ++processed;
        t_state = m[0];
    }
    for (int i = 0; i < 13; ++i) PyBuffer_Release(&b[i]);
    return Py_BuildValue("(nnd)", used, processed, t_state);
}

static int solve3(const double* A_in, const double* b_in, double* x) {   /* partial pivoting */
    double A[9], b[3];
    memcpy(A, A_in, sizeof A);
    memcpy(b, b_in, sizeof b);
    for (int c = 0; c < 3; ++c) {
        int piv = c;
        for (int r = c + 1; r < 3; ++r)
            if (fabs(A[3 * r + c]) > fabs(A[3 * piv + c])) piv = r;
        if (A[3 * piv + c] == 0.0) return -1;
        if (piv != c) {
            for (int j = 0; j < 3; ++j) {
                const double t = A[3 * c + j];
                A[3 * c + j] = A[3 * piv + j];
                A[3 * piv + j] = t;
            }
            const double t = b[c];
            b[c] = b[piv];
            b[piv] = t;
        }
        for (int r = c + 1; r < 3; ++r) {
            const double f = A[3 * r + c] / A[3 * c + c];
            for (int j = c; j < 3; ++j) A[3 * r + j] -= f * A[3 * c + j];
            b[r] -= f * b[c];
        }
    }
    for (int r = 2; r >= 0; --r) {
        double a = b[r];
        for (int j = r + 1; j < 3; ++j) a -= A[3 * r + j] * x[j];
        x[r] = a / A[3 * r + r];
    }
    return 0;
}

static double lm_cost(const double* Rs, const double* ts, const double* zs, Py_ssize_t k, const double* x) {
    double e = 0;
    for (Py_ssize_t i = 0; i < k; ++i) {
        const double* R = Rs + 9 * i;
        const double* t = ts + 3 * i;
        const double h1 = R[0] * x[0] + R[1] * x[1] + R[2] + x[2] * t[0];
        const double h2 = R[3] * x[0] + R[4] * x[1] + R[5] + x[2] * t[1];
        const double h3 = R[6] * x[0] + R[7] * x[1] + R[8] + x[2] * t[2];
        const double a = h1 / h3 - zs[2 * i], b = h2 / h3 - zs[2 * i + 1];
        e += a * a + b * b;
    }
    return e;
}

/* triangulate(Rs (k x 3 x 3), ts (k x 3), zs (k x 2), x0 (3), huber, precision, damping, outer_max, inner_max) -> (alpha, beta, rho)
 * k camera poses relative to the first cam0 (x_ci = R x_c0 + t), their normalized measurements, initial (alpha, beta, rho). */
/* Levenberg-Marquardt on (alpha, beta, rho) over k views x_i = R_i x + t_i (feature_position_initializer.py:31-70);
 * `inner` counts over the whole optimisation, as in the reference.  Returns -1 when the normal equations are singular. */
static int lm_solve(const double* Rs, const double* ts, const double* zs, Py_ssize_t k, double* sol, double huber, double precision,
                    double lambd, int outer_max, int inner_max) {
    int outer = 0, inner = 0, singular = 0;
    double delta_norm = INFINITY, total = lm_cost(Rs, ts, zs, k, sol);
    while (outer < outer_max && delta_norm > precision && !singular) {
        double A[9] = {0}, bb[3] = {0};
        for (Py_ssize_t i = 0; i < k; ++i) {
            const double* R = Rs + 9 * i;
            const double* t = ts + 3 * i;
            const double h1 = R[0] * sol[0] + R[1] * sol[1] + R[2] + sol[2] * t[0];
            const double h2 = R[3] * sol[0] + R[4] * sol[1] + R[5] + sol[2] * t[1];
            const double h3 = R[6] * sol[0] + R[7] * sol[1] + R[8] + sol[2] * t[2];
            const double W[9] = {R[0], R[1], t[0], R[3], R[4], t[1], R[6], R[7], t[2]};
            double J[6];
            for (int j = 0; j < 3; ++j) {
                J[j] = W[j] / h3 - W[6 + j] * h1 / (h3 * h3);
                J[3 + j] = W[3 + j] / h3 - W[6 + j] * h2 / (h3 * h3);
            }
            const double r0 = h1 / h3 - zs[2 * i], r1 = h2 / h3 - zs[2 * i + 1];
            const double e = sqrt(r0 * r0 + r1 * r1);
            const double w2 = (e <= huber) ? 1.0 : (huber / (2 * e)) * (huber / (2 * e));
            for (int a = 0; a < 3; ++a) {
                for (int c = 0; c < 3; ++c) A[3 * a + c] += w2 * (J[a] * J[c] + J[3 + a] * J[3 + c]);
                bb[a] += w2 * (J[a] * r0 + J[3 + a] * r1);
            }
        }
        int reduced = 0;
        while (inner < inner_max && !reduced) {
            double Ad[9], delta[3], ns[3];
            memcpy(Ad, A, sizeof Ad);
            Ad[0] += lambd;
            Ad[4] += lambd;
            Ad[8] += lambd;
            if (solve3(Ad, bb, delta) < 0) {
                singular = 1;
                break;
            }
            for (int j = 0; j < 3; ++j) ns[j] = sol[j] - delta[j];
            delta_norm = sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
            const double nc = lm_cost(Rs, ts, zs, k, ns);
            if (nc < total) {
                reduced = 1;
                memcpy(sol, ns, 3 * sizeof(double));
                total = nc;
                lambd = fmax(lambd / 10.0, 1e-10);
            } else {
                lambd = fmin(lambd * 10.0, 1e12);
            }
            ++inner;
        }
        ++outer;
    }
    return singular ? -1 : 0;
}

static PyObject* py_triangulate(PyObject* self, PyObject* args) {
    PyObject *oR, *ot, *oz, *ox;
    double huber, precision, lambd;
    int outer_max, inner_max;
    if (!PyArg_ParseTuple(args, "OOOOdddii", &oR, &ot, &oz, &ox, &huber, &precision, &lambd, &outer_max, &inner_max)) return NULL;
    Py_buffer bR, bt, bz, bx;
    if (get_buf(oR, &bR, 9, 0, "Rs") < 0) return NULL;
    const Py_ssize_t k = bR.len / 72;
    if (get_buf(ot, &bt, 3 * k, 0, "ts") < 0) {
        PyBuffer_Release(&bR);
        return NULL;
    }
    if (get_buf(oz, &bz, 2 * k, 0, "zs") < 0) {
        PyBuffer_Release(&bR);
        PyBuffer_Release(&bt);
        return NULL;
    }
    if (get_buf(ox, &bx, 3, 0, "x0") < 0) {
        PyBuffer_Release(&bR);
        PyBuffer_Release(&bt);
        PyBuffer_Release(&bz);
        return NULL;
    }
    const double *Rs = bR.buf, *ts = bt.buf, *zs = bz.buf;
    double sol[3];
    memcpy(sol, bx.buf, sizeof sol);
    const int singular = lm_solve(Rs, ts, zs, k, sol, huber, precision, lambd, outer_max, inner_max) < 0;
    PyBuffer_Release(&bR);
    PyBuffer_Release(&bt);
    PyBuffer_Release(&bz);
    PyBuffer_Release(&bx);
    if (singular) {
        PyErr_SetString(PyExc_ArithmeticError, "singular normal equations in feature triangulation");
        return NULL;
    }
    return Py_BuildValue("(ddd)", sol[0], sol[1], sol[2]);
}

/* triangulate_world(Rw (m,3,3), pw (m,3), Z (m,4), R01 (3,3), t01 (3), huber, precision, lambd, outer_max, inner_max)
 *   -> (x, y, z, valid): MSCKF.initialize_position in one call.  Poses of cam0 / cam1 of every observation relative to
 *   the first cam0, the two-view initial guess from the first stereo pair, the LM refinement, the positive-depth check
 *   and the feature position in the world frame (feature_position_initializer.py:6-76, feature_depth_estimator.py:4-14). */
static PyObject* py_triangulate_world(PyObject* self, PyObject* args) {
    PyObject *oR, *op, *oz, *o01, *ot01;
    double huber, precision, lambd;
    int outer_max, inner_max;
    if (!PyArg_ParseTuple(args, "OOOOOdddii", &oR, &op, &oz, &o01, &ot01, &huber, &precision, &lambd, &outer_max, &inner_max)) return NULL;
    Py_buffer b[5];
    PyObject* objs[5] = {oR, op, oz, o01, ot01};
    static const char* names[5] = {"Rw", "pw", "Z", "R01", "t01"};
    int got = 0;
    for (; got < 5; ++got)
        if (get_buf(objs[got], &b[got], 1, 0, names[got]) < 0) break;
    PyObject* ret = NULL;
    if (got == 5) {
        const Py_ssize_t m = b[0].len / 72;
        if (m < 1 || b[0].len != m * 72 || b[1].len != m * 24 || b[2].len != m * 32 || b[3].len != 72 || b[4].len != 24) {
            PyErr_SetString(PyExc_ValueError, "triangulate_world: inconsistent array sizes");
        } else {
            const double *Rw = b[0].buf, *pw = b[1].buf, *Z = b[2].buf, *R01 = b[3].buf, *t01 = b[4].buf;
            double* W = PyMem_Malloc(sizeof(double) * (size_t)(2 * m) * (9 + 3 + 2));
            if (!W) {
                PyErr_NoMemory();
            } else {
                double *Rs = W, *ts = W + 18 * m, *zs = ts + 6 * m;
                const double *R0w = Rw, *p0 = pw;
                for (Py_ssize_t i = 0; i < m; ++i) {
                    double* Rc0 = Rs + 18 * i;
                    double* Rc1 = Rc0 + 9;
                    double *tc0 = ts + 6 * i, *tc1 = tc0 + 3;
                    const double d[3] = {p0[0] - pw[3 * i], p0[1] - pw[3 * i + 1], p0[2] - pw[3 * i + 2]};
                    mat3_mul_bt(Rw + 9 * i, R0w, Rc0);
                    mat3_vec(Rw + 9 * i, d, tc0);
                    mat3_mul(R01, Rc0, Rc1);
                    mat3_vec(R01, tc0, tc1);
                    tc1[0] += t01[0];
                    tc1[1] += t01[1];
                    tc1[2] += t01[2];
                    memcpy(zs + 4 * i, Z + 4 * i, 4 * sizeof(double));
                }
                const double z1[3] = {zs[0], zs[1], 1.0};
                double mm[3];
                mat3_vec(Rs + 9, z1, mm);
                const double a0 = mm[0] - zs[2] * mm[2], a1 = mm[1] - zs[3] * mm[2];
                const double b0 = zs[2] * ts[5] - ts[3], b1 = zs[3] * ts[5] - ts[4];
                const double depth = (a0 * b0 + a1 * b1) / (a0 * a0 + a1 * a1);
                double sol[3] = {zs[0], zs[1], 1.0 / depth};
                if (lm_solve(Rs, ts, zs, 2 * m, sol, huber, precision, lambd, outer_max, inner_max) < 0) {
                    PyErr_SetString(PyExc_ArithmeticError, "singular normal equations in feature triangulation");
                } else {
                    const double fin[3] = {sol[0] / sol[2], sol[1] / sol[2], 1.0 / sol[2]};
                    int valid = 1;
                    for (Py_ssize_t i = 0; i < 2 * m; ++i) {
                        const double* R = Rs + 9 * i;
                        const double dep = (R[6] * fin[0] + R[7] * fin[1] + R[8] * fin[2]) + ts[3 * i + 2];
                        if (!(dep > 0)) valid = 0;
                    }
                    double pos[3];
                    mat3_tvec(R0w, fin, pos);
                    ret = Py_BuildValue("(dddi)", pos[0] + p0[0], pos[1] + p0[1], pos[2] + p0[2], valid);
                }
                PyMem_Free(W);
            }
        }
    }
    for (int i = 0; i < got; ++i) PyBuffer_Release(&b[i]);
    return ret;
}

/* ---- stereo measurement Jacobians of F features x m camera states (msckf.py:443-502 of the reference) ------------------
 * jacobians(R0 (F,m,3,3), pcam (F,m,3), Rn (F,m,3,3), pn (F,m,3), pw (F,3), Z (F,m,4), R01 (3,3), t01 (3), g (3),
 *           Hx out (F,m,4,6), Hf out (F,4m,3), r out (F,4m))
 * Per observation: p_c0 = R0 (p_w - p_cam), p_c1 = R01 p_c0 + t01; H_x = [d(proj)/dp * skew(p_c0) | -d(proj)/dp * R0] for
 * both cameras; the observability constraint removes the component of H_x along u = [R_null g; (p_w - p_null) x g];
 * H_f = -H_x[:, 3:6]; r = z - proj.  About sixty small numpy calls per group of features otherwise. */
static PyObject* py_jacobians(PyObject* self, PyObject* args) {
    PyObject* o[12];
    if (!PyArg_ParseTuple(args, "OOOOOOOOOOOO", &o[0], &o[1], &o[2], &o[3], &o[4], &o[5], &o[6], &o[7], &o[8], &o[9], &o[10], &o[11]))
        return NULL;
    static const char* names[12] = {"R0", "pcam", "Rn", "pn", "pw", "Z", "R01", "t01", "g", "Hx", "Hf", "r"};
    Py_buffer b[12];
    int got = 0;
    for (; got < 12; ++got)
        if (get_buf(o[got], &b[got], 1, got >= 9, names[got]) < 0) break;
    PyObject* ret = NULL;
    if (got == 12) {
        const Py_ssize_t F = b[4].len / 24, fm = b[1].len / 24;
        const Py_ssize_t m = F > 0 ? fm / F : 0;
        if (F <= 0 || m <= 0 || fm != F * m || b[0].len != fm * 72 || b[2].len != fm * 72 || b[3].len != fm * 24 ||
            b[5].len != fm * 32 || b[6].len != 72 || b[7].len != 24 || b[8].len != 24 || b[9].len != fm * 192 ||
            b[10].len != fm * 96 || b[11].len != fm * 32) {
            PyErr_SetString(PyExc_ValueError, "jacobians: inconsistent array sizes");
        } else {
            const double *R0a = b[0].buf, *pca = b[1].buf, *Rna = b[2].buf, *pna = b[3].buf, *pwa = b[4].buf, *Za = b[5].buf;
            const double *R01 = b[6].buf, *t01 = b[7].buf, *g = b[8].buf;
            double *Hxa = b[9].buf, *Hfa = b[10].buf, *ra = b[11].buf;
            for (Py_ssize_t f = 0; f < F; ++f) {
                const double* pw = pwa + 3 * f;
                for (Py_ssize_t k = 0; k < m; ++k) {
                    const Py_ssize_t i = f * m + k;
                    const double *R0 = R0a + 9 * i, *pc = pca + 3 * i, *Rn = Rna + 9 * i, *pn = pna + 3 * i, *z = Za + 4 * i;
                    double* Hx = Hxa + 24 * i;
                    const double d[3] = {pw[0] - pc[0], pw[1] - pc[1], pw[2] - pc[2]};
                    double p0[3], p1[3], sk[9], R01sk[9], R01R0[9];
                    mat3_vec(R0, d, p0);
                    mat3_vec(R01, p0, p1);
                    p1[0] += t01[0];
                    p1[1] += t01[1];
                    p1[2] += t01[2];
                    skew3(p0, sk);
                    mat3_mul(R01, sk, R01sk);
                    mat3_mul(R01, R0, R01R0);
                    const double iz0 = 1.0 / p0[2], iz1 = 1.0 / p1[2];
                    const double J0[6] = {iz0, 0.0, -p0[0] * iz0 * iz0, 0.0, iz0, -p0[1] * iz0 * iz0};
                    const double J1[6] = {iz1, 0.0, -p1[0] * iz1 * iz1, 0.0, iz1, -p1[1] * iz1 * iz1};
                    for (int a = 0; a < 2; ++a)
                        for (int c = 0; c < 3; ++c) {
                            Hx[6 * a + c] = J0[3 * a] * sk[c] + J0[3 * a + 1] * sk[3 + c] + J0[3 * a + 2] * sk[6 + c];
                            Hx[6 * a + 3 + c] = -(J0[3 * a] * R0[c] + J0[3 * a + 1] * R0[3 + c] + J0[3 * a + 2] * R0[6 + c]);
                            Hx[6 * (a + 2) + c] = J1[3 * a] * R01sk[c] + J1[3 * a + 1] * R01sk[3 + c] + J1[3 * a + 2] * R01sk[6 + c];
                            Hx[6 * (a + 2) + 3 + c] =
                                -(J1[3 * a] * R01R0[c] + J1[3 * a + 1] * R01R0[3 + c] + J1[3 * a + 2] * R01R0[6 + c]);
                        }
                    double u[6];
                    mat3_vec(Rn, g, u);
                    const double dn[3] = {pw[0] - pn[0], pw[1] - pn[1], pw[2] - pn[2]};
                    u[3] = dn[1] * g[2] - dn[2] * g[1];
                    u[4] = dn[2] * g[0] - dn[0] * g[2];
                    u[5] = dn[0] * g[1] - dn[1] * g[0];
                    double uu = 0.0;
                    for (int c = 0; c < 6; ++c) uu += u[c] * u[c];
                    for (int a = 0; a < 4; ++a) {
                        double Au = 0.0;
                        for (int c = 0; c < 6; ++c) Au += Hx[6 * a + c] * u[c];
                        for (int c = 0; c < 6; ++c) Hx[6 * a + c] -= Au * (u[c] / uu);
                        double* hf = Hfa + 3 * (4 * i + a);
                        hf[0] = -Hx[6 * a + 3];
                        hf[1] = -Hx[6 * a + 4];
                        hf[2] = -Hx[6 * a + 5];
                    }
                    double* r = ra + 4 * i;
                    r[0] = z[0] - p0[0] / p0[2];
                    r[1] = z[1] - p0[1] / p0[2];
                    r[2] = z[2] - p1[0] / p1[2];
                    r[3] = z[3] - p1[1] / p1[2];
                }
            }
            ret = Py_None;
            Py_INCREF(ret);
        }
    }
    for (int i = 0; i < got; ++i) PyBuffer_Release(&b[i]);
    return ret;
}

/* ---- projection onto the left null space of the feature Jacobian (msckf.py:504-540 of the reference) -------------------
 * null_project(Hx (F,m,4,6), Hf (F,4m,3), r (F,4m), H out (F,4m-3,6m), rp out (F,4m-3))
 * Per feature: three Householder reflectors triangularise H_f (the QR numpy's `qr(..., mode='complete')` does through
 * LAPACK); applied to the block-sparse H_x (observation k fills rows 4k..4k+3, columns 6k..6k+5) and to r, the rows
 * below the third are Q[:, 3:]^T H_x and Q[:, 3:]^T r.  Any orthonormal basis of the null space gives the same gate and
 * the same update. */
static PyObject* py_null_project(PyObject* self, PyObject* args) {
    PyObject* o[5];
    if (!PyArg_ParseTuple(args, "OOOOO", &o[0], &o[1], &o[2], &o[3], &o[4])) return NULL;
    static const char* names[5] = {"Hx", "Hf", "r", "H", "rp"};
    Py_buffer b[5];
    int got = 0;
    for (; got < 5; ++got)
        if (get_buf(o[got], &b[got], 1, got >= 3, names[got]) < 0) break;
    PyObject* ret = NULL;
    if (got == 5) {
        const Py_ssize_t fm = b[0].len / 192;                    /* F * m */
        Py_ssize_t F = 0, m = 0;
        /* F from r and rp: r has 4m per feature, rp 4m - 3:  len(r) - len(rp) = 3 F doubles */
        if (b[2].len >= b[4].len) F = (b[2].len - b[4].len) / 24;
        if (F > 0) m = fm / F;
        const Py_ssize_t rows = 4 * m, cols = 6 * m;
        if (F <= 0 || m <= 0 || fm != F * m || rows < 4 || b[0].len != fm * 192 || b[1].len != F * rows * 24 ||
            b[2].len != F * rows * 8 || b[3].len != F * (rows - 3) * cols * 8 || b[4].len != F * (rows - 3) * 8) {
            PyErr_SetString(PyExc_ValueError, "null_project: inconsistent array sizes");
        } else {
            const double *Hxa = b[0].buf, *Hfa = b[1].buf, *ra = b[2].buf;
            double *Ha = b[3].buf, *rpa = b[4].buf;
            double* W = PyMem_Malloc(sizeof(double) * (size_t)(rows * cols + rows * 3 + rows));
            if (!W) {
                PyErr_NoMemory();
            } else {
                double *D = W, *A = W + rows * cols, *rr = A + rows * 3;
                for (Py_ssize_t f = 0; f < F; ++f) {
                    memset(D, 0, sizeof(double) * (size_t)(rows * cols));
                    for (Py_ssize_t k = 0; k < m; ++k)
                        for (int a = 0; a < 4; ++a)
                            memcpy(D + (4 * k + a) * cols + 6 * k, Hxa + ((f * m + k) * 4 + a) * 6, 6 * sizeof(double));
                    memcpy(A, Hfa + f * rows * 3, sizeof(double) * (size_t)(rows * 3));
                    memcpy(rr, ra + f * rows, sizeof(double) * (size_t)rows);
                    for (int j = 0; j < 3; ++j) {
                        double xn2 = 0.0;
                        for (Py_ssize_t i = j + 1; i < rows; ++i) xn2 += A[i * 3 + j] * A[i * 3 + j];
                        if (xn2 == 0.0) continue;                 /* already triangular in this column: H = I */
                        const double alpha = A[j * 3 + j];
                        const double beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
                        const double tau = (beta - alpha) / beta, sc = 1.0 / (alpha - beta);
                        /* v = (1, A[j+1:, j] * sc); apply I - tau v v^T to the later columns of A, to D and to r */
                        for (int c = j + 1; c < 3; ++c) {
                            double w = A[j * 3 + c];
                            for (Py_ssize_t i = j + 1; i < rows; ++i) w += A[i * 3 + j] * sc * A[i * 3 + c];
                            w *= tau;
                            A[j * 3 + c] -= w;
                            for (Py_ssize_t i = j + 1; i < rows; ++i) A[i * 3 + c] -= A[i * 3 + j] * sc * w;
                        }
                        for (Py_ssize_t c = 0; c < cols; ++c) {
                            double w = D[j * cols + c];
                            for (Py_ssize_t i = j + 1; i < rows; ++i) w += A[i * 3 + j] * sc * D[i * cols + c];
                            if (w == 0.0) continue;
                            w *= tau;
                            D[j * cols + c] -= w;
                            for (Py_ssize_t i = j + 1; i < rows; ++i) D[i * cols + c] -= A[i * 3 + j] * sc * w;
                        }
                        {
                            double w = rr[j];
                            for (Py_ssize_t i = j + 1; i < rows; ++i) w += A[i * 3 + j] * sc * rr[i];
                            w *= tau;
                            rr[j] -= w;
                            for (Py_ssize_t i = j + 1; i < rows; ++i) rr[i] -= A[i * 3 + j] * sc * w;
                        }
                    }
                    memcpy(Ha + f * (rows - 3) * cols, D + 3 * cols, sizeof(double) * (size_t)((rows - 3) * cols));
                    memcpy(rpa + f * (rows - 3), rr + 3, sizeof(double) * (size_t)(rows - 3));
                }
                PyMem_Free(W);
                ret = Py_None;
                Py_INCREF(ret);
            }
        }
    }
    for (int i = 0; i < got; ++i) PyBuffer_Release(&b[i]);
    return ret;
}

/* ---- chi-square gate statistic without the projected Jacobian (msckf.py:605-612 of the reference) ---------------------
 * gate(Hx (F,m,4,6), Hf (F,4m,3), r (F,4m), P (n,n), slots int64 (F,m), sigma, gamma out (F))
 * The reference computes gamma = r_p^T (H_p P H_p^T + s I)^-1 r_p with H_p = A^T H_x, r_p = A^T r, A an orthonormal
 * basis of the left null space of H_f.  With W = H_x P H_x^T + s I (4m x 4m) this is r^T A (A^T W A)^-1 A^T r, and for
 * [A N] orthogonal, range(N) = range(H_f):
 *     A (A^T W A)^-1 A^T = W^-1 - W^-1 H_f (H_f^T W^-1 H_f)^-1 H_f^T W^-1,
 * so one Cholesky of W with four right-hand sides (r and the three columns of H_f) gives gamma.  H_x is block diagonal
 * (observation k: rows 4k.., columns of camera state slots[k]), so W costs ~500 m^2 flops instead of the ~500 m^3 of the
 * dense products. */
static PyObject* py_gate(PyObject* self, PyObject* args) {
    PyObject *oHx, *oHf, *orr, *oP, *osl, *og;
    double sigma;
    if (!PyArg_ParseTuple(args, "OOOOOdO", &oHx, &oHf, &orr, &oP, &osl, &sigma, &og)) return NULL;
    Py_buffer bHx, bHf, br, bP, bs, bg;
    if (get_buf(oHx, &bHx, 1, 0, "Hx") < 0) return NULL;
    if (get_buf(oHf, &bHf, 1, 0, "Hf") < 0) { PyBuffer_Release(&bHx); return NULL; }
    if (get_buf(orr, &br, 1, 0, "r") < 0) { PyBuffer_Release(&bHx); PyBuffer_Release(&bHf); return NULL; }
    if (get_buf(oP, &bP, 1, 0, "P") < 0) { PyBuffer_Release(&bHx); PyBuffer_Release(&bHf); PyBuffer_Release(&br); return NULL; }
    if (get_buf(osl, &bs, 1, 0, "slots") < 0) {
        PyBuffer_Release(&bHx); PyBuffer_Release(&bHf); PyBuffer_Release(&br); PyBuffer_Release(&bP);
        return NULL;
    }
    if (get_buf(og, &bg, 1, 1, "gamma") < 0) {
        PyBuffer_Release(&bHx); PyBuffer_Release(&bHf); PyBuffer_Release(&br); PyBuffer_Release(&bP); PyBuffer_Release(&bs);
        return NULL;
    }
    PyObject* ret = NULL;
    const Py_ssize_t F = bg.len / 8, fm = bs.len / 8;
    const Py_ssize_t m = F > 0 ? fm / F : 0, R = 4 * m;
    Py_ssize_t n = 0;
    while (n * n * 8 < bP.len) ++n;
    if (F <= 0 || m <= 0 || fm != F * m || bHx.len != fm * 192 || bHf.len != F * R * 24 || br.len != F * R * 8 || n * n * 8 != bP.len) {
        PyErr_SetString(PyExc_ValueError, "gate: inconsistent array sizes");
    } else {
        const double *Hxa = bHx.buf, *Hfa = bHf.buf, *ra = br.buf, *P = bP.buf;
        const long long* sl = bs.buf;
        double* gam = bg.buf;
        double* Wk = PyMem_Malloc(sizeof(double) * (size_t)(R * 6 * m + R * R + R * 4));
        int bad = 0;
        if (!Wk) {
            PyErr_NoMemory();
        } else {
            double *HP = Wk, *W = Wk + R * 6 * m, *B = W + R * R;     /* HP (R x 6m), W (R x R), B (R x 4) = [r | Hf] */
            for (Py_ssize_t f = 0; f < F && !bad; ++f) {
                const long long* s = sl + f * m;
                for (Py_ssize_t k = 0; k < m; ++k)
                    if (21 + 6 * s[k] + 6 > n || s[k] < 0) bad = 1;
                if (bad) break;
                /* HP[4k+a, 6l+b] = sum_c Hx[k][a][c] P[21+6 s_k + c, 21 + 6 s_l + b]: rows of P streamed contiguously */
                for (Py_ssize_t k = 0; k < m; ++k) {
                    const double* hx = Hxa + (f * m + k) * 24;
                    for (int a = 0; a < 4; ++a) {
                        double* out = HP + (4 * k + a) * 6 * m;
                        for (Py_ssize_t j = 0; j < 6 * m; ++j) out[j] = 0.0;
                        for (int c = 0; c < 6; ++c) {
                            const double h = hx[6 * a + c];
                            const double* prow = P + (21 + 6 * s[k] + c) * n + 21;
                            for (Py_ssize_t l = 0; l < m; ++l) {
                                const double* pb = prow + 6 * s[l];
                                double* o = out + 6 * l;
                                for (int b2 = 0; b2 < 6; ++b2) o[b2] += h * pb[b2];
                            }
                        }
                    }
                }
                /* W = HP Hx^T + sigma I (lower triangle): W[i, 4l+a] = sum_b HP[i, 6l+b] Hx[l][a][b] */
                for (Py_ssize_t i = 0; i < R; ++i)
                    for (Py_ssize_t l = 0; 4 * l <= i; ++l) {
                        const double* hx = Hxa + (f * m + l) * 24;
                        const double* hp = HP + i * 6 * m + 6 * l;
                        for (int a = 0; a < 4; ++a)
                            W[i * R + 4 * l + a] = ((hp[0] * hx[6 * a] + hp[1] * hx[6 * a + 1]) + (hp[2] * hx[6 * a + 2] + hp[3] * hx[6 * a + 3])) +
                                                   (hp[4] * hx[6 * a + 4] + hp[5] * hx[6 * a + 5]);
                    }
                for (Py_ssize_t i = 0; i < R; ++i) W[i * R + i] += sigma;
                /* Cholesky W = L L^T, right-looking on the lower triangle (the updates are axpys over contiguous rows) */
                int notpd = 0;
                for (Py_ssize_t j = 0; j < R; ++j) {
                    double d = W[j * R + j];
                    if (!(d > 0.0)) { notpd = 1; break; }
                    d = sqrt(d);
                    W[j * R + j] = d;
                    const double id = 1.0 / d;
                    for (Py_ssize_t i = j + 1; i < R; ++i) W[i * R + j] *= id;
                    for (Py_ssize_t i = j + 1; i < R; ++i) {
                        const double lij = W[i * R + j];
                        double* wi = W + i * R;
                        for (Py_ssize_t k = j + 1; k <= i; ++k) wi[k] -= lij * W[k * R + j];
                    }
                }
                /* W not positive definite (the reference's (I - KH) P update does not preserve that on a diverging run): the
                 * reference's gating_test still answers through an LU solve (msckf.py:605-612).  Mark the feature; the caller
                 * recomputes its statistic that way. */
                if (notpd) { gam[f] = NAN; continue; }
                /* Y = L^-1 [r | Hf]  (forward substitution, 4 right-hand sides) */
                for (Py_ssize_t i = 0; i < R; ++i) {
                    double v[4] = {ra[f * R + i], Hfa[(f * R + i) * 3], Hfa[(f * R + i) * 3 + 1], Hfa[(f * R + i) * 3 + 2]};
                    for (Py_ssize_t k = 0; k < i; ++k)
                        for (int c = 0; c < 4; ++c) v[c] -= W[i * R + k] * B[k * 4 + c];
                    for (int c = 0; c < 4; ++c) B[i * 4 + c] = v[c] / W[i * R + i];
                }
                /* with y = L^-1 r, Z = L^-1 Hf:  r^T W^-1 r = y.y,  Hf^T W^-1 r = Z^T y,  Hf^T W^-1 Hf = Z^T Z */
                double yy = 0.0, q[3] = {0, 0, 0}, G[9] = {0};
                for (Py_ssize_t i = 0; i < R; ++i) {
                    const double* bi = B + 4 * i;
                    yy += bi[0] * bi[0];
                    for (int a = 0; a < 3; ++a) {
                        q[a] += bi[1 + a] * bi[0];
                        for (int c = 0; c < 3; ++c) G[3 * a + c] += bi[1 + a] * bi[1 + c];
                    }
                }
                double x[3];
                if (solve3(G, q, x) < 0) { gam[f] = NAN; continue; }
                gam[f] = yy - (q[0] * x[0] + q[1] * x[1] + q[2] * x[2]);
            }
            PyMem_Free(Wk);
            if (bad == 1)
                PyErr_SetString(PyExc_ValueError, "gate: camera-state slot outside the covariance");
            else if (!PyErr_Occurred()) {
                ret = Py_None;
                Py_INCREF(ret);
            }
        }
    }
    PyBuffer_Release(&bHx); PyBuffer_Release(&bHf); PyBuffer_Release(&br); PyBuffer_Release(&bP); PyBuffer_Release(&bs); PyBuffer_Release(&bg);
    return ret;
}

/* ---- correction of every camera state of the window (msckf.py:583-591 of the reference) -----------------------------
 * update_cams(q (cap,4), p (cap,3), R (cap,3,3), p_null (cap,3), dc (n,6)): for the first n states
 *   q <- small_angle_quaternion(dc[:3]) * q  (utils.py:78-97, 61-76), R <- to_rotation(q), p <- p + dc[3:], and
 *   p_null <- p (position_null is an alias of position in the reference from the augmentation on). */
static PyObject* py_update_cams(PyObject* self, PyObject* args) {
    PyObject *oq, *op, *oR, *opn, *od;
    if (!PyArg_ParseTuple(args, "OOOOO", &oq, &op, &oR, &opn, &od)) return NULL;
    Py_buffer bq, bp, bR, bn, bd;
    if (get_buf(oq, &bq, 1, 1, "q") < 0) return NULL;
    if (get_buf(op, &bp, 1, 1, "p") < 0) { PyBuffer_Release(&bq); return NULL; }
    if (get_buf(oR, &bR, 1, 1, "R") < 0) { PyBuffer_Release(&bq); PyBuffer_Release(&bp); return NULL; }
    if (get_buf(opn, &bn, 1, 1, "p_null") < 0) { PyBuffer_Release(&bq); PyBuffer_Release(&bp); PyBuffer_Release(&bR); return NULL; }
    if (get_buf(od, &bd, 1, 0, "dc") < 0) {
        PyBuffer_Release(&bq); PyBuffer_Release(&bp); PyBuffer_Release(&bR); PyBuffer_Release(&bn);
        return NULL;
    }
    PyObject* ret = NULL;
    const Py_ssize_t n = bd.len / 48, cap = bq.len / 32;
    if (bd.len != n * 48 || n > cap || bp.len != cap * 24 || bR.len != cap * 72 || bn.len != cap * 24) {
        PyErr_SetString(PyExc_ValueError, "update_cams: inconsistent array sizes");
    } else {
        double *qa = bq.buf, *pa = bp.buf, *Ra = bR.buf, *pna = bn.buf;
        const double* dca = bd.buf;
        for (Py_ssize_t i = 0; i < n; ++i) {
            const double* dc = dca + 6 * i;
            double* q = qa + 4 * i;
            const double h[3] = {dc[0] / 2.0, dc[1] / 2.0, dc[2] / 2.0};
            const double n2 = (h[0] * h[0] + h[1] * h[1]) + h[2] * h[2];
            double dq[4] = {h[0], h[1], h[2], 1.0};
            if (n2 <= 1.0) {
                dq[3] = sqrt(1.0 - n2);
            } else {
                const double s = sqrt(1.0 + n2);
                for (int k = 0; k < 4; ++k) dq[k] /= s;
            }
            double nd = sqrt(((dq[0] * dq[0] + dq[1] * dq[1]) + dq[2] * dq[2]) + dq[3] * dq[3]);
            for (int k = 0; k < 4; ++k) dq[k] /= nd;
            const double nq = sqrt(((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3]);
            const double q2[4] = {q[0] / nq, q[1] / nq, q[2] / nq, q[3] / nq};
            const double x = dq[0], y = dq[1], z = dq[2], w = dq[3];
            double qn[4] = {((w * q2[0] + z * q2[1]) - y * q2[2]) + x * q2[3], ((-z * q2[0] + w * q2[1]) + x * q2[2]) + y * q2[3],
                            ((y * q2[0] - x * q2[1]) + w * q2[2]) + z * q2[3], ((-x * q2[0] - y * q2[1]) - z * q2[2]) + w * q2[3]};
            const double nn = sqrt(((qn[0] * qn[0] + qn[1] * qn[1]) + qn[2] * qn[2]) + qn[3] * qn[3]);
            for (int k = 0; k < 4; ++k) q[k] = qn[k] / nn;
            to_rotation(q, Ra + 9 * i);
            double* p = pa + 3 * i;
            for (int k = 0; k < 3; ++k) {
                p[k] += dc[3 + k];
                pna[3 * i + k] = p[k];
            }
        }
        ret = Py_None;
        Py_INCREF(ret);
    }
    PyBuffer_Release(&bq); PyBuffer_Release(&bp); PyBuffer_Release(&bR); PyBuffer_Release(&bn); PyBuffer_Release(&bd);
    return ret;
}

static PyMethodDef methods[] = {
    {"update_cams", py_update_cams, METH_VARARGS, "EKF correction of every camera state of the window."},
    {"gate", py_gate, METH_VARARGS, "Chi-square gate statistic of F features from H_x, H_f, r and the covariance."},
    {"null_project", py_null_project, METH_VARARGS, "Projection of H_x and r onto the left null space of H_f."},
    {"jacobians", py_jacobians, METH_VARARGS, "Stereo measurement Jacobians of F features x m camera states."},
    {"propagate", py_propagate, METH_VARARGS, "IMU batch propagation (msckf.py:251-388 of the reference)."},
    {"triangulate", py_triangulate, METH_VARARGS, "Levenberg-Marquardt feature triangulation on inverse depth."},
    {"triangulate_world", py_triangulate_world, METH_VARARGS, "MSCKF.initialize_position in one call: relative poses, guess, LM, depth check."},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_msckfhost", "Scalar inner loops of the host MSCKF.", -1, methods};

PyMODINIT_FUNC PyInit__msckfhost(void) { return PyModule_Create(&moddef); }
