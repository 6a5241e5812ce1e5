// pnp_core.inl — minimal-solver arithmetic of the PnP-RANSAC stage (loop geometric verification,
// reference src/loopclosing.cpp:207-293: cv::solvePnPRansac + pose refinement), written once for the device
// kernel (pnp.cu) and, with SB_HOST_MODEL defined, for a host build that the CPU test-suite checks against
// numpy (tests/pnp_host_model.cpp).  All double precision.
//   pnp_quartic   real roots of a monic quartic (Ferrari through the resolvent cubic, Newton-polished)
//   pnp_p3p       Grunert's three-point pose: up to 4 (R, t) with  x_cam = R x_world + t
#ifndef PNP_CORE_INL
#define PNP_CORE_INL

#include <math.h>

#ifdef SB_HOST_MODEL
#define PNP_HD static inline
#else
#define PNP_HD static __device__
#endif

// largest real root of m^3 + B m^2 + C m + D
PNP_HD double pnp_cubic_largest(double B, double C, double D) {
    const double p = C - B * B / 3.0, q = 2.0 * B * B * B / 27.0 - B * C / 3.0 + D;
    const double disc = q * q / 4.0 + p * p * p / 27.0;
    double t;
    if (disc > 0) {
        const double s = sqrt(disc);
        t = cbrt(-q / 2.0 + s) + cbrt(-q / 2.0 - s);
    } else if (p < 0) {
        const double r = sqrt(-p / 3.0);
        double c = 3.0 * q / (2.0 * p * r);  // = -q / (2 r^3)
        c = c > 1.0 ? 1.0 : c < -1.0 ? -1.0 : c;
        t = 2.0 * r * cos(acos(c) / 3.0);
    } else {
        t = 0.0;
    }
    double m = t - B / 3.0;
    for (int it = 0; it < 3; it++) {  // polish on the original cubic
        const double f = ((m + B) * m + C) * m + D, df = (3.0 * m + 2.0 * B) * m + C;
        if (fabs(df) < 1e-300) break;
        m -= f / df;
    }
    return m;
}

PNP_HD int pnp_quadratic(double b, double c, double *x) {  // x^2 + b x + c
    const double disc = b * b - 4.0 * c;
    if (!(disc >= 0)) return 0;
    const double s = sqrt(disc);
    const double q = -0.5 * (b + (b >= 0 ? s : -s));  // avoids cancellation
    x[0] = q;
    x[1] = q != 0 ? c / q : -b - q;
    return 2;
}

// real roots of x^4 + a x^3 + b x^2 + c x + d
PNP_HD int pnp_quartic(double a, double b, double c, double d, double *x) {
    const double a2 = a * a;
    const double p = b - 3.0 * a2 / 8.0, q = c - a * b / 2.0 + a2 * a / 8.0, r = d - a * c / 4.0 + a2 * b / 16.0 - 3.0 * a2 * a2 / 256.0;
    int n = 0;
    const double scale = fabs(p) + fabs(r) + 1.0;
    if (fabs(q) < 1e-14 * scale) {  // biquadratic: y^4 + p y^2 + r
        double z[2];
        const int nz = pnp_quadratic(p, r, z);
        for (int i = 0; i < nz; i++)
            if (z[i] >= 0) { const double s = sqrt(z[i]); x[n++] = s; x[n++] = -s; }
    } else {
        const double m = pnp_cubic_largest(p, p * p / 4.0 - r, -q * q / 8.0);
        if (!(m > 0)) return 0;
        const double s = sqrt(2.0 * m), h = q / (2.0 * s);
        n += pnp_quadratic(-s, p / 2.0 + m + h, x + n);
        n += pnp_quadratic(s, p / 2.0 + m - h, x + n);
    }
    for (int i = 0; i < n; i++) {
        double v = x[i] - a / 4.0;
        for (int it = 0; it < 3; it++) {  // polish on the original quartic
            const double f = (((v + a) * v + b) * v + c) * v + d, df = ((4.0 * v + 3.0 * a) * v + 2.0 * b) * v + c;
            if (fabs(df) < 1e-300) break;
            v -= f / df;
        }
        x[i] = v;
    }
    return n;
}

PNP_HD void pnp_cross(const double *a, const double *b, double *c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
PNP_HD double pnp_normalize(double *a) {
    const double n = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    if (n > 0) { a[0] /= n; a[1] /= n; a[2] /= n; }
    return n;
}

// Rigid motion that maps the world triangle P onto the camera triangle Q (both 3 x 3, one point per row), from the
// orthonormal frames the two triangles span.  Rt = row-major R, then t.  false: degenerate (collinear) triangle.
PNP_HD bool pnp_align3(const double *P, const double *Q, double *Rt) {
    double e[9], f[9];
    for (int k = 0; k < 2; k++) {
        const double *X = k == 0 ? P : Q;
        double *o = k == 0 ? e : f;
        double d1[3] = {X[3] - X[0], X[4] - X[1], X[5] - X[2]}, d2[3] = {X[6] - X[0], X[7] - X[1], X[8] - X[2]};
        if (!(pnp_normalize(d1) > 1e-12)) return false;
        double n[3];
        pnp_cross(d1, d2, n);
        if (!(pnp_normalize(n) > 1e-12)) return false;
        double m[3];
        pnp_cross(n, d1, m);
        for (int i = 0; i < 3; i++) { o[i] = d1[i]; o[3 + i] = m[i]; o[6 + i] = n[i]; }  // rows: the three axes
    }
    // R = F^T E  (axes as rows of E, F): R e_k = f_k
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Rt[3 * i + j] = f[i] * e[j] + f[3 + i] * e[3 + j] + f[6 + i] * e[6 + j];
    for (int i = 0; i < 3; i++) Rt[9 + i] = Q[i] - (Rt[3 * i] * P[0] + Rt[3 * i + 1] * P[1] + Rt[3 * i + 2] * P[2]);
    return true;
}

// Grunert's P3P (Haralick et al. 1994, eq. for the quartic in v = s3 / s1).  P: 3 world points (rows), j: their unit
// bearing vectors in the camera frame.  Writes up to 4 poses (12 doubles each), returns their number.
PNP_HD int pnp_p3p(const double *P, const double *j, double *Rt_out) {
    double d[3];
    d[0] = P[3] - P[6]; d[1] = P[4] - P[7]; d[2] = P[5] - P[8];
    const double a2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    d[0] = P[0] - P[6]; d[1] = P[1] - P[7]; d[2] = P[2] - P[8];
    const double b2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    d[0] = P[0] - P[3]; d[1] = P[1] - P[4]; d[2] = P[2] - P[5];
    const double c2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    if (!(a2 > 1e-18 && b2 > 1e-18 && c2 > 1e-18)) return 0;
    const double ca = j[3] * j[6] + j[4] * j[7] + j[5] * j[8];
    const double cb = j[0] * j[6] + j[1] * j[7] + j[2] * j[8];
    const double cg = j[0] * j[3] + j[1] * j[4] + j[2] * j[5];
    const double q1 = (a2 - c2) / b2, q2 = (a2 + c2) / b2, q3 = (b2 - c2) / b2, q4 = (b2 - a2) / b2;
    const double A4 = (q1 - 1) * (q1 - 1) - 4 * c2 / b2 * ca * ca;
    const double A3 = 4 * (q1 * (1 - q1) * cb - (1 - q2) * ca * cg + 2 * c2 / b2 * ca * ca * cb);
    const double A2 = 2 * (q1 * q1 - 1 + 2 * q1 * q1 * cb * cb + 2 * q3 * ca * ca - 4 * q2 * ca * cb * cg + 2 * q4 * cg * cg);
    const double A1 = 4 * (-q1 * (1 + q1) * cb + 2 * a2 / b2 * cg * cg * cb - (1 - q2) * ca * cg);
    const double A0 = (1 + q1) * (1 + q1) - 4 * a2 / b2 * cg * cg;
    if (!(fabs(A4) > 1e-14)) return 0;
    double v[4];
    const int nr = pnp_quartic(A3 / A4, A2 / A4, A1 / A4, A0 / A4, v);
    int n = 0;
    for (int i = 0; i < nr; i++) {
        const double vv = v[i];
        if (!(vv > 0) || !isfinite(vv)) continue;
        const double den = 2 * (cg - vv * ca);
        if (!(fabs(den) > 1e-14)) continue;
        const double u = ((q1 - 1) * vv * vv - 2 * q1 * cb * vv + 1 + q1) / den;
        const double s1sq = b2 / (1 + vv * vv - 2 * vv * cb);
        if (!(u > 0) || !(s1sq > 0) || !isfinite(u) || !isfinite(s1sq)) continue;
        const double s1 = sqrt(s1sq), s2 = u * s1, s3 = vv * s1;
        double Q[9];
        for (int k = 0; k < 3; k++) { Q[k] = s1 * j[k]; Q[3 + k] = s2 * j[3 + k]; Q[6 + k] = s3 * j[6 + k]; }
        if (pnp_align3(P, Q, Rt_out + 12 * n)) n++;
    }
    return n;
}

#endif  // PNP_CORE_INL
