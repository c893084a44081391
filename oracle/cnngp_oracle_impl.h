/* Included twice by cnngp_oracle.c with REAL / SFX / R_SQRT / R_ACOS defined.
 * Test infrastructure only -- see the header of cnngp_oracle.c. */

/* kernels.py:43-49.  x [N1,C,P], y [N2,C,P] (P = W*H, contiguous NCHW).
 *   diag == 0: xy [N1*N2, P]   xy[i*N2+j,p] = mean_c x[i,c,p]*y[j,c,p]
 *   diag != 0: xy [N1, P]      (N1 == N2 asserted by the caller, kernels.py:28)
 *   xx [N1,P] = mean_c x^2 ; yy [N2,P] = mean_c y^2
 * torch's mean() is sum / count, evaluated in the tensor dtype. */
void CAT(oracle_init, SFX)(const REAL *x, const REAL *y, int64_t N1, int64_t N2, int64_t C,
                           int64_t P, int diag, REAL *xy, REAL *xx, REAL *yy) {
    const REAL cnt = (REAL)C;
    if (diag) {
#pragma omp parallel for schedule(static)
        for (int64_t n = 0; n < N1; ++n)
            for (int64_t p = 0; p < P; ++p) {
                REAL s = 0;
                for (int64_t c = 0; c < C; ++c) s += x[(n * C + c) * P + p] * y[(n * C + c) * P + p];
                xy[n * P + p] = s / cnt;
            }
    } else {
#pragma omp parallel for schedule(static) collapse(2)
        for (int64_t i = 0; i < N1; ++i)
            for (int64_t j = 0; j < N2; ++j)
                for (int64_t p = 0; p < P; ++p) {
                    REAL s = 0;
                    for (int64_t c = 0; c < C; ++c)
                        s += x[(i * C + c) * P + p] * y[(j * C + c) * P + p];
                    xy[(i * N2 + j) * P + p] = s / cnt;
                }
    }
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N1; ++n)
        for (int64_t p = 0; p < P; ++p) {
            REAL s = 0;
            for (int64_t c = 0; c < C; ++c) { REAL v = x[(n * C + c) * P + p]; s += v * v; }
            xx[n * P + p] = s / cnt;
        }
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N2; ++n)
        for (int64_t p = 0; p < P; ++p) {
            REAL s = 0;
            for (int64_t c = 0; c < C; ++c) { REAL v = y[(n * C + c) * P + p]; s += v * v; }
            yy[n * P + p] = s / cnt;
        }
}

/* kernels.py:92-97 with the buffer of kernels.py:78-88.
 * in [M, Hi, Wi] -> out [M, Ho, Wo], cross-correlation (F.conv2d) with a ke x ke kernel whose
 * taps are all `tap`, except that when zero_first != 0 row 0 and column 0 of the kernel are
 * zero (the even-kernel "same" trick, kernels.py:73-84; then ke = kernel_size + 1).
 * Zero padding `pad` on every side, stride, dilation as in torch:
 *   Ho = floor((Hi + 2 pad - dil (ke-1) - 1) / stride) + 1
 * `tap` must already be the value stored in the reference's buffer, i.e.
 * float32(var_weight / kernel_size**2) (widened to double for a .double() model). */
void CAT(oracle_conv, SFX)(const REAL *in, int64_t M, int64_t Hi, int64_t Wi, int ke,
                           int zero_first, int stride, int pad, int dil, REAL tap, REAL bias,
                           REAL *out, int64_t Ho, int64_t Wo) {
    const int t0 = zero_first ? 1 : 0;
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < M; ++m) {
        const REAL *src = in + m * Hi * Wi;
        REAL *dst = out + m * Ho * Wo;
        for (int64_t yo = 0; yo < Ho; ++yo)
            for (int64_t xo = 0; xo < Wo; ++xo) {
                REAL acc = 0;
                for (int ty = t0; ty < ke; ++ty) {
                    const int64_t yi = yo * stride - pad + (int64_t)dil * ty;
                    if (yi < 0 || yi >= Hi) continue;
                    for (int tx = t0; tx < ke; ++tx) {
                        const int64_t xi = xo * stride - pad + (int64_t)dil * tx;
                        if (xi < 0 || xi >= Wi) continue;
                        acc += tap * src[yi * Wi + xi];
                    }
                }
                dst[yo * Wo + xo] = acc + bias;
            }
    }
}

/* kernels.py:134-165.  xy [Nx*Ny, P] (diag: [Nx, P]), xx [Nx, P], yy [Ny, P]; all updated in
 * place.  Every intermediate is rounded to REAL exactly where the reference's tensor ops round:
 *   xx_yy = xx*yy + f32_tiny ; cos = clamp(xy * rsqrt(xx_yy), -1, 1)
 *   sin   = sqrt(clamp(xx_yy - xy**2, min=0)) ; theta = acos(cos)
 *   xy'   = (sin + (pi - theta)*xy) / (2 pi) ; xx' = xx/2 ; yy' = xx' if same else yy/2
 *   same & diag: xy' = xx' ; same & !diag: xy'[i,i] = xx'[i]   (needs Nx == Ny)
 * f32_tiny = np.finfo(np.float32).tiny is used in f64 mode as well (kernels.py:133). */
void CAT(oracle_relu, SFX)(REAL *xy, REAL *xx, REAL *yy, int64_t Nx, int64_t Ny, int64_t P,
                           int same, int diag) {
    const REAL tiny = (REAL)FLT_MIN;
    const REAL pi = (REAL)3.14159265358979323846;
    const REAL two_pi = (REAL)(2.0 * 3.14159265358979323846);
    const int64_t rows = diag ? Nx : Nx * Ny;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; ++r) {
        const int64_t i = diag ? r : r / Ny;
        const int64_t j = diag ? r : r % Ny;
        for (int64_t p = 0; p < P; ++p) {
            const REAL c = xy[r * P + p];
            const REAL vx = xx[i * P + p];
            const REAL vy = yy[j * P + p];
            REAL res;
            if (same && (diag || i == j)) {
                res = vx / (REAL)2;
            } else {
                REAL m = vx * vy;
                m = m + tiny;
                REAL cs = c * ((REAL)1 / R_SQRT(m));
                if (cs < (REAL)-1) cs = (REAL)-1;
                if (cs > (REAL)1) cs = (REAL)1;
                REAL c2 = c * c;
                REAL d = m - c2;
                if (d < (REAL)0) d = (REAL)0;
                const REAL sn = R_SQRT(d);
                const REAL th = R_ACOS(cs);
                REAL t = pi - th;
                t = t * c;
                t = sn + t;
                res = t / two_pi;
            }
            xy[r * P + p] = res;
        }
    }
    /* halve the variances after every xy entry has consumed the old ones */
    for (int64_t n = 0; n < Nx * P; ++n) xx[n] = xx[n] / (REAL)2;
    if (same) {
        for (int64_t n = 0; n < Ny * P; ++n) yy[n] = xx[n];
    } else {
        for (int64_t n = 0; n < Ny * P; ++n) yy[n] = yy[n] / (REAL)2;
    }
}
