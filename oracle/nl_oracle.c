/* Brute-force periodic neighbour list -- TEST INFRASTRUCTURE (CPU oracle), plain C.
 *
 * Restates the neighbour-list contract of /root/reference/HermNet/data.py:18-24 (ASE
 * primitive_neighbor_list('ijS', pbc=[T,T,T]) [upstream, unverified here]) for sizes the numpy
 * oracle cannot brute-force in seconds.  For every requested centre i it tests every atom j and
 * every integer shift in a conservative range:
 *     D = (double)(float)(pos_j - pos_i) + S.cell      (S.cell in double, row-vector convention)
 *     keep  iff  sqrt(Dx*Dx + Dy*Dy + Dz*Dz) < rc  and not (i == j and S == 0)
 * Built by oracle/build.py into oracle/_build/libnl_oracle.so; only tests/ and bench.py's CPU
 * baseline may load it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static void inv3(const double *c, double *o) {
    double det = c[0] * (c[4] * c[8] - c[5] * c[7]) - c[1] * (c[3] * c[8] - c[5] * c[6]) +
                 c[2] * (c[3] * c[7] - c[4] * c[6]);
    double id = 1.0 / det;
    o[0] = (c[4] * c[8] - c[5] * c[7]) * id; o[1] = (c[2] * c[7] - c[1] * c[8]) * id; o[2] = (c[1] * c[5] - c[2] * c[4]) * id;
    o[3] = (c[5] * c[6] - c[3] * c[8]) * id; o[4] = (c[0] * c[8] - c[2] * c[6]) * id; o[5] = (c[2] * c[3] - c[0] * c[5]) * id;
    o[6] = (c[3] * c[7] - c[4] * c[6]) * id; o[7] = (c[1] * c[6] - c[0] * c[7]) * id; o[8] = (c[0] * c[4] - c[1] * c[3]) * id;
}

/* Returns the number of (i,j,S) entries for the given centres.  Entries are written (up to cap)
 * ordered by centre (in the order given), then j, then S lexicographically. */
int64_t nl_pbc_rows(const float *pos, int64_t n, const float *cell9, double rc,
                    const int64_t *centres, int64_t n_centres, int64_t cap,
                    int64_t *out_i, int64_t *out_j, int64_t *out_s) {
    double c[9], ic[9];
    for (int k = 0; k < 9; ++k) c[k] = (double)cell9[k];
    inv3(c, ic);
    /* perpendicular heights -> conservative image range */
    double vol = fabs(c[0] * (c[4] * c[8] - c[5] * c[7]) - c[1] * (c[3] * c[8] - c[5] * c[6]) +
                      c[2] * (c[3] * c[7] - c[4] * c[6]));
    int m[3];
    for (int a = 0; a < 3; ++a) {
        const double *u = &c[3 * ((a + 1) % 3)], *v = &c[3 * ((a + 2) % 3)];
        double cx = u[1] * v[2] - u[2] * v[1], cy = u[2] * v[0] - u[0] * v[2], cz = u[0] * v[1] - u[1] * v[0];
        double h = vol / sqrt(cx * cx + cy * cy + cz * cz);
        m[a] = (int)ceil(rc / h) + 1;
    }
    /* wrap offsets w = floor(frac) so that the shift loop can be centred on the wrapped images */
    int64_t *w = (int64_t *)malloc(sizeof(int64_t) * 3 * (size_t)n);
    for (int64_t a = 0; a < n; ++a) {
        double p[3] = {pos[3 * a], pos[3 * a + 1], pos[3 * a + 2]};
        for (int k = 0; k < 3; ++k)   /* frac_k = sum_r p_r * inv[r][k] */
            w[3 * a + k] = (int64_t)floor(p[0] * ic[k] + p[1] * ic[3 + k] + p[2] * ic[6 + k]);
    }
    int64_t cnt = 0;
    for (int64_t q = 0; q < n_centres; ++q) {
        int64_t i = centres[q];
        for (int64_t j = 0; j < n; ++j) {
            float dx = pos[3 * j] - pos[3 * i], dy = pos[3 * j + 1] - pos[3 * i + 1], dz = pos[3 * j + 2] - pos[3 * i + 2];
            for (int s0 = -m[0]; s0 <= m[0]; ++s0)
                for (int s1 = -m[1]; s1 <= m[1]; ++s1)
                    for (int s2 = -m[2]; s2 <= m[2]; ++s2) {
                        int64_t S0 = s0 - w[3 * j] + w[3 * i], S1 = s1 - w[3 * j + 1] + w[3 * i + 1],
                                S2 = s2 - w[3 * j + 2] + w[3 * i + 2];
                        if (i == j && S0 == 0 && S1 == 0 && S2 == 0) continue;
                        double ox = ((double)S0 * c[0] + (double)S1 * c[3]) + (double)S2 * c[6];
                        double oy = ((double)S0 * c[1] + (double)S1 * c[4]) + (double)S2 * c[7];
                        double oz = ((double)S0 * c[2] + (double)S1 * c[5]) + (double)S2 * c[8];
                        double X = (double)dx + ox, Y = (double)dy + oy, Z = (double)dz + oz;
                        double d = sqrt((X * X + Y * Y) + Z * Z);
                        if (d < rc) {
                            if (cnt < cap) {
                                out_i[cnt] = i; out_j[cnt] = j;
                                out_s[3 * cnt] = S0; out_s[3 * cnt + 1] = S1; out_s[3 * cnt + 2] = S2;
                            }
                            ++cnt;
                        }
                    }
        }
    }
    free(w);
    return cnt;
}
