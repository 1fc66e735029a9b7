/*
 * rnmf_oracle.c -- CPU restatement of the wattrg/rnmf hot path.  TEST INFRASTRUCTURE ONLY
 * (see rnmf_oracle.h).  Each function names the reference file:line it restates.
 * Arithmetic is written with the reference's exact association; compile with
 * -O2 -ffp-contract=off (no FMA contraction, no fast-math).
 */
#include "rnmf_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------- */
/* geometry / layout                                                                     */
/* ------------------------------------------------------------------------------------- */

/* mesh.rs:36-45: dx[d] = (hi[d]-lo[d])/(n[d] as Real) */
void ora_make_geom(ora_geom* g, int dim, const int64_t* n, const double* lo, const double* hi,
                   int ng, int ncomp)
{
    memset(g, 0, sizeof(*g));
    g->dim = dim; g->ng = ng; g->ncomp = ncomp;
    for (int d = 0; d < 3; ++d) { g->n[d] = 1; g->lo[d] = 0.0; g->dx[d] = 1.0; }
    for (int d = 0; d < dim; ++d) {
        g->n[d] = n[d];
        g->lo[d] = lo[d];
        g->dx[d] = (hi[d] - lo[d]) / (double)n[d];
    }
}

/* dataframe.rs:85  n_grown = n + 2*n_ghost (only for the frame's own dimensions) */
int64_t ora_grown(const ora_geom* g, int d)
{
    return d < g->dim ? g->n[d] + 2 * (int64_t)g->ng : 1;
}

/* dataframe.rs:82 */
int64_t ora_n_nodes(const ora_geom* g)
{
    return ora_grown(g, 0) * ora_grown(g, 1) * ora_grown(g, 2) * g->ncomp;
}

/* cartesian2d/mod.rs:31-37 (get_flat_index_unchecked); 3D: SURVEY 8a extrapolation rule */
int64_t ora_flat_index(const ora_geom* g, int64_t i, int64_t j, int64_t k, int64_t c)
{
    const int64_t ng = g->ng, G0 = ora_grown(g, 0), G1 = ora_grown(g, 1), G2 = ora_grown(g, 2);
    if (g->dim == 2) return (i + ng) + G0 * (j + ng) + G0 * G1 * c;
    return (i + ng) + G0 * ((j + ng) + G1 * ((k + ng) + G2 * c));
}

/* cartesian2d/mod.rs:40-54 get_ij_from_p, quirks kept on purpose:
 *   `p > n_nodes` (not >=), j*dimension[1] where dimension[0] is meant, `% 2` on the component. */
int ora_get_ij_from_p(int64_t p, int64_t dim0, int64_t dim1, int64_t n_nodes, int64_t ng,
                      int64_t* i_out, int64_t* j_out, int64_t* c_out)
{
    if (p > n_nodes) return 0;
    int64_t i = p % dim0;
    int64_t j = (p - i) / dim0 % dim1;
    int64_t d = (p - j * dim1 - i) / (dim0 * dim1) % 2;
    *i_out = i - ng; *j_out = j - ng; *c_out = d;
    return 1;
}

/* dataframe.rs:88-150: the 16-way classification, same test order */
void ora_classify_2d(const ora_geom* g, uint8_t* cell_type)
{
    const int64_t n_nodes = ora_n_nodes(g), G0 = ora_grown(g, 0), G1 = ora_grown(g, 1);
    const int64_t n0 = g->n[0], n1 = g->n[1], ng = g->ng;
    for (int64_t p = 0; p < n_nodes; ++p) {
        int64_t i, j, c;
        ora_get_ij_from_p(p, G0, G1, n_nodes, ng, &i, &j, &c);
        uint8_t t = ORA_LOC_REGULAR; /* Valid(Regular) */
        if (j < 0 && i < 0)                          t = ORA_GHOST + ORA_LOC_SOUTHWEST;
        else if (j < 0 && i >= n0)                   t = ORA_GHOST + ORA_LOC_SOUTHEAST;
        else if (i < 0 && j >= n1)                   t = ORA_GHOST + ORA_LOC_NORTHWEST;
        else if (i >= n0 && j >= n1)                 t = ORA_GHOST + ORA_LOC_NORTHEAST;
        else if (i < 0)                              t = ORA_GHOST + ORA_LOC_WEST;
        else if (j < 0)                              t = ORA_GHOST + ORA_LOC_SOUTH;
        else if (i >= n0)                            t = ORA_GHOST + ORA_LOC_EAST;
        else if (j >= n1)                            t = ORA_GHOST + ORA_LOC_NORTH;
        else if (i < ng && j < ng)                   t = ORA_LOC_SOUTHWEST;
        else if (i >= n0 - ng && j < ng)             t = ORA_LOC_SOUTHEAST;
        else if (j >= n1 - ng && i < ng)             t = ORA_LOC_NORTHWEST;
        else if (j >= n1 - ng && i >= n0 - ng)       t = ORA_LOC_NORTHEAST;
        else if (i < ng)                             t = ORA_LOC_WEST;
        else if (i >= n0 - ng)                       t = ORA_LOC_EAST;
        else if (j < ng)                             t = ORA_LOC_SOUTH;
        else if (j >= n1 - ng)                       t = ORA_LOC_NORTH;
        cell_type[p] = t;
    }
}

static inline int is_valid3(const ora_geom* g, int64_t i, int64_t j, int64_t k)
{
    return i >= 0 && j >= 0 && k >= 0 && i < g->n[0] && j < g->n[1] && k < g->n[2];
}

/* dataframe.rs:181-189 get_valid_cells: valid cells in flat order (all components).
 * (the valid iterator dataframe.rs:451-480 walks the same order.) */
int64_t ora_valid_cells(const ora_geom* g, const double* data, double* out)
{
    const int64_t ng = g->ng, G0 = ora_grown(g, 0), G1 = ora_grown(g, 1), G2 = ora_grown(g, 2);
    const int64_t kg = g->dim == 3 ? ng : 0;
    int64_t m = 0, p = 0;
    for (int64_t c = 0; c < g->ncomp; ++c)
        for (int64_t kk = 0; kk < G2; ++kk)
            for (int64_t jj = 0; jj < G1; ++jj)
                for (int64_t ii = 0; ii < G0; ++ii, ++p)
                    if (is_valid3(g, ii - ng, jj - ng, kk - kg)) out[m++] = data[p];
    return m;
}

/* mesh.rs:69-70: lo + ((idx as Real) + 0.5) * dx */
double ora_cell_centre(const ora_geom* g, int d, int64_t idx)
{
    return g->lo[d] + ((double)idx + 0.5) * g->dx[d];
}

/* mesh.rs:65-74 node_pos: SoA, p over n0*n1*2, (i,j,n)=get_ij(p) with ng=0 */
void ora_cell_centres_2d(const ora_geom* g, double* node_pos)
{
    const int64_t n0 = g->n[0], n1 = g->n[1], n_nodes = n0 * n1 * 2;
    for (int64_t p = 0; p < n_nodes; ++p) {
        int64_t i, j, c;
        ora_get_ij_from_p(p, n0, n1, n_nodes, 0, &i, &j, &c);
        node_pos[p] = c == 0 ? ora_cell_centre(g, 0, i) : ora_cell_centre(g, 1, j);
    }
}

/* dataframe.rs:761-787 (ghost-side iterator; also used on Valid(..) types by
 * collect_typed_cells :273-278): cells whose type is in `sides` and whose component passes,
 * in flat order.  The component of p is get_ij(p).2 (quirky formula kept). */
int64_t ora_ghost_iter_2d(const ora_geom* g, const double* data, const uint8_t* sides, int n_sides,
                          int64_t comp, double* out, int64_t* out_p)
{
    const int64_t n_nodes = ora_n_nodes(g), G0 = ora_grown(g, 0), G1 = ora_grown(g, 1);
    uint8_t* ct = (uint8_t*)malloc((size_t)n_nodes);
    ora_classify_2d(g, ct);
    int64_t m = 0;
    for (int64_t p = 0; p < n_nodes; ++p) {
        int hit = 0;
        for (int s = 0; s < n_sides; ++s) hit |= (ct[p] == sides[s]);
        if (!hit) continue;
        int64_t i, j, c;
        ora_get_ij_from_p(p, G0, G1, n_nodes, g->ng, &i, &j, &c);
        if (comp >= 0 && c != comp) continue;
        if (out) out[m] = data[p];
        if (out_p) out_p[m] = p;
        ++m;
    }
    free(ct);
    return m;
}

/* ------------------------------------------------------------------------------------- */
/* boundary fill                                                                         */
/* ------------------------------------------------------------------------------------- */

/* one Dirichlet/Neumann face.  dataframe.rs:289-355.  The tangential loops run over the VALID
 * cells of the other directions (intended semantics; the reference loops `0..n[i_dim]`
 * (:290,308,325,341), identical on the square/cubic meshes of every config and test).
 * Corners/edges are never written. */
static void fill_face(const ora_geom* g, double* data, int c, int d, int side, int kind, double v)
{
    const int64_t ng = g->ng, nd = g->n[d];
    const double dxd = g->dx[d];
    int d1 = (d + 1) % 3, d2 = (d + 2) % 3;
    const int64_t n1 = g->n[d1], n2 = g->n[d2];
    for (int64_t t2 = 0; t2 < n2; ++t2)
        for (int64_t t1 = 0; t1 < n1; ++t1) {
            int64_t idx[3];
            idx[d1] = t1; idx[d2] = t2;
#define AT(x) (*(idx[d] = (x), &data[ora_flat_index(g, idx[0], idx[1], idx[2], c)]))
            if (kind == ORA_BC_NEUMANN && side == 0) {
                /* :292-295  self[-gi] = self[gi-1] - gradient*(2.0*gi - 1.0)*dx */
                for (int64_t gi = 1; gi < ng + 1; ++gi) {
                    double m = AT(gi - 1);
                    double r = m - v * (2.0 * (double)gi - 1.0) * dxd;
                    AT(-gi) = r;
                }
            } else if (kind == ORA_BC_NEUMANN) {
                /* :310-313  self[hi+gi] = self[hi-1-gi] + gradient*(2.0*(gi+1) - 1.0)*dx */
                for (int64_t gi = 0; gi < ng; ++gi) {
                    double m = AT(nd - 1 - gi);
                    double r = m + v * (2.0 * (double)(gi + 1) - 1.0) * dxd;
                    AT(nd + gi) = r;
                }
            } else if (kind == ORA_BC_DIRICHLET && side == 0) {
                /* :327-329  self[-gi-1] = 2.0*value - self[gi] */
                for (int64_t gi = 0; gi < ng; ++gi) {
                    double m = AT(gi);
                    double r = 2.0 * v - m;
                    AT(-gi - 1) = r;
                }
            } else if (kind == ORA_BC_DIRICHLET) {
                /* :343-346  self[hi+gi] = 2.0*value - self[hi-1-gi] */
                for (int64_t gi = 0; gi < ng; ++gi) {
                    double m = AT(nd - 1 - gi);
                    double r = 2.0 * v - m;
                    AT(nd + gi) = r;
                }
            }
#undef AT
        }
}

/* Prescribed(values): dataframe.rs:378-398,413-428 + :280-284.  In 2D the side lists are
 * {West,NW,SW} / {South,SW,SE} / {East,NE,SE} / {North,NE,NW}, i.e. every grown cell whose
 * coordinate in direction d is below 0 (lo) or >= n (hi), CORNERS INCLUDED, visited in flat
 * order for component c and zipped with `values` (stops at the shorter). 3D: same rule. */
static void fill_prescribed(const ora_geom* g, double* data, int c, int d, int side,
                            const double* vals, int64_t nvals)
{
    const int64_t ng = g->ng, G0 = ora_grown(g, 0), G1 = ora_grown(g, 1), G2 = ora_grown(g, 2);
    const int64_t kg = g->dim == 3 ? ng : 0;
    int64_t m = 0;
    int64_t p = G0 * G1 * G2 * (int64_t)c;
    for (int64_t kk = 0; kk < G2; ++kk)
        for (int64_t jj = 0; jj < G1; ++jj)
            for (int64_t ii = 0; ii < G0; ++ii, ++p) {
                int64_t idx[3] = { ii - ng, jj - ng, kk - kg };
                int on = side == 0 ? idx[d] < 0 : idx[d] >= g->n[d];
                if (!on) continue;
                if (m >= nvals) return;
                data[p] = vals[m++];
            }
}

/* dataframe.rs:361-433 fill_bc: comp -> dim (x, y[, z]) -> low then high */
int ora_fill_bc(const ora_geom* g, double* data, const ora_bc_face* bc)
{
    for (int c = 0; c < g->ncomp; ++c)
        for (int d = 0; d < g->dim; ++d)
            for (int side = 0; side < 2; ++side) {
                const ora_bc_face* f = &bc[(c * g->dim + d) * 2 + side];
                switch (f->kind) {
                case ORA_BC_DIRICHLET:
                case ORA_BC_NEUMANN:   fill_face(g, data, c, d, side, f->kind, f->value); break;
                case ORA_BC_PRESCRIBED: fill_prescribed(g, data, c, d, side, f->prescribed, f->n_prescribed); break;
                case ORA_BC_HALO:      break;             /* internal face: filled by halo exchange */
                case ORA_BC_REFLECT:   return -1;         /* dataframe.rs:375,410 unimplemented!() */
                default:               return -2;
                }
            }
    return 0;
}

/* ------------------------------------------------------------------------------------- */
/* whole-frame operators: over the entire data vector, ghosts included                    */
/* ------------------------------------------------------------------------------------- */
void ora_scale(int64_t n, double a, const double* x, double* out)          /* dataframe.rs:635-637 */
{
    for (int64_t i = 0; i < n; ++i) out[i] = a * x[i];
}
void ora_add(int64_t n, const double* x, const double* y, double* out)     /* dataframe.rs:669-671 */
{
    for (int64_t i = 0; i < n; ++i) out[i] = x[i] + y[i];
}
void ora_mul(int64_t n, const double* x, const double* y, double* out)     /* dataframe.rs:686-688 */
{
    for (int64_t i = 0; i < n; ++i) out[i] = x[i] * y[i];
}

/* ------------------------------------------------------------------------------------- */
/* magnetostatics Poisson pieces                                                         */
/* ------------------------------------------------------------------------------------- */

/* poisson.rs:47-64.  powf(2.0) restated as x*x (what LLVM emits for pow(x, 2.0)). */
double ora_mu(int64_t i, int64_t j, const double* dx, const ora_model* m)
{
    double x_pos = ((double)i + 0.5) * dx[0];
    double y_pos = ((double)j + 0.5) * dx[1];
    double xd = x_pos - m->bubble_centre[0];
    double yd = y_pos - m->bubble_centre[1];
    double x2 = xd * xd, y2 = yd * yd;
    return (x2 + y2 <= m->bubble_radius * m->bubble_radius) ? m->mu[0] : m->mu[1];
}

/* poisson.rs:13-36: variable-mu 5-point operator on valid cells, then Dirichlet(0) ghost fill */
void ora_get_laplace_2d(const ora_geom* g, const double* phi, const ora_model* m, double* lap)
{
    const int64_t nn = ora_n_nodes(g);
    memset(lap, 0, (size_t)nn * sizeof(double));             /* new_from: default-initialised, :155 */
    const double* dx = g->dx;
#define PHI(i, j) phi[ora_flat_index(g, (i), (j), 0, 0)]
    for (int64_t j = 0; j < g->n[1]; ++j)
        for (int64_t i = 0; i < g->n[0]; ++i) {
            double mc = ora_mu(i, j, dx, m);
            double v = PHI(i + 1, j) * (ora_mu(i + 1, j, dx, m) + mc) / 2.0
                     + PHI(i - 1, j) * (ora_mu(i - 1, j, dx, m) + mc) / 2.0
                     + PHI(i, j + 1) * (ora_mu(i, j + 1, dx, m) + mc) / 2.0
                     + PHI(i, j - 1) * (ora_mu(i, j - 1, dx, m) + mc) / 2.0
                     - (PHI(i, j) * (ora_mu(i + 1, j, dx, m) + ora_mu(i - 1, j, dx, m)
                                     + ora_mu(i, j + 1, dx, m) + ora_mu(i, j - 1, dx, m) + 4.0 * mc)) / 2.0;
            lap[ora_flat_index(g, i, j, 0, 0)] = v;
        }
#undef PHI
    ora_bc_face z[4];
    for (int f = 0; f < 4; ++f) { z[f].kind = ORA_BC_DIRICHLET; z[f].value = 0.0; z[f].prescribed = 0; z[f].n_prescribed = 0; }
    ora_geom g1 = *g; g1.ncomp = 1;
    ora_fill_bc(&g1, lap, z);                                 /* :34 */
}

/* poisson.rs:38-44: source = -(1.0 - 1.0/relax)/mu on valid cells (ghosts 0), times lap over ALL data */
void ora_get_source_2d(const ora_geom* g, const double* phi, const ora_model* m, double* src)
{
    const int64_t nn = ora_n_nodes(g);
    double* lap = (double*)malloc((size_t)nn * sizeof(double));
    double* s = (double*)calloc((size_t)nn, sizeof(double));
    for (int64_t j = 0; j < g->n[1]; ++j)
        for (int64_t i = 0; i < g->n[0]; ++i)
            s[ora_flat_index(g, i, j, 0, 0)] = -(1.0 - 1.0 / m->relax) / ora_mu(i, j, g->dx, m);
    ora_get_laplace_2d(g, phi, m, lap);
    ora_mul(nn, s, lap, src);
    free(lap); free(s);
}

/* poisson.rs:72-76 (2D).  3D: SURVEY 8a extrapolation rule:
 *   a = 1/(2(1/dx2+1/dy2+1/dz2)), c_d = a/dx_d^2 ; this file fixes the exact expressions. */
void ora_poisson_coefs(int dim, const double* dx, double* coef)
{
    if (dim == 2) {
        double dx2 = dx[0] * dx[0], dy2 = dx[1] * dx[1];
        coef[0] = dx2 * dy2 / (2.0 * (dx2 + dy2));    /* alpha */
        coef[1] = dy2 / (2.0 * (dx2 + dy2));          /* beta  (E+W) */
        coef[2] = dx2 / (2.0 * (dx2 + dy2));          /* gamma (N+S) */
        coef[3] = 0.0;
    } else {
        double dx2 = dx[0] * dx[0], dy2 = dx[1] * dx[1], dz2 = dx[2] * dx[2];
        double a = 1.0 / (2.0 * (1.0 / dx2 + 1.0 / dy2 + 1.0 / dz2));
        coef[0] = a; coef[1] = a / dx2; coef[2] = a / dy2; coef[3] = a / dz2;
    }
}

/* the update expression, association exactly as poisson.rs:83-85 (3D: + cz*(U+D) appended) */
static inline double upd2(const double* c, double rhs, double e, double w, double n, double s)
{
    return c[0] * rhs + c[1] * (e + w) + c[2] * (n + s);
}
static inline double upd3(const double* c, double rhs, double e, double w, double n, double s, double u, double d)
{
    return c[0] * rhs + c[1] * (e + w) + c[2] * (n + s) + c[3] * (u + d);
}

/* poisson.rs:80-89: n_sub_iter x { for i { for j { in-place update } } ; fill_bc }.
 * In place => lexicographic Gauss-Seidel, i outer / j inner (3D: i, j, k-inner). */
int ora_solve_poisson_gs(const ora_geom* g, double* psi, const double* rhs, const ora_bc_face* bc,
                         int64_t n_sub_iter)
{
    double c[4];
    ora_poisson_coefs(g->dim, g->dx, c);
    const int64_t sx = 1, sy = ora_grown(g, 0), sz = sy * ora_grown(g, 1);
    const int64_t p0 = ora_flat_index(g, 0, 0, 0, 0);
    const int64_t nx = g->n[0], ny = g->n[1], nz = g->n[2];
    const double c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3];
    for (int64_t it = 0; it < n_sub_iter; ++it) {
        if (g->dim == 2) {
            /* poisson.rs:81-86: i outer, j inner (walks memory with stride G0) */
            for (int64_t i = 0; i < nx; ++i) {
                double* q = psi + p0 + i;
                const double* r = rhs + p0 + i;
                for (int64_t j = 0; j < ny; ++j, q += sy, r += sy)
                    *q = c0 * *r + c1 * (q[sx] + q[-sx]) + c2 * (q[sy] + q[-sy]);
            }
        } else {
            for (int64_t i = 0; i < nx; ++i)
                for (int64_t j = 0; j < ny; ++j) {
                    double* q = psi + p0 + i + sy * j;
                    const double* r = rhs + p0 + i + sy * j;
                    for (int64_t k = 0; k < nz; ++k, q += sz, r += sz)
                        *q = c0 * *r + c1 * (q[sx] + q[-sx]) + c2 * (q[sy] + q[-sy]) + c3 * (q[sz] + q[-sz]);
                }
        }
        int rc = ora_fill_bc(g, psi, bc);
        if (rc) return rc;
    }
    return 0;
}

/* ghost value the per-sweep fill would have produced for a 1-cell halo, from the adjacent valid cell */
static inline double ghost_of(const ora_bc_face* f, int side, double mirror, double dxd)
{
    if (f->kind == ORA_BC_DIRICHLET) return 2.0 * f->value - mirror;
    /* Neumann, ghost index 1 (low) / 0 (high): factor (2.0*1 - 1.0) */
    if (side == 0) return mirror - f->value * (2.0 * 1.0 - 1.0) * dxd;
    return mirror + f->value * (2.0 * (double)(0 + 1) - 1.0) * dxd;
}

/* Same result as ora_solve_poisson_gs (2D, ng=1, Dirichlet/Neumann faces) computed on the
 * hyperplanes t = i + j + 2*sweep, in place: cell (i,j) of sweep s needs NEW (i-1,j),(i,j-1)
 * [time t-1] and OLD (i+1,j),(i,j+1) [sweep s-1, also time t-1].  Face ghosts needed in
 * sweep s>=1 are f(own old value) (what fill_bc wrote after sweep s-1); sweep 0 reads the
 * stored ghosts.  This is the schedule the GPU parity kernel (K6) uses; cells of one
 * hyperplane are independent, so they are visited here in an arbitrary (descending) order. */
int ora_solve_poisson_gs_wavefront_2d(const ora_geom* g, double* psi, const double* rhs,
                                      const ora_bc_face* bc, int64_t S)
{
    if (g->dim != 2 || g->ng != 1 || g->ncomp != 1) return -3;
    for (int f = 0; f < 4; ++f)
        if (bc[f].kind != ORA_BC_DIRICHLET && bc[f].kind != ORA_BC_NEUMANN) return -4;
    double c[4];
    ora_poisson_coefs(2, g->dx, c);
    const int64_t nx = g->n[0], ny = g->n[1], sy = ora_grown(g, 0);
    const int64_t t_end = (nx - 1) + (ny - 1) + 2 * (S - 1);
    for (int64_t t = 0; S > 0 && t <= t_end; ++t) {
        for (int64_t s = S - 1; s >= 0; --s) {
            int64_t r = t - 2 * s;                 /* i + j */
            if (r < 0 || r > nx + ny - 2) continue;
            int64_t i_lo = r - (ny - 1) > 0 ? r - (ny - 1) : 0;
            int64_t i_hi = r < nx - 1 ? r : nx - 1;
            for (int64_t i = i_hi; i >= i_lo; --i) {
                int64_t j = r - i;
                int64_t p = ora_flat_index(g, i, j, 0, 0);
                double self = psi[p];
                double e = psi[p + 1], w = psi[p - 1], n = psi[p + sy], so = psi[p - sy];
                if (s > 0) {
                    if (i == 0)      w  = ghost_of(&bc[0], 0, self, g->dx[0]);
                    if (i == nx - 1) e  = ghost_of(&bc[1], 1, self, g->dx[0]);
                    if (j == 0)      so = ghost_of(&bc[2], 0, self, g->dx[1]);
                    if (j == ny - 1) n  = ghost_of(&bc[3], 1, self, g->dx[1]);
                }
                psi[p] = upd2(c, rhs[p], e, w, n, so);
            }
        }
    }
    return S > 0 ? ora_fill_bc(g, psi, bc) : 0;
}

/* Jacobi restatement of the SAME update expression (poisson.rs:83-85), out of place:
 * n_sweeps x { out = f(in) on valid cells ; swap ; fill_bc }.  The result ends in psi.
 * scratch must hold n_nodes doubles; it starts as a copy of psi so that the never-written
 * edge/corner ghosts agree in both buffers.  l1_update = sum |new-old| over valid cells of
 * the LAST sweep (long-double accumulator). */
static int jacobi_impl(const ora_geom* g, double* psi, double* scratch, const double* rhs,
                       const ora_bc_face* bc, int64_t n_sweeps, double* l1_update, int par)
{
    double c[4];
    ora_poisson_coefs(g->dim, g->dx, c);
    const int64_t nn = ora_n_nodes(g);
    const int64_t sy = ora_grown(g, 0), sz = sy * ora_grown(g, 1);
    memcpy(scratch, psi, (size_t)nn * sizeof(double));
    double* in = psi; double* out = scratch;
    (void)par;
    for (int64_t it = 0; it < n_sweeps; ++it) {
        if (g->dim == 2) {
#pragma omp parallel for if (par) schedule(static)
            for (int64_t j = 0; j < g->n[1]; ++j) {
                const int64_t p0 = ora_flat_index(g, 0, j, 0, 0);
                const double* q = in + p0; const double* r = rhs + p0; double* o = out + p0;
                for (int64_t i = 0; i < g->n[0]; ++i)
                    o[i] = upd2(c, r[i], q[i + 1], q[i - 1], q[i + sy], q[i - sy]);
            }
        } else {
#pragma omp parallel for if (par) schedule(static) collapse(2)
            for (int64_t k = 0; k < g->n[2]; ++k)
                for (int64_t j = 0; j < g->n[1]; ++j) {
                    const int64_t p0 = ora_flat_index(g, 0, j, k, 0);
                    const double* q = in + p0; const double* r = rhs + p0; double* o = out + p0;
                    for (int64_t i = 0; i < g->n[0]; ++i)
                        o[i] = upd3(c, r[i], q[i + 1], q[i - 1], q[i + sy], q[i - sy], q[i + sz], q[i - sz]);
                }
        }
        if (l1_update && it == n_sweeps - 1) {
            long double acc = 0.0L;
            for (int64_t k = 0; k < g->n[2]; ++k)
                for (int64_t j = 0; j < g->n[1]; ++j)
                    for (int64_t i = 0; i < g->n[0]; ++i) {
                        int64_t p = ora_flat_index(g, i, j, k, 0);
                        acc += fabsl((long double)out[p] - (long double)in[p]);
                    }
            *l1_update = (double)acc;
        }
        int rc = ora_fill_bc(g, out, bc);
        if (rc) return rc;
        double* t = in; in = out; out = t;
    }
    if (in != psi) memcpy(psi, in, (size_t)nn * sizeof(double));
    return 0;
}

int ora_jacobi_sweeps(const ora_geom* g, double* psi, double* scratch, const double* rhs,
                      const ora_bc_face* bc, int64_t n_sweeps, double* l1_update)
{
    return jacobi_impl(g, psi, scratch, rhs, bc, n_sweeps, l1_update, 0);
}

/* all-core variant: NOT the reference (which is single-threaded); a labelled extra baseline */
int ora_jacobi_sweeps_omp(const ora_geom* g, double* psi, double* scratch, const double* rhs,
                          const ora_bc_face* bc, int64_t n_sweeps, double* l1_update)
{
    return jacobi_impl(g, psi, scratch, rhs, bc, n_sweeps, l1_update, 1);
}

/* poisson.rs:94-100 over the valid iterator dataframe.rs:451-480: naive serial sum, flat order */
double ora_sum_abs_valid(const ora_geom* g, const double* data)
{
    const int64_t ng = g->ng, G0 = ora_grown(g, 0), G1 = ora_grown(g, 1), G2 = ora_grown(g, 2);
    const int64_t kg = g->dim == 3 ? ng : 0;
    double acc = 0.0;
    int64_t p = 0;
    for (int64_t c = 0; c < g->ncomp; ++c)
        for (int64_t kk = 0; kk < G2; ++kk)
            for (int64_t jj = 0; jj < G1; ++jj)
                for (int64_t ii = 0; ii < G0; ++ii, ++p)
                    if (is_valid3(g, ii - ng, jj - ng, kk - kg)) acc += fabs(data[p]);
    return acc;
}

double ora_sum_abs_valid_ld(const ora_geom* g, const double* data)
{
    const int64_t ng = g->ng, G0 = ora_grown(g, 0), G1 = ora_grown(g, 1), G2 = ora_grown(g, 2);
    const int64_t kg = g->dim == 3 ? ng : 0;
    long double acc = 0.0L;
    int64_t p = 0;
    for (int64_t c = 0; c < g->ncomp; ++c)
        for (int64_t kk = 0; kk < G2; ++kk)
            for (int64_t jj = 0; jj < G1; ++jj)
                for (int64_t ii = 0; ii < G0; ++ii, ++p)
                    if (is_valid3(g, ii - ng, jj - ng, kk - kg)) acc += fabsl((long double)data[p]);
    return (double)acc;
}

/* main.rs:50-53: psi(x,y) = -h_far[0]/dx[0]*x - h_far[1]/dx[1]*y at cell centres (valid cells) */
void ora_fill_ic_magnetostatics(const ora_geom* g, double* psi, const ora_model* m)
{
    memset(psi, 0, (size_t)ora_n_nodes(g) * sizeof(double));
    for (int64_t j = 0; j < g->n[1]; ++j)
        for (int64_t i = 0; i < g->n[0]; ++i) {
            double x = ora_cell_centre(g, 0, i), y = ora_cell_centre(g, 1, j);
            psi[ora_flat_index(g, i, j, 0, 0)] = -m->h_far[0] / g->dx[0] * x - m->h_far[1] / g->dx[1] * y;
        }
}

/* main.rs:29-74: BCs Neumann(-h_far/dx) x4, IC, fill_bc, then the under-relaxed outer loop.
 * jacobi=0: reference-faithful (GS inner sweeps); jacobi=1: same loop with Jacobi inner sweeps. */
int ora_magnetostatics_run(const ora_geom* g, double* psi, const ora_model* m, int jacobi,
                           double* trace, int64_t* iters_done)
{
    if (g->dim != 2 || g->ncomp != 1) return -3;
    const int64_t nn = ora_n_nodes(g);
    ora_bc_face bc[4];
    for (int d = 0; d < 2; ++d)
        for (int s = 0; s < 2; ++s) {
            bc[d * 2 + s].kind = ORA_BC_NEUMANN;
            bc[d * 2 + s].value = -m->h_far[d] / g->dx[d];              /* main.rs:40-44 */
            bc[d * 2 + s].prescribed = 0; bc[d * 2 + s].n_prescribed = 0;
        }
    ora_fill_ic_magnetostatics(g, psi, m);
    ora_fill_bc(g, psi, bc);                                             /* main.rs:54 */
    double* source = (double*)malloc((size_t)nn * sizeof(double));
    double* star = (double*)malloc((size_t)nn * sizeof(double));
    double* tmp = (double*)malloc((size_t)nn * sizeof(double));
    double* diff = (double*)malloc((size_t)nn * sizeof(double));
    double sum_diff = 1.0;
    int64_t iter = 0;
    int rc = 0;
    while (iter < m->n_iter && sum_diff > m->tol) {                      /* main.rs:65 */
        ora_get_source_2d(g, psi, m, source);                            /* :66 */
        memcpy(star, psi, (size_t)nn * sizeof(double));                  /* poisson.rs:71 clone */
        rc = jacobi ? ora_jacobi_sweeps(g, star, tmp, source, bc, m->n_sub_iter, 0)
                    : ora_solve_poisson_gs(g, star, source, bc, m->n_sub_iter);
        if (rc) break;
        ora_scale(nn, -1.0, psi, tmp);                                   /* :68 -1.0 * &psi */
        ora_add(nn, star, tmp, diff);                                    /* :68 &psi_star + .. */
        sum_diff = ora_sum_abs_valid(g, diff);                           /* :69 */
        ora_scale(nn, m->relax, diff, tmp);                              /* :70 relax * diff */
        ora_add(nn, psi, tmp, psi);                                      /* :70 &psi + .. */
        ora_fill_bc(g, psi, bc);                                         /* :71 */
        if (trace) trace[iter] = sum_diff;
        ++iter;
    }
    if (iters_done) *iters_done = iter;
    free(source); free(star); free(tmp); free(diff);
    return rc;
}

/* model.rs:72-88 get_h: 2-component frame, h0 = -(psi[i+1]-psi[i-1])/(2.0*dx0), h1 likewise in y */
void ora_get_h_2d(const ora_geom* g, const double* psi, double* h)
{
    ora_geom g2 = *g; g2.ncomp = 2;
    memset(h, 0, (size_t)ora_n_nodes(&g2) * sizeof(double));
    for (int64_t j = 0; j < g->n[1]; ++j)
        for (int64_t i = 0; i < g->n[0]; ++i) {
            h[ora_flat_index(&g2, i, j, 0, 0)] =
                -(psi[ora_flat_index(g, i + 1, j, 0, 0)] - psi[ora_flat_index(g, i - 1, j, 0, 0)]) / (2.0 * g->dx[0]);
            h[ora_flat_index(&g2, i, j, 0, 1)] =
                -(psi[ora_flat_index(g, i, j + 1, 0, 0)] - psi[ora_flat_index(g, i, j - 1, 0, 0)]) / (2.0 * g->dx[1]);
        }
}

int ora_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
