/*
 * rnmf_oracle.h -- CPU restatement of the wattrg/rnmf stencil / ghost-fill / residual path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The CUDA library (rnmf_b200/csrc) never links, includes or calls anything from here.
 *
 * The reference (Rust) cannot be compiled in this image (no cargo/rustc, deps not vendored),
 * so this is a restatement, function by function, of the reference's own files; every
 * function cites the reference file:line it follows (paths relative to /root/reference).
 * Parity pinning: the ghost fill, classification, iteration and gather orders are pinned by
 * the reference's own unit-test vectors (tests/golden/, transcribed from
 * src/mesh/cartesian2d/dataframe.rs:862-1201 and mesh.rs:189-212).  The stencil sweep, the
 * variable-mu operator, the frame operators, the norm and the outer loop have NO reference
 * test -> for those rows the oracle is "parity unpinned" (pinned only by restating
 * examples/magnetostatics/poisson.rs and main.rs).
 *
 * Layout is the reference's dense layout (no pitch):  p = (i+ng) + G0*((j+ng) + G1*((k+ng) + G2*c))
 * with G = n + 2*ng  (src/mesh/cartesian2d/mod.rs:31-37; 3D by the extrapolation rule SURVEY 8a).
 *
 * Build: -O2 -ffp-contract=off, no fast-math: rustc never contracts a*b+c.
 */
#ifndef RNMF_ORACLE_H
#define RNMF_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* boundary_conditions.rs:11-24 (+ HALO: an extension used by the slab decomposition for
 * internal faces; the reference only remarks that internal faces need "modified" BCs,
 * cartesian2d/mod.rs:131) */
enum { ORA_BC_DIRICHLET = 0, ORA_BC_NEUMANN = 1, ORA_BC_REFLECT = 2, ORA_BC_PRESCRIBED = 3, ORA_BC_HALO = 4 };

/* dataframe.rs:11-31  CellType{Valid(Loc),Ghost(Loc)}: code = 16*is_ghost + loc */
enum { ORA_LOC_REGULAR = 0, ORA_LOC_NORTH, ORA_LOC_SOUTH, ORA_LOC_EAST, ORA_LOC_WEST,
       ORA_LOC_NORTHEAST, ORA_LOC_NORTHWEST, ORA_LOC_SOUTHWEST, ORA_LOC_SOUTHEAST };
#define ORA_GHOST 16

typedef struct {
    int32_t dim;        /* 2 or 3 */
    int32_t ng;         /* n_ghost */
    int32_t ncomp;      /* n_comp */
    int32_t _pad;
    int64_t n[3];       /* valid cells per direction (n[2]=1 in 2D) */
    double  lo[3];      /* mesh.rs:36 lo corner */
    double  dx[3];      /* mesh.rs:40  (hi-lo)/n */
} ora_geom;

typedef struct {
    int32_t kind;               /* ORA_BC_* */
    int32_t _pad;
    double  value;              /* Dirichlet value / Neumann gradient */
    const double* prescribed;   /* Prescribed(Vec) */
    int64_t n_prescribed;
} ora_bc_face;                  /* array index: (comp*dim + d)*2 + side, side 0=lo 1=hi */

typedef struct {                /* examples/magnetostatics/model.rs:11-20 */
    double h_far[2];
    double relax;
    int64_t n_iter;
    int64_t n_sub_iter;
    double bubble_centre[2];
    double bubble_radius;
    double mu[2];
    double tol;
} ora_model;

/* ---- geometry / layout -------------------------------------------------------------- */
void    ora_make_geom(ora_geom* g, int dim, const int64_t* n, const double* lo, const double* hi,
                      int ng, int ncomp);                                   /* mesh.rs:36-45 */
int64_t ora_grown(const ora_geom* g, int d);                                /* dataframe.rs:85 */
int64_t ora_n_nodes(const ora_geom* g);                                     /* dataframe.rs:82 */
int64_t ora_flat_index(const ora_geom* g, int64_t i, int64_t j, int64_t k, int64_t c); /* mod.rs:31-37 */
int     ora_get_ij_from_p(int64_t p, int64_t dim0, int64_t dim1, int64_t n_nodes, int64_t ng,
                          int64_t* i, int64_t* j, int64_t* c);              /* mod.rs:40-54 (quirks kept) */
void    ora_classify_2d(const ora_geom* g, uint8_t* cell_type);            /* dataframe.rs:88-150 */
int64_t ora_valid_cells(const ora_geom* g, const double* data, double* out);/* dataframe.rs:181-189, 451-480 */
void    ora_cell_centres_2d(const ora_geom* g, double* node_pos);          /* mesh.rs:65-74 (SoA, n0*n1*2) */
double  ora_cell_centre(const ora_geom* g, int d, int64_t idx);            /* mesh.rs:69-70 */
int64_t ora_ghost_iter_2d(const ora_geom* g, const double* data, const uint8_t* sides, int n_sides,
                          int64_t comp /* -1 = All */, double* out, int64_t* out_p); /* dataframe.rs:761-787 */

/* ---- boundary fill ------------------------------------------------------------------ */
int     ora_fill_bc(const ora_geom* g, double* data, const ora_bc_face* bc);/* dataframe.rs:289-433 */

/* ---- whole-frame operators (ghosts included) ---------------------------------------- */
void ora_scale(int64_t n, double a, const double* x, double* out);          /* dataframe.rs:628-657 */
void ora_add(int64_t n, const double* x, const double* y, double* out);     /* dataframe.rs:662-674 */
void ora_mul(int64_t n, const double* x, const double* y, double* out);     /* dataframe.rs:679-691 */

/* ---- magnetostatics Poisson pieces -------------------------------------------------- */
double ora_mu(int64_t i, int64_t j, const double* dx, const ora_model* m);  /* poisson.rs:47-64 */
void   ora_get_laplace_2d(const ora_geom* g, const double* phi, const ora_model* m, double* lap); /* poisson.rs:13-36 */
void   ora_get_source_2d(const ora_geom* g, const double* phi, const ora_model* m, double* src);  /* poisson.rs:38-44 */
void   ora_poisson_coefs(int dim, const double* dx, double* coef /*[4]: a, cx, cy, cz*/);        /* poisson.rs:72-76; 3D SURVEY 8a */
int    ora_solve_poisson_gs(const ora_geom* g, double* psi, const double* rhs, const ora_bc_face* bc,
                            int64_t n_sub_iter);                            /* poisson.rs:67-92 (in place, lexicographic GS) */
int    ora_solve_poisson_gs_wavefront_2d(const ora_geom* g, double* psi, const double* rhs,
                            const ora_bc_face* bc, int64_t n_sub_iter);     /* same result by anti-diagonals */
int    ora_jacobi_sweeps(const ora_geom* g, double* psi, double* scratch, const double* rhs,
                         const ora_bc_face* bc, int64_t n_sweeps, double* l1_update /*nullable*/);
int    ora_jacobi_sweeps_omp(const ora_geom* g, double* psi, double* scratch, const double* rhs,
                         const ora_bc_face* bc, int64_t n_sweeps, double* l1_update);
double ora_sum_abs_valid(const ora_geom* g, const double* data);            /* poisson.rs:94-100 (naive, flat order) */
double ora_sum_abs_valid_ld(const ora_geom* g, const double* data);         /* same set, long-double accumulator */
void   ora_fill_ic_magnetostatics(const ora_geom* g, double* psi, const ora_model* m); /* main.rs:50-53 */
int    ora_magnetostatics_run(const ora_geom* g, double* psi, const ora_model* m, int jacobi,
                              double* sum_diff_trace /*[n_iter]*/, int64_t* iters_done); /* main.rs:29-74 */
void   ora_get_h_2d(const ora_geom* g, const double* psi, double* h /* 2-comp frame, ng same */); /* model.rs:72-88 */

int    ora_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
