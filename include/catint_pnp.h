/*
 * catint_pnp.h -- C ABI of the B200 batched 1D Poisson-Nernst-Planck solver.
 *
 * Drop-in boundary for CatINT's finite-difference transport hot path.  The
 * reference has no FFI: its boundary is the Python object protocol
 * Transport -> Calculator.run().  Each entry point below replaces
 * one piece of /root/reference/catint/calculator_old.py, batched over the
 * cells (sweep points) that /root/reference/catint/calculator.py:204-226 walks
 * serially:
 *
 *   catint_pnp_rhs_batch       <- integrate_odeint.ode_func        calculator_old.py:827-935
 *                                 + get_potential_and_gradient      calculator_old.py:680-819
 *                                 + Calculator.get_rates            calculator_old.py:159-208
 *   catint_pnp_jacobian_batch  <- (no reference code: odeint's internal finite-difference
 *                                 Jacobian, calculator_old.py:947) analytic block-tridiagonal
 *                                 Jacobian of the same residual in the local (c,g[,phi]) form
 *   catint_pnp_solve_batch     <- the integrator call               calculator_old.py:946-973
 *                                 (scipy odeint = ODEPACK LSODA) : variable-order BDF with a
 *                                 modified-Newton corrector, per-cell adaptive step, same error
 *                                 control (rtol/atol, weighted max norm), block-Thomas factors
 *                                 re-used across steps like LSODA re-uses its Jacobian
 *
 * Plain pointers and sizes only.  Unless stated otherwise every array pointer
 * is a DEVICE pointer owned by the caller; small model tables (CatintPnpShared)
 * are HOST memory.  No hidden allocation: scratch comes from a caller-provided
 * workspace sized by catint_pnp_workspace_bytes().  All functions return 0 on
 * success or a negative CATINT_PNP_E* code (never exit); catint_pnp_last_error()
 * gives the text.  Thread-safe per stream: no static launch state, the error text and the debug hook
 * are per host thread, every call checks that the current device is sm_100.
 *
 * Layouts (doubles):
 *   concentrations  c   [B][nx_max][S]      node-major, species interleaved
 *   local state     y   [B][nx_max][b]      b = S+1 (c..., g=dphi/dx) or S+2 (+phi, Stern mode)
 *   per-node fields g, phi [B][nx_max]
 *   blocks          Lb, Db, Ub [B][nx_max][b][b]   d(row)/d(col), row-major
 *   cell parameters par [B][CATINT_PNP_NPAR(S)]  see CatintPnpCells
 * Nodes >= nx[cell] of a cell are padding and are left untouched.
 */
#ifndef CATINT_PNP_H
#define CATINT_PNP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CATINT_PNP_MAX_SPECIES   14
#define CATINT_PNP_MAX_REACTIONS 12
#define CATINT_PNP_MAX_REACTANTS 4      /* per side of one reaction */
#define CATINT_PNP_MAX_BLOCK     13     /* catint_pnp_solve_batch / _jacobian_batch: unknowns per node, S+1
                                           (default Poisson boundary) or S+2 (Stern) must be 2..13        */
#define CATINT_PNP_MIN_NODES     6      /* every nx[cell] must satisfy MIN_NODES <= nx[cell] <= nx_max     */
#define CATINT_PNP_MAX_OUTPUT_TIMES 4096 /* n_out per catint_pnp_solve_batch call (the output times travel
                                           in the tail of the workspace)                                   */

enum {
    CATINT_PNP_OK = 0,
    CATINT_PNP_EINVAL = -1,      /* bad argument / unsupported size      */
    CATINT_PNP_ENOMEM = -2,      /* workspace too small                  */
    CATINT_PNP_ECUDA = -3,       /* CUDA runtime error                   */
    CATINT_PNP_ENODEV = -4       /* no sm_100 device                     */
};

/* Poisson boundary conditions = the reference's pb_bound combinations (catint/calculator_old.py:776-803; exactly one
   gradient and one potential, or both potentials).  0 is the reference default (catint/transport.py:1278-1322) and,
   with 1 (Stern/Robin wall, phi carried as an unknown; extension), what K1-K3 handle; 2-5 are available in
   catint_pnp_potential_batch and in the fixed-step steppers of catint_pnp_step_batch.  catint_pnp_solve_batch also
   takes 5: its field is the default pair's (backward sum from the bulk gradient, :787-790), the potential is
   integrated from the bulk end (:795-797) out of g_out (pass g_out with phi_out; no flux equations).          */
enum { CATINT_PNP_BC_DIRICHLET_WALL_NEUMANN_BULK = 0, CATINT_PNP_BC_STERN_ROBIN = 1,
       CATINT_PNP_BC_DIRICHLET_BOTH = 2,               /* potential at wall and bulk                           */
       CATINT_PNP_BC_DIRICHLET_WALL_NEUMANN_WALL = 3,  /* potential and gradient at the wall                   */
       CATINT_PNP_BC_DIRICHLET_BULK_NEUMANN_WALL = 4,  /* potential in the bulk, gradient at the wall          */
       CATINT_PNP_BC_DIRICHLET_BULK_NEUMANN_BULK = 5   /* potential and gradient in the bulk                   */ };
enum { CATINT_PNP_STEPPER_FTCS = 0, CATINT_PNP_STEPPER_CRANK_NICOLSON = 1 };
enum { CATINT_PNP_MODE_TRANSIENT = 0, CATINT_PNP_MODE_STEADY = 1,
       CATINT_PNP_MODE_MASK = 0xff,
       /* flag, OR-ed into CatintPnpControl.mode: keep every species in the block system.  By default a steady solve
          with one output time takes PASSIVE species (no charge or migration off, no part in any homogeneous
          reaction, not read by a flux equation) out of the block system -- their equations decouple and their
          discrete steady profile c_bulk + (J/D)(L - x) is written in closed form.                              */
       CATINT_PNP_MODE_KEEP_ALL = 0x100 };

/* per-cell status written by catint_pnp_solve_batch */
enum {
    CATINT_PNP_CELL_CONVERGED = 0,       /* reached t_end (and, steady mode, the Newton polish converged) */
    CATINT_PNP_CELL_MAX_STEPS = 1,
    CATINT_PNP_CELL_CORRECTOR_FAILED = 2,
    CATINT_PNP_CELL_ERROR_TEST_FAILED = 3,
    CATINT_PNP_CELL_NOT_FINITE = 4,
    CATINT_PNP_CELL_POLISH_FAILED = 5,   /* integrated to t_end but Newton on the steady residual did not converge */
    CATINT_PNP_CELL_STEP_UNDERFLOW = 6,  /* t+h == t: the discrete ODE blows up in finite time (LSODA: 'step size too small') */
    CATINT_PNP_CELL_BAD_INPUT = 7        /* nx[cell] outside CATINT_PNP_MIN_NODES..nx_max: the cell was skipped, its outputs untouched */
};

/* Model tables shared by all cells of a batch (HOST memory). */
typedef struct CatintPnpShared {
    int32_t S;                 /* transported species (catint/transport.py:264)                    */
    int32_t nx_max;            /* padded node count of the state arrays                             */
    int32_t R;                 /* homogeneous reactions with rate constants                         */
    int32_t poisson_bc;        /* CATINT_PNP_BC_*  (pb_bound, catint/transport.py:207-210)          */
    int32_t use_migration;     /* system['migration'] (catint/transport.py:316-319)                 */
    int32_t n_mesh;            /* rows of mesh_xi (0: every cell is uniform)                        */
    int32_t z[CATINT_PNP_MAX_SPECIES];                 /* integer charges (transport.py:1240-1276)  */
    /* reaction r:  educts -> products, net_r = kf*prod(c_educt) - kr*prod(c_product);
       species indices, one entry per stoichiometric unit, -1 padded; species that are not
       transported (H2O, e-) are simply absent (catint/calculator_old.py:169-187)                  */
    int32_t educt[CATINT_PNP_MAX_REACTIONS][CATINT_PNP_MAX_REACTANTS];
    int32_t product[CATINT_PNP_MAX_REACTIONS][CATINT_PNP_MAX_REACTANTS];
    double  kf[CATINT_PNP_MAX_REACTIONS];
    double  kr[CATINT_PNP_MAX_REACTIONS];
    /* R_k = sum_r nu[k][r]*net_r.  'summed' semantics: products minus educts multiplicities;
       the reference's legacy overwrite semantics is the same formula with another table
       (SURVEY 0-6), so the switch is data, not code.                                             */
    double  nu[CATINT_PNP_MAX_SPECIES][CATINT_PNP_MAX_REACTIONS];
    const struct CatintPnpFluxEq* flux_eq;             /* HOST pointer or NULL: wall kinetics (below)          */
} CatintPnpShared;

/* Flux equations (HOST): wall kinetics as expressions of the surface concentrations and the surface potential,
 * the reference's species[sp]['flux-equation'] (docs/source/topics/flux_definition.rst:100-156,
 * catint/comsol_model.py:986-1000), which only its COMSOL backend could evaluate.  The inward wall flux becomes
 *     J_k = par[FLUX + k] + sum_e coef[k][e] * E_e(c(0), phi(0); fpar[cell])
 * (the stoichiometric propagation of catint/transport.py:1057-1087 is linear, hence the coefficient table).
 * E_e is a postfix program of int32 words  opcode | operand << 8  evaluated on the device together with its
 * derivatives (they enter the wall block of the Newton matrix):
 *     0 CONST i  push consts[e][i]        4 ADD  5 SUB  6 MUL  7 DIV  8 POW  (binary, pop 2 push 1)
 *     1 PAR i    push fpar[cell][i]       9 NEG 10 EXP 11 LOG 12 SQRT 13 LOG10 14 TANH (unary)
 *     2 CONC k   push c_k(0)
 *     3 PHI      push phi(0) (the wall potential with the default Poisson boundary, an unknown with Stern)  */
#define CATINT_PNP_MAX_FLUX_EQ     4
#define CATINT_PNP_MAX_FLUX_CODE   96
#define CATINT_PNP_MAX_FLUX_CONST  32
#define CATINT_PNP_MAX_FLUX_PAR    24
#define CATINT_PNP_MAX_FLUX_STACK  16
typedef struct CatintPnpFluxEq {
    int32_t n_eq;                                      /* 0: no flux equations                                 */
    int32_t n_par;                                     /* doubles per cell in CatintPnpCells.fpar              */
    int32_t n_code[CATINT_PNP_MAX_FLUX_EQ];
    int32_t code[CATINT_PNP_MAX_FLUX_EQ][CATINT_PNP_MAX_FLUX_CODE];
    double  consts[CATINT_PNP_MAX_FLUX_EQ][CATINT_PNP_MAX_FLUX_CONST];
    double  coef[CATINT_PNP_MAX_SPECIES][CATINT_PNP_MAX_FLUX_EQ];
} CatintPnpFluxEq;

/* offsets inside one cell-parameter record of NPAR = 3*S+8 doubles */
#define CATINT_PNP_NPAR(S)      (3 * (S) + 8)
#define CATINT_PNP_P_CBULK(S)   0             /* [S] bulk concentrations, mol/m^3                    */
#define CATINT_PNP_P_FLUX(S)    (S)           /* [S] INWARD wall flux = +species[sp]['flux']          */
#define CATINT_PNP_P_DIFF(S)    (2 * (S))     /* [S] diffusion coefficients                           */
#define CATINT_PNP_P_BETA(S)    (3 * (S) + 0) /* 1/(R T)                                              */
#define CATINT_PNP_P_EPS(S)     (3 * (S) + 1) /* eps_r*eps_0                                          */
#define CATINT_PNP_P_PHIWALL(S) (3 * (S) + 2) /* Dirichlet wall potential   | Stern: phiM - phiPZC    */
#define CATINT_PNP_P_GBULK(S)   (3 * (S) + 3) /* Neumann bulk gradient      | Stern: unused           */
#define CATINT_PNP_P_CSTERN(S)  (3 * (S) + 4) /* Stern capacitance F/m^2                              */
#define CATINT_PNP_P_SCALE(S)   (3 * (S) + 5) /* uniform mesh: dx ; mesh table: x = scale*xi          */
#define CATINT_PNP_P_PHIBULK(S) (3 * (S) + 6) /* Dirichlet bulk potential (BC modes 2, 4, 5)                */
#define CATINT_PNP_P_GWALL(S)   (3 * (S) + 7) /* Neumann wall gradient    (BC modes 3, 4)                   */

/* Per-cell structure-of-arrays (DEVICE pointers). */
typedef struct CatintPnpCells {
    const double*  par;        /* [B][NPAR]                                                          */
    const int32_t* nx;         /* [B] nodes of each cell (ragged: 101 or 102 for nx=100, SURVEY C-6) */
    const int32_t* mesh_id;    /* [B] row of mesh_xi, or -1 = uniform mesh x_i = i*scale             */
    const double*  mesh_xi;    /* [n_mesh][nx_max] normalised node positions (may be NULL)           */
    const double*  fpar;       /* [B][flux_eq->n_par] per-cell parameters of the flux equations, or NULL */
    const int32_t* order;      /* [B] optional launch order for catint_pnp_solve_batch: a permutation of 0..B-1,
                                  expensive cells first (launch slot k works on cell order[k]); NULL = 0..B-1.
                                  A scheduling hint only: results and their layout do not depend on it.  */
} CatintPnpCells;

/* Integrator control (HOST). */
typedef struct CatintPnpControl {
    int32_t mode;              /* CATINT_PNP_MODE_*                                                  */
    int32_t max_steps;         /* per cell                                                           */
    int32_t n_out;             /* number of output times (>=1); last one is t_end                    */
    int32_t polish_max_iter;   /* steady mode: Newton iterations on the steady residual              */
    double  rtol, atol;        /* scipy odeint defaults 1.49012e-8 (calculator_old.py:947)           */
    double  h0;                /* first step; <=0: chosen from the initial rate of change            */
    double  polish_rtol;       /* steady mode: |delta| <= polish_rtol*|y| + tiny                     */
    const double* t_out;       /* HOST [n_out] increasing output times                               */
} CatintPnpControl;

int catint_pnp_version(void);
const char* catint_pnp_last_error(void);

/* number of visible CUDA devices of compute capability 10.x (0 if none) */
int catint_pnp_device_count(void);

/* Debug hook: device buffer of int64 [n_cells][8] that the next catint_pnp_solve_batch calls fill with
 * SM cycle counts per phase (factor, residual, forward, backward, norms, history, correction, total);
 * NULL (default) switches the instrumentation off. */
void catint_pnp_debug_profile_buffer(void* dev_ptr);

size_t catint_pnp_workspace_bytes(const CatintPnpShared* sh, int64_t n_cells);

/* K1: dc/dt for every cell.  g_out/phi_out may be NULL. */
int catint_pnp_rhs_batch(const CatintPnpShared* sh, const CatintPnpCells* cells, int64_t n_cells,
                         const double* c, double* dcdt, double* g_out, double* phi_out,
                         void* cuda_stream);

/* K2: residual F(y) and blocks of dF/dy in the local form (F and any block pointer may be NULL). */
int catint_pnp_jacobian_batch(const CatintPnpShared* sh, const CatintPnpCells* cells, int64_t n_cells,
                              const double* y, double* F, double* Lb, double* Db, double* Ub,
                              void* cuda_stream);

/* K3: time integration / steady state.
 *   y0        optional initial concentrations [B][nx_max][S] (NULL: bulk everywhere, the reference's
 *             c0, catint/transport.py:1396-1412)
 *   c_out     [n_out][B][nx_max][S]   phi_out, g_out [n_out][B][nx_max]
 *   flux_out  [B][S] discrete wall flux of the final state
 *   status, n_steps, n_newton  [B]; n_setups [B] (block factorisations, may be NULL)
 */
int catint_pnp_solve_batch(const CatintPnpShared* sh, const CatintPnpCells* cells, int64_t n_cells,
                           const double* y0, const CatintPnpControl* ctl,
                           double* c_out, double* phi_out, double* g_out, double* flux_out,
                           int32_t* status, int32_t* n_steps, int32_t* n_newton, int32_t* n_setups,
                           void* workspace, size_t workspace_bytes, void* cuda_stream);

/* get_potential_and_gradient (catint/calculator_old.py:680-819) for every pb_bound combination:
 *   c [B][nx_max][S] -> v, grad_v, lapl_v [B][nx_max] (any output may be NULL).  Uniform meshes.               */
int catint_pnp_potential_batch(const CatintPnpShared* sh, const CatintPnpCells* cells, int64_t n_cells,
                               const double* c, double* v, double* grad_v, double* lapl_v, void* cuda_stream);

/* K4: the reference's fixed-step steppers, integrate_FTCS (catint/calculator_old.py:976-1029) and
 * integrate_Crank_Nicolson (:457-564), dispatch :1121-1140: nt steps of size dt from c0 [B][nx_max][S] (NULL: bulk
 * everywhere), the state after step itout[k] (DEVICE int32 [n_out], increasing; FTCS counts steps from 0,
 * Crank-Nicolson from 1 as the reference does) goes to c_out [n_out][B][nx_max][S]; phi_out / g_out
 * [n_out][B][nx_max] optional (the field the step was taken with).  Wall potential of the Robin condition
 * ("vzeta") = CATINT_PNP_P_PHIWALL.  Uniform meshes, every pb_bound combination but Stern.                     */
int catint_pnp_step_batch(const CatintPnpShared* sh, const CatintPnpCells* cells, int64_t n_cells,
                          const double* c0, int32_t stepper, int32_t lax_friedrich, double dt, int32_t nt,
                          const int32_t* itout, int32_t n_out,
                          double* c_out, double* phi_out, double* g_out, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* CATINT_PNP_H */
