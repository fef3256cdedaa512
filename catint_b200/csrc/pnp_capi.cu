// pnp_capi.cu -- the C ABI of include/catint_pnp.h: argument checks, model-table
// conversion and kernel launches.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "pnp_kernels.cuh"
#include "pnp_rhs.cuh"
#include "pnp_explicit.cuh"

namespace catint {
#define X(NB) extern template int launch_bdf<NB, false>(SolveParams&, cudaStream_t); \
              extern template int launch_jac<NB, false>(JacParams&, cudaStream_t);
X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13)
#undef X
#define X(NB) extern template int launch_bdf<NB, true>(SolveParams&, cudaStream_t);
X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13)
#undef X
}

// ===========================================================================
// C ABI
// ===========================================================================
using namespace catint;

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, const char* a = "") {
    snprintf(g_err, sizeof(g_err), fmt, a);
    return code;
}

static thread_local long long* g_prof = nullptr;      // debug hook (per host thread), see catint_pnp_debug_profile_buffer
extern "C" int catint_pnp_version(void) { return 101; }
extern "C" void catint_pnp_debug_profile_buffer(void* dev_ptr) { g_prof = reinterpret_cast<long long*>(dev_ptr); }
extern "C" const char* catint_pnp_last_error(void) { return g_err; }

extern "C" int catint_pnp_device_count(void) {
    // cudaGetDeviceProperties costs milliseconds: query once per process
    static int cached = -1;
    if (cached >= 0) return cached;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++ok;
    }
    cached = ok;
    return ok;
}

static int build_tables(const CatintPnpShared* sh, DevTables& tb) {
    memset(&tb, 0, sizeof(tb));
    if (!sh) return fail(CATINT_PNP_EINVAL, "shared tables are NULL");
    if (sh->S < 1 || sh->S > CATINT_PNP_MAX_SPECIES) return fail(CATINT_PNP_EINVAL, "S out of range");
    if (sh->R < 0 || sh->R > CATINT_PNP_MAX_REACTIONS) return fail(CATINT_PNP_EINVAL, "R out of range");
    if (sh->nx_max < 4) return fail(CATINT_PNP_EINVAL, "nx_max must be >= 4");
    tb.S = sh->S; tb.R = sh->R; tb.nx_max = sh->nx_max;
    tb.S_full = sh->S; tb.npas = 0;
    for (int k = 0; k < MAXS; ++k) tb.cmap[k] = (int8_t)k;
    tb.stern = sh->poisson_bc == CATINT_PNP_BC_STERN_ROBIN;
    tb.use_migration = sh->use_migration != 0;
    for (int k = 0; k < sh->S; ++k) tb.z[k] = (double)sh->z[k];
    for (int r = 0; r < sh->R; ++r) {
        int ne = 0, np = 0;
        for (int e = 0; e < CATINT_PNP_MAX_REACTANTS; ++e) {
            if (sh->educt[r][e] >= 0) {
                if (sh->educt[r][e] >= sh->S) return fail(CATINT_PNP_EINVAL, "educt index out of range");
                tb.ed[r][ne++] = (int8_t)sh->educt[r][e];
            }
            if (sh->product[r][e] >= 0) {
                if (sh->product[r][e] >= sh->S) return fail(CATINT_PNP_EINVAL, "product index out of range");
                tb.pr[r][np++] = (int8_t)sh->product[r][e];
            }
        }
        tb.ned[r] = (int8_t)ne; tb.npr[r] = (int8_t)np;
        tb.kf[r] = sh->kf[r]; tb.kr[r] = sh->kr[r];
        for (int k = 0; k < sh->S; ++k) tb.nu[r][k] = sh->nu[k][r];
    }
    // derivative terms, grouped by the species the derivative is taken with respect to
    int T = 0;
    for (int j = 0; j < sh->S; ++j) {
        tb.tbeg[j] = (int8_t)T;
        for (int r = 0; r < sh->R; ++r) {
            for (int side = 0; side < 2; ++side) {
                const int8_t* lst = side == 0 ? tb.ed[r] : tb.pr[r];
                const int cnt = side == 0 ? tb.ned[r] : tb.npr[r];
                for (int p = 0; p < cnt; ++p) {
                    if (lst[p] != j) continue;
                    if (T >= MAXT) return fail(CATINT_PNP_EINVAL, "too many reaction derivative terms");
                    int8_t others[3] = {-1, -1, -1};
                    int no = 0;
                    for (int p2 = 0; p2 < cnt; ++p2) if (p2 != p) others[no++] = lst[p2];
                    tb.tr[T] = (int8_t)r;
                    tb.ti1[T] = others[0]; tb.ti2[T] = others[1]; tb.ti3[T] = others[2];
                    tb.tcoef[T] = side == 0 ? tb.kf[r] : -tb.kr[r];
                    ++T;
                }
            }
        }
    }
    tb.tbeg[sh->S] = (int8_t)T;
    for (int j = sh->S + 1; j <= MAXS; ++j) tb.tbeg[j] = (int8_t)T;
    tb.T = T;
    // flux equations: copy and validate the programs (operands in range, stack never under- or overflows)
    tb.fpar = nullptr;
    if (sh->flux_eq && sh->flux_eq->n_eq > 0) {
        const CatintPnpFluxEq& fq = *sh->flux_eq;
        if (fq.n_eq > CATINT_PNP_MAX_FLUX_EQ || fq.n_par < 0 || fq.n_par > CATINT_PNP_MAX_FLUX_PAR)
            return fail(CATINT_PNP_EINVAL, "flux equations: n_eq / n_par out of range");
        for (int e = 0; e < fq.n_eq; ++e) {
            if (fq.n_code[e] < 1 || fq.n_code[e] > CATINT_PNP_MAX_FLUX_CODE)
                return fail(CATINT_PNP_EINVAL, "flux equations: program length out of range");
            int depth = 0;
            for (int k = 0; k < fq.n_code[e]; ++k) {
                const int op = fq.code[e][k] & 0xff, arg = fq.code[e][k] >> 8;
                if (op > 14) return fail(CATINT_PNP_EINVAL, "flux equations: unknown opcode");
                if (op <= 3) {
                    if ((op == 0 && (arg < 0 || arg >= CATINT_PNP_MAX_FLUX_CONST)) || (op == 1 && (arg < 0 || arg >= fq.n_par)) ||
                        (op == 2 && (arg < 0 || arg >= sh->S)))
                        return fail(CATINT_PNP_EINVAL, "flux equations: operand out of range");
                    ++depth;
                } else if (op <= 8) {
                    if (depth < 2) return fail(CATINT_PNP_EINVAL, "flux equations: malformed program (stack underflow)");
                    --depth;
                } else if (depth < 1) {
                    return fail(CATINT_PNP_EINVAL, "flux equations: malformed program (stack underflow)");
                }
                if (depth > CATINT_PNP_MAX_FLUX_STACK) return fail(CATINT_PNP_EINVAL, "flux equations: stack too deep");
            }
            if (depth != 1) return fail(CATINT_PNP_EINVAL, "flux equations: malformed program");
        }
        tb.fq = fq;
    }
    return CATINT_PNP_OK;
}

// Steady-mode elimination of PASSIVE species.  A species is passive when nothing depends on its concentration
// but its own diffusion: no charge (or migration switched off), no part in any homogeneous reaction (neither
// stoichiometry nor rate law), not read by any flux equation.  Its equations c_t = D c_xx decouple from the block
// system, and at the steady state of a finite boundary layer (Dirichlet bulk node, flux wall) their discrete
// solution is exactly linear -- the kernel writes it in closed form (pnp_bdf_kernel, final outputs).  The block
// system shrinks from S_full+1 to S+1 unknowns per node (C2: 9 -> 7, C4: 11 -> 8), which is ~b^2 in the solve
// sweeps and ~b^3 in the factorisation.  Only in CATINT_PNP_MODE_STEADY with one output time (transient
// outputs need the passive species' history); CATINT_PNP_MODE_KEEP_ALL switches it off.
static bool reduce_passive(const DevTables& full, DevTables& out) {
    bool passive[MAXS];
    const int SF = full.S;
    for (int k = 0; k < SF; ++k) passive[k] = (full.z[k] == 0.0) || !full.use_migration;
    for (int r = 0; r < full.R; ++r) {
        for (int e = 0; e < full.ned[r]; ++e) passive[full.ed[r][e]] = false;
        for (int e = 0; e < full.npr[r]; ++e) passive[full.pr[r][e]] = false;
        for (int k = 0; k < SF; ++k) if (full.nu[r][k] != 0.0) passive[k] = false;
    }
    for (int e = 0; e < full.fq.n_eq; ++e)
        for (int w = 0; w < full.fq.n_code[e]; ++w)
            if ((full.fq.code[e][w] & 0xff) == 2) passive[full.fq.code[e][w] >> 8] = false;
    int nc = 0, np = 0, idx[MAXS];
    for (int k = 0; k < SF; ++k) { idx[k] = passive[k] ? -1 : nc; if (passive[k]) ++np; else ++nc; }
    if (np == 0 || nc == 0) return false;
    out = full;
    out.S = nc; out.S_full = SF; out.npas = np;
    memset(out.z, 0, sizeof(out.z)); memset(out.nu, 0, sizeof(out.nu));
    memset(out.fq.coef, 0, sizeof(out.fq.coef)); memset(out.pcoef, 0, sizeof(out.pcoef));
    int ip = 0;
    for (int k = 0; k < SF; ++k) {
        if (passive[k]) {
            out.pmap[ip] = (int8_t)k;
            for (int e = 0; e < full.fq.n_eq; ++e) out.pcoef[ip][e] = full.fq.coef[k][e];
            ++ip;
            continue;
        }
        const int kc = idx[k];
        out.cmap[kc] = (int8_t)k;
        out.z[kc] = full.z[k];
        for (int r = 0; r < full.R; ++r) out.nu[r][kc] = full.nu[r][k];
        for (int e = 0; e < full.fq.n_eq; ++e) out.fq.coef[kc][e] = full.fq.coef[k][e];
        out.tbeg[kc] = full.tbeg[k];              // passive species own no derivative terms: the ranges stay contiguous
    }
    for (int j = nc; j <= MAXS; ++j) out.tbeg[j] = (int8_t)full.T;
    for (int r = 0; r < full.R; ++r) {
        for (int e = 0; e < full.ned[r]; ++e) out.ed[r][e] = (int8_t)idx[full.ed[r][e]];
        for (int e = 0; e < full.npr[r]; ++e) out.pr[r][e] = (int8_t)idx[full.pr[r][e]];
    }
    for (int t = 0; t < full.T; ++t) {
        if (full.ti1[t] >= 0) out.ti1[t] = (int8_t)idx[full.ti1[t]];
        if (full.ti2[t] >= 0) out.ti2[t] = (int8_t)idx[full.ti2[t]];
        if (full.ti3[t] >= 0) out.ti3[t] = (int8_t)idx[full.ti3[t]];
    }
    for (int e = 0; e < full.fq.n_eq; ++e)
        for (int w = 0; w < full.fq.n_code[e]; ++w)
            if ((full.fq.code[e][w] & 0xff) == 2) out.fq.code[e][w] = 2 | (idx[full.fq.code[e][w] >> 8] << 8);
    return true;
}

// per-cell parameters of the flux equations
static int attach_fpar(DevTables& tb, const CatintPnpCells* cells) {
    if (tb.fq.n_eq > 0 && tb.fq.n_par > 0 && !cells->fpar)
        return fail(CATINT_PNP_EINVAL, "flux equations need cells->fpar");
    tb.fpar = cells->fpar;
    return CATINT_PNP_OK;
}

static int block_size_of(const CatintPnpShared* sh) {
    return sh->S + (sh->poisson_bc == CATINT_PNP_BC_STERN_ROBIN ? 2 : 1);
}

static size_t ws_doubles_per_cell(const CatintPnpShared* sh) {
    const size_t NB = (size_t)block_size_of(sh), nxm = (size_t)sh->nx_max;
    // zn[LMAX][N] + ewt[N] + inv[nx][NB][NBP] + la[nx][NB][4] + V0,W1,dJ[NB][NBP] + (y,psi,zb)[N] (used only when the state
    // does not fit in shared memory, always reserved so that the size query is stateless)
    const size_t NBP = NB + (NB & 1);
    const size_t REC = NB * NBP + NB * 4;
    size_t d = align4((size_t)LMAX * nxm * NB) + align4(nxm * NB) + align4(nxm * REC) + 3 * align4(NB * NBP) +
               3 * align4(nxm * NB);
    return (d + 15) & ~size_t(15);
}

extern "C" size_t catint_pnp_workspace_bytes(const CatintPnpShared* sh, int64_t n_cells) {
    if (!sh || n_cells <= 0) return 0;
    return ws_doubles_per_cell(sh) * sizeof(double) * (size_t)n_cells + CATINT_PNP_MAX_OUTPUT_TIMES * sizeof(double);
}

static int check_cuda(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
        return CATINT_PNP_ECUDA;
    }
    return CATINT_PNP_OK;
}

#define DISPATCH_NB(NBVAL, CALL, STV, ...)                          \
    switch (NBVAL) {                                                \
        case 2: rc = CALL<2, false>(__VA_ARGS__); break;            \
        case 3: rc = CALL<3, STV>(__VA_ARGS__); break;              \
        case 4: rc = CALL<4, STV>(__VA_ARGS__); break;              \
        case 5: rc = CALL<5, STV>(__VA_ARGS__); break;              \
        case 6: rc = CALL<6, STV>(__VA_ARGS__); break;              \
        case 7: rc = CALL<7, STV>(__VA_ARGS__); break;              \
        case 8: rc = CALL<8, STV>(__VA_ARGS__); break;              \
        case 9: rc = CALL<9, STV>(__VA_ARGS__); break;              \
        case 10: rc = CALL<10, STV>(__VA_ARGS__); break;            \
        case 11: rc = CALL<11, STV>(__VA_ARGS__); break;            \
        case 12: rc = CALL<12, STV>(__VA_ARGS__); break;            \
        case 13: rc = CALL<13, STV>(__VA_ARGS__); break;            \
        default: rc = fail(CATINT_PNP_EINVAL, "unsupported block size (2..13 unknowns per node)"); \
    }

static int check_common(const CatintPnpShared* sh, const CatintPnpCells* cells, int64_t n_cells) {
    if (!sh || !cells) return fail(CATINT_PNP_EINVAL, "NULL argument");
    if (n_cells <= 0) return fail(CATINT_PNP_EINVAL, "n_cells must be positive");
    if (!cells->par || !cells->nx) return fail(CATINT_PNP_EINVAL, "cells->par / cells->nx are NULL");
    if (sh->n_mesh > 0 && !cells->mesh_xi) return fail(CATINT_PNP_EINVAL, "mesh table missing");
    if (catint_pnp_device_count() <= 0) return fail(CATINT_PNP_ENODEV, "no sm_100 CUDA device visible");
    // the kernels are sm_100a code: the CURRENT device (the one the caller's stream and pointers live on) must be one
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
        cudaGetLastError();
        return fail(CATINT_PNP_ECUDA, "cannot query the current CUDA device");
    }
    if (major != 10) return fail(CATINT_PNP_ENODEV, "the current CUDA device is not sm_100");
    return CATINT_PNP_OK;
}

extern "C" int catint_pnp_rhs_batch(const CatintPnpShared* sh, const CatintPnpCells* cells, int64_t n_cells,
                                    const double* c, double* dcdt, double* g_out, double* phi_out,
                                    void* cuda_stream) {
    int rc = check_common(sh, cells, n_cells);
    if (rc) return rc;
    if (sh->poisson_bc != CATINT_PNP_BC_DIRICHLET_WALL_NEUMANN_BULK)
        return fail(CATINT_PNP_EINVAL, "catint_pnp_rhs_batch: only the default Poisson boundary (wall potential + bulk gradient); "
                                       "Stern: catint_pnp_solve_batch, other pb_bound pairs: catint_pnp_potential_batch / _step_batch");
    if (!c || !dcdt) return fail(CATINT_PNP_EINVAL, "c / dcdt are NULL");
    RhsParams P;
    rc = build_tables(sh, P.tb);
    if (rc) return rc;
    rc = attach_fpar(P.tb, cells);
    if (rc) return rc;
    P.par = cells->par; P.nx = cells->nx; P.mesh_id = cells->mesh_id; P.mesh_xi = cells->mesh_xi;
    P.c = c; P.n_cells = n_cells; P.dcdt = dcdt; P.g_out = g_out; P.phi_out = phi_out;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    switch (sh->S) {
#define X(SV) case SV: rc = launch_rhs<SV>(P, st); break;
        X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14)
#undef X
        default: rc = CATINT_PNP_EINVAL;
    }
    if (rc == CATINT_PNP_EINVAL) return fail(rc, "pnp_rhs_kernel: grid too large for shared memory (nx_max*S)");
    if (rc == CATINT_PNP_ECUDA) return fail(rc, "pnp_rhs_kernel launch failed");
    return rc;
}

extern "C" int catint_pnp_jacobian_batch(const CatintPnpShared* sh, const CatintPnpCells* cells, int64_t n_cells,
                                         const double* y, double* F, double* Lb, double* Db, double* Ub,
                                         void* cuda_stream) {
    int rc = check_common(sh, cells, n_cells);
    if (rc) return rc;
    if (sh->poisson_bc != CATINT_PNP_BC_DIRICHLET_WALL_NEUMANN_BULK)
        return fail(CATINT_PNP_EINVAL, "catint_pnp_jacobian_batch: only the default Poisson boundary");
    if (!y) return fail(CATINT_PNP_EINVAL, "y is NULL");
    JacParams P;
    rc = build_tables(sh, P.tb);
    if (rc) return rc;
    rc = attach_fpar(P.tb, cells);
    if (rc) return rc;
    P.par = cells->par; P.nx = cells->nx; P.mesh_id = cells->mesh_id; P.mesh_xi = cells->mesh_xi;
    P.y = y; P.n_cells = n_cells; P.F = F; P.Lb = Lb; P.Db = Db; P.Ub = Ub;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    DISPATCH_NB(block_size_of(sh), launch_jac, false, P, st);
    if (rc == CATINT_PNP_ECUDA) return fail(rc, "pnp_jacobian_kernel launch failed");
    return rc;
}

extern "C" int catint_pnp_solve_batch(const CatintPnpShared* sh, const CatintPnpCells* cells, int64_t n_cells,
                                      const double* y0, const CatintPnpControl* ctl,
                                      double* c_out, double* phi_out, double* g_out, double* flux_out,
                                      int32_t* status, int32_t* n_steps, int32_t* n_newton, int32_t* n_setups,
                                      void* workspace, size_t workspace_bytes, void* cuda_stream) {
    int rc = check_common(sh, cells, n_cells);
    if (rc) return rc;
    if (sh->nx_max < CATINT_PNP_MIN_NODES) return fail(CATINT_PNP_EINVAL, "nx_max must be >= CATINT_PNP_MIN_NODES");
    // the bulk/bulk pair has the field of the default pair (backward sum from the bulk gradient); only its potential
    // is integrated from the bulk end: same kernel, potential rewritten from g_out afterwards
    const bool bulk_pair = sh->poisson_bc == CATINT_PNP_BC_DIRICHLET_BULK_NEUMANN_BULK;
    if (sh->poisson_bc != CATINT_PNP_BC_DIRICHLET_WALL_NEUMANN_BULK && sh->poisson_bc != CATINT_PNP_BC_STERN_ROBIN && !bulk_pair)
        return fail(CATINT_PNP_EINVAL, "catint_pnp_solve_batch: Poisson boundary must be the default pair, the bulk/bulk "
                                       "pair or Stern (the other pb_bound pairs run in catint_pnp_step_batch)");
    if (bulk_pair && phi_out && !g_out)
        return fail(CATINT_PNP_EINVAL, "catint_pnp_solve_batch: the bulk/bulk pair derives phi_out from g_out; pass both");
    if (bulk_pair && sh->flux_eq && sh->flux_eq->n_eq > 0)
        return fail(CATINT_PNP_EINVAL, "catint_pnp_solve_batch: flux equations read the wall potential; not with the bulk/bulk pair");
    if (!ctl || !ctl->t_out || ctl->n_out < 1) return fail(CATINT_PNP_EINVAL, "control / t_out missing");
    if (!c_out || !status || !n_steps || !n_newton) return fail(CATINT_PNP_EINVAL, "output pointers are NULL");
    if (!(ctl->rtol >= 0.0) || !(ctl->atol > 0.0)) return fail(CATINT_PNP_EINVAL, "need rtol >= 0 and atol > 0");
    for (int k = 0; k < ctl->n_out; ++k)
        if (!(ctl->t_out[k] > 0.0) || (k > 0 && !(ctl->t_out[k] > ctl->t_out[k - 1])))
            return fail(CATINT_PNP_EINVAL, "t_out must be positive and increasing");
    const size_t need = catint_pnp_workspace_bytes(sh, n_cells);
    if (!workspace || workspace_bytes < need) return fail(CATINT_PNP_ENOMEM, "workspace too small");
    if (ctl->n_out > CATINT_PNP_MAX_OUTPUT_TIMES)
        return fail(CATINT_PNP_EINVAL, "too many output times per call (CATINT_PNP_MAX_OUTPUT_TIMES)");

    SolveParams P;
    rc = build_tables(sh, P.tb);
    if (rc) return rc;
    rc = attach_fpar(P.tb, cells);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int mode = ctl->mode & CATINT_PNP_MODE_MASK;
    int nb = block_size_of(sh);
    if (mode == CATINT_PNP_MODE_STEADY && ctl->n_out == 1 && !(ctl->mode & CATINT_PNP_MODE_KEEP_ALL)) {
        DevTables red;
        if (reduce_passive(P.tb, red)) { P.tb = red; nb -= red.npas; }
    }
    P.par = cells->par; P.nx = cells->nx; P.mesh_id = cells->mesh_id; P.mesh_xi = cells->mesh_xi;
    P.y0 = y0; P.n_cells = n_cells; P.order = cells->order;
    P.mode = mode; P.max_steps = ctl->max_steps > 0 ? ctl->max_steps : 100000;
    P.n_out = ctl->n_out; P.polish_max_iter = ctl->polish_max_iter > 0 ? ctl->polish_max_iter : 8;
    P.rtol = ctl->rtol; P.atol = ctl->atol; P.h0 = ctl->h0;
    P.polish_rtol = ctl->polish_rtol > 0.0 ? ctl->polish_rtol : 1e-10;
    // the output times travel in the tail of the workspace
    double* wsd = reinterpret_cast<double*>(workspace);
    double* t_dev = wsd + ws_doubles_per_cell(sh) * (size_t)n_cells;
    if (cudaMemcpyAsync(t_dev, ctl->t_out, sizeof(double) * ctl->n_out, cudaMemcpyHostToDevice, st) != cudaSuccess)
        return check_cuda("copy of t_out");
    P.t_out = t_dev;
    P.c_out = c_out; P.phi_out = phi_out; P.g_out = g_out; P.flux_out = flux_out;
    P.status = status; P.n_steps = n_steps; P.n_newton = n_newton; P.n_setups = n_setups;
    P.ws = wsd; P.ws_stride = (long long)ws_doubles_per_cell(sh);
    P.state_in_smem = 1;
    P.prof = g_prof;
    if (sh->poisson_bc == CATINT_PNP_BC_STERN_ROBIN) {
        DISPATCH_NB(nb, launch_bdf, true, P, st);
    } else {
        DISPATCH_NB(nb, launch_bdf, false, P, st);
    }
    if (rc == CATINT_PNP_ECUDA) return fail(rc, "pnp_bdf_kernel launch failed");
    if (rc == CATINT_PNP_OK && bulk_pair && phi_out) {
        BulkPotentialParams Q;
        Q.par = cells->par; Q.nx = cells->nx; Q.mesh_id = cells->mesh_id; Q.mesh_xi = cells->mesh_xi;
        Q.g = g_out; Q.v = phi_out; Q.n_cells = n_cells; Q.rows = (long long)ctl->n_out * n_cells;
        Q.S = sh->S; Q.nx_max = sh->nx_max;
        if (launch_bulk_potential(Q, st) != 0) return fail(CATINT_PNP_ECUDA, "pnp_bulk_potential_kernel launch failed");
    }
    return rc;
}

static int check_explicit(const CatintPnpShared* sh, const CatintPnpCells* cells) {
    if (sh->poisson_bc == CATINT_PNP_BC_STERN_ROBIN || sh->poisson_bc < 0 || sh->poisson_bc > CATINT_PNP_BC_DIRICHLET_BULK_NEUMANN_BULK)
        return fail(CATINT_PNP_EINVAL, "Poisson boundary must be one of the reference's pb_bound pairs (not Stern)");
    if (cells->mesh_id || sh->n_mesh > 0) return fail(CATINT_PNP_EINVAL, "the reference's Poisson routine and fixed-step steppers need a uniform mesh");
    if (sh->flux_eq && sh->flux_eq->n_eq > 0) return fail(CATINT_PNP_EINVAL, "flux equations are not available in the fixed-step steppers");
    return CATINT_PNP_OK;
}

extern "C" int catint_pnp_potential_batch(const CatintPnpShared* sh, const CatintPnpCells* cells, int64_t n_cells,
                                          const double* c, double* v, double* grad_v, double* lapl_v, void* cuda_stream) {
    int rc = check_common(sh, cells, n_cells);
    if (rc) return rc;
    rc = check_explicit(sh, cells);
    if (rc) return rc;
    if (!c) return fail(CATINT_PNP_EINVAL, "c is NULL");
    PotentialParams P;
    rc = build_tables(sh, P.tb);
    if (rc) return rc;
    P.par = cells->par; P.nx = cells->nx; P.c = c; P.n_cells = n_cells;
    P.v = v; P.grad = grad_v; P.lapl = lapl_v; P.bc = sh->poisson_bc;
    rc = launch_potential(P, (cudaStream_t)cuda_stream);
    if (rc == CATINT_PNP_EINVAL) return fail(rc, "pnp_potential_kernel: grid too large for shared memory");
    if (rc == CATINT_PNP_ECUDA) return fail(rc, "pnp_potential_kernel launch failed");
    return rc;
}

extern "C" int catint_pnp_step_batch(const CatintPnpShared* sh, const CatintPnpCells* cells, int64_t n_cells,
                                     const double* c0, int32_t stepper, int32_t lax_friedrich, double dt, int32_t nt,
                                     const int32_t* itout, int32_t n_out,
                                     double* c_out, double* phi_out, double* g_out, void* cuda_stream) {
    int rc = check_common(sh, cells, n_cells);
    if (rc) return rc;
    rc = check_explicit(sh, cells);
    if (rc) return rc;
    if (!c0 || !c_out || !itout) return fail(CATINT_PNP_EINVAL, "c0 / c_out / itout are NULL");
    if (stepper != CATINT_PNP_STEPPER_FTCS && stepper != CATINT_PNP_STEPPER_CRANK_NICOLSON)
        return fail(CATINT_PNP_EINVAL, "unknown stepper");
    if (!(dt > 0.0) || nt < 1 || n_out < 1) return fail(CATINT_PNP_EINVAL, "need dt > 0, nt >= 1, n_out >= 1");
    ExplicitParams P;
    rc = build_tables(sh, P.tb);
    if (rc) return rc;
    P.par = cells->par; P.nx = cells->nx; P.c0 = c0; P.n_cells = n_cells;
    P.method = stepper; P.lax_friedrich = lax_friedrich; P.nt = nt; P.n_out = n_out; P.dt = dt;
    P.itout = itout; P.c_out = c_out; P.phi_out = phi_out; P.g_out = g_out; P.bc = sh->poisson_bc;
    rc = launch_explicit(P, (cudaStream_t)cuda_stream);
    if (rc == CATINT_PNP_EINVAL) return fail(rc, "pnp_explicit_kernel: grid too large for shared memory ((3S+3)*nx_max doubles per cell)");
    if (rc == CATINT_PNP_ECUDA) return fail(rc, "pnp_explicit_kernel launch failed");
    return rc;
}
