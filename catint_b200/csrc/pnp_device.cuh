// pnp_device.cuh -- device-side model tables and per-node arithmetic of the
// 1D finite-difference Poisson-Nernst-Planck system in its local
// (block-tridiagonal) form.
//
// Discrete equations restated from /root/reference/catint/calculator_old.py
// (SURVEY Appendix A): transport rows :884-927, wall stencil :902-909 (with the
// inward-flux sign of docs/source/topics/flux_definition.rst:60-65), Poisson
// cumulative sums :753-761,793-800 carried as the algebraic unknown g=dphi/dx,
// homogeneous mass-action rates :159-208 (as R_k = sum_r nu[k][r]*net_r).
//
// Unknowns per node  y_i = (c_0..c_{S-1}, g);  NB = S+1 = block size.
// "Newton matrix"    A = Mass/gamma - dF/dy   with blocks A_L, A_D, A_U (and
// the extra wall block A_E = coupling of node 0 to node 2).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "pnp_fluxeq.cuh"

namespace catint {

constexpr int MAXS = 14;
constexpr int MAXR = 12;
constexpr int MAXRT = 4;      // reactants per side
constexpr int MAXT = 96;      // derivative terms: sum over reactions of (#educts+#products)
constexpr double UNIT_F = 96485.33289;   // catint/units.py:16
constexpr unsigned FULL = 0xffffffffu;

// Model tables, passed by value as a kernel parameter and copied to shared memory.
struct DevTables {
    int S, R, T, stern, use_migration, nx_max;
    double z[MAXS];
    int8_t ned[MAXR], npr[MAXR];
    int8_t ed[MAXR][MAXRT], pr[MAXR][MAXRT];
    double kf[MAXR], kr[MAXR];
    double nu[MAXR][MAXS];          // nu[r][k]
    // d net_r / d c_j = sum over terms t of species j of tcoef[t]*c[ti1[t]]*c[ti2[t]]*c[ti3[t]]
    // (index -1: factor 1); terms sorted by j, species j owns [tbeg[j], tbeg[j+1])
    int8_t tr[MAXT], ti1[MAXT], ti2[MAXT], ti3[MAXT];
    double tcoef[MAXT];
    int8_t tbeg[MAXS + 1];
    // Steady-mode elimination of PASSIVE species (pnp_capi.cu: reduce_passive): the tables above describe the
    // S coupled species the kernel integrates; the caller's model has S_full >= S species.  cmap: kernel species
    // -> caller's species; pmap: the npas eliminated species (caller's indices); pcoef (below): their flux-equation
    // coefficients.  Without elimination S_full == S and cmap is the identity.
    int S_full, npas;
    int8_t cmap[MAXS], pmap[MAXS];
    // wall kinetics (flux equations): programs + coefficient table (by value), per-cell parameters (device).
    // LAST members: a kernel copies only tables_prefix_bytes() to shared memory when there are no flux equations.
    const double* fpar;
    CatintPnpFluxEq fq;
    double pcoef[MAXS][CATINT_PNP_MAX_FLUX_EQ];
};

// bytes of the tables a kernel needs in shared memory: without flux equations everything up to fq.n_eq / n_par
__host__ __device__ inline size_t tables_prefix_bytes(const DevTables& tb) {
    return tb.fq.n_eq > 0 ? sizeof(DevTables) : (offsetof(DevTables, fq) + 16);
}


// Per-cell parameters held in registers by every lane (uniform across the warp).
struct CellScalars {
    int n;              // nodes of this cell
    int uniform;        // 1: x_i = i*dx
    double dx;          // uniform: spacing; else: scale of the normalised mesh
    double beta, eps;
    double phi_wall, g_bulk, cstern;
    const double* xi;   // normalised mesh row (non-uniform) or nullptr
    const double* fpar; // this cell's flux-equation parameters or nullptr
    double u_am, u_ac;  // cached uniform-mesh stencil weights 1/dx^2, 1/(2dx)
    double u_sg;        // cached uniform-mesh g-row scale eps/(dx*F)
};

// Per-cell, per-species parameters in shared memory (one copy per warp).
struct CellSpecies {
    double D[MAXS], q[MAXS], bq[MAXS], cb[MAXS], J[MAXS], qe[MAXS];   // qe = q/eps
    double Jfix[MAXS];     // the fixed part of the inward wall flux (J = Jfix + flux equations, refreshed per residual)
};

struct NodeCoef {
    double am, ap, ac;   // second-derivative weights of c_{i-1}, c_{i+1}; central first-derivative weight
    double hi, him;      // h_i = x_{i+1}-x_i ; h_{i-1}
};

// graded mesh: three fp64 divisions per node.  Kept out of line: the division sequences would
// otherwise be replicated at every call site (instruction-cache footprint).
static __device__ __noinline__ NodeCoef graded_coef(const double* xi, double dx, int i) {
    NodeCoef k;
    const double xm = dx * xi[i - 1], x0 = dx * xi[i], xp = dx * xi[i + 1];
    const double hm = x0 - xm, hp = xp - x0;
    k.am = 2.0 / (hm * (hm + hp));
    k.ap = 2.0 / (hp * (hm + hp));
    k.ac = 1.0 / (hm + hp);
    k.hi = hp;
    k.him = hm;
    return k;
}

__device__ __forceinline__ NodeCoef interior_coef(const CellScalars& cs, int i) {
    if (cs.uniform) {
        NodeCoef k;
        k.am = cs.u_am; k.ap = cs.u_am; k.ac = cs.u_ac; k.hi = cs.dx; k.him = cs.dx;
        return k;
    }
    return graded_coef(cs.xi, cs.dx, i);
}

struct WallCoef { double w0, ih0, ext; };   // 1/(h0+h1), 1/h0, h0/h1

static __device__ __noinline__ WallCoef graded_wall_coef(const double* xi, double dx) {
    WallCoef w;
    const double h0 = dx * (xi[1] - xi[0]), h1 = dx * (xi[2] - xi[1]);
    w.w0 = 1.0 / (h0 + h1);
    w.ih0 = 1.0 / h0;
    w.ext = h0 / h1;
    return w;
}

__device__ __forceinline__ WallCoef wall_coef(const CellScalars& cs) {
    if (cs.uniform) {
        WallCoef w;
        w.w0 = cs.u_ac;
        w.ih0 = 2.0 * cs.u_ac;
        w.ext = 1.0;
        return w;
    }
    return graded_wall_coef(cs.xi, cs.dx);
}

// scale of the algebraic g-row of an interior node: eps/(h_i*F) turns the charge entries
// q_j*h/eps (~1e8) into the integer charges z_j, so that partial pivoting rarely swaps
__device__ __forceinline__ double grow_scale(const CellScalars& cs, double hi) {
    return cs.uniform ? cs.u_sg : cs.eps / (hi * UNIT_F);
}

// net rate of reaction r at a node whose concentrations are c[0..S)
__device__ __forceinline__ double net_rate(const DevTables& tb, int r, const double* c) {
    double f = tb.kf[r];
#pragma unroll 1
    for (int e = 0; e < tb.ned[r]; ++e) f *= c[tb.ed[r][e]];
    double b = tb.kr[r];
#pragma unroll 1
    for (int e = 0; e < tb.npr[r]; ++e) b *= c[tb.pr[r][e]];
    return f - b;
}

// cell set-up shared by all kernels
__device__ __forceinline__ void load_cell(const DevTables& tb, const double* par, const int* nx,
                                          const int* mesh_id, const double* mesh_xi, long long cell,
                                          int lane, CellScalars& cs, CellSpecies* sp) {
    const int S = tb.S_full;                 // layout of the caller's parameter record
    const int NPAR = 3 * S + 8;
    const double* p = par + (size_t)cell * NPAR;
    cs.n = nx[cell];
    const int mid = mesh_id ? mesh_id[cell] : -1;
    cs.uniform = mid < 0;
    cs.xi = mid < 0 ? nullptr : mesh_xi + (size_t)mid * tb.nx_max;
    cs.fpar = (tb.fq.n_eq > 0 && tb.fpar) ? tb.fpar + (size_t)cell * tb.fq.n_par : nullptr;
    cs.beta = p[3 * S + 0];
    cs.eps = p[3 * S + 1];
    cs.phi_wall = p[3 * S + 2];
    cs.g_bulk = p[3 * S + 3];
    cs.cstern = p[3 * S + 4];
    cs.dx = p[3 * S + 5];
    cs.u_am = 1.0 / (cs.dx * cs.dx);
    cs.u_ac = 1.0 / (2.0 * cs.dx);
    cs.u_sg = cs.eps / (cs.dx * UNIT_F);
    if (lane < tb.S) {
        const int k = tb.cmap[lane];         // caller's index of kernel species `lane`
        sp->cb[lane] = p[k];
        sp->J[lane] = p[S + k];
        sp->Jfix[lane] = p[S + k];
        sp->D[lane] = p[2 * S + k];
        const double q = tb.z[lane] * UNIT_F;
        sp->q[lane] = q;
        sp->bq[lane] = tb.use_migration ? cs.beta * q : 0.0;
        sp->qe[lane] = q / cs.eps;
    }
    __syncwarp();
}

// max over the warp of NON-NEGATIVE values (norms): IEEE doubles with a clear sign bit order like
// their bit patterns, so two integer warp reductions (REDUX) replace five shuffle+max rounds; +Inf
// and NaN patterns sort above every finite value, i.e. a poisoned norm stays poisoned.
__device__ __forceinline__ double warp_max(double v) {
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_max_sync(FULL, hi);
    const unsigned ml = __reduce_max_sync(FULL, hi == mh ? lo : 0u);
    return __hiloint2double((int)mh, (int)ml);
}

}  // namespace catint
