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
};

// Per-cell parameters held in registers by every lane (uniform across the warp).
struct CellScalars {
    int n;              // nodes of this cell
    int uniform;        // 1: x_i = i*dx
    double dx;          // uniform: spacing; else: scale of the normalised mesh
    double beta, eps;
    double phi_wall, g_bulk, cstern;
    const double* xi;   // normalised mesh row (non-uniform) or nullptr
    double u_am, u_ac;  // cached uniform-mesh stencil weights 1/dx^2, 1/(2dx)
};

// Per-cell, per-species parameters in shared memory (one copy per warp).
struct CellSpecies {
    double D[MAXS], q[MAXS], bq[MAXS], cb[MAXS], J[MAXS], qe[MAXS];   // qe = q/eps
};

struct NodeCoef {
    double am, ap, ac;   // second-derivative weights of c_{i-1}, c_{i+1}; central first-derivative weight
    double hi, him;      // h_i = x_{i+1}-x_i ; h_{i-1}
};

__device__ __forceinline__ NodeCoef interior_coef(const CellScalars& cs, int i) {
    NodeCoef k;
    if (cs.uniform) {
        k.am = cs.u_am; k.ap = cs.u_am; k.ac = cs.u_ac; k.hi = cs.dx; k.him = cs.dx;
    } else {
        const double xm = cs.dx * cs.xi[i - 1], x0 = cs.dx * cs.xi[i], xp = cs.dx * cs.xi[i + 1];
        const double hm = x0 - xm, hp = xp - x0;
        k.am = 2.0 / (hm * (hm + hp));
        k.ap = 2.0 / (hp * (hm + hp));
        k.ac = 1.0 / (hm + hp);
        k.hi = hp;
        k.him = hm;
    }
    return k;
}

struct WallCoef { double w0, ih0, ext; };   // 1/(h0+h1), 1/h0, h0/h1

__device__ __forceinline__ WallCoef wall_coef(const CellScalars& cs) {
    WallCoef w;
    if (cs.uniform) {
        w.w0 = cs.u_ac;
        w.ih0 = 2.0 * cs.u_ac;
        w.ext = 1.0;
    } else {
        const double h0 = cs.dx * (cs.xi[1] - cs.xi[0]), h1 = cs.dx * (cs.xi[2] - cs.xi[1]);
        w.w0 = 1.0 / (h0 + h1);
        w.ih0 = 1.0 / h0;
        w.ext = h0 / h1;
    }
    return w;
}

// scale of the algebraic g-row of an interior node: eps/(h_i*F) turns the charge entries
// q_j*h/eps (~1e8) into the integer charges z_j, so that partial pivoting rarely swaps
__device__ __forceinline__ double grow_scale(const CellScalars& cs, double hi) {
    return cs.eps / (hi * UNIT_F);
}

// net rate of reaction r at a node whose concentrations are c[0..S)
__device__ __forceinline__ double net_rate(const DevTables& tb, int r, const double* c) {
    double f = tb.kf[r];
    for (int e = 0; e < tb.ned[r]; ++e) f *= c[tb.ed[r][e]];
    double b = tb.kr[r];
    for (int e = 0; e < tb.npr[r]; ++e) b *= c[tb.pr[r][e]];
    return f - b;
}

// scalars of a cell (every lane keeps a copy in registers)
__device__ __forceinline__ void load_cell_scalars(const DevTables& tb, const double* par, const int* nx,
                                                  const int* mesh_id, const double* mesh_xi, long long cell,
                                                  CellScalars& cs) {
    const int S = tb.S;
    const double* p = par + (size_t)cell * (3 * S + 8);
    cs.n = nx[cell];
    const int mid = mesh_id ? mesh_id[cell] : -1;
    cs.uniform = mid < 0;
    cs.xi = mid < 0 ? nullptr : mesh_xi + (size_t)mid * tb.nx_max;
    cs.beta = p[3 * S + 0];
    cs.eps = p[3 * S + 1];
    cs.phi_wall = p[3 * S + 2];
    cs.g_bulk = p[3 * S + 3];
    cs.cstern = p[3 * S + 4];
    cs.dx = p[3 * S + 5];
    cs.u_am = 1.0 / (cs.dx * cs.dx);
    cs.u_ac = 1.0 / (2.0 * cs.dx);
}

// cell set-up shared by all kernels
__device__ __forceinline__ void load_cell(const DevTables& tb, const double* par, const int* nx,
                                          const int* mesh_id, const double* mesh_xi, long long cell,
                                          int lane, CellScalars& cs, CellSpecies* sp) {
    const int S = tb.S;
    const int NPAR = 3 * S + 8;
    const double* p = par + (size_t)cell * NPAR;
    cs.n = nx[cell];
    const int mid = mesh_id ? mesh_id[cell] : -1;
    cs.uniform = mid < 0;
    cs.xi = mid < 0 ? nullptr : mesh_xi + (size_t)mid * tb.nx_max;
    cs.beta = p[3 * S + 0];
    cs.eps = p[3 * S + 1];
    cs.phi_wall = p[3 * S + 2];
    cs.g_bulk = p[3 * S + 3];
    cs.cstern = p[3 * S + 4];
    cs.dx = p[3 * S + 5];
    cs.u_am = 1.0 / (cs.dx * cs.dx);
    cs.u_ac = 1.0 / (2.0 * cs.dx);
    if (lane < S) {
        sp->cb[lane] = p[lane];
        sp->J[lane] = p[S + lane];
        sp->D[lane] = p[2 * S + lane];
        const double q = tb.z[lane] * UNIT_F;
        sp->q[lane] = q;
        sp->bq[lane] = tb.use_migration ? cs.beta * q : 0.0;
        sp->qe[lane] = q / cs.eps;
    }
    __syncwarp();
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

}  // namespace catint
