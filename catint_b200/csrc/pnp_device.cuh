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
// Unknowns per node  y_i = (c_0..c_{S-1}, g [, phi]);  NB = block size.
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

// Model tables, passed by value as a kernel parameter and copied to shared memory.
struct DevTables {
    int S, R, T, stern, use_migration, nx_max;
    double z[MAXS];
    int8_t ned[MAXR], npr[MAXR];
    int8_t ed[MAXR][MAXRT], pr[MAXR][MAXRT];
    double kf[MAXR], kr[MAXR];
    double nu[MAXR][MAXS];          // nu[r][k]
    // d net_r / d c_j = sum over terms t with tj[t]==j of tcoef[t]*c[ti1[t]]*c[ti2[t]]*c[ti3[t]]
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
};

// Per-cell, per-species parameters in shared memory (one copy per warp).
struct CellSpecies {
    double D[MAXS], q[MAXS], bq[MAXS], cb[MAXS], J[MAXS];
};

struct NodeCoef {
    double am, ap, ac;   // second-derivative weights of c_{i-1}, c_{i+1}; central first-derivative weight
    double hi, him;      // h_i = x_{i+1}-x_i ; h_{i-1}
};

__device__ __forceinline__ NodeCoef interior_coef(const CellScalars& cs, int i) {
    NodeCoef k;
    if (cs.uniform) {
        const double dx = cs.dx;
        k.am = 1.0 / (dx * dx);
        k.ap = k.am;
        k.ac = 1.0 / (2.0 * dx);
        k.hi = dx;
        k.him = dx;
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
        w.w0 = 1.0 / (2.0 * cs.dx);
        w.ih0 = 1.0 / cs.dx;
        w.ext = 1.0;
    } else {
        const double h0 = cs.dx * (cs.xi[1] - cs.xi[0]), h1 = cs.dx * (cs.xi[2] - cs.xi[1]);
        w.w0 = 1.0 / (h0 + h1);
        w.ih0 = 1.0 / h0;
        w.ext = h0 / h1;
    }
    return w;
}

// net rate of reaction r at a node whose concentrations are c[0..S)
__device__ __forceinline__ double net_rate(const DevTables& tb, int r, const double* c) {
    double f = tb.kf[r];
    for (int e = 0; e < tb.ned[r]; ++e) f *= c[tb.ed[r][e]];
    double b = tb.kr[r];
    for (int e = 0; e < tb.npr[r]; ++e) b *= c[tb.pr[r][e]];
    return f - b;
}

// homogeneous source R_k = sum_r nu[r][k]*net_r
__device__ __forceinline__ double reaction_source(const DevTables& tb, int k, const double* c) {
    double s = 0.0;
    for (int r = 0; r < tb.R; ++r) s += tb.nu[r][k] * net_rate(tb, r, c);
    return s;
}

}  // namespace catint
