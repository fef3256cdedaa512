// pnp_rhs.cuh -- K1: right-hand side of the reference ODE (streaming stencil).
#pragma once
#include "pnp_kernels.cuh"

namespace catint {

// ===========================================================================
// K1: dc/dt of the reference ODE.  One warp per cell, lanes stride the nodes.
// Pass 1: charge density and its suffix sum -> g (warp scan), kept in shared
// memory; pass 2: stencil + reactions.  c is [B][nx_max][S].
// ===========================================================================
struct RhsParams {
    DevTables tb;
    const double* par; const int* nx; const int* mesh_id; const double* mesh_xi;
    const double* c; long long n_cells;
    double* dcdt; double* g_out; double* phi_out;
};

__global__ void __launch_bounds__(128) pnp_rhs_kernel(RhsParams P) {
    constexpr int WARPS = 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long cell = (long long)blockIdx.x * WARPS + warp;
    DevTables* tb = reinterpret_cast<DevTables*>(smem_raw);
    {
        const int words = (int)(sizeof(DevTables) / 4);
        const int* src = reinterpret_cast<const int*>(&P.tb);
        int* dst = reinterpret_cast<int*>(tb);
        for (int w = threadIdx.x; w < words; w += blockDim.x) dst[w] = src[w];
    }
    __syncthreads();
    const int nxm = P.tb.nx_max, S = P.tb.S;
    size_t off = (sizeof(DevTables) + 15) & ~size_t(15);
    const size_t per_warp = ((sizeof(CellSpecies) + 15) & ~size_t(15)) + (size_t)nxm * sizeof(double);
    unsigned char* mine = smem_raw + off + (size_t)warp * per_warp;
    if (cell >= P.n_cells) return;
    CellSpecies* sp = reinterpret_cast<CellSpecies*>(mine);
    double* gs = reinterpret_cast<double*>(mine + ((sizeof(CellSpecies) + 15) & ~size_t(15)));
    CellScalars cs;
    load_cell(*tb, P.par, P.nx, P.mesh_id, P.mesh_xi, cell, lane, cs, sp);
    const int n = cs.n;
    const double* c = P.c + (size_t)cell * nxm * S;
    double* out = P.dcdt + (size_t)cell * nxm * S;

    // ---- pass 1: g_i = g_bulk - sum_{j=i}^{n-2} lapl_j*h_j  (i=1..n-2), lapl = -sum q c/eps
    if (tb->use_migration) {
        double carry = 0.0;           // sum over nodes already processed (towards the bulk)
        for (int base = n - 2; base >= 1; base -= 32) {
            const int i = base - lane;
            double term = 0.0;
            if (i >= 1) {
                double lapl = 0.0;
                for (int s = 0; s < S; ++s) lapl -= sp->q[s] * c[(size_t)i * S + s] / cs.eps;
                const double hi = cs.uniform ? cs.dx : cs.dx * (cs.xi[i + 1] - cs.xi[i]);
                term = lapl * hi;
            }
            // inclusive scan over lanes (lane 0 = node closest to the bulk)
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double v = __shfl_up_sync(FULL, term, o);
                if (lane >= o) term += v;
            }
            if (i >= 1) gs[i] = cs.g_bulk - (carry + term);
            carry += __shfl_sync(FULL, term, 31);
        }
        if (lane == 0) {
            gs[n - 1] = cs.g_bulk;
        }
        __syncwarp();
        if (lane == 0) {
            const WallCoef w = wall_coef(cs);
            gs[0] = gs[1] + (gs[1] - gs[2]) * w.ext;
        }
    } else {
        for (int i = lane; i < n; i += 32) gs[i] = 0.0;
    }
    __syncwarp();

    // ---- pass 2: stencil
    for (int i = lane; i < n; i += 32) {
        const double* c0 = c + (size_t)i * S;
        if (i == n - 1) {
            for (int s = 0; s < S; ++s) out[(size_t)i * S + s] = 0.0;        // frozen bulk node (:886)
        } else if (i == 0) {
            const WallCoef w = wall_coef(cs);
            const double* c1 = c + S; const double* c2 = c + 2 * S;
            for (int s = 0; s < S; ++s) {
                const double bq = tb->use_migration ? sp->bq[s] : 0.0;
                out[s] = (sp->D[s] * ((c2[s] - c0[s]) * w.w0 + bq * c1[s] * gs[1]) + sp->J[s]) * w.ih0;
            }
        } else {
            const NodeCoef k = interior_coef(cs, i);
            const double* cm = c0 - S; const double* cp = c0 + S;
            const double gm = gs[i - 1], gp = gs[i + 1];
            double net[MAXR];
            for (int r = 0; r < tb->R; ++r) net[r] = net_rate(*tb, r, c0);
            for (int s = 0; s < S; ++s) {
                const double bq = tb->use_migration ? sp->bq[s] : 0.0;
                double R = 0.0;
                for (int r = 0; r < tb->R; ++r) R += tb->nu[r][s] * net[r];
                double d2, dcg;
                if (cs.uniform) {
                    d2 = (cp[s] - 2.0 * c0[s] + cm[s]) / (cs.dx * cs.dx);               // :890
                    dcg = (cp[s] * gp - cm[s] * gm) / (2.0 * cs.dx);                     // :892
                } else {
                    d2 = k.am * cm[s] - (k.am + k.ap) * c0[s] + k.ap * cp[s];
                    dcg = (cp[s] * gp - cm[s] * gm) * k.ac;
                }
                out[(size_t)i * S + s] = sp->D[s] * (d2 + bq * dcg) + R;                  // :920-927
            }
        }
    }
    if (P.g_out) for (int i = lane; i < n; i += 32) P.g_out[(size_t)cell * nxm + i] = gs[i];
    if (P.phi_out && lane == 0) {
        double* po = P.phi_out + (size_t)cell * nxm;
        double v = cs.phi_wall, vm1 = v, vm2 = v;
        po[0] = v;
        for (int i = 1; i <= n - 2; ++i) {
            const double him = cs.uniform ? cs.dx : cs.dx * (cs.xi[i] - cs.xi[i - 1]);
            v = v + gs[i] * him;
            po[i] = v; vm2 = vm1; vm1 = v;
        }
        if (n >= 3) {
            const double ratio = cs.uniform ? 1.0 : (cs.xi[n - 1] - cs.xi[n - 2]) / (cs.xi[n - 2] - cs.xi[n - 3]);
            po[n - 1] = vm1 + (vm1 - vm2) * ratio;
        }
    }
}

}  // namespace catint
