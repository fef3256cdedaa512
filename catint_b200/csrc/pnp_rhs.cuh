// pnp_rhs.cuh -- K1: right-hand side of the reference ODE, dc/dt = f(c)
// (/root/reference/catint/calculator_old.py:827-935, Poisson :680-819, rates :159-208).
//
// Pure streaming stencil: 16*S*n algorithmic bytes per cell (read c, write dc/dt), ~1.4 flop/B,
// HBM bound.  One warp per cell:
//   1. the cell's interleaved state [node][species] is read with coalesced 16-byte loads and
//      transposed into shared memory [species][node] (conflict-free lane-per-node access);
//   2. charge density per node, warp-scan suffix sum from the bulk -> g (field gradient);
//   3. stencil + mass-action rates, lane per node; each lane writes its node record (S doubles,
//      16-byte stores; the partial sectors merge in the write-back L2).  No output staging in
//      shared memory: occupancy matters more than store coalescing for this latency-bound body.
#pragma once
#include "pnp_kernels.cuh"

namespace catint {

struct RhsParams {
    DevTables tb;
    const double* par; const int* nx; const int* mesh_id; const double* mesh_xi;
    const double* c; long long n_cells;
    double* dcdt; double* g_out; double* phi_out;
};

constexpr int RHS_WARPS = 4;

// shared-memory doubles per warp: state [S][NP] + g [NP]
__host__ __device__ inline size_t rhs_smem_doubles(int S, int nxm) {
    const size_t NP = (size_t)(nxm | 1) + 2;      // odd pitch
    return ((size_t)S + 1) * NP;
}

template <int S>
__global__ void __launch_bounds__(RHS_WARPS * 32) pnp_rhs_kernel(RhsParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    DevTables* tb = reinterpret_cast<DevTables*>(smem_raw);
    {
        const int words = (int)(sizeof(DevTables) / 4);
        const int* src = reinterpret_cast<const int*>(&P.tb);
        int* dst = reinterpret_cast<int*>(tb);
        for (int w = threadIdx.x; w < words; w += blockDim.x) dst[w] = src[w];
    }
    __syncthreads();
    const int nxm = P.tb.nx_max;
    const int NP = (nxm | 1) + 2;
    const size_t off = (sizeof(DevTables) + 15) & ~size_t(15);
    const size_t per_warp = ((sizeof(CellSpecies) + 15) & ~size_t(15)) + rhs_smem_doubles(S, nxm) * sizeof(double);
    unsigned char* mine = smem_raw + off + (size_t)warp * per_warp;
    CellSpecies* sp = reinterpret_cast<CellSpecies*>(mine);
    double* cs_ = reinterpret_cast<double*>(mine + ((sizeof(CellSpecies) + 15) & ~size_t(15)));   // [S][NP]
    double* gs = cs_ + (size_t)S * NP;                                                           // [NP]

    // persistent over cells: grid = multiple of the SM count
    for (long long cell = (long long)blockIdx.x * RHS_WARPS + warp; cell < P.n_cells;
         cell += (long long)gridDim.x * RHS_WARPS) {
        CellScalars cs;
        load_cell(*tb, P.par, P.nx, P.mesh_id, P.mesh_xi, cell, lane, cs, sp);
        const int n = cs.n;
        if (n < 4 || n > nxm) continue;                 // bad nx[cell]: skip (warp-uniform), outputs untouched
        const double* c = P.c + (size_t)cell * nxm * S;
        double* dst = P.dcdt + (size_t)cell * nxm * S;
        const bool mig = tb->use_migration;

        // ---- 1. coalesced load + transpose
        const int total = n * S;                       // doubles of this cell
        if (((size_t)c & 15) == 0) {
            const double2* c2 = reinterpret_cast<const double2*>(c);
            for (int e2 = lane; e2 < total / 2; e2 += 32) {
                const double2 v = c2[e2];
                const int e = 2 * e2;
                const int i = e / S, s = e - i * S;
                cs_[s * NP + i] = v.x;
                if (s + 1 < S) cs_[(s + 1) * NP + i] = v.y; else cs_[i + 1] = v.y;
            }
            if ((total & 1) && lane == 0) { const int e = total - 1; cs_[(e % S) * NP + e / S] = c[e]; }
        } else {
            for (int e = lane; e < total; e += 32) cs_[(e % S) * NP + e / S] = c[e];
        }
        __syncwarp();

        // ---- 2. g_i = g_bulk - sum_{j=i}^{n-2} lapl_j*h_j (i=1..n-2), lapl = -sum q c/eps
        if (mig) {
            double carry = 0.0;
            for (int base = n - 2; base >= 1; base -= 32) {
                const int i = base - lane;
                double term = 0.0;
                if (i >= 1) {
                    double lapl = 0.0;
#pragma unroll
                    for (int s = 0; s < S; ++s) lapl = fma(-sp->qe[s], cs_[s * NP + i], lapl);   // :767-771
                    const double hi = cs.uniform ? cs.dx : cs.dx * (cs.xi[i + 1] - cs.xi[i]);
                    term = lapl * hi;
                }
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double v = __shfl_up_sync(FULL, term, o);
                    if (lane >= o) term += v;
                }
                if (i >= 1) gs[i] = cs.g_bulk - (carry + term);
                carry += __shfl_sync(FULL, term, 31);
            }
            if (lane == 0) gs[n - 1] = cs.g_bulk;
            __syncwarp();
            if (lane == 0) {
                const WallCoef w = wall_coef(cs);
                gs[0] = gs[1] + (gs[1] - gs[2]) * w.ext;
            }
        } else {
            for (int i = lane; i < n; i += 32) gs[i] = 0.0;
        }
        __syncwarp();

        // ---- 3. stencil, lane per node
        for (int i = lane; i < n; i += 32) {
            double res[S];
            if (i == n - 1) {
#pragma unroll
                for (int s = 0; s < S; ++s) res[s] = 0.0;                                // frozen bulk node (:886)
            } else if (i == 0) {
                const WallCoef w = wall_coef(cs);
                const double g1 = gs[1];
#pragma unroll
                for (int s = 0; s < S; ++s)
                    res[s] = (sp->D[s] * ((cs_[s * NP + 2] - cs_[s * NP]) * w.w0 + sp->bq[s] * cs_[s * NP + 1] * g1)
                              + sp->J[s]) * w.ih0;
            } else {
                const NodeCoef k = interior_coef(cs, i);
                const double gm = gs[i - 1], gp = gs[i + 1];
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    const double cm = cs_[s * NP + i - 1], cp = cs_[s * NP + i + 1], c0 = cs_[s * NP + i];
                    double d2, dcg;
                    if (cs.uniform) {
                        d2 = (cp - 2.0 * c0 + cm) * cs.u_am;                            // :890 (reciprocal hoisted)
                        dcg = (cp * gp - cm * gm) * cs.u_ac;                            // :892
                    } else {
                        d2 = k.am * cm - (k.am + k.ap) * c0 + k.ap * cp;
                        dcg = (cp * gp - cm * gm) * k.ac;
                    }
                    res[s] = sp->D[s] * (d2 + sp->bq[s] * dcg);
                }
                for (int rr = 0; rr < tb->R; ++rr) {
                    const unsigned ew = *reinterpret_cast<const unsigned*>(tb->ed[rr]);
                    const unsigned pw = *reinterpret_cast<const unsigned*>(tb->pr[rr]);
                    const int ne = tb->ned[rr], np = tb->npr[rr];
                    double f = tb->kf[rr], b = tb->kr[rr];
#pragma unroll
                    for (int e = 0; e < MAXRT; ++e) {
                        if (e < ne) f *= cs_[((ew >> (8 * e)) & 0xff) * NP + i];
                        if (e < np) b *= cs_[((pw >> (8 * e)) & 0xff) * NP + i];
                    }
                    const double net = f - b;
                    const double* nur = tb->nu[rr];
#pragma unroll
                    for (int s = 0; s < S; ++s) res[s] = fma(nur[s], net, res[s]);       // :920-927
                }
            }
            double* o = dst + (size_t)i * S;
            if ((((size_t)o) & 15) == 0) {
#pragma unroll
                for (int s = 0; s + 1 < S; s += 2) *reinterpret_cast<double2*>(o + s) = make_double2(res[s], res[s + 1]);
                if (S & 1) o[S - 1] = res[S - 1];
            } else {
#pragma unroll
                for (int s = 0; s < S; ++s) o[s] = res[s];
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
        __syncwarp();
    }
}

template <int S>
int launch_rhs(RhsParams& P, cudaStream_t st) {
    const size_t base = ((sizeof(DevTables) + 15) & ~size_t(15));
    const size_t per_warp = ((sizeof(CellSpecies) + 15) & ~size_t(15)) + rhs_smem_doubles(S, P.tb.nx_max) * sizeof(double);
    const size_t smem = base + RHS_WARPS * per_warp;
    // per-device launch geometry, queried on every call (a few microseconds; no static state to race on,
    // correct when one process drives several GPUs)
    int dev = 0, sms = 148, max_optin = 0, per_sm = 1;
    if (cudaGetDevice(&dev) != cudaSuccess) return CATINT_PNP_ECUDA;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (smem > (size_t)max_optin) return CATINT_PNP_EINVAL;
    cudaFuncSetAttribute(pnp_rhs_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pnp_rhs_kernel<S>, RHS_WARPS * 32, smem);
    if (per_sm < 1) per_sm = 1;
    const long long cached_blocks = (long long)sms * per_sm;
    const long long cap = cached_blocks;                      // one resident wave, persistent loop inside
    const long long want = (P.n_cells + RHS_WARPS - 1) / RHS_WARPS;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    pnp_rhs_kernel<S><<<grid, RHS_WARPS * 32, smem, st>>>(P);
    return cudaGetLastError() == cudaSuccess ? 0 : CATINT_PNP_ECUDA;
}

}  // namespace catint
