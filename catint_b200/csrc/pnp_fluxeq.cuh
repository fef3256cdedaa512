// pnp_fluxeq.cuh -- wall kinetics given as expressions ("flux equations") of the surface concentrations and the
// surface potential: device-side evaluation of the postfix programs of include/catint_pnp.h (CatintPnpFluxEq),
// value and derivative by forward-mode differentiation.
//
// Reference: species[sp]['flux-equation'] (/root/reference/docs/source/topics/flux_definition.rst:100-156), handed
// to COMSOL as 'RF*flux_factor*(expr)' with [[name]] -> surface concentration (catint/comsol_model.py:986-1000);
// the reference's finite-difference solvers cannot evaluate them, this backend does:
//     J_k(y_0) = J_k^fixed + sum_e coef[k][e] * E_e(c(0), phi(0))
// The derivative dJ_k/dy_0 enters the wall block of the Newton matrix (factor_nodes, wall node).
#pragma once
#include <math.h>
#include "../../include/catint_pnp.h"

namespace catint {

// E_e and dE_e/d(component `comp`): comp < S -> d/dc_comp(0), comp == S -> d/dphi(0), other -> value only.
// c0: wall concentrations; par: this cell's parameter vector.  The stacks live in local memory (dynamic index).
static __device__ __noinline__ double fluxeq_eval(const CatintPnpFluxEq* fq, int e, const double* par, const double* c0,
                                           double phi0, int comp, int S, double* grad_out) {
    double val[CATINT_PNP_MAX_FLUX_STACK], grd[CATINT_PNP_MAX_FLUX_STACK];
    int sp = 0;
    const int nc = fq->n_code[e];
    const int* code = fq->code[e];
    for (int k = 0; k < nc; ++k) {
        const int w = code[k];
        const int op = w & 0xff, arg = w >> 8;
        if (op <= 3) {
            double v, g = 0.0;
            if (op == 0) v = fq->consts[e][arg];
            else if (op == 1) v = par[arg];
            else if (op == 2) { v = c0[arg]; g = comp == arg ? 1.0 : 0.0; }
            else { v = phi0; g = comp == S ? 1.0 : 0.0; }
            if (sp < CATINT_PNP_MAX_FLUX_STACK) { val[sp] = v; grd[sp] = g; }
            ++sp;
        } else if (op <= 8) {
            if (sp < 2) continue;                    // malformed programs are rejected on the host; stay in bounds
            --sp;
            const double a = val[sp - 1], ga = grd[sp - 1], b = val[sp], gb = grd[sp];
            double v, g;
            if (op == 4) { v = a + b; g = ga + gb; }
            else if (op == 5) { v = a - b; g = ga - gb; }
            else if (op == 6) { v = a * b; g = ga * b + a * gb; }
            else if (op == 7) { v = a / b; g = (ga - v * gb) / b; }
            else {
                v = pow(a, b);
                g = 0.0;
                if (ga != 0.0) g += b * pow(a, b - 1.0) * ga;
                if (gb != 0.0) g += v * log(a) * gb;
            }
            val[sp - 1] = v; grd[sp - 1] = g;
        } else {
            if (sp < 1) continue;
            const double a = val[sp - 1], ga = grd[sp - 1];
            double v, g;
            if (op == 9) { v = -a; g = -ga; }
            else if (op == 10) { v = exp(a); g = v * ga; }
            else if (op == 11) { v = log(a); g = ga / a; }
            else if (op == 12) { v = sqrt(a); g = ga / (2.0 * v); }
            else if (op == 13) { v = log10(a); g = ga / (a * 2.302585092994046); }
            else { v = tanh(a); g = (1.0 - v * v) * ga; }
            val[sp - 1] = v; grd[sp - 1] = g;
        }
    }
    if (grad_out) *grad_out = grd[0];
    return val[0];
}

}  // namespace catint
