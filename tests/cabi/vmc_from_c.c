/* The C ABI used from plain C (no Python, no torch): PsiRBM + Heisenberg ring + MonteCarlo -> TDVP::eval_F -> CG solve.
 * Built and run by tests/test_cabi_from_c.py:  gcc vmc_from_c.c -I include -L annongpu_b200 -langpu -lm
 * Prints "E <re> <im> acc <rate> cg <iterations> <rel_residual>"; exit code 0 on success. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "angpu.h"

#define OK(call) do { if((call) != 0) { fprintf(stderr, "%s failed: %s\n", #call, angpu_last_error()); return 1; } } while(0)

int main(void) {
    enum { N = 12, M = 24, CHAINS = 2048 };
    OK(angpu_init(0));
    /* W_ij = 0.05 (sin(0.37 k + 0.1) + i cos(0.11 k)), k = i M + j */
    double* W = (double*)malloc(sizeof(double) * 2 * N * M);
    for(int k = 0; k < N * M; k++) { W[2 * k] = 0.05 * sin(0.37 * k + 0.1); W[2 * k + 1] = 0.05 * cos(0.11 * k); }
    const double fw[2] = {2.0, 0.0}, lp[2] = {0.0, 0.0};
    angpu_psi_t psi; OK(angpu_rbm_create(N, M, W, fw, lp, &psi));
    /* Heisenberg ring: X_i X_j + Y_i Y_j + Z_i Z_j; masks: X = (a 1, b 0), Y = (0, 1), Z = (1, 1) */
    double coef[2 * 3 * N]; uint64_t a[3 * N], b[3 * N];
    for(int i = 0; i < N; i++) {
        const uint64_t m = (1ull << i) | (1ull << ((i + 1) % N));
        for(int t = 0; t < 3; t++) { coef[2 * (3 * i + t)] = 1.0; coef[2 * (3 * i + t) + 1] = 0.0; }
        a[3 * i] = m; b[3 * i] = 0;  a[3 * i + 1] = 0; b[3 * i + 1] = m;  a[3 * i + 2] = m; b[3 * i + 2] = m;
    }
    angpu_operator_t H; OK(angpu_operator_create(3 * N, coef, a, b, 1, &H));
    angpu_ensemble_t mc; OK(angpu_mc_create(CHAINS, 1, 10, CHAINS, 42ull, &mc));
    unsigned P = 0; OK(angpu_psi_num_params(psi, &P));
    angpu_tdvp_t tdvp; OK(angpu_tdvp_create(P, &tdvp));
    OK(angpu_tdvp_eval_F(tdvp, H, psi, mc));
    double s[5]; OK(angpu_tdvp_get_scalars(tdvp, s));
    unsigned long long ar[2]; OK(angpu_mc_acceptance(mc, ar));
    double* x = (double*)malloc(sizeof(double) * 2 * P);
    const double phase[2] = {1.0, 0.0};
    double rr = 0.0; unsigned it = 0;
    OK(angpu_tdvp_solve_cg(tdvp, 1e-8, 500, 0.0, 1e-3, phase, x, &it, &rr));
    printf("E %.12f %.12f acc %.4f cg %u %.3e\n", s[0], s[1], (double)ar[0] / (double)(ar[0] + ar[1]), it, rr);
    const int good = isfinite(s[0]) && fabs(s[1]) < 1.0 && s[2] >= 0.0 && it > 0 && rr <= 1e-8 && ar[0] > 0;
    free(x); free(W);
    OK(angpu_tdvp_destroy(tdvp)); OK(angpu_ensemble_destroy(mc)); OK(angpu_operator_destroy(H)); OK(angpu_psi_destroy(psi));
    return good ? 0 : 2;
}
