/* Two ranks from plain C (no Python, no torch, no MPI): the process forks before touching CUDA, rank 0 creates the
 * communicator id and hands it to rank 1 through a pipe, each rank runs its half of the Markov chains on its own GPU
 * (ensembles inherit the shard from the communicator) and libangpu sums the partial results with ncclAllReduce.
 * Rank 0 then repeats the run unsharded and compares.  Built and run by tests/test_cabi_from_c.py (needs 2 GPUs):
 *   gcc vmc_two_ranks.c -I include -L annongpu_b200 -langpu -lm
 * Prints "E2 <re> <im> E1 <re> <im> dF <max rel deviation of F>"; exit code 0 when they agree to 1e-10. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <sys/wait.h>
#include <unistd.h>

#include "angpu.h"

#define OK(call) do { if((call) != 0) { fprintf(stderr, "rank %d: %s failed: %s\n", rank, #call, angpu_last_error()); return 1; } } while(0)

enum { N = 12, M = 24, CHAINS = 2048 };

static int run(int rank, int fd_read, int fd_write) {
    unsigned char id[ANGPU_COMM_ID_BYTES];
    OK(angpu_init(rank));
    if(rank == 0) { OK(angpu_comm_unique_id(id)); if(write(fd_write, id, sizeof id) != (ssize_t)sizeof id) return 1; }
    else if(read(fd_read, id, sizeof id) != (ssize_t)sizeof id) return 1;
    OK(angpu_comm_init(id, rank, 2));
    double* W = (double*)malloc(sizeof(double) * 2 * N * M);
    for(int k = 0; k < N * M; k++) { W[2 * k] = 0.05 * sin(0.37 * k + 0.1); W[2 * k + 1] = 0.05 * cos(0.11 * k); }
    const double fw[2] = {2.0, 0.0}, lp[2] = {0.0, 0.0};
    angpu_psi_t psi; OK(angpu_rbm_create(N, M, W, fw, lp, &psi));
    double coef[2 * 3 * N]; uint64_t a[3 * N], b[3 * N];
    for(int i = 0; i < N; i++) {
        const uint64_t m = (1ull << i) | (1ull << ((i + 1) % N));
        for(int t = 0; t < 3; t++) { coef[2 * (3 * i + t)] = 1.0; coef[2 * (3 * i + t) + 1] = 0.0; }
        a[3 * i] = m; b[3 * i] = 0;  a[3 * i + 1] = 0; b[3 * i + 1] = m;  a[3 * i + 2] = m; b[3 * i + 2] = m;
    }
    angpu_operator_t H; OK(angpu_operator_create(3 * N, coef, a, b, 1, &H));
    unsigned P = 0; OK(angpu_psi_num_params(psi, &P));
    angpu_ensemble_t mc; OK(angpu_mc_create(CHAINS, 1, 10, CHAINS, 42ull, &mc));     /* owns chains [rank*1024, rank*1024 + 1024) */
    unsigned long long local = 0; OK(angpu_ensemble_local_steps(mc, &local));
    angpu_tdvp_t tdvp; OK(angpu_tdvp_create(P, &tdvp));
    OK(angpu_tdvp_eval_F(tdvp, H, psi, mc));
    double s2[5]; OK(angpu_tdvp_get_scalars(tdvp, s2));
    double* F2 = (double*)malloc(sizeof(double) * 2 * P); OK(angpu_tdvp_get_F(tdvp, F2));
    int status = (local == CHAINS / 2) ? 0 : 2;
    if(rank == 0) {
        angpu_ensemble_t one; OK(angpu_mc_create(CHAINS, 1, 10, CHAINS, 42ull, &one));
        OK(angpu_ensemble_set_shard(one, 0, 1));                                      /* all chains on this rank: never reduced */
        angpu_tdvp_t t1; OK(angpu_tdvp_create(P, &t1));
        OK(angpu_tdvp_eval_F(t1, H, psi, one));
        double s1[5]; OK(angpu_tdvp_get_scalars(t1, s1));
        double* F1 = (double*)malloc(sizeof(double) * 2 * P); OK(angpu_tdvp_get_F(t1, F1));
        double dF = 0.0, nF = 0.0;
        for(unsigned k = 0; k < 2 * P; k++) { dF = fmax(dF, fabs(F1[k] - F2[k])); nF = fmax(nF, fabs(F1[k])); }
        printf("E2 %.14f %.14f E1 %.14f %.14f dF %.3e\n", s2[0], s2[1], s1[0], s1[1], dF / nF);
        if(!(fabs(s2[0] - s1[0]) <= 1e-10 * fabs(s1[0]) && dF <= 1e-10 * nF)) status = 2;
        free(F1); OK(angpu_tdvp_destroy(t1)); OK(angpu_ensemble_destroy(one));
    }
    free(F2); free(W);
    OK(angpu_tdvp_destroy(tdvp)); OK(angpu_ensemble_destroy(mc)); OK(angpu_operator_destroy(H)); OK(angpu_psi_destroy(psi));
    OK(angpu_comm_destroy());
    return status;
}

int main(void) {
    int fds[2];
    if(pipe(fds) != 0) return 1;
    const pid_t child = fork();                       /* before any CUDA call: each process creates its own context */
    if(child < 0) return 1;
    if(child == 0) return run(1, fds[0], -1);
    const int r0 = run(0, -1, fds[1]);
    int st = 0;
    waitpid(child, &st, 0);
    return (r0 == 0 && WIFEXITED(st) && WEXITSTATUS(st) == 0) ? 0 : 3;
}
