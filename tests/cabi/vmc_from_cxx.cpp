// The C++ face (include/angpu.hpp: RAII classes with the reference's names) used from a plain C++ program:
// PsiRBM + Heisenberg ring + MonteCarloSpins -> TDVP::eval_F_vector -> solve_cg -> apply_update, and value semantics.
// Built and run by tests/test_cabi_from_c.py.  Prints "E <re> <im> cg <iterations> <rel_residual> copy <max |diff|>".
#include <cmath>
#include <cstdio>

#include "angpu.hpp"

using namespace ann_on_gpu_b200;

int main() {
    try {
        const unsigned N = 12, M = 24, CHAINS = 2048;
        setDevice(0);
        std::vector<complex_t> W(N * M);
        for(unsigned k = 0; k < N * M; k++) W[k] = 0.05 * complex_t(std::sin(0.37 * k + 0.1), std::cos(0.11 * k));
        PsiRBM psi(N, M, W, 2.0);
        std::vector<complex_t> coef; std::vector<uint64_t> a, b;
        for(unsigned i = 0; i < N; i++) {
            const uint64_t m = (1ull << i) | (1ull << ((i + 1) % N));
            coef.insert(coef.end(), 3, complex_t(1.0));
            a.push_back(m); b.push_back(0);  a.push_back(0); b.push_back(m);  a.push_back(m); b.push_back(m);
        }
        Operator H(coef, a, b);
        MonteCarloSpins mc(CHAINS, 1, 10, CHAINS, 42ull);
        TDVP tdvp(psi.num_params());
        tdvp.eval_F_vector(H, psi, mc);
        const complex_t E = tdvp.E_local();
        auto cg = tdvp.solve_cg(1e-8, 500);
        PsiRBM before = psi;                                   // deep copy (Array<T> semantics)
        tdvp.apply_update(psi, complex_t(-0.01, 0.0));
        const auto p0 = before.get_params(), p1 = psi.get_params();
        double dev = 0.0;
        for(size_t k = 0; k < p0.size(); k++) dev = std::fmax(dev, std::abs(p1[k] - (p0[k] - 0.01 * cg.x[k])));
        std::printf("E %.12f %.12f cg %u %.3e copy %.3e\n", E.real(), E.imag(), cg.iterations, cg.rel_residual, dev);
        bool threw = false;
        try { psi.set_params(std::vector<complex_t>(3)); } catch(const std::runtime_error&) { threw = true; }
        return (std::isfinite(E.real()) && cg.rel_residual <= 1e-8 && dev <= 1e-14 && threw) ? 0 : 2;
    } catch(const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
