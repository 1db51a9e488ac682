// ThreadSanitizer harness for the host LM solver helper pool (host_solve.cpp): two "contexts" arm / solve / disarm concurrently;
// every step must equal the serial result bit for bit and ThreadSanitizer must stay silent.  Built by tests/test_host_api.py.
#include <cstdio>
#include <cstring>
#include <random>
#include <thread>
#include <vector>
bool dmsa_host_lu_inverse(const std::vector<double>& A, int n, std::vector<double>& inv);
bool dmsa_host_lm_step(const double* hg, int n, double lambda, double alpha, double* step);
void dmsa_host_solver_arm();
void dmsa_host_solver_disarm();
int main() {
    const int n = 114;
    std::mt19937_64 rng(1);
    std::normal_distribution<double> nd;
    std::vector<double> J((size_t)3 * n * n), hg((size_t)n * n + n + 1, 0.0);
    for (auto& v : J) v = nd(rng);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0;
            for (int r = 0; r < 3 * n; ++r) s += J[(size_t)r * n + i] * J[(size_t)r * n + j];
            hg[(size_t)i * n + j] = s;
        }
    for (int i = 0; i < n; ++i) hg[(size_t)n * n + i] = nd(rng);
    std::vector<double> ref(n), out(n);
    dmsa_host_lm_step(hg.data(), n, 1e-5, 0.2, ref.data());  // unarmed: serial
    int bad = 0;
    auto worker = [&](int reps) {
        std::vector<double> o(n);
        for (int k = 0; k < reps; ++k) {
            dmsa_host_solver_arm();
            dmsa_host_lm_step(hg.data(), n, 1e-5, 0.2, o.data());
            dmsa_host_solver_disarm();
            if (memcmp(o.data(), ref.data(), n * sizeof(double)) != 0) __atomic_fetch_add(&bad, 1, __ATOMIC_RELAXED);
        }
    };
    std::thread t1(worker, 200), t2(worker, 200);  // two "contexts" solving concurrently: one gets the helpers, the other runs serially
    t1.join();
    t2.join();
    printf("mismatches %d\n", bad);
    return bad != 0;
}
