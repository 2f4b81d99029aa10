"""CPU unit test of the numeric core of the device path: the exponent-extended f64 arithmetic ("XF", dbgphmm_b200/csrc/common.cuh)
that replaces the reference's log-space Prob (prob.rs:181-253) in every kernel.  The functions are `__host__ __device__`, so the
same source is compiled here for the host and checked against log-space arithmetic done the reference's way."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include "common.cuh"
#define CHECK(cond, code) do { if (!(cond)) { std::fprintf(stderr, "check failed (%d): %s\n", code, #cond); return code; } } while (0)
// Prob + Prob of the reference (prob.rs:181-197)
static double padd(double a, double b) {
    const double x = a >= b ? a : b, y = a >= b ? b : a;
    if (y == -INFINITY) return x;
    if (x == y) return x + std::log(2.0);
    return x + std::log1p(std::exp(y - x));
}
static bool close(double got, double want, double rel) { return (std::isinf(want) && got == want) || std::fabs(got - want) <= rel * std::fmax(1.0, std::fabs(want)); }

int main() {
    // 1. inside the normal range the arithmetic IS f64 arithmetic, whatever the split between mantissa and exponent
    std::mt19937_64 rng(1);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    for (int t = 0; t < 20000; t++) {
        const double a = U(rng) * std::ldexp(1.0, (int)(rng() % 60) - 30), b = U(rng) * std::ldexp(1.0, (int)(rng() % 60) - 30);
        const int ea = (int)(rng() % 200) - 100, eb = (int)(rng() % 200) - 100;
        XF s = xadd(xf(std::ldexp(a, -ea), ea), xf(std::ldexp(b, -eb), eb));     // the same two numbers, differently split
        CHECK(std::ldexp(s.v, s.e) == a + b, 1);
        CHECK(xgt(xf(std::ldexp(a, -ea), ea), xf(std::ldexp(b, -eb), eb)) == (a > b), 2);
        XF n = xnorm(s);
        CHECK((s.v == 0.0 && n.v == 0.0) || (n.v >= 1.0 && n.v < 2.0 && std::ldexp(n.v, n.e) == a + b && xexp(s) == n.e), 3);
    }
    // 2. zero is an identity of +, absorbs *, and exports as -inf (prob.rs:187-189, test p(0) + p(1) == p(1))
    XF z = xf_zero(), one = xf(1.0, 0);
    CHECK(xadd(z, one).v == 1.0 && xadd(one, z).e == 0 && xmul(z, 0.25).v == 0.0 && xlog(z) == -INFINITY && xlog(xadd(z, z)) == -INFINITY, 4);
    CHECK(!xgt(z, z) && xgt(one, z) && !xgt(z, one), 5);
    // 3. no underflow: 20,000 steps of a forward-like recurrence m' = p_match * p_MM * m + p_II * p_random * i (a 20 kbp read, C5)
    //    stay exact in the exponent; log-space does the same sums with logaddexp.  ln P ~ -3.6e4, far below f64's 1e-308.
    {
        const double pm = 0.999 * 0.99799, pi = 0.001 * 0.25, pmi = 0.001, pim = 0.99799;
        XF m = xf(0.1, 0), i = xf_zero();
        double lm = std::log(0.1), li = -INFINITY;
        for (int s = 0; s < 20000; s++) {
            XF m2 = xadd(xmul(m, pm * 0.17), xmul(i, pim * 0.17));      // (0.17: an emission-like factor that drives the value down)
            XF i2 = xadd(xmul(m, pmi * 0.25), xmul(i, pi));
            if ((s & 63) == 0) { m2 = xnorm(m2); i2 = xnorm(i2); }        // kernels renormalise when they pack a row
            const double lm2 = padd(lm + std::log(pm * 0.17), li + std::log(pim * 0.17));
            const double li2 = padd(lm + std::log(pmi * 0.25), li + std::log(pi));
            m = m2; i = i2; lm = lm2; li = li2;
        }
        CHECK(lm < -30000.0 && close(xlog(m), lm, 1e-9) && close(xlog(i), li, 1e-9), 6);
    }
    // 4. terms more than 2^1022 apart: the small one is dropped, exactly where log-space f64 drops it too
    {
        XF big = xf(1.5, 0), tiny = xf(1.25, -1100);
        CHECK(xadd(big, tiny).v == 1.5 && xadd(tiny, big).e == 0, 7);
        CHECK(padd(std::log(1.5), std::log(1.25) - 1100 * std::log(2.0)) == std::log(1.5), 8);
        XF near = xadd(xf(1.0, 0), xf(1.0, -52));                        // ... and nothing is dropped that f64 keeps
        CHECK(near.v == 1.0 + std::ldexp(1.0, -52), 9);
    }
    // 5. a DP cell shares one exponent (28 B/cell): states keep their value, a state > 2^1022 below the largest is flushed
    {
        Cell c = cell_pack(xf(1.5, -700), xf(1.25, -710), xf(1.75, -2000));
        CHECK(c.e == -700 && c.m == 1.5 && c.i == std::ldexp(1.25, -10) && c.d == 0.0, 10);
        Cell e0 = cell_pack(xf_zero(), xf_zero(), xf_zero());
        CHECK(e0.m == 0.0 && e0.i == 0.0 && e0.d == 0.0, 11);
        Cell d = cell_pack(xf(std::ldexp(1.0, -1060), 0), xf_zero(), xf_zero());   // a denormal mantissa still reports its exponent
        CHECK(d.m > 0.0 && std::fabs(xlog(xf(d.m, d.e)) - (-1060 * std::log(2.0))) < 1e-9, 12);
    }
    // 6. the ABI export: ln v + e ln 2
    CHECK(close(xlog(xf(1.5, -123456)), std::log(1.5) - 123456 * std::log(2.0), 1e-15), 13);
    CHECK(sizeof(XF) == 16 && sparse_row_bytes(3, 1) == 104, 14);
    std::puts("xf ok");
    return 0;
}
'''


def test_xf_arithmetic_on_the_host(tmp_path):
    src = tmp_path / "xf.cpp"
    src.write_text(SRC)
    exe = tmp_path / "xf"
    env = {k: v for k, v in os.environ.items() if k not in ("CXX", "CC")}
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-ffp-contract=off", "-Wall", "-I", os.path.join(ROOT, "dbgphmm_b200", "csrc"),
                           "-I", "/usr/local/cuda/include", str(src), "-o", str(exe)], env=env)
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and "xf ok" in p.stdout, (p.returncode, p.stdout, p.stderr)
