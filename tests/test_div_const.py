"""CPU: the FMA-corrected reciprocal used by the EV kernels for divisions by constants
(csrc/ev.cu::div_c) returns exactly the IEEE quotient the reference's Python arithmetic produces.
The same three-operation sequence is compiled here with gcc (fma() is correctly rounded)."""
import os
import subprocess

SRC = r"""
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
static inline double divc(double a, double b, double inv) {
    double q = a * inv; double r = fma(-q, b, a); return fma(r, inv, q);
}
int main(void) {
    const double apers = ((1.0 / 60.0) * (208.0 / 1000.0)) * 5.0;
    const double consts[] = {1000.0, 208.0, 100.0, 1.0 - 0.8, 5.0 / 60.0, apers};
    long bad = 0;
    srand48(7);
    for (int c = 0; c < 6; ++c) {
        const double b = consts[c], inv = 1.0 / b;
        for (long i = 0; i < 3000000; ++i) {
            double a;
            switch (i & 3) {
                case 0: a = drand48() * 100.0; break;
                case 1: a = drand48() * 7000.0; break;
                case 2: a = ldexp(drand48(), (int)(drand48() * 40) - 30); break;
                default: a = (double)(float)(drand48() * 100.0);
            }
            if (divc(a, b, inv) != a / b) ++bad;
        }
    }
    printf("%ld\n", bad);
    return 0;
}
"""


def test_markstein_division_is_exact(tmp_path):
    c = tmp_path / "divc.c"
    c.write_text(SRC)
    exe = tmp_path / "divc"
    subprocess.check_call(["gcc", "-O2", "-mfma", "-ffp-contract=off", str(c), "-o", str(exe), "-lm"])
    assert int(subprocess.check_output([str(exe)]).strip()) == 0
