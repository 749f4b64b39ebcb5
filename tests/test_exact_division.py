"""CPU: the stereo kernel replaces the reference's float64 true-divide (d - min) / (max - min)
(src/stereoimage_generation.py:81) by a reciprocal multiply + two FMAs (csrc/stereo.cu, NdSrc::get).  This compiles a small
C program that checks the identity for EVERY pair 0 <= a <= den <= 65535 (2.1e9 pairs, ~2 s with hardware FMA)."""
import os
import shutil
import subprocess
import tempfile

import pytest

SRC = r'''
#include <math.h>
#include <stdio.h>
int main(void) {
    long long bad = 0, total = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : bad, total)
    for (int b = 1; b <= 65535; ++b) {
        const double db = (double)b, r = 1.0 / db;
        for (int a = 0; a <= b; ++a) {
            const double da = (double)a;
            const double q0 = da * r;
            const double rem = fma(-q0, db, da);
            const double q = fma(rem, r, q0);
            if (q != da / db) ++bad;
            ++total;
        }
    }
    printf("pairs %lld mismatches %lld\n", total, bad);
    return bad != 0;
}
'''


def test_reciprocal_division_is_exact_for_all_u16_pairs():
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    if "fma" not in open("/proc/cpuinfo").read():
        pytest.skip("no hardware FMA on this host (the software fma would take minutes)")
    with tempfile.TemporaryDirectory() as d:
        c, exe = os.path.join(d, "m.c"), os.path.join(d, "m")
        open(c, "w").write(SRC)
        subprocess.run(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-fopenmp", c, "-o", exe, "-lm"], check=True)
        r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout
        assert "pairs 2147516415 mismatches 0" in r.stdout
