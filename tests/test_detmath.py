"""include/mctsb200_detmath.h on the host: det_sincos returns det_sin and det_cos bit for bit (GaussianBoxDev::gen relies on it:
the oracle calls the two separate functions, the kernels the fused one)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r"""
#include <cstdio>
#include <cstdint>
#include <cstring>
#include "mctsb200_detmath.h"
int main() {
    uint64_t st = 88172645463325252ULL, bad = 0;
    const double special[] = {0.0, -0.0, 1e-300, 0.7853981633974483, 1.5707963267948966, 3.141592653589793, 6.283185307179586, 1048576.0, 1048577.0, -1048576.0};
    for (long i = 0; i < 3000000 + 10; ++i) {
        double x;
        if (i < 10) x = special[i];
        else {
            st ^= st << 13; st ^= st >> 7; st ^= st << 17;
            x = ((double)(st >> 11) / 9007199254740992.0 - 0.5) * ((i % 7 == 0) ? 2.2e6 : (i % 3 == 0 ? 100.0 : 13.0));
        }
        double s, c;
        det_sincos(x, &s, &c);
        const double s2 = det_sin(x), c2 = det_cos(x);
        if (memcmp(&s, &s2, 8) || memcmp(&c, &c2, 8)) ++bad;
    }
    printf("%llu\n", (unsigned long long)bad);
    return 0;
}
"""


def test_det_sincos_equals_det_sin_and_det_cos(tmp_path):
    src = tmp_path / "sincos.cpp"
    src.write_text(SRC)
    exe = tmp_path / "sincos"
    cxx = os.environ.get("CXX", "g++")
    subprocess.run([cxx, "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip()
    assert out == "0", f"{out} of 3 000 010 inputs differ"
