"""The division shortcut of the leaf tests (rtnw_device.cuh, div_by_recip), restated in C and checked on the host: from
y = RN(1/d), two residual corrections must give the bits of the IEEE quotient x/d.  The device self-test
(tests/test_gpu_parity.py::test_division_by_reciprocal) runs the kernel's own code; this one pins the ALGORITHM where no GPU
is needed, with glibc's correctly rounded fmaf as the FMA."""
import shutil
import subprocess
import textwrap

import pytest

SRC = textwrap.dedent(r"""
    #include <math.h>
    #include <stdint.h>
    #include <stdio.h>
    #include <stdlib.h>
    #include <string.h>
    static float asf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
    static uint32_t asu(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
    static uint64_t s = 88172645463325252ull;
    static uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
    static float div_by_recip(float x, float d, float y) {
        float q = x * y;
        q = fmaf(fmaf(-q, d, x), y, q);
        return fmaf(fmaf(-q, d, x), y, q);
    }
    int main(int argc, char** argv) {
        long n = atol(argv[1]), bad = 0;
        for (long i = 0; i < n; i++) {
            uint64_t r = rnd();
            uint32_t mx = (uint32_t)(r & 0x7fffff), md = (uint32_t)((r >> 23) & 0x7fffff);
            int mode = (r >> 46) & 7;  /* mantissas near all-ones / near 1.0 are the hard cases of reciprocal-based division */
            if (mode == 0) md = 0x7fffff - (md & 0xff); else if (mode == 1) md &= 0xff;
            else if (mode == 2) mx = 0x7fffff - (mx & 0xff); else if (mode == 3) mx &= 0xff;
            /* the guard of the kernels: |d| within 2^+-40, |x| <= 2^80, and a quotient of at least 2^-31 */
            int ed = 127 - 40 + (int)((r >> 49) % 81), ex = ed - 31 + (int)((r >> 56) % 100);
            if (ex > 127 + 80) ex = 127 + 80;
            if (ex < 1) ex = 1;
            float x = asf(((uint32_t)ex << 23) | mx | (uint32_t)((r >> 62) & 1) << 31), d = asf(((uint32_t)ed << 23) | md | (uint32_t)(r >> 63) << 31);
            if (fabsf(x) > 0x1p80f) continue;
            if (asu(div_by_recip(x, d, 1.0f / d)) != asu(x / d)) { if (bad++ < 5) printf("x=%a d=%a\n", x, d); }
        }
        printf("bad=%ld\n", bad);
        return bad != 0;
    }
""")


@pytest.mark.skipif(shutil.which("gcc") is None, reason="no gcc")
def test_two_residual_corrections_give_the_ieee_quotient(tmp_path):
    c = tmp_path / "recip.c"
    c.write_text(SRC)
    exe = tmp_path / "recip"
    # -ffp-contract=off: only the explicit fmaf calls fuse, as in the kernels (--fmad=false)
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), str(c), "-lm"], check=True)
    r = subprocess.run([str(exe), "20000000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "bad=0" in r.stdout, r.stdout[-400:]
