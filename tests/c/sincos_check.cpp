// Host-side check of ferox_b200/csrc/fx_math.cuh: fx_sincosf (the glibc 2.39 sinf/cosf algorithm
// restated for the device integrator) against the libm of this machine, and the odd/even symmetry
// the narrow phase relies on (xf_unrotate).  Built and run by tests/test_sincos.py.
#include <cstdio>
#include <cstring>

#include "../../ferox_b200/csrc/fx_math.cuh"

static unsigned bits(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }

extern "C" long check_range(float lo, float hi, long *symmetry_violations, long *normalize_violations) {
    long bad = 0, sym = 0;
    for (unsigned u = bits(lo); u <= bits(hi); ++u) {
        float f;
        std::memcpy(&f, &u, 4);
        SinCos r = fx_sincosf(f);
        float s = sinf(f), c = cosf(f);
        if (bits(r.s) != bits(s) || bits(r.c) != bits(c)) ++bad;
        if (bits(sinf(-f)) != bits(-s) || bits(cosf(-f)) != bits(c)) ++sym;
    }
    *symmetry_violations = sym;
    *normalize_violations = 0;
    return bad;
}
