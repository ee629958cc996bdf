/* Measurement helper for bench.py (`e2e.ferox_h_loop`): the loop an existing ferox program runs, written against
   the UNCHANGED public API only (reference include/ferox.h:404-592) -- per step one frApplyForceToBody per body,
   frStepWorld, one frGetBodyPosition per body -- linked against libferox.so like any user program.
   Built by ferox_b200/csrc/Makefile into ferox_b200/_build/libfx_e2e_loop.so. */
#define _POSIX_C_SOURCE 199309L
#include <time.h>

#include "../../include/ferox_b200.h"

double fx_e2e_ferox_h_loop(frWorld *w, int steps, float dt, float fx, float fy, double *checksum) {
    struct timespec t0, t1;
    const frVector2 f = { fx, fy };
    double acc = 0.0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int s = 0; s < steps; ++s) {
        const int n = frGetBodyCountInWorld(w);
        for (int i = 0; i < n; ++i) {
            frBody *b = frGetBodyInWorld(w, i);
            frApplyForceToBody(b, frGetBodyPosition(b), f);
        }
        frStepWorld(w, dt);
        for (int i = 0; i < n; ++i) acc += frGetBodyPosition(frGetBodyInWorld(w, i)).y;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (checksum) *checksum = acc;
    return (double) (t1.tv_sec - t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
}
