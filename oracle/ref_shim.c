/*
 * TEST INFRASTRUCTURE ONLY -- never linked into, imported by or shipped with the product library.
 *
 * Instrumentation shim linked into oracle/_ref/libferox_ref*.so next to the UNMODIFIED reference
 * objects (compiled where they lie under /root/reference/src by oracle/Makefile).  The link uses
 * `-Wl,--wrap=frComputeCollision`, so the one call the reference world step makes into the narrow
 * phase (reference src/world.c:347, inside frPreStepHashQueryCallback) lands here first.  We record
 * (first body, second body, hit?) in call order and forward to the real function.  This gives the
 * parity tests the reference's candidate pair SEQUENCE of a step without touching reference sources.
 */
#include <stdbool.h>
#include <stdlib.h>

typedef struct frBody_ frBody;
typedef struct frCollision_ frCollision;

bool __real_frComputeCollision(frBody *b1, frBody *b2, frCollision *collision);

typedef struct { frBody *first, *second; int hit; } refshim_call;

static refshim_call *g_calls;
static size_t g_len, g_cap;
static int g_enabled;

void refshim_enable(int on) { g_enabled = on; }
void refshim_reset(void) { g_len = 0; }
size_t refshim_count(void) { return g_len; }
const refshim_call *refshim_calls(void) { return g_calls; }

bool __wrap_frComputeCollision(frBody *b1, frBody *b2, frCollision *collision) {
    bool hit = __real_frComputeCollision(b1, b2, collision);
    if (g_enabled) {
        if (g_len == g_cap) {
            g_cap = g_cap ? g_cap * 2 : 4096;
            g_calls = realloc(g_calls, g_cap * sizeof *g_calls);
        }
        g_calls[g_len++] = (refshim_call) { b1, b2, hit };
    }
    return hit;
}
