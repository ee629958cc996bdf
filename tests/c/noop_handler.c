/* A collision handler that does nothing, for bench.py's "handler installed" number: with any handler
   installed the step has to hand every live manifold to the host and take the (possibly edited)
   manifolds back (reference src/world.c:209-214, 254-259), whatever the callbacks do.
   Built by ferox_b200/csrc/Makefile into ferox_b200/_build/libfx_noop_handler.so. */
typedef struct { void *first, *second; } BodyPair;
static unsigned long long calls;
void fx_noop_collision_event(BodyPair key, void *value) { (void) key; (void) value; ++calls; }
unsigned long long fx_noop_collision_event_calls(void) { return calls; }
