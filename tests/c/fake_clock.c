/* Test infrastructure: a deterministic frGetCurrentTime, LD_PRELOADed in front of either library, so that
   frUpdateWorld (wall-clock accumulator, reference src/world.c:268-285) can be compared bit for bit.
   Call k returns 1 + sum of the first k increments; the increments cycle through a pattern, in units of 1/60 s,
   that makes a call run 0, 1 or 2 steps. */
static const float pattern[] = { 1.0f, 1.0f, 1.0f, 0.4f, 2.3f, 1.0f, 0.2f, 0.9f, 3.1f, 1.0f };
static float now = 1.0f;
static int calls = 0;
float frGetCurrentTime(void) {
    float r = now;
    now += pattern[calls % 10] * (1.0f / 60.0f);
    ++calls;
    return r;
}
