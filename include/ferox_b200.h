/*
 * ferox_b200.h -- C ABI of the B200-native ferox world-step library (libferox.so / libferox.a).
 *
 * DROP-IN BOUNDARY.  The library exports, symbol for symbol and layout for layout, the public C API
 * of c-krit/ferox v0.9.7 (reference include/ferox.h).  An existing program keeps compiling against
 * the reference's own ferox.h and only relinks:   cc app.c -I<ferox>/include -L. -lferox -lm
 * Part 1 of this header restates that ABI (each group cites the reference lines it replaces) and is
 * skipped when the reference header was included first (its FEROX_H guard), so both can coexist.
 * Part 2 declares the `frx*` extension entry points: batch/array access, capacity, counters and the
 * multi-GPU strip set-up that ferox.h has no vocabulary for.  Plain pointers and sizes only.
 *
 * Everything that the reference runs inside frStepWorld (reference src/world.c:204-262) executes in
 * hand-written sm_100a CUDA kernels; there is no CPU fallback: without a CUDA device the step entry
 * points print a diagnostic to stderr and leave the world untouched (frxGetLastError() != 0).
 */
#ifndef FEROX_B200_H
#define FEROX_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ============================ Part 1: the ferox.h ABI ============================ */
#ifndef FEROX_H

/* compile-time configuration of the reference (ferox.h:43-76); the library is built with the same
   values.  FR_WORLD_MAX_OBJECT_COUNT is only the DEFAULT capacity here, see frxSetWorldCapacity. */
#ifndef FR_GEOMETRY_MAX_VERTEX_COUNT
#define FR_GEOMETRY_MAX_VERTEX_COUNT 8
#endif
#ifndef FR_WORLD_BAUMGARTE_FACTOR
#define FR_WORLD_BAUMGARTE_FACTOR 0.2f
#endif
#ifndef FR_WORLD_BAUMGARTE_SLOP
#define FR_WORLD_BAUMGARTE_SLOP 0.01f
#endif
#ifndef FR_WORLD_ITERATION_COUNT
#define FR_WORLD_ITERATION_COUNT 10
#endif
#ifndef FR_WORLD_MAX_OBJECT_COUNT
#define FR_WORLD_MAX_OBJECT_COUNT 2048
#endif

/* plain-old-data that crosses the boundary by value or pointer (ferox.h:122-271) */
typedef struct frVector2_ { float x, y; } frVector2;                         /* 8 B  */
typedef struct frAABB_ { float x, y, width, height; } frAABB;               /* 16 B */
typedef struct frContextNode_ { int id; void *ctx; } frContextNode;         /* 16 B */
typedef struct frContact_ {                                                  /* 36 B */
    int id;
    float depth, timestamp;
    frVector2 point;
    struct { float normalMass, normalScalar, tangentMass, tangentScalar; } cache;
} frContact;
typedef struct frCollision_ {                                                /* 92 B */
    int count;
    frVector2 direction;
    frContact contacts[2];
    float friction, restitution;
} frCollision;
typedef struct frMaterial_ { float density, friction, restitution; } frMaterial;
typedef struct frVertices_ { frVector2 data[FR_GEOMETRY_MAX_VERTEX_COUNT]; int count; } frVertices;
typedef struct frTransform_ {                                                /* 20 B */
    frVector2 position;
    struct { float sin_, cos_; } rotation;
    float angle;
} frTransform;
typedef struct frRay_ { frVector2 origin, direction; float maxDistance; } frRay;

typedef enum frShapeType_ { FR_SHAPE_UNKNOWN, FR_SHAPE_CIRCLE, FR_SHAPE_POLYGON } frShapeType;
typedef enum frBodyType_ { FR_BODY_UNKNOWN, FR_BODY_STATIC, FR_BODY_KINEMATIC, FR_BODY_DYNAMIC } frBodyType;
typedef enum frBodyFlag_ { FR_FLAG_NONE, FR_FLAG_INFINITE_MASS, FR_FLAG_INFINITE_INERTIA } frBodyFlag;
typedef unsigned int frBodyFlags;

/* opaque handles (ferox.h:138-155): the host objects behind them are ours to design */
typedef struct frShape_ frShape;
typedef struct frBody_ frBody;
typedef struct frWorld_ frWorld;
typedef struct frSpatialHash_ frSpatialHash;

typedef struct frRaycastHit_ {                                               /* 32 B */
    frBody *body;
    frVector2 point, normal;
    float distance;
    bool inside;
} frRaycastHit;
typedef struct frBodyPair_ { frBody *first, *second; } frBodyPair;

typedef bool (*frHashQueryFunc)(frContextNode ctxNode);
typedef void (*frCollisionEventFunc)(frBodyPair key, frCollision *value);
typedef struct frCollisionHandler_ { frCollisionEventFunc preStep, postStep; } frCollisionHandler;
typedef void (*frRaycastQueryFunc)(frRaycastHit raycastHit, void *ctx);

/* --- world: the hot path and its container (replaces ferox.h:544-592, src/world.c) ------------- */
frWorld *frCreateWorld(frVector2 gravity, float cellSize);
void frReleaseWorld(frWorld *w);
void frClearWorld(frWorld *w);
bool frAddBodyToWorld(frWorld *w, frBody *b);         /* deferred: applied in the next post-step */
bool frRemoveBodyFromWorld(frWorld *w, frBody *b);    /* deferred, swap-remove renumbering       */
bool frIsBodyInWorld(const frWorld *w, frBody *b);
frBody *frGetBodyInWorld(const frWorld *w, int i);
int frGetBodyCountInWorld(const frWorld *w);
frVector2 frGetWorldGravity(const frWorld *w);
void frSetWorldCollisionHandler(frWorld *w, frCollisionHandler handler);
void frSetWorldGravity(frWorld *w, frVector2 gravity);
void frStepWorld(frWorld *w, float dt);               /* ferox.h:577  src/world.c:204-262  (GPU) */
void frUpdateWorld(frWorld *w, float dt);             /* ferox.h:583  src/world.c:268-285  (GPU) */
void frComputeWorldRaycast(frWorld *w, frRay ray, frRaycastQueryFunc func, void *userData);

/* --- narrow phase (replaces ferox.h:307-310, src/collision.c) ---------------------------------- */
bool frComputeCollision(frBody *b1, frBody *b2, frCollision *collision);          /* (GPU) */
bool frComputeRaycast(const frBody *b, frRay ray, frRaycastHit *raycastHit);

/* --- solver and integrator entry points (replaces ferox.h:499-531, src/rigid_body.c:292-551) --- */
void frClearBodyForces(frBody *b);
void frApplyForceToBody(frBody *b, frVector2 point, frVector2 force);
void frApplyGravityToBody(frBody *b, frVector2 g);
void frApplyImpulseToBody(frBody *b, frVector2 point, frVector2 impulse);
void frApplyAccumulatedImpulses(frBody *b1, frBody *b2, frCollision *collision);  /* (GPU) */
void frIntegrateForBodyVelocity(frBody *b, float dt);                             /* (GPU) */
void frIntegrateForBodyPosition(frBody *b, float dt);                             /* (GPU) */
void frResolveCollision(frBody *b1, frBody *b2, frCollision *collision, float inverseDt); /* (GPU) */

/* --- rigid bodies (replaces ferox.h:404-497, src/rigid_body.c:67-289) -------------------------- */
frBody *frCreateBody(frBodyType type, frVector2 position);
frBody *frCreateBodyFromShape(frBodyType type, frVector2 position, frShape *s);
void frReleaseBody(frBody *b);
frBodyType frGetBodyType(const frBody *b);
frBodyFlags frGetBodyFlags(const frBody *b);
frShape *frGetBodyShape(const frBody *b);
frTransform frGetBodyTransform(const frBody *b);
frVector2 frGetBodyPosition(const frBody *b);
float frGetBodyAngle(const frBody *b);
float frGetBodyMass(const frBody *b);
float frGetBodyInverseMass(const frBody *b);
float frGetBodyInertia(const frBody *b);
float frGetBodyInverseInertia(const frBody *b);
float frGetBodyGravityScale(const frBody *b);
frVector2 frGetBodyVelocity(const frBody *b);
float frGetBodyAngularVelocity(const frBody *b);
frVector2 frGetBodyForce(const frBody *b);
float frGetBodyTorque(const frBody *b);
frAABB frGetBodyAABB(const frBody *b);
void *frGetBodyUserData(const frBody *b);
void frSetBodyType(frBody *b, frBodyType type);
void frSetBodyFlags(frBody *b, frBodyFlags flags);
void frSetBodyShape(frBody *b, frShape *s);
void frSetBodyTransform(frBody *b, frTransform tx);   /* declared at ferox.h:476, defined nowhere in
                                                         the reference; defined here */
void frSetBodyPosition(frBody *b, frVector2 position);
void frSetBodyAngle(frBody *b, float angle);
void frSetBodyGravityScale(frBody *b, float scale);
void frSetBodyVelocity(frBody *b, frVector2 v);
void frSetBodyAngularVelocity(frBody *b, float angularVelocity);
void frSetBodyUserData(frBody *b, void *userData);
bool frBodyContainsPoint(const frBody *b, frVector2 point);

/* --- shapes (replaces ferox.h:315-399, src/geometry.c; set-up time, host) ---------------------- */
frShape *frCreateCircle(frMaterial material, float radius);
frShape *frCreateRectangle(frMaterial material, float width, float height);
frShape *frCreatePolygon(frMaterial material, const frVertices *vertices);
void frReleaseShape(frShape *s);
frShapeType frGetShapeType(const frShape *s);
frMaterial frGetShapeMaterial(const frShape *s);
float frGetShapeDensity(const frShape *s);
float frGetShapeFriction(const frShape *s);
float frGetShapeRestitution(const frShape *s);
float frGetShapeArea(const frShape *s);
float frGetShapeMass(const frShape *s);
float frGetShapeInertia(const frShape *s);
frAABB frGetShapeAABB(const frShape *s, frTransform tx);
float frGetCircleRadius(const frShape *s);
frVector2 frGetPolygonVertex(const frShape *s, int i);
const frVertices *frGetPolygonVertices(const frShape *s);
frVector2 frGetPolygonNormal(const frShape *s, int i);
const frVertices *frGetPolygonNormals(const frShape *s);
void frSetShapeType(frShape *s, frShapeType type);
void frSetShapeMaterial(frShape *s, frMaterial material);
void frSetShapeDensity(frShape *s, float density);
void frSetShapeFriction(frShape *s, float friction);
void frSetShapeRestitution(frShape *s, float restitution);
void frSetCircleRadius(frShape *s, float radius);
void frSetRectangleDimensions(frShape *s, float width, float height);
void frSetPolygonVertices(frShape *s, const frVertices *vertices);

/* --- stand-alone spatial hash (replaces ferox.h:281-299, src/broad_phase.c; query API, host) --- */
frSpatialHash *frCreateSpatialHash(float cellSize);
void frReleaseSpatialHash(frSpatialHash *sh);
void frClearSpatialHash(frSpatialHash *sh);
float frGetSpatialHashCellSize(const frSpatialHash *sh);
void frInsertIntoSpatialHash(frSpatialHash *sh, frAABB key, int value);
void frQuerySpatialHash(frSpatialHash *sh, frAABB aabb, frHashQueryFunc func, void *userData);

/* --- clock (replaces ferox.h:536, src/timer.c); weak, so a program may interpose its own ------- */
float frGetCurrentTime(void);

#endif /* !FEROX_H */

/* ============================ Part 2: frx* extensions ============================ */

/* Solver ordering.  FRX_SOLVER_COLORED is the production path (graph-coloured Gauss-Seidel);
   FRX_SOLVER_SERIAL visits manifolds in the reference's contact-cache array order (one thread),
   which makes a step bit-comparable with the reference CPU step.  Also settable with the
   environment variable FEROX_SOLVER=colored|serial read at frCreateWorld. */
enum { FRX_SOLVER_COLORED = 0, FRX_SOLVER_SERIAL = 1 };

/* per-step work counters of the last completed step */
typedef struct frxStepCounters_ {
    long long bodies;      /* N                                                    */
    long long cellEntries; /* K: sum over bodies of cells covered                   */
    long long candidates;  /* C: unique pairs that reached the narrow phase         */
    long long manifolds;   /* M: contact-cache entries solved                       */
    long long contacts;    /* Q: contact points (sum of counts)                     */
    long long colors;      /* colours used by the solver (0 in serial mode)         */
    long long steps;       /* frStepWorld calls completed on this world             */
    long long overflow;    /* != 0: a device buffer overflowed (work was dropped)   */
    long long colorRounds; /* rounds the greedy colouring needed                    */
} frxStepCounters;

/* CUDA device selection for worlds created afterwards by this thread (default: device 0, or the
   environment variable FEROX_DEVICE).  Returns 0 on success. */
int frxSetDevice(int device);
/* 0 = no error; otherwise the sticky code of the first CUDA/capacity failure.  frxGetLastErrorString
   returns a static description. */
int frxGetLastError(void);
void frxClearLastError(void);      /* the error is sticky: the first one is kept until cleared */
const char *frxGetLastErrorString(void);

/* Capacity: the reference fixes the body count at compile time (FR_WORLD_MAX_OBJECT_COUNT,
   ferox.h:73-76, used at src/world.c:113-115,149).  Here it is the default; this call raises the
   limit (and the add/remove queue, rounded up to a power of two with one unusable slot as in
   src/external/ferox_utils.h:172-205) for one world.  FEROX_MAX_OBJECT_COUNT in the environment
   changes the default for worlds created afterwards. */
bool frxSetWorldCapacity(frWorld *w, int maxObjectCount);
int frxGetWorldCapacity(const frWorld *w);

/* Batched independent worlds (RL-style: thousands of small worlds): ONE device pipeline and one
   launch set per step for all of them.  Bodies of different sub-worlds never become candidates (the
   sub-world index is folded into the broadphase cell key); each sub-world evolves exactly as a
   separate frWorld with the same bodies in the same order would.  frxAddBodyToBatch is deferred like
   frAddBodyToWorld (src/world.c:147-155); frStepWorld(batch, dt) steps every sub-world. */
frWorld *frxCreateWorldBatch(frVector2 gravity, float cellSize, int nWorlds);
bool frxAddBodyToBatch(frWorld *batch, int worldIndex, frBody *b);
int frxGetBatchWorldCount(const frWorld *batch);

void frxSetWorldSolverMode(frWorld *w, int mode);
int frxGetWorldSolverMode(const frWorld *w);
/* world clock used for contact staleness (src/world.c:222-235, 362); frStepWorld never advances it */
void frxSetWorldTimestamp(frWorld *w, float timestamp);
float frxGetWorldTimestamp(const frWorld *w);

/* n frStepWorld(w, dt) calls back to back without returning to the host in between (no collision
   handler must be installed). */
void frxStepWorldN(frWorld *w, float dt, int n);
/* block until all device work of `w` has finished (getters do this implicitly) */
void frxSynchronizeWorld(frWorld *w);
void frxGetStepCounters(frWorld *w, frxStepCounters *out);
/* `overflow` accumulates (bitwise or) over all steps since the world was created or this was last called */
void frxResetStepCounters(frWorld *w);

/* Batch access to the bodies of a world in world-index order; every array is optional (NULL).
   Layouts: position float[n][2], angle float[n], sincos float[n][2], velocity float[n][2],
   angularVelocity float[n], aabb float[n][4], force float[n][2], torque float[n]. */
int frxGetBodyStates(frWorld *w, float *position, float *angle, float *sincos, float *velocity,
                     float *angularVelocity, float *aabb);
/* Zero-copy variant: downloads the state once (if the device is ahead) and returns READ-ONLY views of
   the library's pinned host mirror, valid until the next call that steps or edits the world:
   posRot float[n][4] = (x, y, sin, cos), angle float[n], velocity float[n][4] = (vx, vy, omega,
   1/mass), aabb float[n][4] = (x, y, w, h).  This is what the per-body getters read.  Any of the four
   may be NULL: only the arrays asked for are downloaded (the others follow when a getter needs them).
   Returns n. */
int frxViewBodyStates(frWorld *w, const float **posRot, const float **angle, const float **velocity, const float **aabb);
/* Writable view of the force accumulators in the pinned host mirror: force float[n][4] = (fx, fy, torque, 0),
   zero after every step (src/world.c:481-482); whatever the caller ADDS here before the next step is uploaded
   with it, exactly as frApplyForceToBody would have accumulated it (the caller skips bodies that cannot move:
   velocity[i][3], the inverse mass, is 0 for them).  Returns n. */
int frxMapBodyForces(frWorld *w, float **force);
/* adds force/torque to every body with inverse mass > 0, like n frApplyForceToBody calls at the
   body's centre plus a pure torque (src/rigid_body.c:299-306) */
int frxApplyForces(frWorld *w, const float *force, const float *torque, int n);
int frxSetBodyVelocities(frWorld *w, const float *velocity, const float *angularVelocity, int n);

/* Bulk creation: n bodies created, attached to shapes[shapeIndex[i]] and queued/added exactly as n
   frCreateBodyFromShape + frSetBodyAngle + frAddBodyToWorld calls would be; outBodies (optional)
   receives the handles.  Returns the number of bodies accepted by the add queue. */
int frxCreateBodies(frWorld *w, int n, const int *bodyType, const int *shapeIndex, frShape *const *shapes,
                    const float *position, const float *angle, frBody **outBodies);

/* As frxCreateBodies, plus optional initial velocities and, for a batch, the sub-world of each body. */
int frxCreateBodiesEx(frWorld *w, int n, const int *bodyType, const int *shapeIndex, frShape *const *shapes,
                      const float *position, const float *angle, const float *velocity,
                      const float *angularVelocity, const int *worldIndex, frBody **outBodies);

/* Strip decomposition of ONE large world over several GPUs (BASELINE config 5).
   Each strip is a frWorld that owns the bodies added to it; frxStripConfigure (on the still-empty
   world) gives it the x-interval [cutLo, cutHi) it is responsible for (use -INFINITY / INFINITY at the
   ends), the border `margin` (bodies whose AABB reaches within `margin` of a cut are mirrored on the
   neighbour as ghosts: at least one cell plus the largest body extent plus the expected drift), the
   ghost capacity per side and this strip's slice [uidBase, ...) of a global id space of uidSpace ids.
   Static bodies that span cuts must be added to every strip.  Ghost rows carry shape-table indices:
   call frxStripRegisterShapes with the complete shape list, in the same order, on every strip.  Neighbours are linked either in-process (frxStripLinkLocal: tests, or one
   process driving several GPUs) or across processes with NCCL (frxStripGetNcclUniqueId on rank 0,
   broadcast the 128 bytes, frxStripInitNccl everywhere).  frxStripStep advances all strips of this
   process by one step; frStepWorld is refused on a strip.  Ownership is static in this version. */
int frxStripConfigure(frWorld *w, float cutLo, float cutHi, float margin, int ghostCapacity, unsigned int uidBase,
                      unsigned int uidSpace);
int frxStripRegisterShapes(frWorld *w, frShape *const *shapes, int n);   /* same list, same order, on every strip */
int frxStripLinkLocal(frWorld *left, frWorld *right);
int frxStripGetNcclUniqueId(void *out128);
int frxStripInitNccl(frWorld *w, int rank, int nranks, const void *uniqueId128);
void frxStripStep(frWorld **worlds, int nLocal, float dt);
/* how this strip talks to its neighbours: 0 in-process links (copies between worlds of this process), 1 NCCL send/recv,
   2 peer-direct (the default across processes when every rank could map its neighbours' receive buffers through CUDA
   IPC: messages are stored straight into the neighbour's buffer over NVLink by the sender's kernel and announced with a
   flag; FEROX_STRIP_PEER=0 keeps NCCL); -1 not a strip */
int frxStripGetTransport(const frWorld *w);
/* Ownership migration.  Every non-static body whose position has left its strip's x-interval by more than a quarter of
   the margin changes owner: it is removed from its strip and added to the neighbour on that side, with its state
   (transform, velocities, accumulated forces) intact; its contacts are found again by the new owner's next step (their
   warm-start impulses start from zero once).  Collective: every process calls it at the same step with its strips in
   the order it gives frxStripStep; call it every few dozen steps (it synchronises the strips and re-sizes their buffers).
   In-process neighbours hand over the frBody itself.  Across processes (NCCL) the body is RE-CREATED on the receiving
   rank -- reach it through frGetBodyInWorld -- and its handle on the sending rank is released: do not use it again.
   Returns the number of bodies that left strips of this process, -1 on error. */
int frxStripMigrate(frWorld **worlds, int nLocal);

/* Introspection for parity tests: the candidate pairs (world indices, first < second, plus hit flag)
   of the last step, and the contact cache as (first index, second index, frCollision) records in
   the order the solver visits them.  Each returns the number of records available; at most `max`
   are written. */
int frxGetCandidatePairs(frWorld *w, int *outTriples, int max);
typedef struct frxManifold_ { int first, second; frCollision value; } frxManifold;   /* 100 B */
int frxGetManifolds(frWorld *w, frxManifold *out, int max);

/* World ray casts on the DEVICE state, many rays per call (the sensor path of batched RL worlds): for ray r up to
   maxHitsPerRay hits are written to hits[r * maxHitsPerRay ...] in ascending world-index order -- exactly the
   bodies, the order and the values frComputeWorldRaycast's callback would see for that ray (src/world.c:291-320) --
   and counts[r] receives the number of hits found (it may exceed maxHitsPerRay: the rest were dropped).  Returns
   the total number of hits found.  frComputeWorldRaycast itself runs through this path. */
int frxRaycastBatch(frWorld *w, const frRay *rays, int nRays, frRaycastHit *hits, int *counts, int maxHitsPerRay);

/* Point queries on the device state: for point p the world indices of the bodies that contain it -- frBodyContainsPoint
   (src/rigid_body.c:261-289) evaluated for every body, ascending -- are written to bodyIndices[p * maxPerPoint ...],
   counts[p] receives how many there are (it may exceed maxPerPoint).  Returns the total. */
int frxPointQueryBatch(frWorld *w, const frVector2 *points, int nPoints, int *bodyIndices, int *counts, int maxPerPoint);

/* Batched narrow phase on bodies of a world: pairs[2*i], pairs[2*i+1] are world indices; out[i] is
   written only on a hit (like frComputeCollision); hit[i] receives 0/1.  Returns hits. */
int frxComputeCollisions(frWorld *w, const int *pairs, int nPairs, frCollision *out, int *hit);

/* Measurement hooks.  frxSetProfiling(w, 1) brackets every stage of every following step with CUDA
   events on the world's stream; frxGetStageTimes sums them into out[5] (ms): broadphase (+ velocity
   integration), narrow phase, cache update, solve, position integration, and returns the number of
   profiled steps.  frxRecordEvent / frxEventElapsedMs time arbitrary spans on the same stream
   (slots 0..3).  frxGetLaunchCount: kernels launched by the library so far (all worlds). */
void frxSetProfiling(frWorld *w, int on);
int frxGetStageTimes(frWorld *w, double *outMs5);
void frxRecordEvent(frWorld *w, int slot);
float frxEventElapsedMs(frWorld *w, int fromSlot, int toSlot);
long long frxGetLaunchCount(void);
/* steps of this world that were served by replaying the captured CUDA graph of the launch sequence (the graph
   survives changes of dt and of the world clock: frUpdateWorld replays it call after call) */
long long frxGetGraphReplays(const frWorld *w);
/* which kernel solved the last production-mode step: 0 = k_solve_colored on the whole grid (large worlds, strips),
   -1 = k_solve_local (one CTA per sub-world / small world), 8 or 16 = k_sweep_cluster in one thread-block cluster of
   that many CTAs (mid-size worlds; FEROX_CLUSTER=0 disables it, FEROX_CLUSTER=8|16 pins the size) */
int frxGetWorldSolverKernel(const frWorld *w);
/* manifolds per colour of the last production-mode step: out[c] for c < 64; returns the number of colours */
int frxGetColorHistogram(frWorld *w, int *out64);
/* FEROX_SOLVE_TRACE=1 in the environment when the world is created: global-timer stamps (ns) of the coloured solver's
   stages in the last step -- [0] start, [1] body state cleared, [2] inherited colours validated, [3] colouring rounds
   done, [4] colours counted, [5] manifolds ordered by colour, [6] records built + warm start, [7 + i] sweep i done.
   Returns the slots written (0: tracing is off). */
int frxGetSolveTrace(frWorld *w, unsigned long long *out, int n);
/* cudaProfilerStart (1) / cudaProfilerStop (0): lets `ncu --profile-from-start off` capture a span */
void frxProfilerRange(int start);
/* n frStepWorld calls, each bracketed by its own CUDA-event pair on the world's stream; with flushL2
   a 256 MiB scratch buffer is overwritten before every step, outside the bracket (cold L2).
   outMs[n] = per-step device milliseconds.  Returns n. */
int frxTimedSteps(frWorld *w, float dt, int n, int flushL2, float *outMs);

/* Device sinf/cosf (the glibc 2.39 algorithm restated for the position integrator,
   src/rigid_body.c:234-235): evaluates n angles on the GPU; used by the parity tests. */
int frxDeviceSinCos(const float *angles, int n, float *outSin, float *outCos);

#ifdef __cplusplus
}
#endif
#endif /* FEROX_B200_H */
