// fx_device.h -- device-resident world layout and the host-callable launch surface.
//
// Data layout in HBM (all struct-of-arrays, 16-byte vectors so a warp's 32 bodies are 512 B
// contiguous per array):
//   bodies    posrot[i] = {x, y, sin, cos}      angle[i]
//             vel[i]    = {vx, vy, omega, invMass}
//             mprop[i]  = {mass, invInertia, gravityScale, as_float(type | flags<<8)}
//             force[i]  = {fx, fy, torque, 0}    aabb[i] = {x, y, w, h}
//             shape[i]  = {shape-table index (-1: none), shape type}      uid[i]
//   shapes    160-byte records (type, nv, radius, friction, restitution, verts[8], normals[8])
//   manifolds (the contact cache, dense, double-buffered prev/cur):
//             m_key {uidA, uidB}   m_body {idxA, idxB}   m_hdr {dir.x, dir.y, friction, restitution}
//             m_meta {count, colour}
//             per contact slot (2*m+c): c_geo {point.x, point.y, depth, timestamp}
//                                       c_imp {normalMass, normalScalar, tangentMass, tangentScalar}
//                                       c_id
//   hash map  open addressing over 4-byte slots holding an index into the dense manifold array; the key
//             (uidA<<32|uidB) is compared through that array, so the table stays small enough for L2
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "fx_math.cuh"

#define FX_MAX_VERTS 8
#define FX_MAX_COLORS 64

enum { FX_SHAPE_UNKNOWN = 0, FX_SHAPE_CIRCLE = 1, FX_SHAPE_POLYGON = 2 };
enum { FX_BODY_UNKNOWN = 0, FX_BODY_STATIC = 1, FX_BODY_KINEMATIC = 2, FX_BODY_DYNAMIC = 3 };

struct __align__(16) ShapeDev {
    int type, nv;
    float radius, friction;
    float restitution, bound, pad1, pad2;   // bound: radius of a circle around the body position that holds the shape (slightly inflated)
    V2 verts[FX_MAX_VERTS];
    V2 normals[FX_MAX_VERTS];
};
static_assert(sizeof(ShapeDev) == 160, "ShapeDev layout");

// Public frCollision layout (reference include/ferox.h:163-183), used for the API-shaped I/O of the
// single-pair entry points and the manifold dump.
struct ContactPod {
    int id;
    float depth, timestamp;
    V2 point;
    float normalMass, normalScalar, tangentMass, tangentScalar;
};
struct CollisionPod {
    int count;
    V2 direction;
    ContactPod contacts[2];
    float friction, restitution;
};
static_assert(sizeof(CollisionPod) == 92, "frCollision layout");
// one record of the collision-handler bridge = include/ferox_b200.h frxManifold
struct ManifoldRec {
    int first, second;
    CollisionPod value;
};
static_assert(sizeof(ManifoldRec) == 100, "frxManifold layout");

// what the narrow phase produces for one pair (the part of frCollision it writes)
struct Manifold {
    int count;       // 0 = miss
    int has_ids;     // polygon/polygon writes feature ids; the circle cases leave ids untouched
    V2 dir;
    V2 p[2];
    float depth[2];
    int id[2];
};

struct ManifoldSet {
    uint2 *key;      // {uidA, uidB}
    int2 *body;      // {idxA, idxB}
    float4 *hdr;     // {dir.x, dir.y, friction, restitution}
    int2 *meta;      // {count, colour}
    float4 *geo;     // [2*cap] {px, py, depth, timestamp}
    float4 *imp;     // [2*cap] {normalMass, normalScalar, tangentMass, tangentScalar}
    int *cid;        // [2*cap]
};

// Solver constraint records, colour-major, written once per step by the prepare phase and streamed
// by every Gauss-Seidel sweep.
#define FX_TRACE_SLOTS 32     // 0 start, 1 body state, 2 colours validated, 3 rounds done, 4 counted, 5 ordered, 6 records + warm start, 7.. after sweep i, 30 write-back, 31 rounds
struct SolverSet {
    int2 *body;      // {idxA, idxB}
    float4 *nt;      // {n.x, n.y, t.x, t.y}                            (full records)
    float2 *n;       // contact normal; the tangent is recomputed from it   (lean records, fx_solve.cu load_nt)
    float4 *mat;     // {friction, restitution, invInertiaA, invInertiaB}
    int *count_src;  // count | (source manifold index << 2)
    float4 *r;       // [2][cap] contact-major: {r1.x, r1.y, r2.x, r2.y}
    float4 *u;       // [2][cap] {u1.x, u1.y, u2.x, u2.y}   unit left-normals of r1, r2   (full records only)
    float4 *k;       // [2][cap] {bias, normalMass, tangentMass, 0}
    float2 *lam;     // [2][cap] {normalScalar, tangentScalar}
};

struct StepCountersDev {
    unsigned int n_entries;      // K
    unsigned int n_cand;         // C
    unsigned int n_hits;         // manifolds produced by this step's narrow phase
    unsigned int n_manifolds;    // M = hits + carried-over stale entries
    unsigned int n_contacts;     // Q
    unsigned int n_colors;
    unsigned int overflow;       // bit 0: entries, 1: candidates, 2: manifolds, 3: colours, 4: key range (non-finite extent / sort passes)
    unsigned int sort_passes;
    int cell_min_x, cell_min_y, cell_max_x, cell_max_y;
    unsigned int grid_w, grid_h, key_bits, n_large;
    unsigned int uncolored;      // scratch for the colouring loop
    unsigned int pad[3];
};

// One body row as it travels between neighbouring strips of a decomposed world (96 B).
struct __align__(16) GhostRow {
    float4 posrot, vel, mprop, force, aabb;
    float angle;
    int shape_idx, shape_type;
    unsigned int uid;
};
static_assert(sizeof(GhostRow) == 96, "GhostRow layout");
// message = 16-byte header {count, 0, 0, 0} + capacity rows (fixed size: no host round trip)
#define FX_MSG_HEADER 16

// Per-step scalars.  They live in device memory (DevWorld::scal) and not in the by-value kernel argument, so
// that a captured CUDA graph of the step stays valid when the time step or the world clock changes
// (frUpdateWorld moves the clock on every call): the host rewrites this block with a one-thread kernel ahead of
// the (replayed) launch sequence whenever a value changed.
struct StepScalars {
    float gx, gy;
    float timestamp;
    float dt, inv_dt;
    float pad[3];
};

struct DevWorld {
    int n;               // body rows the kernels sweep (owned bodies + ghost rows in strip mode)
    int n_owned;         // rows [0, n_owned) belong to this world; the rest are ghosts of neighbours
    int n_shapes;
    float cell, inv_cell;
    const StepScalars *scal;   // gravity, world clock, dt (see above)
    int solver_mode;
    int sort_launch;     // radix-sort passes the host launches this step (k_plan refuses a world whose keys need more)
    int color_squeeze;   // how many of last step's top colours are re-coloured every step (0: never)
    int narrow_mode;     // 0: one lane per pair (default), 1: 8 lanes per pair with shuffle arg-max over axes
    // bodies
    float4 *posrot, *vel, *mprop, *force, *aabb;
    float *angle;
    int2 *shape;
    unsigned int *uid;
    const int *group;    // batched worlds: sub-world index per body (nullptr: a single world)
    int n_groups;
    const int2 *grange;  // per sub-world {first body index, body count} (host-maintained)
    unsigned int *gcnt, *goff, *gfill;   // per sub-world manifold count / offset / fill cursor
    unsigned int *gorder;                // manifold indices grouped by sub-world
    int *idx_of_uid;     // uid -> world index, -1 when the body left the world
    const ShapeDev *shapes;
    // broadphase scratch
    int4 *rect;          // per body inclusive cell rectangle {x0, y0, x1, y1}
    unsigned int *cell_cnt, *cell_off;   // per body entry count / exclusive offset
    unsigned int *large_list;            // bodies whose rectangle is expanded by a whole block
    unsigned int *keys[2], *vals[2];     // (cell key, packed body) entries, ping-pong for the sort
    unsigned int *keys_hi[2];            // bits 32..63 of the cell keys; touched only in steps whose key needs them
    unsigned int *hist;                  // radix histograms [256][tiles]
    unsigned int cap_entries, cap_cand, cap_manifolds, cap_hash, cap_large;
    int2 *cand;          // candidate pairs (i < j)
    unsigned char *cand_hit;
    // contact cache
    ManifoldSet man[2];  // double buffer: man[*cur_flip] is written this step, the other one is
                         // last step's cache (read-only); k_finish_step flips the selector
    int *cur_flip;
    unsigned char *touched;              // per prev entry: probed (hit or miss) by this step
    unsigned int *h_slot;                // [cap_hash] index into the dense array the table was built over, ~0u: empty
    // solver
    SolverSet sol;
    int sol_lean;        // solver records without what a sweep can recompute (see SolverSet)
    unsigned int sol_stride, sol_mult;   // per-contact record (slot s, contact c) lives at c * sol_stride + s * sol_mult
    unsigned int *order;                 // manifold indices grouped by colour
    unsigned int *color_off;             // [FX_MAX_COLORS + 1]
    unsigned long long *body_mask;       // colours already used by each body
    unsigned long long *body_dup;        // colours seen more than once on a body (inherited clashes)
    unsigned int *winner;                // colouring: per-body claim, [2][n] (alternating by round)
    int2 *cbody;                         // per manifold: the bodies that constrain its colour (-1: none)
    unsigned int *color_scratch;         // [4][FX_MAX_COLORS]: counts, fill cursors, per-round remaining, flags
    unsigned long long *solve_trace;     // FEROX_SOLVE_TRACE=1: [FX_TRACE_SLOTS] globaltimer stamps of k_solve_colored's stages (else nullptr)
    // single-pass ordered compaction (decoupled look-back)
    StepCountersDev *ctr;                // current step
    unsigned int *prev_count;            // [0] number of entries in the prev manifold set, [1] colours used by the last solve
};
