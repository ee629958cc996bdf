// fx_host.h -- host-side objects behind the opaque ferox.h handles.
//
// frShape / frBody / frWorld / frSpatialHash are opaque in the public API (reference
// include/ferox.h:138-155), so their layout is ours.  A body keeps its slowly-changing properties
// itself; its DYNAMIC state (transform, velocity, force accumulator, AABB) lives in the packed rows
// below -- inside the body while it is outside a world, inside the world's pinned host mirror of the
// device struct-of-arrays while it is in one.  The device copy is authoritative after a step;
// getters pull it back lazily (one bulk D2H per step at most), setters mark rows dirty for the next
// bulk H2D.
#pragma once
#include <cstdint>
#include <unordered_map>
#include <vector>

#include "../../include/ferox_b200.h"
#include "fx_device.h"
#include "fx_kernels.h"

struct frShape_ {
    frShapeType type;
    frMaterial material;
    float area;
    float radius;
    frVertices vertices, normals;
};

// packed dynamic state of one body; the same packing as the device arrays (fx_device.h)
struct BodyRow {
    float4 posrot;   // x, y, sin, cos
    float angle;
    float4 vel;      // vx, vy, omega, invMass
    float4 mprop;    // mass, invInertia, gravityScale, bits(type | flags << 8)
    float4 force;    // fx, fy, torque, 0
    float4 aabb;     // x, y, w, h
};

struct frBody_ {
    frBodyType type;
    frBodyFlags flags;
    frShape *shape;
    float mass, inv_mass, inertia, inv_inertia, gravity_scale;
    void *ctx;
    frWorld *world;      // non-null while the body is a member of world->bodies
    int index;           // world index while in a world
    uint32_t uid;        // stable id while in a world (contact-cache key)
    int group;           // sub-world index inside a batch (0 in a plain world)
    BodyRow loc;         // dynamic state while outside a world
};

enum { FX_OP_ADD = 1, FX_OP_REMOVE = 2 };   // reference src/world.c:35-39
struct WorldOp {
    int id;
    frBody *body;
};

// dirty bits of the host mirror (which arrays must be uploaded before the next step)
enum {
    FX_DIRTY_POSROT = 1, FX_DIRTY_ANGLE = 2, FX_DIRTY_VEL = 4, FX_DIRTY_MPROP = 8, FX_DIRTY_FORCE = 16,
    FX_DIRTY_AABB = 32, FX_DIRTY_SHAPE = 64, FX_DIRTY_UID = 128, FX_DIRTY_ALL = 255
};

struct HostMirror {     // pinned host arrays, capacity `cap`
    int cap = 0;
    float4 *posrot = nullptr, *vel = nullptr, *mprop = nullptr, *force = nullptr, *aabb = nullptr;
    float *angle = nullptr;
    int2 *shape = nullptr;
    unsigned int *uid = nullptr;
    int *group = nullptr;
};

struct frWorld_ {
    int device = 0;
    cudaStream_t stream = nullptr;
    frVector2 gravity{ 0.f, 0.f };
    float cell = 0.f;
    std::vector<frBody *> bodies;
    // deferred add/remove queue: ring of pow2(capacity) slots, one unusable
    // (reference src/external/ferox_utils.h:172-219, src/world.c:113-115)
    std::vector<WorldOp> ring;
    int ring_head = 0, ring_tail = 0;
    int capacity = FR_WORLD_MAX_OBJECT_COUNT;
    float accumulator = 0.f, timestamp = 0.f;
    frCollisionHandler handler{ nullptr, nullptr };
    int solver_mode = FRX_SOLVER_COLORED;
    void *bridge = nullptr;              // collision-handler bridge buffers (fx_world.cpp HandlerBridge)
    void *mh_cache = nullptr;            // pinned host copy of the manifold arrays (handlers, frxGetManifolds)
    int n_groups = 1;                // > 1: a batch of independent sub-worlds sharing this pipeline
    int *d_group = nullptr;
    // block-local solve (fx_solve.cu k_solve_local): per sub-world body ranges, host-maintained
    std::vector<int2> grange_host;
    int2 *d_grange = nullptr;
    unsigned int *d_gcnt = nullptr, *d_goff = nullptr, *d_gfill = nullptr, *d_gorder = nullptr;
    int d_groups_cap = 0;
    unsigned int d_gorder_cap = 0;
    int local_max_bodies = 0;        // 0: not eligible (some sub-world's bodies are not one short index range)
    int use_local = 1;
    int cluster_size = 0;            // > 0: the solve runs in one thread-block cluster of this many CTAs (fx_solve.cu k_sweep_cluster)
    int cluster_bodies = -1;         // body count cluster_size was decided for
    int force_large = 0;             // FEROX_FORCE_LARGE=1: the large-world kernel variants at any size (parity tests)

    HostMirror h;
    unsigned dirty = 0;
    bool device_newer = false;       // device state advanced past the host mirror
    bool force_clear_pending = false;   // a completed step cleared the device accumulators; the mirror's are zeroed at the next pull
    bool extent_unknown = false;     // a setter moved a body: launch every sort pass until a step has re-measured the key width
    long long extent_mark_step = 0;  // w->steps when that happened
    long long snapshot_step = -1;    // index of the step the pending / last snapshot belongs to
    long long snapshot_read_step = 0;  // w->steps when a snapshot was last read (the host never runs further ahead than FX_MAX_RUN_AHEAD steps)
    StepScalars *d_scal = nullptr;   // device block behind DevWorld::scal
    StepScalars scal_dev{};          // what that block holds
    bool scal_valid = false;
    unsigned stale = 0;              // while device_newer was being consumed piecewise: FX_DIRTY_* bits of mirror arrays not yet downloaded
    bool needs_sizing = true;        // run the sizing pre-pass before the next step
    bool device_ok = false;

    // shape table
    std::vector<frShape *> shape_list;
    std::unordered_map<const frShape *, int> shape_index;
    std::vector<ShapeDev> shape_host;
    ShapeDev *d_shapes = nullptr;
    int d_shapes_cap = 0;
    uint64_t shape_epoch_seen = 0;
    bool shape_rows_stale = false;   // a member body changed its shape: every row's table index is looked up again (adds / removes keep their own rows)

    // uid allocator
    std::vector<int> uid_index;          // uid -> world index or -1
    std::vector<uint32_t> uid_free, uid_pending_free;
    int *d_idx_of_uid = nullptr;
    int d_uid_cap = 0;
    bool uid_dirty = true;

    DevWorld dw{};
    FxLaunchCtx cx{};
    int body_cap = 0;
    // serial-mode extras
    FxLexScratch lex{};
    SerialOrder so{};
    unsigned int so_cap = 0, lex_cap = 0;

    StepCountersDev *d_snapshot = nullptr;
    StepCountersDev *h_snapshot = nullptr;      // pinned
    cudaEvent_t snapshot_ready = nullptr;
    bool snapshot_pending = false;
    StepCountersDev last{};                     // last snapshot read back
    long long steps = 0;
    long long sticky_overflow = 0;
    long long manifold_hint = 0;

    // strip decomposition of one large world (fx_strip.cu, frxStrip* in fx_world.cpp)
    struct Strip {
        bool enabled = false;
        float cut_lo = 0.f, cut_hi = 0.f, margin = 0.f;
        unsigned int ghost_cap = 0;                  // ghost rows per side
        unsigned int uid_base = 0, uid_space = 0;    // this rank's slice of the global uid space
        unsigned int uid_slice = 0;                  // ids this strip may hand out (0: up to the end of the space)
        unsigned char *send_full = nullptr;          // my left-border bodies  -> left neighbour
        unsigned char *recv_full = nullptr;          // right neighbour's left-border bodies = my ghosts
        float4 *send_vel = nullptr, *recv_vel = nullptr;       // border velocities: -> left / <- right
        float4 *send_delta = nullptr, *recv_delta = nullptr;   // ghost velocity changes: -> right / <- left
        float4 *ghost_before = nullptr, *border_before = nullptr;
        unsigned int *ghost_split = nullptr, *border_split = nullptr;   // rows that are mass-split this step
        unsigned int *border_idx = nullptr;
        unsigned int *ghost_cnt = nullptr;           // device word: ghost rows in use this step
        frWorld *peer[2] = { nullptr, nullptr };     // in-process neighbours (left, right)
        void *nccl_comm = nullptr;                   // multi-process neighbours
        int rank = 0, nranks = 1;
        cudaEvent_t sent = nullptr;                  // my copies into the neighbours' receive buffers are done
        cudaEvent_t consumed = nullptr;              // I have read my receive buffers (neighbours may overwrite them)
        bool ghosts_primed = false;
        // Receive side: ONE allocation ("inbox") = flag words + the ghost-row buffer + two parities of the per-stage
        // buffers; recv_full / recv_vel / recv_delta above point into it.  With NCCL neighbours the inbox is also
        // exported through a CUDA IPC handle and the neighbours' inboxes are mapped here (peer-direct transport).
        unsigned char *inbox = nullptr;
        float4 *recv_vel_buf[2] = { nullptr, nullptr }, *recv_delta_buf[2] = { nullptr, nullptr };
        unsigned char *peer_inbox[2] = { nullptr, nullptr };   // [left, right] neighbours' inboxes, mapped (nullptr: NCCL / in-process)
        bool peer_direct = false;
        unsigned int epoch_rows = 0, epoch_stage = 0;          // exchange rounds so far (the same on every rank)
        unsigned int *push_done = nullptr;                     // [4] "last block" counters of the push kernels
        // ownership migration (frxStripMigrate): rows that change rank travel through these (NCCL neighbours only)
        unsigned char *mig_send[2] = { nullptr, nullptr }, *mig_recv[2] = { nullptr, nullptr };      // device, [left, right]
        unsigned char *mig_hsend[2] = { nullptr, nullptr }, *mig_hrecv[2] = { nullptr, nullptr };    // pinned host
        long long migrated_out = 0, migrated_in = 0;
    } strip;

    // CUDA graph of the fixed launch sequence of a step (production mode, no handler, no profiling).
    // The kernels take DevWorld by value, so the graph is keyed on that struct + the cooperative
    // grid size; it is (re)captured only when the key was stable for two consecutive steps.
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    DevWorld graph_key{};
    DevWorld last_key{};
    int graph_blocks = 0, last_blocks = 0;
    long long graph_launches = 0;
    long long graph_replays = 0;     // steps served by cudaGraphLaunch (frxGetGraphReplays)
    bool graph_valid = false;
    int use_graph = 1;

    // optional per-stage timing (frxSetProfiling): CUDA events on the world's stream around each
    // stage of every step; frxGetStageTimes sums the elapsed times
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;   // FX_STAGES + 1 marks per profiled step
    size_t prof_used = 0;
    cudaEvent_t user_events[4] = { nullptr, nullptr, nullptr, nullptr };
};

enum { FX_STAGE_BROADPHASE = 0, FX_STAGE_NARROWPHASE, FX_STAGE_CACHE, FX_STAGE_SOLVE, FX_STAGE_INTEGRATE, FX_STAGES };
extern long long g_fx_launches;   // kernels launched by this library (all worlds)

struct frSpatialHash_;

// ---- internal helpers shared between the host translation units ----
uint64_t fx_shape_epoch();
void fx_shape_touch();
frAABB fx_shape_aabb(const frShape *s, frTransform tx);
void fx_body_recompute_mass(frBody *b);
BodyRow *fx_row(frBody *b);                 // current dynamic row for writing (pulls + marks dirty)
const BodyRow fx_row_read(const frBody *b); // current dynamic row (pulls if the device is newer)
void fx_world_pull(frWorld *w);
void fx_world_pull_some(frWorld *w, unsigned want);   // want: FX_DIRTY_POSROT | _ANGLE | _VEL | _AABB
void fx_world_mark(frWorld *w, unsigned bits);
void fx_world_mark_body(frWorld *w, int index, unsigned bits);   // as fx_world_mark, for a setter on ONE body (its mirror row is up to date)
void fx_set_error(int code, const char *what);
bool fx_cuda_ok(cudaError_t e, const char *what);

// kernels' extra launchers (fx_bodies.cu)
void fx_launch_finish_step(const DevWorld &w, const FxLaunchCtx &cx, StepCountersDev *snapshot);
void fx_launch_single_body(const DevWorld &w, cudaStream_t s, int i, int op, float dt);
void fx_launch_single_pair_solve(const DevWorld &w, cudaStream_t s, CollisionPod *col, int op, float inv_dt);
void fx_launch_sincos(const float *a, int n, float *s, float *c, cudaStream_t st);
