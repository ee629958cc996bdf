// fx_world.cpp -- the world container, host<->device state management and the step driver.
//
// Replaces reference src/world.c: frCreateWorld ... frUpdateWorld (:107-285), the deferred add/remove
// queue (:147-167, :450-479) and the collision-handler sweeps (:209-214, :254-259).  frStepWorld
// launches the kernel sequence of fx_broadphase.cu / fx_contacts.cu / fx_solve.cu / fx_bodies.cu on
// the world's stream; nothing of the step runs on the host, and a world without a usable CUDA
// device refuses to step (sticky error + stderr), it never falls back to the CPU.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <new>

#include <cuda_profiler_api.h>

#include "fx_host.h"

// ---------------------------------------------------------------------------------------------
// errors, device selection
// ---------------------------------------------------------------------------------------------
long long g_fx_launches = 0;
// the sticky error is process-wide and may be set from any thread that steps a world (different worlds may be
// stepped from different threads): first error wins, code and text change together under a lock
static std::atomic<int> g_error{ 0 };
static std::mutex g_error_lock;
static char g_error_text[256] = "no error";
static thread_local int g_device = -1;

void fx_set_error(int code, const char *what) {
    {
        std::lock_guard<std::mutex> hold(g_error_lock);
        if (g_error.load(std::memory_order_relaxed) == 0) {
            std::snprintf(g_error_text, sizeof g_error_text, "%s", what);
            g_error.store(code, std::memory_order_release);
        }
    }
    std::fprintf(stderr, "[ferox-b200] ERROR %d: %s\n", code, what);
}
bool fx_cuda_ok(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return true;
    char buf[256];
    std::snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    fx_set_error(1000 + (int) e, buf);
    return false;
}
extern "C" int frxGetLastError(void) { return g_error.load(std::memory_order_acquire); }
extern "C" void frxClearLastError(void) {
    std::lock_guard<std::mutex> hold(g_error_lock);
    g_error.store(0, std::memory_order_release);
    g_error_text[0] = 0;
}
extern "C" const char *frxGetLastErrorString(void) {
    static thread_local char copy[256];
    std::lock_guard<std::mutex> hold(g_error_lock);
    std::snprintf(copy, sizeof copy, "%s", g_error_text);
    return copy;
}
extern "C" int frxSetDevice(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return -1;
    g_device = device;
    return 0;
}
static int current_device() {
    if (g_device >= 0) return g_device;
    const char *e = std::getenv("FEROX_DEVICE");
    return e ? std::atoi(e) : 0;
}

// ---------------------------------------------------------------------------------------------
// small allocation helpers
// ---------------------------------------------------------------------------------------------
template <typename T>
static bool dev_realloc(T *&p, size_t old_count, size_t new_count, cudaStream_t s, bool keep) {
    T *q = nullptr;
    if (!fx_cuda_ok(cudaMalloc((void **) &q, (new_count ? new_count : 1) * sizeof(T)), "cudaMalloc")) return false;
    if (keep && p && old_count)
        fx_cuda_ok(cudaMemcpyAsync(q, p, (old_count < new_count ? old_count : new_count) * sizeof(T), cudaMemcpyDeviceToDevice, s), "grow copy");
    if (p) { cudaStreamSynchronize(s); cudaFree(p); }
    p = q;
    return true;
}
template <typename T>
static bool pin_realloc(T *&p, size_t old_count, size_t new_count) {
    T *q = nullptr;
    if (!fx_cuda_ok(cudaMallocHost((void **) &q, (new_count ? new_count : 1) * sizeof(T)), "cudaMallocHost")) return false;
    std::memset(q, 0, (new_count ? new_count : 1) * sizeof(T));
    if (p && old_count) std::memcpy(q, p, (old_count < new_count ? old_count : new_count) * sizeof(T));
    if (p) cudaFreeHost(p);
    p = q;
    return true;
}
static unsigned int pow2_at_least(unsigned long long v) {
    unsigned long long p = 1;
    while (p < v) p <<= 1;
    return (unsigned int) (p > 0x80000000ull ? 0x80000000ull : p);
}

// ---------------------------------------------------------------------------------------------
// world life cycle
// ---------------------------------------------------------------------------------------------
static int default_capacity() {
    const char *e = std::getenv("FEROX_MAX_OBJECT_COUNT");
    int v = e ? std::atoi(e) : 0;
    return v > 0 ? v : FR_WORLD_MAX_OBJECT_COUNT;
}
static void ring_resize(frWorld *w, int capacity) {
    unsigned int len = pow2_at_least((unsigned long long) (capacity > 1 ? capacity : 1));
    w->ring.assign(len, WorldOp{ 0, nullptr });
    w->ring_head = w->ring_tail = 0;
}

static bool world_device_init(frWorld *w) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        fx_set_error(1, "no CUDA device: the ferox world step has no CPU fallback");
        return false;
    }
    w->device = current_device();
    if (w->device >= n) w->device = 0;
    if (!fx_cuda_ok(cudaSetDevice(w->device), "cudaSetDevice")) return false;
    if (!fx_cuda_ok(cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking), "cudaStreamCreate")) return false;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, w->device);
    w->cx.stream = w->stream;
    w->cx.max_blocks = sms * 8;
    w->cx.num_sms = sms;
    w->cx.coop_blocks = fx_query_coop_blocks(w->device);
    { const char *e = std::getenv("FEROX_SOLVE_BLOCKS"); w->cx.solve_blocks_cap = e ? std::atoi(e) : 0; }
    fx_init_kernel_attributes();     // function attributes are per device
    bool ok = fx_cuda_ok(cudaMalloc((void **) &w->d_snapshot, sizeof(StepCountersDev)), "cudaMalloc");
    ok = fx_cuda_ok(cudaMallocHost((void **) &w->h_snapshot, sizeof(StepCountersDev)), "cudaMallocHost") && ok;
    if (w->h_snapshot) std::memset(w->h_snapshot, 0, sizeof(StepCountersDev));
    ok = fx_cuda_ok(cudaEventCreateWithFlags(&w->snapshot_ready, cudaEventDisableTiming), "cudaEventCreate") && ok;
    ok = fx_cuda_ok(cudaMalloc((void **) &w->dw.ctr, sizeof(StepCountersDev)), "cudaMalloc") && ok;
    ok = fx_cuda_ok(cudaMalloc((void **) &w->dw.prev_count, 2 * sizeof(unsigned int)), "cudaMalloc") && ok;
    ok = fx_cuda_ok(cudaMalloc((void **) &w->dw.cur_flip, sizeof(int)), "cudaMalloc") && ok;
    ok = fx_cuda_ok(cudaMalloc((void **) &w->dw.color_off, (FX_MAX_COLORS + 1) * sizeof(unsigned int)), "cudaMalloc") && ok;
    ok = fx_cuda_ok(cudaMalloc((void **) &w->dw.color_scratch, 4 * FX_MAX_COLORS * sizeof(unsigned int)), "cudaMalloc") && ok;
    w->dw.solve_trace = nullptr;
    if (const char *e = std::getenv("FEROX_SOLVE_TRACE")) if (std::atoi(e) > 0) {
        ok = fx_cuda_ok(cudaMalloc((void **) &w->dw.solve_trace, FX_TRACE_SLOTS * sizeof(unsigned long long)), "cudaMalloc") && ok;
        if (w->dw.solve_trace) cudaMemset(w->dw.solve_trace, 0, FX_TRACE_SLOTS * sizeof(unsigned long long));
    }
    ok = fx_cuda_ok(cudaMalloc((void **) &w->d_scal, sizeof(StepScalars)), "cudaMalloc") && ok;
    if (!ok) return false;
    w->dw.scal = w->d_scal;
    cudaMemsetAsync(w->d_scal, 0, sizeof(StepScalars), w->stream);
    cudaMemsetAsync(w->dw.ctr, 0, sizeof(StepCountersDev), w->stream);
    cudaMemsetAsync(w->dw.prev_count, 0, 2 * sizeof(unsigned int), w->stream);
    cudaMemsetAsync(w->dw.cur_flip, 0, sizeof(int), w->stream);
    return fx_cuda_ok(cudaStreamSynchronize(w->stream), "world init");
}

extern "C" frWorld *frCreateWorld(frVector2 gravity, float cellSize) {
    frWorld *w = new (std::nothrow) frWorld_();
    if (!w) return nullptr;
    w->gravity = gravity;
    w->cell = cellSize;
    w->capacity = default_capacity();
    ring_resize(w, w->capacity);
    if (const char *g = std::getenv("FEROX_GRAPH")) w->use_graph = std::atoi(g);
    if (const char *g = std::getenv("FEROX_LOCAL_SOLVE")) w->use_local = std::atoi(g);
    // parity knob (read per world): contact-major solver records, k_emit_pairs<4>, k_narrowphase_tiled<4>,
    // k_carry_over<8> and k_solve_colored<2> -- the variants worlds of >= 400k manifolds run -- at any size
    if (const char *g = std::getenv("FEROX_FORCE_LARGE")) { w->force_large = std::atoi(g); if (w->force_large) w->use_local = 0; }
    const char *mode = std::getenv("FEROX_SOLVER");
    if (mode && std::strcmp(mode, "serial") == 0) w->solver_mode = FRX_SOLVER_SERIAL;
    w->device_ok = world_device_init(w);
    return w;
}

void fx_strip_release_comm(frWorld *w);
static void free_manifold_set(ManifoldSet &m) {
    cudaFree(m.key); cudaFree(m.body); cudaFree(m.hdr); cudaFree(m.meta); cudaFree(m.geo); cudaFree(m.imp); cudaFree(m.cid);
    std::memset(&m, 0, sizeof m);
}
static void world_free_device(frWorld *w) {
    if (w->stream) cudaStreamSynchronize(w->stream);
    DevWorld &d = w->dw;
    cudaFree(d.posrot); cudaFree(d.vel); cudaFree(d.mprop); cudaFree(d.force); cudaFree(d.aabb); cudaFree(d.angle);
    cudaFree(d.shape); cudaFree(d.uid); cudaFree(w->d_idx_of_uid); cudaFree(w->d_shapes);
    cudaFree(d.cell_cnt); cudaFree(d.cell_off); cudaFree(d.large_list);
    cudaFree(d.keys[0]); cudaFree(d.keys[1]); cudaFree(d.vals[0]); cudaFree(d.vals[1]); cudaFree(d.keys_hi[0]); cudaFree(d.keys_hi[1]);
    cudaFree(d.cand); cudaFree(d.cand_hit);
    free_manifold_set(d.man[0]); free_manifold_set(d.man[1]);
    cudaFree(d.touched); cudaFree(d.h_slot);
    cudaFree(d.sol.body); cudaFree(d.sol.nt); cudaFree(d.sol.n); cudaFree(d.sol.mat); cudaFree(d.sol.count_src);
    cudaFree(d.sol.r); cudaFree(d.sol.u); cudaFree(d.sol.k); cudaFree(d.sol.lam);
    cudaFree(d.order); cudaFree(d.color_off); cudaFree(d.body_mask); cudaFree(d.body_dup); cudaFree(d.winner); cudaFree(d.cbody);
    cudaFree(d.solve_trace); cudaFree(d.color_scratch); cudaFree(d.ctr); cudaFree(d.prev_count); cudaFree(d.cur_flip); cudaFree(w->d_scal);
    cudaFree(w->cx.zero_base);
    cudaFree(w->lex.keys[0]); cudaFree(w->lex.keys[1]); cudaFree(w->lex.vals[0]); cudaFree(w->lex.vals[1]);
    cudaFree(w->lex.tmp); cudaFree(w->lex.zero_base);
    cudaFree(w->so.r_key); cudaFree(w->so.r_len); cudaFree(w->so.t_key); cudaFree(w->so.t_val);
    for (int side = 0; side < 2; ++side)
        if (w->strip.peer_inbox[side]) cudaIpcCloseMemHandle(w->strip.peer_inbox[side]);
    cudaFree(w->strip.send_full); cudaFree(w->strip.inbox); cudaFree(w->strip.send_vel); cudaFree(w->strip.push_done);
    cudaFree(w->strip.send_delta); cudaFree(w->strip.ghost_before); cudaFree(w->strip.border_before);
    cudaFree(w->strip.ghost_split); cudaFree(w->strip.border_split);
    cudaFree(w->strip.border_idx); cudaFree(w->strip.ghost_cnt);
    for (int side = 0; side < 2; ++side) {
        cudaFree(w->strip.mig_send[side]); cudaFree(w->strip.mig_recv[side]);
        if (w->strip.mig_hsend[side]) cudaFreeHost(w->strip.mig_hsend[side]);
        if (w->strip.mig_hrecv[side]) cudaFreeHost(w->strip.mig_hrecv[side]);
    }
    fx_strip_release_comm(w);
    if (w->strip.sent) cudaEventDestroy(w->strip.sent);
    if (w->strip.consumed) cudaEventDestroy(w->strip.consumed);
    cudaFree(w->d_snapshot);
    if (w->h_snapshot) cudaFreeHost(w->h_snapshot);
    HostMirror &h = w->h;
    if (h.posrot) cudaFreeHost(h.posrot);
    if (h.vel) cudaFreeHost(h.vel);
    if (h.mprop) cudaFreeHost(h.mprop);
    if (h.force) cudaFreeHost(h.force);
    if (h.aabb) cudaFreeHost(h.aabb);
    if (h.angle) cudaFreeHost(h.angle);
    if (h.shape) cudaFreeHost(h.shape);
    if (h.uid) cudaFreeHost(h.uid);
    if (h.group) cudaFreeHost(h.group);
    cudaFree(w->d_group); cudaFree(w->d_grange); cudaFree(w->d_gcnt); cudaFree(w->d_goff); cudaFree(w->d_gfill); cudaFree(w->d_gorder);
    if (w->graph_valid) { cudaGraphExecDestroy(w->graph_exec); cudaGraphDestroy(w->graph); }
    if (w->snapshot_ready) cudaEventDestroy(w->snapshot_ready);
    for (cudaEvent_t e : w->prof_events) cudaEventDestroy(e);
    for (cudaEvent_t e : w->user_events) if (e) cudaEventDestroy(e);
    if (w->stream) cudaStreamDestroy(w->stream);
}

static void detach_body(frWorld *w, frBody *b);
static void attach_body(frWorld *w, frBody *b);
static void strip_map_neighbour_inboxes(frWorld *w);
void fx_strip_release_comm(frWorld *w);
static void fx_strip_abort_comm(frWorld *w);

void fx_release_manifold_host(frWorld *w);
void fx_release_bridge(frWorld *w);
extern "C" void frReleaseWorld(frWorld *w) {
    if (!w) return;
    if (w->device_ok) { cudaSetDevice(w->device); fx_world_pull(w); }
    // frees every body that is currently a member (src/world.c:124-125), not the queued ones
    for (frBody *b : w->bodies) std::free(b);
    w->bodies.clear();
    if (w->device_ok || w->stream) world_free_device(w);
    fx_release_manifold_host(w);
    fx_release_bridge(w);
    delete w;
}

extern "C" void frClearWorld(frWorld *w) {
    if (!w) return;
    fx_world_pull(w);
    // src/world.c:138-144: the bodies are neither freed nor is the contact cache cleared
    while (!w->bodies.empty()) detach_body(w, w->bodies.back());
    w->needs_sizing = true;
}

static bool ring_push(frWorld *w, WorldOp op) {
    const int mask = (int) w->ring.size() - 1;
    if (((w->ring_head + 1) & mask) == w->ring_tail) return false;
    w->ring[w->ring_head] = op;
    w->ring_head = (w->ring_head + 1) & mask;
    return true;
}
extern "C" bool frAddBodyToWorld(frWorld *w, frBody *b) {
    if (!w || !b || (int) w->bodies.size() >= w->capacity) return false;
    return ring_push(w, WorldOp{ FX_OP_ADD, b });
}
// ---- batched independent worlds (BASELINE config 4) ----
extern "C" frWorld *frxCreateWorldBatch(frVector2 gravity, float cellSize, int nWorlds) {
    if (nWorlds <= 0) return nullptr;
    frWorld *w = frCreateWorld(gravity, cellSize);
    if (w) w->n_groups = nWorlds;
    return w;
}
extern "C" int frxGetBatchWorldCount(const frWorld *w) { return w ? w->n_groups : 0; }
extern "C" bool frxAddBodyToBatch(frWorld *w, int worldIndex, frBody *b) {
    if (!w || !b || worldIndex < 0 || worldIndex >= w->n_groups) return false;
    if (b->world == nullptr) b->group = worldIndex;
    return frAddBodyToWorld(w, b);
}

extern "C" bool frRemoveBodyFromWorld(frWorld *w, frBody *b) {
    if (!w || !b) return false;
    return ring_push(w, WorldOp{ FX_OP_REMOVE, b });
}
extern "C" bool frIsBodyInWorld(const frWorld *w, frBody *b) {
    if (!w || !b) return false;
    return b->world == w && b->index >= 0 && b->index < (int) w->bodies.size() && w->bodies[b->index] == b;
}
extern "C" frBody *frGetBodyInWorld(const frWorld *w, int i) {
    if (!w || i < 0 || i >= (int) w->bodies.size()) return nullptr;
    return w->bodies[i];
}
extern "C" int frGetBodyCountInWorld(const frWorld *w) { return w ? (int) w->bodies.size() : 0; }
extern "C" frVector2 frGetWorldGravity(const frWorld *w) {
    frVector2 z = { 0.f, 0.f };
    return w ? w->gravity : z;
}
extern "C" void frSetWorldCollisionHandler(frWorld *w, frCollisionHandler handler) { if (w) w->handler = handler; }
extern "C" void frSetWorldGravity(frWorld *w, frVector2 gravity) { if (w) w->gravity = gravity; }

extern "C" bool frxSetWorldCapacity(frWorld *w, int maxObjectCount) {
    if (!w || maxObjectCount <= 0 || w->ring_head != w->ring_tail) return false;
    w->capacity = maxObjectCount;
    ring_resize(w, maxObjectCount);
    return true;
}
extern "C" int frxGetWorldCapacity(const frWorld *w) { return w ? w->capacity : 0; }
extern "C" void frxSetWorldSolverMode(frWorld *w, int mode) {
    if (w && (mode == FRX_SOLVER_COLORED || mode == FRX_SOLVER_SERIAL)) w->solver_mode = mode;
}
extern "C" int frxGetWorldSolverMode(const frWorld *w) { return w ? w->solver_mode : 0; }
extern "C" void frxSetWorldTimestamp(frWorld *w, float t) { if (w) w->timestamp = t; }
extern "C" float frxGetWorldTimestamp(const frWorld *w) { return w ? w->timestamp : 0.0f; }

// ---------------------------------------------------------------------------------------------
// host mirror <-> device
// ---------------------------------------------------------------------------------------------
static bool rebuild_zero_region(frWorld *w);

static bool ensure_body_capacity(frWorld *w, int need) {
    if (need <= w->body_cap) return true;
    int cap = w->body_cap ? w->body_cap : 1024;
    while (cap < need) cap *= 2;
    const size_t o = (size_t) w->body_cap, n = (size_t) cap;
    HostMirror &h = w->h;
    bool ok = pin_realloc(h.posrot, o, n) && pin_realloc(h.vel, o, n) && pin_realloc(h.mprop, o, n) &&
              pin_realloc(h.force, o, n) && pin_realloc(h.aabb, o, n) && pin_realloc(h.angle, o, n) &&
              pin_realloc(h.shape, o, n) && pin_realloc(h.uid, o, n) && pin_realloc(h.group, o, n);
    h.cap = cap;
    DevWorld &d = w->dw;
    cudaStream_t s = w->stream;
    ok = ok && dev_realloc(d.posrot, o, n, s, true) && dev_realloc(d.vel, o, n, s, true) && dev_realloc(d.mprop, o, n, s, true) &&
         dev_realloc(d.force, o, n, s, true) && dev_realloc(d.aabb, o, n, s, true) && dev_realloc(d.angle, o, n, s, true) &&
         dev_realloc(d.shape, o, n, s, true) && dev_realloc(d.uid, o, n, s, true) && dev_realloc(w->d_group, o, n, s, true) &&
         dev_realloc(d.cell_cnt, 0, n, s, false) && dev_realloc(d.cell_off, 0, n, s, false) &&
         dev_realloc(d.large_list, 0, n, s, false) && dev_realloc(d.body_mask, 0, n, s, false) && dev_realloc(d.body_dup, 0, n, s, false) &&
         dev_realloc(d.winner, 0, 2 * (size_t) n, s, false);
    d.cap_large = (unsigned int) n;
    w->body_cap = cap;
    // the look-back status words of the per-body sweep are sized by the body capacity
    if (ok && w->cx.zero_base != nullptr) ok = rebuild_zero_region(w);
    return ok;
}

void fx_world_mark(frWorld *w, unsigned bits) {
    w->dirty |= bits;
    if (bits & FX_DIRTY_SHAPE) w->shape_rows_stale = true;
    // A moved body can change the world's cell extent (hence the key width) but rarely the buffer sizes: no sizing
    // pre-pass (a blocking extra broadphase) -- the next step launches every sort pass instead, the buffers keep their
    // 2x head room, and a real overflow re-sizes the step after (read_snapshot).  Membership changes still re-size.
    if (bits & (FX_DIRTY_POSROT | FX_DIRTY_AABB | FX_DIRTY_SHAPE)) { w->extent_unknown = true; w->extent_mark_step = w->steps; }
}

// A setter on one body: the key width only becomes unknown when the body's new cell rectangle leaves the cell bounds
// the last completed step measured (a kinematic platform nudged every frame stays inside them: no spare sort passes,
// no change of the graph key).
static void read_snapshot(frWorld *w, bool wait);
void fx_world_mark_body(frWorld *w, int index, unsigned bits) {
    if ((bits & (FX_DIRTY_POSROT | FX_DIRTY_AABB | FX_DIRTY_SHAPE)) && w->snapshot_pending) read_snapshot(w, false);   // (a poll; it waits only when the host is FX_MAX_RUN_AHEAD steps ahead of the device)
    if ((bits & (FX_DIRTY_POSROT | FX_DIRTY_AABB | FX_DIRTY_SHAPE)) && !w->extent_unknown && w->last.key_bits != 0u &&
        index >= 0 && index < (int) w->bodies.size() && w->cell > 0.0f && !w->snapshot_pending) {
        const float4 a = w->h.aabb[index];
        const float inv = 1.0f / w->cell;
        const int x0 = cell_coord(a.x, inv), y0 = cell_coord(a.y, inv), x1 = cell_coord(a.x + a.z, inv), y1 = cell_coord(a.y + a.w, inv);
        const StepCountersDev &c = w->last;
        if (x0 >= c.cell_min_x && y0 >= c.cell_min_y && x1 <= c.cell_max_x && y1 <= c.cell_max_y) {
            w->dirty |= bits;
            if (bits & FX_DIRTY_SHAPE) w->shape_rows_stale = true;
            return;
        }
    }
    fx_world_mark(w, bits);
}

// bring the pinned host mirror up to date with the device (one bulk D2H per dynamic array)
// Downloads the arrays of the host mirror that are stale AND wanted.  A step makes all four state arrays stale
// (and the mirror's force accumulators, which the step cleared on the device).
void fx_world_pull_some(frWorld *w, unsigned want) {
    if (!w) return;
    if (w->device_newer) {
        w->device_newer = false;
        w->stale = FX_DIRTY_POSROT | FX_DIRTY_ANGLE | FX_DIRTY_VEL | FX_DIRTY_AABB;
    }
    if (w->force_clear_pending) {
        // only a COMPLETED step clears the accumulators (src/world.c:481-482); a pull from inside a pre-step
        // handler must still see -- and keep -- what the user applied before frStepWorld
        w->force_clear_pending = false;
        const size_t n0 = w->bodies.size();
        if (n0) std::memset(w->h.force, 0, n0 * sizeof(float4));
    }
    const unsigned get = w->stale & want;
    const size_t n = w->bodies.size();
    if (!get) return;
    w->stale &= ~get;
    if (n == 0 || !w->device_ok) return;
    cudaSetDevice(w->device);
    cudaStream_t s = w->stream;
    if (get & FX_DIRTY_POSROT) cudaMemcpyAsync(w->h.posrot, w->dw.posrot, n * sizeof(float4), cudaMemcpyDeviceToHost, s);
    if (get & FX_DIRTY_ANGLE) cudaMemcpyAsync(w->h.angle, w->dw.angle, n * sizeof(float), cudaMemcpyDeviceToHost, s);
    if (get & FX_DIRTY_VEL) cudaMemcpyAsync(w->h.vel, w->dw.vel, n * sizeof(float4), cudaMemcpyDeviceToHost, s);
    if (get & FX_DIRTY_AABB) cudaMemcpyAsync(w->h.aabb, w->dw.aabb, n * sizeof(float4), cudaMemcpyDeviceToHost, s);
    fx_cuda_ok(cudaStreamSynchronize(s), "state download");
}
void fx_world_pull(frWorld *w) { fx_world_pull_some(w, FX_DIRTY_POSROT | FX_DIRTY_ANGLE | FX_DIRTY_VEL | FX_DIRTY_AABB); }

// Per sub-world body index range (a plain world is one group).  The block-local solver stages a
// group's velocities in shared memory, which needs the group's bodies in ONE short index range.
static void update_group_ranges(frWorld *w) {
    const int n = (int) w->bodies.size(), G = w->n_groups;
    if (w->graph_valid) {                       // the captured sequence may pick another solver kernel now
        cudaGraphExecDestroy(w->graph_exec);
        cudaGraphDestroy(w->graph);
        w->graph_valid = false;
    }
    w->local_max_bodies = 0;
    if (G <= 1) {
        if (n > 0 && n <= FX_LOCAL_MAX_BODIES) w->local_max_bodies = n;
        return;
    }
    std::vector<int> lo((size_t) G, 0x7fffffff), hi((size_t) G, -1);
    for (int i = 0; i < n; ++i) {
        const int g = w->h.group[i];
        if (i < lo[g]) lo[g] = i;
        if (i > hi[g]) hi[g] = i;
    }
    w->grange_host.assign((size_t) G, make_int2(0, 0));
    int worst = 0;
    for (int g = 0; g < G; ++g) {
        if (hi[g] < 0) continue;
        w->grange_host[g] = make_int2(lo[g], hi[g] - lo[g] + 1);
        if (hi[g] - lo[g] + 1 > worst) worst = hi[g] - lo[g] + 1;
    }
    if (G > w->d_groups_cap) {
        dev_realloc(w->d_grange, 0, (size_t) G, w->stream, false);
        dev_realloc(w->d_gcnt, 0, (size_t) G, w->stream, false);
        dev_realloc(w->d_goff, 0, (size_t) G, w->stream, false);
        dev_realloc(w->d_gfill, 0, (size_t) G, w->stream, false);
        w->d_groups_cap = G;
    }
    cudaMemcpyAsync(w->d_grange, w->grange_host.data(), (size_t) G * sizeof(int2), cudaMemcpyHostToDevice, w->stream);
    cudaStreamSynchronize(w->stream);
    if (worst > 0 && worst <= FX_LOCAL_MAX_BODIES) w->local_max_bodies = worst;
}

static void world_push(frWorld *w) {
    const size_t n = w->bodies.size();
    cudaStream_t s = w->stream;
    const unsigned d = w->dirty;
    if (n && d) {
        if (d & FX_DIRTY_POSROT) cudaMemcpyAsync(w->dw.posrot, w->h.posrot, n * sizeof(float4), cudaMemcpyHostToDevice, s);
        if (d & FX_DIRTY_ANGLE) cudaMemcpyAsync(w->dw.angle, w->h.angle, n * sizeof(float), cudaMemcpyHostToDevice, s);
        if (d & FX_DIRTY_VEL) cudaMemcpyAsync(w->dw.vel, w->h.vel, n * sizeof(float4), cudaMemcpyHostToDevice, s);
        if (d & FX_DIRTY_MPROP) cudaMemcpyAsync(w->dw.mprop, w->h.mprop, n * sizeof(float4), cudaMemcpyHostToDevice, s);
        if (d & FX_DIRTY_FORCE) cudaMemcpyAsync(w->dw.force, w->h.force, n * sizeof(float4), cudaMemcpyHostToDevice, s);
        if (d & FX_DIRTY_AABB) cudaMemcpyAsync(w->dw.aabb, w->h.aabb, n * sizeof(float4), cudaMemcpyHostToDevice, s);
        if (d & FX_DIRTY_SHAPE) cudaMemcpyAsync(w->dw.shape, w->h.shape, n * sizeof(int2), cudaMemcpyHostToDevice, s);
        if (d & FX_DIRTY_UID) {
            cudaMemcpyAsync(w->dw.uid, w->h.uid, n * sizeof(unsigned int), cudaMemcpyHostToDevice, s);
            cudaMemcpyAsync(w->d_group, w->h.group, n * sizeof(int), cudaMemcpyHostToDevice, s);
        }
    }
    w->dirty = 0;
    if (w->uid_dirty) update_group_ranges(w);
    if (w->uid_dirty) {
        const size_t m = w->uid_index.size();
        // strip mode: uids are global (uid_base + local), the map covers the whole uid space
        const size_t need = w->strip.enabled ? (size_t) w->strip.uid_space : m;
        if ((long long) need > (long long) w->d_uid_cap) {
            long long cap = w->d_uid_cap ? w->d_uid_cap : 1024;
            while (cap < (long long) need) cap *= 2;
            dev_realloc(w->d_idx_of_uid, 0, (size_t) cap, s, false);
            w->d_uid_cap = (int) cap;
        }
        if (w->strip.enabled) {
            cudaMemsetAsync(w->d_idx_of_uid, 0xff, (size_t) w->d_uid_cap * sizeof(int), s);
            w->strip.ghosts_primed = false;          // ghost rows are re-registered by the next unpack
        }
        if (m) {
            // pageable source: the copy is staged synchronously, which is what we want here
            cudaMemcpyAsync(w->d_idx_of_uid + w->strip.uid_base, w->uid_index.data(), m * sizeof(int), cudaMemcpyHostToDevice, s);
            cudaStreamSynchronize(s);
        }
        w->uid_dirty = false;
    }
}

// ---- shape table ----
static void fill_shape_dev(ShapeDev &d, const frShape *s) {
    std::memset(&d, 0, sizeof d);
    d.type = (int) s->type;
    d.nv = (s->type == FR_SHAPE_POLYGON) ? s->vertices.count : 0;
    d.radius = s->radius;
    d.friction = s->material.friction;
    d.restitution = s->material.restitution;
    float far2 = 0.0f;
    for (int k = 0; k < d.nv && k < FX_MAX_VERTS; ++k) {
        d.verts[k] = v2(s->vertices.data[k].x, s->vertices.data[k].y);
        d.normals[k] = v2(s->normals.data[k].x, s->normals.data[k].y);
        const float l2 = d.verts[k].x * d.verts[k].x + d.verts[k].y * d.verts[k].y;
        if (l2 > far2) far2 = l2;
    }
    // conservative bounding radius for the narrow phase's early miss (fx_contacts.cu): never smaller than the shape
    const float reach = (s->type == FR_SHAPE_POLYGON) ? std::sqrt(far2) : std::fabs(s->radius);
    d.bound = reach * 1.0001f + 1e-5f;
}
static int shape_slot(frWorld *w, frShape *s) {
    if (!s) return -1;
    auto it = w->shape_index.find(s);
    if (it != w->shape_index.end()) return it->second;
    int idx = (int) w->shape_list.size();
    w->shape_list.push_back(s);
    w->shape_index.emplace(s, idx);
    w->shape_host.emplace_back();
    fill_shape_dev(w->shape_host.back(), s);
    w->shape_epoch_seen = 0;   // force an upload
    return idx;
}
static void refresh_body_shape_rows(frWorld *w) {
    for (size_t i = 0; i < w->bodies.size(); ++i) {
        frBody *b = w->bodies[i];
        int slot = shape_slot(w, b->shape);
        w->h.shape[i] = make_int2(slot, b->shape ? (int) b->shape->type : 0);
    }
    w->dirty |= FX_DIRTY_SHAPE;
}
static bool sync_shapes(frWorld *w) {
    if (w->shape_epoch_seen == fx_shape_epoch() && !w->shape_rows_stale) return true;
    // a shape was created/edited somewhere, or a body changed its shape: rebuild the rows.  (A body that merely enters
    // or leaves keeps the rows valid -- attach_body / detach_body maintain them -- so spawning and killing bodies every
    // frame does not walk the whole world: [measured] 1 M bodies, 50 ms per membership change before.)
    w->shape_rows_stale = false;
    refresh_body_shape_rows(w);
    for (size_t i = 0; i < w->shape_list.size(); ++i) fill_shape_dev(w->shape_host[i], w->shape_list[i]);
    const int n = (int) w->shape_list.size();
    if (n > w->d_shapes_cap) {
        int cap = w->d_shapes_cap ? w->d_shapes_cap : 64;
        while (cap < n) cap *= 2;
        if (!dev_realloc(w->d_shapes, 0, (size_t) cap, w->stream, false)) return false;
        w->d_shapes_cap = cap;
    }
    if (n) {
        cudaMemcpyAsync(w->d_shapes, w->shape_host.data(), (size_t) n * sizeof(ShapeDev), cudaMemcpyHostToDevice, w->stream);
        cudaStreamSynchronize(w->stream);
    }
    w->dw.shapes = w->d_shapes;
    w->dw.n_shapes = n;
    w->shape_epoch_seen = fx_shape_epoch();
    return true;
}

// ---- membership ----
// a strip hands out ids from its own slice of the global id space only: ids are contact-cache keys and ghost rows carry
// them across ranks, so two bodies must never share one
static bool strip_uid_available(const frWorld *w) {
    if (!w->strip.enabled || !w->uid_free.empty()) return true;
    const unsigned long long limit = w->strip.uid_slice ? w->strip.uid_slice
                                                         : (w->strip.uid_space > w->strip.uid_base ? w->strip.uid_space - w->strip.uid_base : 0u);
    return (unsigned long long) w->uid_index.size() < limit;
}
static void attach_body(frWorld *w, frBody *b) {
    fx_world_pull(w);
    const int i = (int) w->bodies.size();
    ensure_body_capacity(w, i + 1);
    uint32_t uid;
    if (!w->uid_free.empty()) { uid = w->uid_free.back(); w->uid_free.pop_back(); }
    else { uid = (uint32_t) w->uid_index.size(); w->uid_index.push_back(-1); }
    w->uid_index[uid] = i;
    w->uid_dirty = true;
    HostMirror &h = w->h;
    h.posrot[i] = b->loc.posrot; h.angle[i] = b->loc.angle; h.vel[i] = b->loc.vel; h.mprop[i] = b->loc.mprop;
    // a body enters in the queue drain of frPostStepWorld, and the force clear that follows the drain covers it
    // (src/world.c:450-482): whatever was applied to it while it was outside the world is gone by its first step
    h.force[i] = make_float4(0.f, 0.f, 0.f, 0.f); h.aabb[i] = b->loc.aabb;
    h.shape[i] = make_int2(shape_slot(w, b->shape), b->shape ? (int) b->shape->type : 0);
    h.uid[i] = w->strip.uid_base + uid;
    h.group[i] = b->group;
    b->world = w; b->index = i; b->uid = uid;
    w->bodies.push_back(b);
    w->dirty = FX_DIRTY_ALL;
    w->needs_sizing = true;
}
static void detach_body(frWorld *w, frBody *b) {
    fx_world_pull(w);
    const int i = b->index, last = (int) w->bodies.size() - 1;
    HostMirror &h = w->h;
    b->loc.posrot = h.posrot[i]; b->loc.angle = h.angle[i]; b->loc.vel = h.vel[i]; b->loc.mprop = h.mprop[i];
    b->loc.force = h.force[i]; b->loc.aabb = h.aabb[i];
    w->uid_index[b->uid] = -1;
    w->uid_pending_free.push_back(b->uid);
    w->uid_dirty = true;
    if (i != last) {   // swap-with-last renumbering, src/world.c:463-469
        frBody *m = w->bodies[last];
        h.posrot[i] = h.posrot[last]; h.angle[i] = h.angle[last]; h.vel[i] = h.vel[last]; h.mprop[i] = h.mprop[last];
        h.force[i] = h.force[last]; h.aabb[i] = h.aabb[last]; h.shape[i] = h.shape[last]; h.uid[i] = h.uid[last];
        h.group[i] = h.group[last];
        w->bodies[i] = m;
        m->index = i;
        w->uid_index[m->uid] = i;
    }
    w->bodies.pop_back();
    b->world = nullptr; b->index = -1;
    w->dirty = FX_DIRTY_ALL;
    w->needs_sizing = true;
}

// frPostStepWorld's queue drain, src/world.c:450-479
static void drain_queue(frWorld *w) {
    const int mask = (int) w->ring.size() - 1;
    while (w->ring_head != w->ring_tail) {
        WorldOp op = w->ring[w->ring_tail];
        w->ring_tail = (w->ring_tail + 1) & mask;
        if (op.id == FX_OP_ADD) {
            if (op.body->world != nullptr) continue;   // duplicate adds / foreign bodies are not supported
            if (!strip_uid_available(w)) { fx_set_error(8, "strip id slice exhausted: configure a larger uidSpace"); continue; }
            attach_body(w, op.body);
        } else if (op.id == FX_OP_REMOVE) {
            if (op.body->world == w) detach_body(w, op.body);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// work buffers
// ---------------------------------------------------------------------------------------------
static bool alloc_manifold_set(ManifoldSet &m, size_t old_cap, size_t cap, cudaStream_t s, bool keep) {
    return dev_realloc(m.key, old_cap, cap, s, keep) && dev_realloc(m.body, old_cap, cap, s, keep) &&
           dev_realloc(m.hdr, old_cap, cap, s, keep) && dev_realloc(m.meta, old_cap, cap, s, keep) &&
           dev_realloc(m.geo, 2 * old_cap, 2 * cap, s, keep) && dev_realloc(m.imp, 2 * old_cap, 2 * cap, s, keep) &&
           dev_realloc(m.cid, 2 * old_cap, 2 * cap, s, keep);
}

static bool rebuild_zero_region(frWorld *w) {
    // status words of the ordered sweeps + tickets + sort histograms + per-digit chain status
    DevWorld &d = w->dw;
    FxLaunchCtx &cx = w->cx;
    size_t tiles_body = (size_t) w->body_cap / 256 + 2, tiles_ent = (size_t) d.cap_entries / 256 + 2;
    size_t tiles_cand = (size_t) d.cap_cand / 128 + 2, tiles_man = (size_t) d.cap_manifolds / 128 + 2;
    size_t sort_tiles = (size_t) d.cap_entries / 2048 + 2;
    cx.status_words[0] = tiles_body; cx.status_words[1] = tiles_ent; cx.status_words[2] = tiles_cand; cx.status_words[3] = tiles_man;
    cx.digit_status_stride = sort_tiles * 256;
    size_t bytes = (tiles_body + tiles_ent + tiles_cand + tiles_man) * sizeof(unsigned long long) + 16 * sizeof(unsigned int) +
                   FX_SORT_MAX_PASSES * 256 * sizeof(unsigned int) + FX_SORT_MAX_PASSES * cx.digit_status_stride * sizeof(unsigned int);
    if (cx.zero_base) { cudaStreamSynchronize(w->stream); cudaFree(cx.zero_base); cx.zero_base = nullptr; }
    if (!fx_cuda_ok(cudaMalloc(&cx.zero_base, bytes), "cudaMalloc scratch")) return false;
    cx.zero_bytes = bytes;
    char *p = (char *) cx.zero_base;
    for (int k = 0; k < 4; ++k) { cx.status[k] = (unsigned long long *) p; p += cx.status_words[k] * sizeof(unsigned long long); }
    cx.tickets = (unsigned int *) p; p += 16 * sizeof(unsigned int);
    cx.sort_hist = (unsigned int *) p; p += FX_SORT_MAX_PASSES * 256 * sizeof(unsigned int);
    cx.digit_status = (unsigned int *) p;
    cx.zero_fixed_bytes = (size_t) (p - (char *) cx.zero_base);
    return true;
}

static bool ensure_work_capacity(frWorld *w, unsigned long long ent, unsigned long long cand, unsigned long long man) {
    DevWorld &d = w->dw;
    cudaStream_t s = w->stream;
    bool changed = false, ok = true;
    if (ent > d.cap_entries) {
        unsigned int cap = (unsigned int) (ent < 0x7fffffffull ? ent : 0x7fffffffull);
        ok = ok && dev_realloc(d.keys[0], 0, cap, s, false) && dev_realloc(d.keys[1], 0, cap, s, false) &&
             dev_realloc(d.vals[0], 0, cap, s, false) && dev_realloc(d.vals[1], 0, cap, s, false) &&
             dev_realloc(d.keys_hi[0], 0, cap, s, false) && dev_realloc(d.keys_hi[1], 0, cap, s, false);
        d.cap_entries = cap;
        changed = true;
    }
    if (cand > d.cap_cand) {
        unsigned int cap = (unsigned int) (cand < 0x7fffffffull ? cand : 0x7fffffffull);
        ok = ok && dev_realloc(d.cand, 0, cap, s, false) && dev_realloc(d.cand_hit, 0, cap, s, false);
        d.cap_cand = cap;
        changed = true;
    }
    if (man > d.cap_manifolds) {
        const size_t o = d.cap_manifolds;
        unsigned int cap = (unsigned int) (man < 0x3ffffffull ? man : 0x3ffffffull);   // 26-bit colouring tickets
        ok = ok && alloc_manifold_set(d.man[0], o, cap, s, true) && alloc_manifold_set(d.man[1], o, cap, s, true);
        ok = ok && dev_realloc(d.touched, 0, cap, s, false) && dev_realloc(d.cbody, 0, cap, s, false) &&
             dev_realloc(d.order, 0, cap, s, false);
        // (+ 8: the cluster solver's bulk copies round their end up to 16 bytes)
        ok = ok && dev_realloc(d.sol.body, 0, (size_t) cap + 8, s, false) && dev_realloc(d.sol.nt, 0, cap, s, false) && dev_realloc(d.sol.n, 0, cap, s, false) &&
             dev_realloc(d.sol.mat, 0, cap, s, false) && dev_realloc(d.sol.count_src, 0, (size_t) cap + 8, s, false) &&
             dev_realloc(d.sol.r, 0, 2 * (size_t) cap, s, false) && dev_realloc(d.sol.u, 0, 2 * (size_t) cap, s, false) &&
             dev_realloc(d.sol.k, 0, 2 * (size_t) cap, s, false) && dev_realloc(d.sol.lam, 0, 2 * (size_t) cap, s, false);
        // the hash map indexes the PREVIOUS dense array: rebuild it at the new size from that array
        unsigned int hcap = pow2_at_least(2ull * cap + 16);
        const unsigned int old_h = d.cap_hash;
        ok = ok && dev_realloc(d.h_slot, 0, hcap, s, false);
        d.cap_hash = hcap;
        d.cap_manifolds = cap;
        cudaMemsetAsync(d.h_slot, 0xff, (size_t) hcap * sizeof(unsigned int), s);
        if (old_h) fx_launch_rehash_prev(d, w->cx);
        changed = true;
    }
    if (changed) ok = ok && rebuild_zero_region(w);
    return ok;
}

static bool ensure_serial_scratch(frWorld *w) {
    DevWorld &d = w->dw;
    cudaStream_t s = w->stream;
    bool ok = true;
    if (w->lex_cap < d.cap_cand) {
        const unsigned int cap = d.cap_cand;
        FxLexScratch &x = w->lex;
        ok = dev_realloc(x.keys[0], 0, cap, s, false) && dev_realloc(x.keys[1], 0, cap, s, false) &&
             dev_realloc(x.vals[0], 0, cap, s, false) && dev_realloc(x.vals[1], 0, cap, s, false) &&
             dev_realloc(x.tmp, 0, cap, s, false);
        x.dstatus_stride = ((size_t) cap / 2048 + 2) * 256;
        size_t bytes = 2 * 4 * 256 * sizeof(unsigned int) + 2 * 4 * x.dstatus_stride * sizeof(unsigned int) + 8 * sizeof(unsigned int);
        if (x.zero_base) { cudaStreamSynchronize(s); cudaFree(x.zero_base); }
        ok = ok && fx_cuda_ok(cudaMalloc(&x.zero_base, bytes), "cudaMalloc lex");
        x.zero_bytes = bytes;
        char *p = (char *) x.zero_base;
        x.ghist = (unsigned int *) p; p += 2 * 4 * 256 * sizeof(unsigned int);
        x.dstatus = (unsigned int *) p; p += 2 * 4 * x.dstatus_stride * sizeof(unsigned int);
        x.tickets = (unsigned int *) p;
        w->lex_cap = cap;
    }
    if (w->so_cap < d.cap_manifolds) {
        // grow the reference-order key array, keeping its content, and rebuild its index
        const unsigned int cap = d.cap_manifolds, tcap = pow2_at_least(2ull * cap + 16);
        const bool first = (w->so.r_len == nullptr);
        ok = ok && dev_realloc(w->so.r_key, w->so_cap, cap, s, true);
        if (first) {
            ok = ok && dev_realloc(w->so.r_len, 0, 1, s, false);
            cudaMemsetAsync(w->so.r_len, 0, sizeof(unsigned int), s);
        }
        ok = ok && dev_realloc(w->so.t_key, 0, tcap, s, false) && dev_realloc(w->so.t_val, 0, tcap, s, false);
        w->so.t_cap = tcap;
        cudaMemsetAsync(w->so.t_key, 0xff, (size_t) tcap * sizeof(unsigned long long), s);
        if (!first) fx_launch_serial_reindex(w->cx, w->so);
        w->so_cap = cap;
    }
    return ok;
}

// ---------------------------------------------------------------------------------------------
// the step
// ---------------------------------------------------------------------------------------------
static void fill_dev_params(frWorld *w, float dt) {
    DevWorld &d = w->dw;
    d.n_owned = (int) w->bodies.size();
    d.n = d.n_owned + (w->strip.enabled ? (int) w->strip.ghost_cap : 0);
    d.cell = w->cell;
    d.inv_cell = (w->cell > 0.0f) ? 1.0f / w->cell : 0.0f;     // src/broad_phase.c:62-63
    {
        StepScalars sc{};
        sc.gx = w->gravity.x;
        sc.gy = w->gravity.y;
        sc.timestamp = w->timestamp;
        sc.dt = dt;
        sc.inv_dt = 1.0f / dt;                                  // src/world.c:242
        d.scal = w->d_scal;
        if (!w->scal_valid || std::memcmp(&sc, &w->scal_dev, sizeof sc) != 0) {
            fx_launch_set_scalars(w->d_scal, sc, w->stream);    // outside any captured graph: the graph stays valid
            w->scal_dev = sc;
            w->scal_valid = true;
        }
    }
    d.solver_mode = w->solver_mode;
    {
        // digit passes of the broadphase sort: all eight until a step has told us the key width, then what that width
        // plus one whole digit needs: the world's cell extent may grow 256-fold from one step to the next (free motion
        // grows it by a few cells) before plan_keys has to refuse a step for want of launched passes
        const unsigned int bits = w->last.key_bits;
        d.sort_launch = (bits == 0u || w->needs_sizing || w->extent_unknown) ? FX_SORT_MAX_PASSES : (int) ((bits + 8u + 7u) / 8u);
        if (d.sort_launch > FX_SORT_MAX_PASSES) d.sort_launch = FX_SORT_MAX_PASSES;
    }
    { static const int sq = [] { const char *e = std::getenv("FEROX_COLOR_SQUEEZE"); return e ? std::atoi(e) : 2; }(); d.color_squeeze = sq; }
    {
        // solver record layout (fx_solve.cu SOL_AT): contact-major once the sweeps are bandwidth-bound
        const bool contact_major = w->force_large || w->manifold_hint >= 400000;
        d.sol_stride = contact_major ? d.cap_manifolds : 1u;
        d.sol_mult = contact_major ? 1u : 2u;
        // lean records where the sweeps are bandwidth-bound: large worlds that k_solve_colored takes (fx_solve.cu load_nt)
        // (the block-local and the cluster solver read full records)
        d.sol_lean = (contact_major && !(w->use_local && w->local_max_bodies > 0) && (w->force_large || fx_cluster_size_for(d.n) == 0)) ? 1 : 0;
    }
    {
        static const int narrow = [] { const char *e = std::getenv("FEROX_NARROWPHASE"); return !e ? 0 : (std::strcmp(e, "group8") == 0 ? 1 : (std::strcmp(e, "lane") == 0 ? 2 : 0)); }();
        d.narrow_mode = narrow;
    }
    d.idx_of_uid = w->d_idx_of_uid;
    d.shapes = w->d_shapes;
    d.group = (w->n_groups > 1) ? w->d_group : nullptr;
    d.n_groups = w->n_groups;
    d.grange = w->d_grange;
    d.gcnt = w->d_gcnt; d.goff = w->d_goff; d.gfill = w->d_gfill;
    if (w->n_groups > 1 && w->d_gorder_cap < d.cap_manifolds) {
        dev_realloc(w->d_gorder, 0, (size_t) d.cap_manifolds, w->stream, false);
        w->d_gorder_cap = d.cap_manifolds;
    }
    d.gorder = w->d_gorder;
}

// The step is sized from the counters of the last snapshot the host has SEEN (sort passes for its key width plus one
// digit, kernel variants for its manifold count).  A caller that enqueues steps faster than the device runs them (a
// loop of frStepWorld, frxStepWorldN) would see none until the queue drains -- the one event always belongs to the newest
// step -- and a world whose extent explodes meanwhile would have every remaining step of the batch refused for want of
// sort passes ([measured] the config-3 column once it bursts: bench.py `config3`).  So the host never runs more than
// FX_MAX_RUN_AHEAD steps past the last snapshot it read: then it waits for the device once (a bubble of one step's
// launch latency per 16 steps).
#define FX_MAX_RUN_AHEAD 16
static void read_snapshot(frWorld *w, bool wait) {
    if (!w->snapshot_pending) return;
    if (!wait && w->steps - w->snapshot_read_step >= FX_MAX_RUN_AHEAD) wait = true;
    if (wait) cudaEventSynchronize(w->snapshot_ready);
    else if (cudaEventQuery(w->snapshot_ready) != cudaSuccess) return;
    w->snapshot_pending = false;
    w->snapshot_read_step = w->steps;
    w->last = *w->h_snapshot;
    w->manifold_hint = w->last.n_manifolds;
    // bit 128 (a body outside its strip, static ownership) is a condition the caller reads from the counters, not lost work
    w->sticky_overflow |= w->last.overflow;
    if (w->last.overflow & (1u | 2u | 4u | 16u)) w->needs_sizing = true;      // next step: every sort pass, buffers re-measured
    if (w->extent_unknown && w->snapshot_step >= w->extent_mark_step) w->extent_unknown = false;   // key width re-measured since
    if (w->last.overflow & ~128u) {
        char buf[260];
        std::snprintf(buf, sizeof buf,
                      "device buffer overflow in step %lld (mask 0x%x: 1 cell entries, 2 candidates, 4 manifolds, 8 colours, 16 "
                      "cell-key range, 32 ghost rows, 64 a neighbouring strip never delivered a message); work was dropped",
                      w->steps, w->last.overflow);
        fx_set_error(2, buf);
    }
}

// Run the broadphase only, read K and C back and size every downstream buffer with head room.
static bool sizing_prepass(frWorld *w, float dt) {
    DevWorld &d = w->dw;
    const unsigned long long n = w->bodies.size();
    if (!ensure_work_capacity(w, 8 * n + 4096, 16 * n + 4096, 4 * n + 4096)) return false;
    for (int attempt = 0; attempt < 6; ++attempt) {
        fill_dev_params(w, dt);
        fx_launch_begin_step(d, w->cx);
        if (w->cell > 0.0f && n > 0) fx_launch_broadphase(d, w->cx, false);
        StepCountersDev c;
        cudaMemcpyAsync(w->h_snapshot, d.ctr, sizeof c, cudaMemcpyDeviceToHost, w->stream);
        if (!fx_cuda_ok(cudaStreamSynchronize(w->stream), "sizing pre-pass")) return false;
        c = *w->h_snapshot;
        if (c.overflow & 16u) {
            fx_set_error(3, "non-finite world extent (> 2^62 broadphase cells), or a body covers > 2^24 cells: enlarge cellSize");
            return false;
        }
        // n_entries / n_cand are clamped to the capacity on overflow: grow and measure again
        if (c.overflow & 3u) {
            if (!ensure_work_capacity(w, (c.overflow & 1u) ? 4ull * d.cap_entries : 0, (c.overflow & 2u) ? 4ull * d.cap_cand : 0, 0)) return false;
            continue;
        }
        unsigned long long want_e = 2ull * c.n_entries + 4096, want_c = 2ull * c.n_cand + 4096;
        unsigned long long want_m = (unsigned long long) c.n_cand + 2ull * w->last.n_manifolds + 4096;
        if (want_m > want_c) want_m = want_c + w->last.n_manifolds;
        w->last.key_bits = c.key_bits;       // measured on this very state: the step sizes its sort passes from it
        return ensure_work_capacity(w, want_e, want_c, want_m);
    }
    fx_set_error(4, "could not size the broadphase buffers");
    return false;
}

static void grow_if_crowded(frWorld *w) {
    const StepCountersDev &c = w->last;
    DevWorld &d = w->dw;
    unsigned long long e = 0, k = 0, m = 0;
    if (2ull * c.n_entries > d.cap_entries) e = 2ull * d.cap_entries;
    if (2ull * c.n_cand > d.cap_cand) k = 2ull * d.cap_cand;
    if (2ull * c.n_manifolds > d.cap_manifolds) m = 2ull * d.cap_manifolds;
    if (e || k || m) ensure_work_capacity(w, e, k, m);
}

static void run_handler(frWorld *w, frCollisionEventFunc fn);
static void mh_mark_stale(frWorld *w);

static void prof_mark(frWorld *w) {
    if (!w->profiling) return;
    if (w->prof_used == w->prof_events.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        w->prof_events.push_back(e);
    }
    cudaEventRecord(w->prof_events[w->prof_used++], w->stream);
}

// the kernel / memset / memcpy sequence of one step on the world's stream (also what gets captured)
static void launch_step_sequence(frWorld *w, bool collide, bool has_handler, bool serial, long long hint) {
    DevWorld &d = w->dw;
    mh_mark_stale(w);
    prof_mark(w);
    fx_launch_begin_step(d, w->cx);
    if (serial && w->lex.zero_base) cudaMemsetAsync(w->lex.zero_base, 0, w->lex.zero_bytes, w->stream);
    if (collide) {
        fx_launch_broadphase(d, w->cx, !has_handler);
        if (serial) {
            int bits = 1;
            while ((1ll << bits) < (long long) w->bodies.size() + 1) ++bits;
            fx_launch_lexsort_candidates(d, w->cx, w->lex, bits);
        }
        prof_mark(w);
        fx_launch_narrowphase(d, w->cx);
    } else {
        if (!has_handler) fx_launch_integrate_velocity(d, w->cx);
        prof_mark(w);
    }
    prof_mark(w);
    fx_launch_cache_update(d, w->cx);
    if (serial) fx_launch_serial_order(d, w->cx, w->so);
    prof_mark(w);
    if (has_handler) {
        // The pre-step sweep runs BEFORE gravity and the velocity integration (src/world.c:209-220): forces applied
        // before frStepWorld are still in the mirror's accumulators (only a completed step clears them) and forces
        // the handler applies are uploaded with them.
        if (w->handler.preStep) run_handler(w, w->handler.preStep);      // src/world.c:209-214
        world_push(w);                                                    // setters called by the handler
        fx_launch_integrate_velocity(d, w->cx, true);                     // src/world.c:216-220; accumulators keep the gravity term
        if (w->handler.postStep && !w->bodies.empty())                    // frGetBodyForce inside postStep: what the reference holds then
            cudaMemcpyAsync(w->h.force, d.force, w->bodies.size() * sizeof(float4), cudaMemcpyDeviceToHost, w->stream);
    }
    if (serial) fx_launch_solve_serial_with(d, w->cx, w->so, FR_WORLD_ITERATION_COUNT);
    else if (w->use_local && w->local_max_bodies > 0) fx_launch_solve_local(d, w->cx, FR_WORLD_ITERATION_COUNT, w->local_max_bodies);
    else if (w->cluster_size > 0) fx_launch_solve_cluster(d, w->cx, FR_WORLD_ITERATION_COUNT, hint, w->cluster_size);
    else fx_launch_solve_colored(d, w->cx, FR_WORLD_ITERATION_COUNT, hint);
    if (!fx_cuda_ok(cudaGetLastError(), "solver launch")) return;        // no contact solve: do not integrate positions on top of it
    prof_mark(w);
    fx_launch_integrate_position(d, w->cx);
    prof_mark(w);
    w->device_newer = true;
    if (has_handler) {
        if (w->handler.postStep) run_handler(w, w->handler.postStep);    // src/world.c:254-259
        // frPostStepWorld clears every accumulator AFTER the sweep (src/world.c:481-482): forces applied inside a
        // post-step handler are wiped with the rest (the device copy was cleared by the position kernel)
        w->dirty &= ~(unsigned) FX_DIRTY_FORCE;
        if (!w->bodies.empty()) std::memset(w->h.force, 0, w->bodies.size() * sizeof(float4));
        world_push(w);
    } else {
        w->force_clear_pending = true;
    }
    fx_launch_finish_step(d, w->cx, w->d_snapshot);
    cudaMemcpyAsync(w->h_snapshot, w->d_snapshot, sizeof(StepCountersDev), cudaMemcpyDeviceToHost, w->stream);
}

static void step_once(frWorld *w, float dt) {
    DevWorld &d = w->dw;
    const bool has_handler = (w->handler.preStep != nullptr || w->handler.postStep != nullptr);
    const bool serial = (w->solver_mode == FRX_SOLVER_SERIAL);
    const bool collide = (w->cell > 0.0f) && !w->bodies.empty();   // NULL hash: no pairs at all
    if (w->strip.enabled) {
        fx_set_error(5, "frStepWorld on a strip of a decomposed world: use frxStripStep");
        return;
    }

    ensure_body_capacity(w, (int) w->bodies.size());
    sync_shapes(w);
    read_snapshot(w, false);
    if (w->needs_sizing) {
        world_push(w);
        if (!sizing_prepass(w, dt)) return;
        w->needs_sizing = false;
    } else {
        grow_if_crowded(w);
    }
    if (serial && !ensure_serial_scratch(w)) return;
    world_push(w);
    fill_dev_params(w, dt);

    long long hint = w->manifold_hint > (long long) w->bodies.size() ? w->manifold_hint : (long long) w->bodies.size();
    if (w->force_large && hint < 400000) hint = 400000;
    // mid-size worlds solve inside one thread-block cluster (velocities in distributed shared memory); small ones in
    // one CTA (block-local solver), large ones on the whole grid
    if (w->cluster_bodies != d.n) {
        static const int cluster_max = [] { const char *e = std::getenv("FEROX_CLUSTER_MAX_BODIES"); return e ? std::atoi(e) : 131072; }();
        w->cluster_bodies = d.n;
        w->cluster_size = (!w->force_large && d.n <= cluster_max) ? fx_cluster_size_for(d.n) : 0;   // 0 unless FEROX_CLUSTER=8|16
    }
    const int coop_blocks = fx_solve_colored_blocks(w->cx, hint) + (w->cluster_size << 20);   // (part of the graph key)
    // ---- fast path: replay the captured launch sequence ----
    const bool graphable = w->use_graph && !has_handler && !serial && !w->profiling;
    if (graphable) {
        const bool same_as_last = std::memcmp(&d, &w->last_key, sizeof d) == 0 && coop_blocks == w->last_blocks;
        w->last_key = d;
        w->last_blocks = coop_blocks;
        if (w->graph_valid && (std::memcmp(&d, &w->graph_key, sizeof d) != 0 || coop_blocks != w->graph_blocks)) {
            cudaGraphExecDestroy(w->graph_exec);
            cudaGraphDestroy(w->graph);
            w->graph_valid = false;
        }
        if (!w->graph_valid && same_as_last) {
            const long long before = g_fx_launches;
            if (cudaStreamBeginCapture(w->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                launch_step_sequence(w, collide, has_handler, serial, hint);
                cudaGraph_t g = nullptr;
                if (cudaStreamEndCapture(w->stream, &g) == cudaSuccess && g != nullptr &&
                    cudaGraphInstantiate(&w->graph_exec, g, 0) == cudaSuccess) {
                    w->graph = g;
                    w->graph_key = d;
                    w->graph_blocks = coop_blocks;
                    w->graph_launches = g_fx_launches - before;
                    w->graph_valid = true;
                } else {
                    if (g) cudaGraphDestroy(g);
                    (void) cudaGetLastError();
                    w->use_graph = 0;                     // capture unsupported here: direct launches from now on
                }
                g_fx_launches = before;
            }
        }
    }
    if (graphable && w->graph_valid) {
        cudaGraphLaunch(w->graph_exec, w->stream);
        g_fx_launches += w->graph_launches;
        w->graph_replays++;
        w->device_newer = true;
        w->force_clear_pending = true;
    } else {
        launch_step_sequence(w, collide, has_handler, serial, hint);
    }
    cudaEventRecord(w->snapshot_ready, w->stream);
    w->snapshot_pending = true;
    w->snapshot_step = w->steps;
    w->steps++;
    fx_cuda_ok(cudaGetLastError(), "step launch");

    // frPostStepWorld: queue drain (forces were cleared by the position kernel)
    if (!w->uid_pending_free.empty()) {
        // uids freed before this step are recyclable now: their cache entries were dropped by it
        w->uid_free.insert(w->uid_free.end(), w->uid_pending_free.begin(), w->uid_pending_free.end());
        w->uid_pending_free.clear();
    }
    if (w->ring_head != w->ring_tail) drain_queue(w);
}

extern "C" void frStepWorld(frWorld *w, float dt) {
    if (w == nullptr || dt <= 0.0f) return;
    if (!w->device_ok) {
        fx_set_error(1, "frStepWorld: world has no CUDA device (no CPU fallback)");
        return;
    }
    cudaSetDevice(w->device);
    step_once(w, dt);
}

extern "C" void frxStepWorldN(frWorld *w, float dt, int n) {
    for (int i = 0; i < n; ++i) frStepWorld(w, dt);
}

// src/world.c:268-285
extern "C" void frUpdateWorld(frWorld *w, float dt) {
    if (w == nullptr || dt <= 0.0f) return;
    float now = frGetCurrentTime();
    if (w->timestamp <= 0.0f) {
        w->timestamp = now;
        return;
    }
    float elapsed = now - w->timestamp;
    w->timestamp = now, w->accumulator += elapsed;
    for (; w->accumulator >= dt; w->accumulator -= dt) frStepWorld(w, dt);
}

extern "C" void frxSetProfiling(frWorld *w, int on) {
    if (!w) return;
    w->profiling = on != 0;
    w->prof_used = 0;
}
// sums, over the steps run since frxSetProfiling(w, 1), the device time of each stage in ms:
// out[0..4] = broadphase (+ velocity integration), narrow phase, cache update, solve, position
// integration; returns the number of profiled steps
extern "C" int frxGetStageTimes(frWorld *w, double *out) {
    if (!w || !w->device_ok || !out) return 0;
    cudaSetDevice(w->device);
    cudaStreamSynchronize(w->stream);
    const size_t per = FX_STAGES + 1, steps = w->prof_used / per;
    for (int k = 0; k < FX_STAGES; ++k) out[k] = 0.0;
    for (size_t s = 0; s < steps; ++s)
        for (int k = 0; k < FX_STAGES; ++k) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, w->prof_events[s * per + k], w->prof_events[s * per + k + 1]);
            out[k] += ms;
        }
    return (int) steps;
}
extern "C" void frxRecordEvent(frWorld *w, int slot) {
    if (!w || !w->device_ok || slot < 0 || slot >= 4) return;
    cudaSetDevice(w->device);
    if (!w->user_events[slot]) cudaEventCreate(&w->user_events[slot]);
    cudaEventRecord(w->user_events[slot], w->stream);
}
extern "C" float frxEventElapsedMs(frWorld *w, int from, int to) {
    if (!w || from < 0 || to < 0 || from >= 4 || to >= 4 || !w->user_events[from] || !w->user_events[to]) return -1.0f;
    cudaSetDevice(w->device);
    cudaEventSynchronize(w->user_events[to]);
    float ms = -1.0f;
    cudaEventElapsedTime(&ms, w->user_events[from], w->user_events[to]);
    return ms;
}
extern "C" long long frxGetLaunchCount(void) { return g_fx_launches; }
extern "C" long long frxGetGraphReplays(const frWorld *w) { return w ? w->graph_replays : 0; }
extern "C" int frxGetColorHistogram(frWorld *w, int *out64) {
    if (!w || !w->device_ok || !out64) return 0;
    cudaSetDevice(w->device);
    unsigned int off[FX_MAX_COLORS + 1];
    cudaMemcpyAsync(off, w->dw.color_off, sizeof off, cudaMemcpyDeviceToHost, w->stream);
    if (!fx_cuda_ok(cudaStreamSynchronize(w->stream), "colour histogram")) return 0;
    int n = 0;
    for (int c = 0; c < FX_MAX_COLORS; ++c) {
        out64[c] = (int) (off[c + 1] - off[c]);
        if (out64[c] > 0) n = c + 1;
    }
    return n;
}
// FEROX_SOLVE_TRACE=1 (read when the world is created): global-timer stamps (ns) the coloured solver left behind its
// stages in the LAST step -- [0] start, [1] body state cleared, [2] inherited colours validated, [3] colouring rounds
// done, [4] colours counted, [5] manifolds ordered by colour, [6] records built + warm start, [7 + i] sweep i done.
// Returns the number of slots written to `out` (0: tracing is off).
extern "C" int frxGetSolveTrace(frWorld *w, unsigned long long *out, int n) {
    if (!w || !w->device_ok || !out || !w->dw.solve_trace) return 0;
    if (n > FX_TRACE_SLOTS) n = FX_TRACE_SLOTS;
    cudaSetDevice(w->device);
    cudaMemcpyAsync(out, w->dw.solve_trace, (size_t) n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, w->stream);
    if (!fx_cuda_ok(cudaStreamSynchronize(w->stream), "solve trace")) return 0;
    return n;
}
extern "C" int frxGetWorldSolverKernel(const frWorld *w) {
    if (!w) return 0;
    if (w->use_local && w->local_max_bodies > 0) return -1;
    return w->cluster_size;
}
extern "C" void frxProfilerRange(int start) { if (start) cudaProfilerStart(); else cudaProfilerStop(); }

// n steps, each bracketed by its own pair of CUDA events on the world's stream; when flushL2 != 0 a
// 256 MiB scratch buffer is overwritten before every step (outside the bracket) so that no step
// starts with its inputs resident in the 126 MB L2.  outMs[n] receives the per-step device times.
static void *g_flush_buf[16] = { nullptr };
static const size_t FLUSH_BYTES = 256ull << 20;
extern "C" int frxTimedSteps(frWorld *w, float dt, int n, int flushL2, float *outMs) {
    if (!w || !w->device_ok || n <= 0 || dt <= 0.0f) return 0;
    cudaSetDevice(w->device);
    void *&buf = g_flush_buf[w->device & 15];
    if (flushL2 && !buf && !fx_cuda_ok(cudaMalloc(&buf, FLUSH_BYTES), "cudaMalloc L2 flush")) return 0;
    std::vector<cudaEvent_t> ev(2 * (size_t) n);
    for (auto &e : ev) cudaEventCreate(&e);
    for (int k = 0; k < n; ++k) {
        if (flushL2) cudaMemsetAsync(buf, k & 0xff, FLUSH_BYTES, w->stream);
        cudaEventRecord(ev[2 * k], w->stream);
        step_once(w, dt);
        cudaEventRecord(ev[2 * k + 1], w->stream);
    }
    fx_cuda_ok(cudaStreamSynchronize(w->stream), "timed steps");
    for (int k = 0; k < n; ++k) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev[2 * k], ev[2 * k + 1]);
        if (outMs) outMs[k] = ms;
    }
    for (auto &e : ev) cudaEventDestroy(e);
    return n;
}

extern "C" void frxSynchronizeWorld(frWorld *w) {
    if (!w || !w->device_ok) return;
    cudaSetDevice(w->device);
    fx_cuda_ok(cudaStreamSynchronize(w->stream), "synchronize");
    read_snapshot(w, true);
}

extern "C" void frxResetStepCounters(frWorld *w) {
    if (!w) return;
    frxSynchronizeWorld(w);
    w->sticky_overflow = 0;
}
extern "C" void frxGetStepCounters(frWorld *w, frxStepCounters *out) {
    if (!w || !out) return;
    frxSynchronizeWorld(w);
    out->bodies = (long long) w->bodies.size();
    out->cellEntries = w->last.n_entries;
    out->candidates = w->last.n_cand;
    out->manifolds = w->last.n_manifolds;
    out->contacts = w->last.n_contacts;
    out->colors = w->last.n_colors;
    out->colorRounds = w->last.uncolored;
    out->steps = w->steps;
    out->overflow = w->sticky_overflow;
}

// ---------------------------------------------------------------------------------------------
// strip decomposition of one large world (BASELINE config 5); device side in
// fx_strip.cu.  frxStripStep drives the strips this process holds through one step:
//   pack border bodies -> exchange -> unpack ghosts -> broadphase .. warm start
//   -> 10 x (one Gauss-Seidel sweep -> exchange border velocities) -> write-back, integrate
// Neighbours are either worlds of the same process (device-to-device / peer copies, used by the
// tests on one GPU and by a single process driving several GPUs) or other ranks (NCCL send/recv over
// NVLink, one group per exchange, stream-ordered, no host synchronisation anywhere in the step).
// Ownership is static in this version: bodies do not migrate between strips, so `margin` has to
// cover how far border bodies drift between two re-partitionings.
// ---------------------------------------------------------------------------------------------
#include <dlfcn.h>
namespace {
struct NcclApi {
    struct Id { char b[128]; };                      // ncclUniqueId: 128 opaque bytes, passed by value
    void *lib = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, Id, int) = nullptr;
    int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*CommAbort)(void *) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
bool nccl_load() {
    if (g_nccl.lib) return true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { fx_set_error(6, "libnccl.so.2 not found (strip decomposition over several processes needs NCCL)"); return false; }
    g_nccl.lib = h;
    *(void **) &g_nccl.GetUniqueId = dlsym(h, "ncclGetUniqueId");
    *(void **) &g_nccl.CommInitRank = dlsym(h, "ncclCommInitRank");
    *(void **) &g_nccl.Send = dlsym(h, "ncclSend");
    *(void **) &g_nccl.Recv = dlsym(h, "ncclRecv");
    *(void **) &g_nccl.GroupStart = dlsym(h, "ncclGroupStart");
    *(void **) &g_nccl.GroupEnd = dlsym(h, "ncclGroupEnd");
    *(void **) &g_nccl.CommDestroy = dlsym(h, "ncclCommDestroy");
    *(void **) &g_nccl.CommAbort = dlsym(h, "ncclCommAbort");
    *(void **) &g_nccl.AllReduce = dlsym(h, "ncclAllReduce");
    *(void **) &g_nccl.GetErrorString = dlsym(h, "ncclGetErrorString");
    return g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.Send && g_nccl.Recv && g_nccl.GroupStart && g_nccl.GroupEnd;
}
bool nccl_ok(int rc, const char *what) {
    if (rc == 0) return true;
    char buf[200];
    std::snprintf(buf, sizeof buf, "%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "NCCL error");
    fx_set_error(2000 + rc, buf);
    return false;
}
size_t full_bytes(const frWorld *w) { return FX_MSG_HEADER + (size_t) w->strip.ghost_cap * sizeof(GhostRow); }
size_t vel_bytes(const frWorld *w) { return (size_t) w->strip.ghost_cap * sizeof(float4); }
// inbox layout (identical on every strip of a decomposition: same ghost capacity)
enum { INBOX_FLAG_ROWS = 0, INBOX_FLAG_VEL = 1, INBOX_FLAG_DELTA = 2, INBOX_FLAGS_BYTES = 256 };
size_t inbox_off_full(const frWorld *) { return INBOX_FLAGS_BYTES; }
size_t inbox_off_vel(const frWorld *w, int parity) { return INBOX_FLAGS_BYTES + full_bytes(w) + (size_t) parity * vel_bytes(w); }
size_t inbox_off_delta(const frWorld *w, int parity) { return INBOX_FLAGS_BYTES + full_bytes(w) + (size_t) (2 + parity) * vel_bytes(w); }
size_t inbox_bytes(const frWorld *w) { return INBOX_FLAGS_BYTES + full_bytes(w) + 4 * vel_bytes(w); }
}   // namespace

extern "C" int frxStripGetNcclUniqueId(void *out128) {
    if (!out128 || !nccl_load()) return -1;
    return nccl_ok(g_nccl.GetUniqueId(out128), "ncclGetUniqueId") ? 0 : -1;
}

extern "C" int frxStripConfigure(frWorld *w, float cutLo, float cutHi, float margin, int ghostCapacity, unsigned int uidBase,
                                 unsigned int uidSpace) {
    if (!w || !w->device_ok || ghostCapacity <= 0 || !w->bodies.empty() || w->ring_head != w->ring_tail) return -1;
    cudaSetDevice(w->device);
    frWorld_::Strip &st = w->strip;
    st.enabled = true;
    st.cut_lo = cutLo; st.cut_hi = cutHi; st.margin = margin;
    st.ghost_cap = (unsigned int) ghostCapacity;
    st.uid_base = uidBase; st.uid_space = uidSpace;
    bool ok = fx_cuda_ok(cudaMalloc((void **) &st.send_full, full_bytes(w)), "cudaMalloc strip") &&
              fx_cuda_ok(cudaMalloc((void **) &st.inbox, inbox_bytes(w)), "cudaMalloc strip") &&
              fx_cuda_ok(cudaMalloc((void **) &st.push_done, 8 * sizeof(unsigned int)), "cudaMalloc strip") &&   // [0..3] counters, [4..7] an empty message header
              fx_cuda_ok(cudaMalloc((void **) &st.send_vel, vel_bytes(w)), "cudaMalloc strip") &&
              fx_cuda_ok(cudaMalloc((void **) &st.send_delta, vel_bytes(w)), "cudaMalloc strip") &&
              fx_cuda_ok(cudaMalloc((void **) &st.ghost_before, vel_bytes(w)), "cudaMalloc strip") &&
              fx_cuda_ok(cudaMalloc((void **) &st.border_before, vel_bytes(w)), "cudaMalloc strip") &&
              fx_cuda_ok(cudaMalloc((void **) &st.ghost_split, (size_t) st.ghost_cap * sizeof(unsigned int)), "cudaMalloc strip") &&
              fx_cuda_ok(cudaMalloc((void **) &st.border_split, (size_t) st.ghost_cap * sizeof(unsigned int)), "cudaMalloc strip") &&
              fx_cuda_ok(cudaMalloc((void **) &st.border_idx, (size_t) st.ghost_cap * sizeof(unsigned int)), "cudaMalloc strip") &&
              fx_cuda_ok(cudaMalloc((void **) &st.ghost_cnt, sizeof(unsigned int)), "cudaMalloc strip");
    if (ok) {
        st.recv_full = st.inbox + inbox_off_full(w);
        for (int par = 0; par < 2; ++par) {
            st.recv_vel_buf[par] = reinterpret_cast<float4 *>(st.inbox + inbox_off_vel(w, par));
            st.recv_delta_buf[par] = reinterpret_cast<float4 *>(st.inbox + inbox_off_delta(w, par));
        }
        st.recv_vel = st.recv_vel_buf[0];
        st.recv_delta = st.recv_delta_buf[0];
        cudaMemsetAsync(st.send_full, 0, full_bytes(w), w->stream);
        cudaMemsetAsync(st.inbox, 0, inbox_bytes(w), w->stream);        // flags 0, ghost count 0 until a neighbour writes
        cudaMemsetAsync(st.push_done, 0, 8 * sizeof(unsigned int), w->stream);
        cudaMemsetAsync(st.send_vel, 0, vel_bytes(w), w->stream);
        cudaMemsetAsync(st.send_delta, 0, vel_bytes(w), w->stream);
        cudaMemsetAsync(st.ghost_cnt, 0, sizeof(unsigned int), w->stream);
        cudaMemsetAsync(st.ghost_split, 0, (size_t) st.ghost_cap * sizeof(unsigned int), w->stream);
        cudaMemsetAsync(st.border_split, 0, (size_t) st.ghost_cap * sizeof(unsigned int), w->stream);
    }
    cudaEventCreateWithFlags(&st.sent, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&st.consumed, cudaEventDisableTiming);
    cudaEventRecord(st.consumed, w->stream);
    w->uid_dirty = true;
    w->use_graph = 0;
    w->use_local = 0;
    return ok ? 0 : -1;
}

// Ghost rows carry shape-table indices, so every strip must number its shapes identically: register
// the complete shape list, in the same order, on every strip before adding bodies.
extern "C" int frxStripRegisterShapes(frWorld *w, frShape *const *shapes, int n) {
    if (!w || !w->strip.enabled || !shapes || n < 0 || !w->shape_list.empty()) return -1;
    for (int i = 0; i < n; ++i)
        if (shapes[i] == nullptr || shape_slot(w, shapes[i]) != i) return -1;
    return 0;
}

extern "C" int frxStripLinkLocal(frWorld *left, frWorld *right) {
    if (!left || !right || !left->strip.enabled || !right->strip.enabled || left->strip.ghost_cap != right->strip.ghost_cap) return -1;
    left->strip.peer[1] = right;
    right->strip.peer[0] = left;
    if (right->strip.uid_base > left->strip.uid_base) left->strip.uid_slice = right->strip.uid_base - left->strip.uid_base;
    if (left->device != right->device) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, left->device, right->device);
        if (can) {
            cudaSetDevice(left->device); cudaDeviceEnablePeerAccess(right->device, 0);
            cudaSetDevice(right->device); cudaDeviceEnablePeerAccess(left->device, 0);
            (void) cudaGetLastError();
        }
    }
    return 0;
}

void fx_strip_release_comm(frWorld *w) {
    if (w->strip.nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(w->strip.nccl_comm);
    w->strip.nccl_comm = nullptr;
}
static void fx_strip_abort_comm(frWorld *w) {
    if (w->strip.nccl_comm && g_nccl.CommAbort) g_nccl.CommAbort(w->strip.nccl_comm);
    w->strip.nccl_comm = nullptr;
    fx_set_error(7, "strip step failed after its first exchange: the NCCL communicator was aborted");
}

// Peer-direct transport set-up (collective over the communicator): every rank exports its inbox through a CUDA IPC handle,
// trades handles with its neighbours over NCCL, maps theirs, and the ranks agree (all-reduce MIN) whether EVERY mapping
// worked; only then do the per-stage exchanges bypass NCCL.  FEROX_STRIP_PEER=0 keeps NCCL send/recv.
static void strip_map_neighbour_inboxes(frWorld *w) {
    frWorld_::Strip &st = w->strip;
    const char *env = std::getenv("FEROX_STRIP_PEER");
    const bool want = !(env && std::atoi(env) == 0) && g_nccl.AllReduce != nullptr;
    cudaStream_t s = w->stream;
    const int left = st.rank - 1, right = st.rank + 1;
    unsigned char *dbuf = nullptr;                       // [mine 64][from left 64][from right 64][ok 4][all ok 4]
    if (!fx_cuda_ok(cudaMalloc((void **) &dbuf, 256), "cudaMalloc ipc")) return;
    cudaIpcMemHandle_t mine, theirs[2];
    std::memset(&mine, 0, sizeof mine);
    int ok = (want && cudaIpcGetMemHandle(&mine, st.inbox) == cudaSuccess) ? 1 : 0;
    (void) cudaGetLastError();
    cudaMemcpyAsync(dbuf, &mine, 64, cudaMemcpyHostToDevice, s);
    g_nccl.GroupStart();
    if (left >= 0) { g_nccl.Send(dbuf, 64, 0, left, st.nccl_comm, s); g_nccl.Recv(dbuf + 64, 64, 0, left, st.nccl_comm, s); }
    if (right < st.nranks) { g_nccl.Send(dbuf, 64, 0, right, st.nccl_comm, s); g_nccl.Recv(dbuf + 128, 64, 0, right, st.nccl_comm, s); }
    nccl_ok(g_nccl.GroupEnd(), "ncclGroupEnd (ipc handles)");
    cudaMemcpyAsync(&theirs[0], dbuf + 64, 64, cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(&theirs[1], dbuf + 128, 64, cudaMemcpyDeviceToHost, s);
    if (!fx_cuda_ok(cudaStreamSynchronize(s), "ipc handle exchange")) ok = 0;
    void *mapped[2] = { nullptr, nullptr };
    if (ok) {
        if (left >= 0 && cudaIpcOpenMemHandle(&mapped[0], theirs[0], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = 0;
        if (right < st.nranks && cudaIpcOpenMemHandle(&mapped[1], theirs[1], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = 0;
        (void) cudaGetLastError();
    }
    int all_ok = 0;
    if (g_nccl.AllReduce != nullptr) {
        cudaMemcpyAsync(dbuf + 192, &ok, sizeof(int), cudaMemcpyHostToDevice, s);
        nccl_ok(g_nccl.AllReduce(dbuf + 192, dbuf + 196, 1, /* ncclInt32 */ 2, /* ncclMin */ 3, st.nccl_comm, s), "ncclAllReduce (peer-direct vote)");
        cudaMemcpyAsync(&all_ok, dbuf + 196, sizeof(int), cudaMemcpyDeviceToHost, s);
        if (!fx_cuda_ok(cudaStreamSynchronize(s), "peer-direct vote")) all_ok = 0;
    }
    cudaFree(dbuf);
    if (all_ok) {
        st.peer_inbox[0] = (unsigned char *) mapped[0];
        st.peer_inbox[1] = (unsigned char *) mapped[1];
        st.peer_direct = true;
    } else {
        for (int side = 0; side < 2; ++side)
            if (mapped[side]) cudaIpcCloseMemHandle(mapped[side]);
        (void) cudaGetLastError();
    }
}

extern "C" int frxStripGetTransport(const frWorld *w) {
    if (!w || !w->strip.enabled) return -1;
    return w->strip.nccl_comm ? (w->strip.peer_direct ? 2 : 1) : 0;
}

extern "C" int frxStripInitNccl(frWorld *w, int rank, int nranks, const void *uniqueId128) {
    if (!w || !w->strip.enabled || !uniqueId128 || !nccl_load()) return -1;
    cudaSetDevice(w->device);
    NcclApi::Id id;
    std::memcpy(&id, uniqueId128, sizeof id);
    w->strip.rank = rank;
    w->strip.nranks = nranks;
    if (!nccl_ok(g_nccl.CommInitRank(&w->strip.nccl_comm, nranks, id, rank), "ncclCommInitRank")) return -1;
    if (nranks > 0) w->strip.uid_slice = w->strip.uid_space / (unsigned int) nranks;     // equal slices, rank r starts at r * slice
    strip_map_neighbour_inboxes(w);
    return 0;
}

// one exchange round for all local strips.  kind 0: border bodies go LEFT.  kind 1: border
// velocities go LEFT and ghost velocity changes go RIGHT.
static void strip_exchange(frWorld **ws, int n, int kind) {
    for (int k = 0; k < n; ++k) {
        frWorld *w = ws[k];
        frWorld_::Strip &st = w->strip;
        cudaSetDevice(w->device);
        const size_t bytes = kind == 0 ? full_bytes(w) : vel_bytes(w);
        const void *to_left = kind == 0 ? (const void *) st.send_full : (const void *) st.send_vel;
        void *from_right = kind == 0 ? (void *) st.recv_full : (void *) st.recv_vel;
        if (st.nccl_comm && st.peer_direct) {
            // stores over NVLink straight into the neighbours' inboxes, count-sized, then their flags; my own readers wait
            // for the flags my neighbours raise here
            const bool has_left = st.rank > 0, has_right = st.rank + 1 < st.nranks;
            unsigned int *my_flags = reinterpret_cast<unsigned int *>(st.inbox);
            const unsigned int *border_count = reinterpret_cast<const unsigned int *>(st.send_full);
            if (kind == 0) {
                const unsigned int e = ++st.epoch_rows;
                if (has_left)
                    fx_launch_strip_push(w->cx, st.send_full, st.peer_inbox[0] + inbox_off_full(w), border_count, FX_MSG_HEADER / 16,
                                         (unsigned int) (sizeof(GhostRow) / 16), st.ghost_cap, st.push_done + 0,
                                         reinterpret_cast<unsigned int *>(st.peer_inbox[0]) + INBOX_FLAG_ROWS, e);
                if (has_right) fx_launch_strip_wait(w->dw, w->cx, my_flags + INBOX_FLAG_ROWS, e, nullptr, 0u, st.push_done + 3);
            } else {
                const unsigned int e = ++st.epoch_stage;
                const int par = (int) (e & 1u);
                st.recv_vel = st.recv_vel_buf[par];
                st.recv_delta = st.recv_delta_buf[par];
                if (has_left)
                    fx_launch_strip_push(w->cx, st.send_vel, st.peer_inbox[0] + inbox_off_vel(w, par), border_count, 0u, 1u, st.ghost_cap,
                                         st.push_done + 1, reinterpret_cast<unsigned int *>(st.peer_inbox[0]) + INBOX_FLAG_VEL, e);
                if (has_right)
                    fx_launch_strip_push(w->cx, st.send_delta, st.peer_inbox[1] + inbox_off_delta(w, par), st.ghost_cnt, 0u, 1u, st.ghost_cap,
                                         st.push_done + 2, reinterpret_cast<unsigned int *>(st.peer_inbox[1]) + INBOX_FLAG_DELTA, e);
                fx_launch_strip_wait(w->dw, w->cx, has_right ? my_flags + INBOX_FLAG_VEL : nullptr, e, has_left ? my_flags + INBOX_FLAG_DELTA : nullptr, e,
                                     st.push_done + 3);
            }
        } else if (st.nccl_comm) {
            const int left = st.rank - 1, right = st.rank + 1;
            g_nccl.GroupStart();
            if (left >= 0) g_nccl.Send(to_left, bytes, /* ncclChar */ 0, left, st.nccl_comm, w->stream);
            if (right < st.nranks) g_nccl.Recv(from_right, bytes, 0, right, st.nccl_comm, w->stream);
            if (kind == 1) {
                if (right < st.nranks) g_nccl.Send(st.send_delta, bytes, 0, right, st.nccl_comm, w->stream);
                if (left >= 0) g_nccl.Recv(st.recv_delta, bytes, 0, left, st.nccl_comm, w->stream);
            }
            nccl_ok(g_nccl.GroupEnd(), "ncclGroupEnd");
        } else {
            for (int side = 0; side < 2; ++side)
                if (st.peer[side]) cudaStreamWaitEvent(w->stream, st.peer[side]->strip.consumed, 0);   // previous message was read
            if (frWorld *p = st.peer[0])
                cudaMemcpyPeerAsync(kind == 0 ? (void *) p->strip.recv_full : (void *) p->strip.recv_vel, p->device, to_left, w->device, bytes, w->stream);
            if (kind == 1)
                if (frWorld *p = st.peer[1]) cudaMemcpyPeerAsync(p->strip.recv_delta, p->device, st.send_delta, w->device, bytes, w->stream);
            cudaEventRecord(st.sent, w->stream);
        }
    }
    // in-process neighbours: a strip may only read its receive buffers after both neighbours wrote them
    for (int k = 0; k < n; ++k) {
        frWorld *w = ws[k];
        if (w->strip.nccl_comm) continue;
        cudaSetDevice(w->device);
        for (int side = 0; side < 2; ++side)
            if (w->strip.peer[side]) cudaStreamWaitEvent(w->stream, w->strip.peer[side]->strip.sent, 0);
    }
}

// after a solver stage: trade velocity changes / fresh border velocities and reconcile both copies
static void strip_stage_sync(frWorld **ws, int n) {
    for (int k = 0; k < n; ++k) {
        frWorld *w = ws[k];
        cudaSetDevice(w->device);
        frWorld_::Strip &st = w->strip;
        fx_launch_strip_stage_pack(w->dw, w->cx, st.ghost_before, st.ghost_split, st.send_delta, st.ghost_cnt, st.send_full, st.border_idx,
                                   st.border_before, st.border_split, st.send_vel, st.ghost_cap);
    }
    strip_exchange(ws, n, 1);
    for (int k = 0; k < n; ++k) {
        frWorld *w = ws[k];
        cudaSetDevice(w->device);
        frWorld_::Strip &st = w->strip;
        fx_launch_strip_stage_apply(w->dw, w->cx, st.send_full, st.border_idx, st.recv_delta, st.recv_vel, st.send_delta, st.ghost_cnt,
                                    st.ghost_before, st.border_before, st.ghost_cap);
        cudaEventRecord(st.consumed, w->stream);
    }
}

extern "C" void frxStripStep(frWorld **ws, int n, float dt) {
    if (!ws || n <= 0 || dt <= 0.0f) return;
    for (int k = 0; k < n; ++k)
        if (!ws[k] || !ws[k]->strip.enabled || !ws[k]->device_ok) { fx_set_error(5, "frxStripStep: world is not a configured strip"); return; }
    // ---- host-side preparation, then pack the bodies at the left cut
    for (int k = 0; k < n; ++k) {
        frWorld *w = ws[k];
        cudaSetDevice(w->device);
        frWorld_::Strip &st = w->strip;
        const int rows = (int) w->bodies.size() + (int) st.ghost_cap;
        ensure_body_capacity(w, rows);
        sync_shapes(w);
        read_snapshot(w, false);
        world_push(w);
        fill_dev_params(w, dt);
        if (!st.ghosts_primed) {
            // make every ghost row inert once (and after every rebuild of the uid map): unpack an EMPTY message.  (Not the
            // receive buffer with its count cleared: with the peer-direct transport the right neighbour may already have
            // stored this step's rows there.)
            fx_launch_strip_unpack(w->dw, w->cx, reinterpret_cast<const unsigned char *>(st.push_done + 4), w->dw.n_owned, st.ghost_cap, st.ghost_cnt);
            st.ghosts_primed = true;
        }
        // (nothing has been exchanged yet: a failure here leaves no neighbour waiting inside a send/recv group)
        if (w->cx.zero_base == nullptr && !ensure_work_capacity(w, 8ull * rows + 4096, 16ull * rows + 4096, 4ull * rows + 4096)) return;
        fill_dev_params(w, dt);
        const bool has_left = st.peer[0] != nullptr || (st.nccl_comm && st.rank > 0);
        // without a left neighbour the cut is -inf: nothing is packed, but the kernel still checks for stray bodies
        fx_launch_strip_pack(w->dw, w->cx, 0, has_left ? st.cut_lo : -INFINITY, st.margin, st.cut_lo - st.margin, st.cut_hi + st.margin,
                             st.send_full, st.border_idx, st.ghost_cap);
    }
    strip_exchange(ws, n, 0);
    // ---- ghosts in, then everything up to and including the warm start
    for (int k = 0; k < n; ++k) {
        frWorld *w = ws[k];
        cudaSetDevice(w->device);
        frWorld_::Strip &st = w->strip;
        DevWorld &d = w->dw;
        fx_launch_strip_clear_ghost_map(d, w->cx, d.n_owned, (int) st.ghost_cap);
        fx_launch_strip_unpack(d, w->cx, st.recv_full, d.n_owned, st.ghost_cap, st.ghost_cnt);
        cudaEventRecord(st.consumed, w->stream);
        if (w->needs_sizing) {
            if (!sizing_prepass(w, dt)) {
                // the neighbours are already inside this step's exchanges: tear the communicator down so that they
                // fail loudly instead of waiting for a message that will never come
                fx_strip_abort_comm(w);
                return;
            }
            w->needs_sizing = false;
        } else {
            grow_if_crowded(w);
        }
        fill_dev_params(w, dt);
        fx_launch_begin_step(d, w->cx);
        if (w->cell > 0.0f && d.n > 0) {
            fx_launch_broadphase(d, w->cx, true);
            fx_launch_narrowphase(d, w->cx);
        } else {
            fx_launch_integrate_velocity(d, w->cx);
        }
        fx_launch_cache_update(d, w->cx);
        // tell the owner which of its border bodies got a manifold here (flags ride in the delta message)
        fx_launch_strip_mark_used(d, w->cx, st.send_delta, st.ghost_cap);
    }
    strip_exchange(ws, n, 1);
    for (int k = 0; k < n; ++k) {
        frWorld *w = ws[k];
        cudaSetDevice(w->device);
        frWorld_::Strip &st = w->strip;
        DevWorld &d = w->dw;
        // mass splitting: both copies of a body with a cross-cut manifold get twice the inverse mass / inertia
        fx_launch_strip_scale_mass(d, w->cx, st.send_delta, st.ghost_split, st.ghost_cnt, st.recv_delta, st.border_split, st.border_idx,
                                   st.send_full, st.ghost_cap, 2.0f, 1);
        cudaEventRecord(st.consumed, w->stream);
        const long long hint = w->manifold_hint > (long long) d.n ? w->manifold_hint : (long long) d.n;
        fx_launch_strip_stage_begin(d, w->cx, st.ghost_before, st.ghost_cnt, st.send_full, st.border_idx, st.border_before, st.ghost_cap);
        fx_launch_solve_colored(d, w->cx, 0, hint, 1 /* set-up + warm start */);
    }
    strip_stage_sync(ws, n);
    // ---- Gauss-Seidel sweeps, both copies of every border body reconciled after each
    for (int it = 0; it < FR_WORLD_ITERATION_COUNT; ++it) {
        for (int k = 0; k < n; ++k) {
            frWorld *w = ws[k];
            cudaSetDevice(w->device);
            const long long hint = w->manifold_hint > (long long) w->dw.n ? w->manifold_hint : (long long) w->dw.n;
            fx_launch_solve_colored(w->dw, w->cx, 1, hint, 2 /* one sweep */);      // "before" values: written by the last stage_apply
        }
        strip_stage_sync(ws, n);
    }
    // ---- write-back, position integration of the owned bodies, bookkeeping
    for (int k = 0; k < n; ++k) {
        frWorld *w = ws[k];
        cudaSetDevice(w->device);
        DevWorld &d = w->dw;
        const long long hint = w->manifold_hint > (long long) d.n ? w->manifold_hint : (long long) d.n;
        fx_launch_solve_colored(d, w->cx, 0, hint, 4 /* write-back */);
        frWorld_::Strip &st = w->strip;
        fx_launch_strip_scale_mass(d, w->cx, st.send_delta, st.ghost_split, st.ghost_cnt, st.recv_delta, st.border_split, st.border_idx,
                                   st.send_full, st.ghost_cap, 0.5f, 0);          // restore the true inverse masses
        fx_launch_integrate_position(d, w->cx);
        w->device_newer = true;
        w->force_clear_pending = true;
        fx_launch_finish_step(d, w->cx, w->d_snapshot);
        cudaMemcpyAsync(w->h_snapshot, w->d_snapshot, sizeof(StepCountersDev), cudaMemcpyDeviceToHost, w->stream);
        cudaEventRecord(w->snapshot_ready, w->stream);
        w->snapshot_pending = true;
        w->steps++;
        fx_cuda_ok(cudaGetLastError(), "strip step launch");
        if (!w->uid_pending_free.empty()) {
            // ids freed before this step are recyclable now: their cache entries were dropped by it
            w->uid_free.insert(w->uid_free.end(), w->uid_pending_free.begin(), w->uid_pending_free.end());
            w->uid_pending_free.clear();
        }
        if (w->ring_head != w->ring_tail) drain_queue(w);
    }
}


// ---------------------------------------------------------------------------------------------
// ownership migration between strips
// ---------------------------------------------------------------------------------------------
namespace {
struct MigRow {                         // one body on its way to another rank (112 B)
    float4 posrot, vel, force, aabb;
    float angle, gravity_scale;
    int type, shape_idx;
    unsigned int flags, pad[7];
};
static_assert(sizeof(MigRow) == 112, "MigRow layout");
size_t mig_bytes(const frWorld *w) { return FX_MSG_HEADER + (size_t) w->strip.ghost_cap * sizeof(MigRow); }

bool mig_reserve(frWorld *w) {
    frWorld_::Strip &st = w->strip;
    if (st.mig_send[0]) return true;
    bool ok = true;
    for (int side = 0; side < 2 && ok; ++side)
        ok = fx_cuda_ok(cudaMalloc((void **) &st.mig_send[side], mig_bytes(w)), "cudaMalloc migrate") &&
             fx_cuda_ok(cudaMalloc((void **) &st.mig_recv[side], mig_bytes(w)), "cudaMalloc migrate") &&
             fx_cuda_ok(cudaMallocHost((void **) &st.mig_hsend[side], mig_bytes(w)), "cudaMallocHost migrate") &&
             fx_cuda_ok(cudaMallocHost((void **) &st.mig_hrecv[side], mig_bytes(w)), "cudaMallocHost migrate");
    return ok;
}
}   // namespace

extern "C" int frxStripMigrate(frWorld **ws, int n) {
    if (!ws || n <= 0) return -1;
    for (int k = 0; k < n; ++k)
        if (!ws[k] || !ws[k]->strip.enabled || !ws[k]->device_ok) { fx_set_error(5, "frxStripMigrate: world is not a configured strip"); return -1; }
    long long moved = 0;
    // ---- who leaves: positions on the host, emigrants per side
    std::vector<std::vector<frBody *>> out_left((size_t) n), out_right((size_t) n);
    for (int k = 0; k < n; ++k) {
        frWorld *w = ws[k];
        cudaSetDevice(w->device);
        fx_world_pull_some(w, FX_DIRTY_POSROT);          // the positions decide; the rest follows only if somebody leaves
        const frWorld_::Strip &st = w->strip;
        const float h = 0.25f * st.margin;
        const bool has_left = st.peer[0] != nullptr || (st.nccl_comm && st.rank > 0);
        const bool has_right = st.peer[1] != nullptr || (st.nccl_comm && st.rank + 1 < st.nranks);
        // (the scan reads the packed mirror only -- positions and the type bits -- not a million body objects)
        const float4 *posrot = w->h.posrot, *mprop = w->h.mprop;
        const float lo = st.cut_lo - h, hi = st.cut_hi + h;
        for (size_t i = 0, nb = w->bodies.size(); i < nb; ++i) {
            const float px = posrot[i].x;
            if (!((has_left && px < lo) || (has_right && px >= hi))) continue;
            uint32_t bits;
            std::memcpy(&bits, &mprop[i].w, sizeof bits);
            if ((bits & 0xffu) == (uint32_t) FR_BODY_STATIC) continue;   // static bodies are replicated, never exchanged
            if (has_left && px < lo) out_left[(size_t) k].push_back(w->bodies[i]);
            else out_right[(size_t) k].push_back(w->bodies[i]);
        }
    }
    // ---- in-process neighbours: the body itself changes world
    for (int k = 0; k < n; ++k) {
        frWorld *w = ws[k];
        for (int side = 0; side < 2; ++side) {
            frWorld *p = w->strip.peer[side];
            if (!p) continue;
            std::vector<frBody *> &list = side == 0 ? out_left[(size_t) k] : out_right[(size_t) k];
            for (frBody *b : list) {
                if ((int) p->bodies.size() >= p->capacity) p->capacity = (int) p->bodies.size() + 1;   // a move inside one world, not an add
                if (!strip_uid_available(p)) { fx_set_error(8, "strip id slice exhausted: configure a larger uidSpace"); break; }
                // every strip registered the same shape list in the same order (frxStripRegisterShapes): the body takes
                // the neighbour's copy of its shape, so that shape-table indices keep meaning the same thing everywhere
                auto it = w->shape_index.find(b->shape);
                if (b->shape != nullptr && (it == w->shape_index.end() || it->second >= (int) p->shape_list.size())) continue;
                cudaSetDevice(w->device);
                detach_body(w, b);
                if (b->shape != nullptr) b->shape = p->shape_list[(size_t) it->second];
                cudaSetDevice(p->device);
                attach_body(p, b);
                // (attach_body clears the accumulators as a queue drain would; a migrant keeps what it carried)
                p->h.force[b->index] = b->loc.force;
                ++moved;
                ++w->strip.migrated_out;
                ++p->strip.migrated_in;
            }
            list.clear();
        }
    }
    // ---- NCCL neighbours: rows travel, bodies are re-created on arrival
    for (int k = 0; k < n; ++k) {
        frWorld *w = ws[k];
        frWorld_::Strip &st = w->strip;
        if (!st.nccl_comm) continue;
        cudaSetDevice(w->device);
        if (!mig_reserve(w)) return -1;
        if (!out_left[(size_t) k].empty() || !out_right[(size_t) k].empty()) fx_world_pull(w);   // the emigrants' full rows
        cudaStream_t s = w->stream;
        const int left = st.rank - 1, right = st.rank + 1;
        std::vector<frBody *> sent[2];
        for (int side = 0; side < 2; ++side) {
            std::vector<frBody *> &list = side == 0 ? out_left[(size_t) k] : out_right[(size_t) k];
            unsigned int cnt = 0;
            MigRow *rows = reinterpret_cast<MigRow *>(st.mig_hsend[side] + FX_MSG_HEADER);
            for (frBody *b : list) {
                if (cnt >= st.ghost_cap) break;                 // the rest goes with the next call
                const int i = b->index;
                MigRow r;
                std::memset(&r, 0, sizeof r);
                r.posrot = w->h.posrot[i]; r.vel = w->h.vel[i]; r.force = w->h.force[i]; r.aabb = w->h.aabb[i];
                r.angle = w->h.angle[i];
                r.gravity_scale = b->gravity_scale;
                r.type = (int) b->type;
                r.flags = (unsigned int) b->flags;
                r.shape_idx = w->h.shape[i].x;
                rows[cnt++] = r;
                sent[side].push_back(b);
            }
            std::memset(st.mig_hsend[side], 0, FX_MSG_HEADER);
            *reinterpret_cast<unsigned int *>(st.mig_hsend[side]) = cnt;
            cudaMemcpyAsync(st.mig_send[side], st.mig_hsend[side], FX_MSG_HEADER + (size_t) cnt * sizeof(MigRow), cudaMemcpyHostToDevice, s);
        }
        g_nccl.GroupStart();
        if (left >= 0) {
            g_nccl.Send(st.mig_send[0], mig_bytes(w), 0, left, st.nccl_comm, s);
            g_nccl.Recv(st.mig_recv[0], mig_bytes(w), 0, left, st.nccl_comm, s);
        }
        if (right < st.nranks) {
            g_nccl.Send(st.mig_send[1], mig_bytes(w), 0, right, st.nccl_comm, s);
            g_nccl.Recv(st.mig_recv[1], mig_bytes(w), 0, right, st.nccl_comm, s);
        }
        if (!nccl_ok(g_nccl.GroupEnd(), "ncclGroupEnd (migrate)")) return -1;
        for (int side = 0; side < 2; ++side)
            if ((side == 0 && left >= 0) || (side == 1 && right < st.nranks))
                cudaMemcpyAsync(st.mig_hrecv[side], st.mig_recv[side], mig_bytes(w), cudaMemcpyDeviceToHost, s);
        if (!fx_cuda_ok(cudaStreamSynchronize(s), "strip migration exchange")) return -1;
        // the emigrants leave (their handles are released: the body now lives on another rank) ...
        for (int side = 0; side < 2; ++side)
            for (frBody *b : sent[side]) {
                detach_body(w, b);
                std::free(b);
                ++moved;
                ++st.migrated_out;
            }
        // ... and the immigrants arrive
        for (int side = 0; side < 2; ++side) {
            if (!((side == 0 && left >= 0) || (side == 1 && right < st.nranks))) continue;
            unsigned int cnt = *reinterpret_cast<const unsigned int *>(st.mig_hrecv[side]);
            if (cnt > st.ghost_cap) cnt = st.ghost_cap;
            const MigRow *rows = reinterpret_cast<const MigRow *>(st.mig_hrecv[side] + FX_MSG_HEADER);
            for (unsigned int r = 0; r < cnt; ++r) {
                const MigRow &m = rows[r];
                frShape *shape = (m.shape_idx >= 0 && m.shape_idx < (int) w->shape_list.size()) ? w->shape_list[(size_t) m.shape_idx] : nullptr;
                frVector2 pos = { m.posrot.x, m.posrot.y };
                frBody *b = frCreateBodyFromShape((frBodyType) m.type, pos, shape);
                if (!b) continue;
                if (m.flags) frSetBodyFlags(b, (frBodyFlags) m.flags);
                if (m.gravity_scale != 1.0f) frSetBodyGravityScale(b, m.gravity_scale);
                b->loc.posrot = m.posrot;
                b->loc.angle = m.angle;
                b->loc.vel = make_float4(m.vel.x, m.vel.y, m.vel.z, b->inv_mass);
                b->loc.aabb = m.aabb;
                b->loc.force = m.force;
                if ((int) w->bodies.size() >= w->capacity) w->capacity = (int) w->bodies.size() + 1;       // a move inside one world, not an add
                if (!strip_uid_available(w)) { fx_set_error(8, "strip id slice exhausted: configure a larger uidSpace (an immigrant was dropped)"); std::free(b); continue; }
                attach_body(w, b);
                w->h.force[b->index] = m.force;
                ++st.migrated_in;
            }
        }
    }
    return (int) moved;
}

// ---------------------------------------------------------------------------------------------
// manifold download / upload (introspection and the collision-handler bridge)
// ---------------------------------------------------------------------------------------------
// grow-only pinned host array (D2H into pageable std::vector storage ran at a few GB/s and paid a page-fault
// storm per step: [measured] 15 ms per 65k-body step with a no-op handler installed)
template <typename T>
struct PinnedBuf {
    T *p = nullptr;
    size_t cap = 0, n = 0;
    void resize(size_t k) {
        if (k > cap) {
            if (p) cudaFreeHost(p);
            cap = k + k / 4 + 1024;
            if (cudaMallocHost((void **) &p, cap * sizeof(T)) != cudaSuccess) { p = nullptr; cap = 0; k = 0; }
        }
        n = k;
    }
    T *data() { return p; }
    const T *data() const { return p; }
    T &operator[](size_t i) { return p[i]; }
    const T &operator[](size_t i) const { return p[i]; }
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
};
struct ManifoldHost {
    PinnedBuf<uint2> key;
    PinnedBuf<int2> body, meta;
    PinnedBuf<float4> hdr, geo, imp;
    PinnedBuf<int> cid;
    std::vector<unsigned long long> order_keys;
    std::vector<unsigned int> order;
    unsigned int count = 0;
    int set = 0;
    bool fresh = false;          // holds this step's live set (only the impulses can have changed since)
};
static ManifoldHost &world_mh(frWorld *w) {
    if (!w->mh_cache) w->mh_cache = new ManifoldHost();
    return *static_cast<ManifoldHost *>(w->mh_cache);
}
static void mh_mark_stale(frWorld *w) {
    if (w->mh_cache) static_cast<ManifoldHost *>(w->mh_cache)->fresh = false;
}
void fx_release_manifold_host(frWorld *w) {
    delete static_cast<ManifoldHost *>(w->mh_cache);
    w->mh_cache = nullptr;
}

// `after_finish`: the step has completed (cur_flip already toggled), so the live set is flip ^ 1
static bool download_manifolds(frWorld *w, ManifoldHost &mh, bool after_finish) {
    DevWorld &d = w->dw;
    cudaStream_t s = w->stream;
    int flip = 0;
    unsigned int m = 0;
    cudaMemcpyAsync(&flip, d.cur_flip, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (after_finish) cudaMemcpyAsync(&m, d.prev_count, sizeof(unsigned int), cudaMemcpyDeviceToHost, s);
    else cudaMemcpyAsync(&m, &d.ctr->n_manifolds, sizeof(unsigned int), cudaMemcpyDeviceToHost, s);
    if (!fx_cuda_ok(cudaStreamSynchronize(s), "manifold count")) return false;
    const int set = after_finish ? (flip ^ 1) : flip;
    if (mh.fresh && mh.set == set && mh.count == m && m > 0) {
        // second sweep of the same step (post-step handler): only the solver's impulses are new
        cudaMemcpyAsync(mh.imp.data(), d.man[set].imp, 2 * (size_t) m * sizeof(float4), cudaMemcpyDeviceToHost, s);
        return fx_cuda_ok(cudaStreamSynchronize(s), "manifold download");
    }
    mh.set = set;
    mh.count = m;
    if (m == 0 || d.cap_manifolds == 0) { mh.count = 0; return true; }
    const ManifoldSet &ms = d.man[set];
    mh.key.resize(m); mh.body.resize(m); mh.meta.resize(m); mh.hdr.resize(m);
    mh.geo.resize(2 * (size_t) m); mh.imp.resize(2 * (size_t) m); mh.cid.resize(2 * (size_t) m);
    cudaMemcpyAsync(mh.key.data(), ms.key, m * sizeof(uint2), cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(mh.body.data(), ms.body, m * sizeof(int2), cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(mh.meta.data(), ms.meta, m * sizeof(int2), cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(mh.hdr.data(), ms.hdr, m * sizeof(float4), cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(mh.geo.data(), ms.geo, 2 * (size_t) m * sizeof(float4), cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(mh.imp.data(), ms.imp, 2 * (size_t) m * sizeof(float4), cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(mh.cid.data(), ms.cid, 2 * (size_t) m * sizeof(int), cudaMemcpyDeviceToHost, s);
    mh.order_keys.clear();
    if (w->solver_mode == FRX_SOLVER_SERIAL && w->so.r_len) {
        unsigned int len = 0;
        cudaMemcpyAsync(&len, w->so.r_len, sizeof(unsigned int), cudaMemcpyDeviceToHost, s);
        cudaStreamSynchronize(s);
        mh.order_keys.resize(len);
        if (len) cudaMemcpyAsync(mh.order_keys.data(), w->so.r_key, len * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s);
    }
    return fx_cuda_ok(cudaStreamSynchronize(s), "manifold download");
}

static void manifold_to_pod(const ManifoldHost &mh, unsigned int i, frCollision *c) {
    std::memset(c, 0, sizeof *c);
    c->count = mh.meta[i].x;
    c->direction.x = mh.hdr[i].x; c->direction.y = mh.hdr[i].y;
    c->friction = mh.hdr[i].z; c->restitution = mh.hdr[i].w;
    for (int k = 0; k < 2; ++k) {
        const float4 g = mh.geo[2 * i + k], im = mh.imp[2 * i + k];
        frContact &ct = c->contacts[k];
        ct.id = mh.cid[2 * i + k];
        ct.point.x = g.x; ct.point.y = g.y; ct.depth = g.z; ct.timestamp = g.w;
        ct.cache.normalMass = im.x; ct.cache.normalScalar = im.y; ct.cache.tangentMass = im.z; ct.cache.tangentScalar = im.w;
    }
}
static void pod_to_manifold(ManifoldHost &mh, unsigned int i, const frCollision *c) {
    mh.meta[i].x = c->count < 0 ? 0 : (c->count > 2 ? 2 : c->count);
    mh.hdr[i] = make_float4(c->direction.x, c->direction.y, c->friction, c->restitution);
    for (int k = 0; k < 2; ++k) {
        const frContact &ct = c->contacts[k];
        mh.cid[2 * i + k] = ct.id;
        mh.geo[2 * i + k] = make_float4(ct.point.x, ct.point.y, ct.depth, ct.timestamp);
        mh.imp[2 * i + k] = make_float4(ct.cache.normalMass, ct.cache.normalScalar, ct.cache.tangentMass, ct.cache.tangentScalar);
    }
}

// visiting order of the dense array: reference array order in serial mode, dense order otherwise
static void visiting_order(const ManifoldHost &mh, std::vector<unsigned int> &order) {
    order.clear();
    if (!mh.order_keys.empty() || mh.count == 0) {
        std::unordered_map<unsigned long long, unsigned int> pos;
        pos.reserve(mh.count * 2);
        for (unsigned int i = 0; i < mh.count; ++i) pos[((unsigned long long) mh.key[i].x << 32) | mh.key[i].y] = i;
        for (unsigned long long k : mh.order_keys) {
            auto it = pos.find(k);
            if (it != pos.end()) order.push_back(it->second);
        }
        if (!mh.order_keys.empty()) return;
    }
    for (unsigned int i = 0; i < mh.count; ++i) order.push_back(i);
}

// collision-handler sweep (src/world.c:209-214 / :254-259): every entry with count > 0, in solver
// order, on the caller's thread; the handler may edit *value and enqueue add/remove requests.
// The live manifolds come over as ONE array of public 100-byte records packed on the device in visiting order
// (fx_solve.cu k_pack_collisions); the callbacks get pointers INTO that pinned array (no per-call copy, no
// compare), and the whole array goes back in one upload that a kernel scatters into the cache
// ([measured] 65k bodies, 123 k manifolds, no-op C handlers: 5.9 ms per step with per-array downloads and a
// host-side gather / copy / compare per callback).
struct HandlerBridge {
    frxManifold *host = nullptr;         // pinned
    ManifoldRec *dev = nullptr;
    int4 *head_host = nullptr, *head_dev = nullptr;   // {first, second, count, 0}: what the callback loop itself reads
    unsigned int *idx = nullptr;         // reference-order mode: dense index per visited entry
    size_t cap = 0;
    cudaStream_t up = nullptr;           // uploads of finished chunks run beside the downloads of later ones
    cudaEvent_t up_done = nullptr;
    std::vector<cudaEvent_t> landed;     // one per chunk: its download has completed
    ~HandlerBridge() {
        if (host) cudaFreeHost(host);
        if (head_host) cudaFreeHost(head_host);
        cudaFree(dev);
        cudaFree(head_dev);
        cudaFree(idx);
        for (cudaEvent_t e : landed) cudaEventDestroy(e);
        if (up_done) cudaEventDestroy(up_done);
        if (up) cudaStreamDestroy(up);
    }
};
#define FX_BRIDGE_CHUNK 32768u          // records per pipelined chunk (3.2 MB)
static bool bridge_reserve(HandlerBridge &hb, size_t n) {
    if (n <= hb.cap) return true;
    if (hb.host) cudaFreeHost(hb.host);
    if (hb.head_host) cudaFreeHost(hb.head_host);
    cudaFree(hb.dev);
    cudaFree(hb.head_dev);
    cudaFree(hb.idx);
    hb.host = nullptr; hb.dev = nullptr; hb.idx = nullptr; hb.head_host = nullptr; hb.head_dev = nullptr;
    hb.cap = n + n / 4 + 1024;
    const bool ok = fx_cuda_ok(cudaMallocHost((void **) &hb.host, hb.cap * sizeof(frxManifold)), "cudaMallocHost bridge") &&
                    fx_cuda_ok(cudaMallocHost((void **) &hb.head_host, hb.cap * sizeof(int4)), "cudaMallocHost bridge") &&
                    fx_cuda_ok(cudaMalloc((void **) &hb.dev, hb.cap * sizeof(ManifoldRec)), "cudaMalloc bridge") &&
                    fx_cuda_ok(cudaMalloc((void **) &hb.head_dev, hb.cap * sizeof(int4)), "cudaMalloc bridge") &&
                    fx_cuda_ok(cudaMalloc((void **) &hb.idx, hb.cap * sizeof(unsigned int)), "cudaMalloc bridge");
    if (!ok) hb.cap = 0;
    return ok;
}
static void run_handler(frWorld *w, frCollisionEventFunc fn) {
    static_assert(sizeof(frxManifold) == sizeof(ManifoldRec), "bridge record layout");
    if (!w->bridge) w->bridge = new HandlerBridge();
    HandlerBridge &hb = *static_cast<HandlerBridge *>(w->bridge);
    DevWorld &d = w->dw;
    cudaStream_t s = w->stream;
    const bool serial = (w->solver_mode == FRX_SOLVER_SERIAL) && w->so.r_len != nullptr;
    unsigned int n = 0;
    cudaMemcpyAsync(&n, serial ? w->so.r_len : &d.ctr->n_manifolds, sizeof(unsigned int), cudaMemcpyDeviceToHost, s);
    if (!fx_cuda_ok(cudaStreamSynchronize(s), "manifold count") || n == 0 || d.cap_manifolds == 0) return;
    if (!bridge_reserve(hb, n)) return;
    fx_world_pull(w);   // getters inside the handler must see current state (before the chunk copies queue up behind it)
    if (!hb.up) {
        cudaStreamCreateWithFlags(&hb.up, cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&hb.up_done, cudaEventDisableTiming);
    }
    const unsigned int nchunks = (n + FX_BRIDGE_CHUNK - 1u) / FX_BRIDGE_CHUNK;
    while (hb.landed.size() < nchunks) {
        cudaEvent_t e;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        hb.landed.push_back(e);
    }
    if (serial) fx_launch_visit_index(d, w->cx, w->so, hb.idx, n);
    fx_launch_pack_collisions(d, w->cx, serial ? hb.idx : nullptr, n, hb.dev, hb.head_dev);
    cudaMemcpyAsync(hb.head_host, hb.head_dev, (size_t) n * sizeof(int4), cudaMemcpyDeviceToHost, s);
    // pipeline: chunk k is handed to the callbacks while chunk k+1.. still download and chunk k-1 uploads
    for (unsigned int k = 0; k < nchunks; ++k) {
        const size_t lo = (size_t) k * FX_BRIDGE_CHUNK, cnt = (k + 1u == nchunks) ? n - lo : FX_BRIDGE_CHUNK;
        cudaMemcpyAsync(hb.host + lo, hb.dev + lo, cnt * sizeof(frxManifold), cudaMemcpyDeviceToHost, s);
        cudaEventRecord(hb.landed[k], s);
    }
    const int nb = (int) w->bodies.size();
    frBody *const *bodies = w->bodies.data();
    for (unsigned int k = 0; k < nchunks; ++k) {
        const size_t lo = (size_t) k * FX_BRIDGE_CHUNK, cnt = (k + 1u == nchunks) ? n - lo : FX_BRIDGE_CHUNK;
        if (!fx_cuda_ok(cudaEventSynchronize(hb.landed[k]), "manifold download")) return;
        for (size_t i = lo; i < lo + cnt; ++i) {
            const int4 h = hb.head_host[i];
            if (h.z <= 0) continue;
            frBodyPair pair;
            pair.first = (h.x >= 0 && h.x < nb) ? bodies[h.x] : nullptr;
            pair.second = (h.y >= 0 && h.y < nb) ? bodies[h.y] : nullptr;
            fn(pair, &hb.host[i].value);
        }
        cudaMemcpyAsync(hb.dev + lo, hb.host + lo, cnt * sizeof(frxManifold), cudaMemcpyHostToDevice, hb.up);
    }
    cudaEventRecord(hb.up_done, hb.up);
    cudaStreamWaitEvent(s, hb.up_done, 0);
    fx_launch_unpack_collisions(d, w->cx, serial ? hb.idx : nullptr, n, hb.dev);
}
void fx_release_bridge(frWorld *w) {
    delete static_cast<HandlerBridge *>(w->bridge);
    w->bridge = nullptr;
}

extern "C" int frxGetManifolds(frWorld *w, frxManifold *out, int max) {
    if (!w || !w->device_ok) return 0;
    cudaSetDevice(w->device);
    ManifoldHost &mh = world_mh(w);
    mh.fresh = false;
    if (!download_manifolds(w, mh, true)) return 0;
    std::vector<unsigned int> &order = mh.order;
    visiting_order(mh, order);
    int n = 0;
    for (unsigned int i : order) {
        if (out && n < max) {
            out[n].first = mh.body[i].x;
            out[n].second = mh.body[i].y;
            manifold_to_pod(mh, i, &out[n].value);
        }
        ++n;
    }
    return n;
}

extern "C" int frxGetCandidatePairs(frWorld *w, int *out, int max) {
    if (!w || !w->device_ok) return 0;
    cudaSetDevice(w->device);
    frxSynchronizeWorld(w);
    const unsigned int c = w->last.n_cand;
    if (!out || max <= 0 || c == 0) return (int) c;
    std::vector<int2> pairs(c);
    std::vector<unsigned char> hit(c);
    cudaMemcpy(pairs.data(), w->dw.cand, c * sizeof(int2), cudaMemcpyDeviceToHost);
    cudaMemcpy(hit.data(), w->dw.cand_hit, c, cudaMemcpyDeviceToHost);
    for (unsigned int i = 0; i < c && (int) i < max; ++i) {
        out[3 * i] = pairs[i].x; out[3 * i + 1] = pairs[i].y; out[3 * i + 2] = hit[i];
    }
    return (int) c;
}

// ---------------------------------------------------------------------------------------------
// batch state access
// ---------------------------------------------------------------------------------------------
extern "C" int frxGetBodyStates(frWorld *w, float *position, float *angle, float *sincos, float *velocity,
                                float *angularVelocity, float *aabb) {
    if (!w) return 0;
    if (w->device_ok) cudaSetDevice(w->device);
    fx_world_pull(w);
    const int n = (int) w->bodies.size();
    const float4 *__restrict__ hp = w->h.posrot, *__restrict__ hv = w->h.vel;
    if (position || sincos)
        for (int i = 0; i < n; ++i) {
            const float4 p = hp[i];
            if (position) { position[2 * i] = p.x; position[2 * i + 1] = p.y; }
            if (sincos) { sincos[2 * i] = p.z; sincos[2 * i + 1] = p.w; }
        }
    if (velocity || angularVelocity)
        for (int i = 0; i < n; ++i) {
            const float4 v = hv[i];
            if (velocity) { velocity[2 * i] = v.x; velocity[2 * i + 1] = v.y; }
            if (angularVelocity) angularVelocity[i] = v.z;
        }
    if (angle) std::memcpy(angle, w->h.angle, (size_t) n * sizeof(float));
    if (aabb) std::memcpy(aabb, w->h.aabb, (size_t) n * sizeof(float4));
    return n;
}

extern "C" int frxViewBodyStates(frWorld *w, const float **posRot, const float **angle, const float **velocity, const float **aabb) {
    if (!w) return 0;
    if (w->device_ok) cudaSetDevice(w->device);
    // only the arrays asked for are downloaded (the others follow when a getter needs them)
    unsigned want = 0u;
    if (posRot) want |= FX_DIRTY_POSROT;
    if (angle) want |= FX_DIRTY_ANGLE;
    if (velocity) want |= FX_DIRTY_VEL;
    if (aabb) want |= FX_DIRTY_AABB;
    fx_world_pull_some(w, want);
    if (posRot) *posRot = reinterpret_cast<const float *>(w->h.posrot);
    if (angle) *angle = w->h.angle;
    if (velocity) *velocity = reinterpret_cast<const float *>(w->h.vel);
    if (aabb) *aabb = reinterpret_cast<const float *>(w->h.aabb);
    return (int) w->bodies.size();
}

extern "C" int frxMapBodyForces(frWorld *w, float **force) {
    if (!w || !force) return 0;
    if (w->device_ok) cudaSetDevice(w->device);
    fx_world_pull_some(w, 0u);             // no download: only clears the mirror's accumulators if a step has run since
    *force = reinterpret_cast<float *>(w->h.force);
    w->dirty |= FX_DIRTY_FORCE;
    return (int) w->bodies.size();
}

extern "C" int frxApplyForces(frWorld *w, const float *force, const float *torque, int n) {
    if (!w) return 0;
    fx_world_pull(w);
    if (n > (int) w->bodies.size()) n = (int) w->bodies.size();
    float4 *__restrict__ hf = w->h.force;
    const float4 *__restrict__ hv = w->h.vel;
    for (int i = 0; i < n; ++i) {
        if (hv[i].w <= 0.0f) continue;                         // inverse mass, src/rigid_body.c:300
        float4 f = hf[i];
        if (force) { f.x = f.x + force[2 * i]; f.y = f.y + force[2 * i + 1]; }
        if (torque) f.z += torque[i];
        hf[i] = f;
    }
    w->dirty |= FX_DIRTY_FORCE;
    return n;
}

extern "C" int frxSetBodyVelocities(frWorld *w, const float *velocity, const float *angularVelocity, int n) {
    if (!w) return 0;
    fx_world_pull(w);
    if (n > (int) w->bodies.size()) n = (int) w->bodies.size();
    for (int i = 0; i < n; ++i) {
        if (velocity) { w->h.vel[i].x = velocity[2 * i]; w->h.vel[i].y = velocity[2 * i + 1]; }
        if (angularVelocity) w->h.vel[i].z = angularVelocity[i];
    }
    w->dirty |= FX_DIRTY_VEL;
    return n;
}

extern "C" int frxCreateBodies(frWorld *w, int n, const int *bodyType, const int *shapeIndex, frShape *const *shapes,
                               const float *position, const float *angle, frBody **outBodies) {
    if (!w || n <= 0 || !bodyType || !position) return 0;
    int accepted = 0;
    for (int i = 0; i < n; ++i) {
        frVector2 p = { position[2 * i], position[2 * i + 1] };
        frShape *s = (shapes && shapeIndex && shapeIndex[i] >= 0) ? shapes[shapeIndex[i]] : nullptr;
        frBody *b = frCreateBodyFromShape((frBodyType) bodyType[i], p, s);
        if (!b) { if (outBodies) outBodies[i] = nullptr; continue; }
        if (angle && angle[i] != 0.0f) frSetBodyAngle(b, angle[i]);
        if (outBodies) outBodies[i] = b;
        if (frAddBodyToWorld(w, b)) ++accepted;
    }
    return accepted;
}

extern "C" int frxCreateBodiesEx(frWorld *w, int n, const int *bodyType, const int *shapeIndex, frShape *const *shapes,
                                 const float *position, const float *angle, const float *velocity,
                                 const float *angularVelocity, const int *worldIndex, frBody **outBodies) {
    if (!w || n <= 0 || !bodyType || !position) return 0;
    int accepted = 0;
    for (int i = 0; i < n; ++i) {
        frVector2 p = { position[2 * i], position[2 * i + 1] };
        frShape *s = (shapes && shapeIndex && shapeIndex[i] >= 0) ? shapes[shapeIndex[i]] : nullptr;
        frBody *b = frCreateBodyFromShape((frBodyType) bodyType[i], p, s);
        if (!b) { if (outBodies) outBodies[i] = nullptr; continue; }
        if (angle && angle[i] != 0.0f) frSetBodyAngle(b, angle[i]);
        if (velocity && (velocity[2 * i] != 0.0f || velocity[2 * i + 1] != 0.0f)) {
            frVector2 v = { velocity[2 * i], velocity[2 * i + 1] };
            frSetBodyVelocity(b, v);
        }
        if (angularVelocity && angularVelocity[i] != 0.0f) frSetBodyAngularVelocity(b, angularVelocity[i]);
        if (outBodies) outBodies[i] = b;
        bool ok = worldIndex ? frxAddBodyToBatch(w, worldIndex[i], b) : frAddBodyToWorld(w, b);
        if (ok) ++accepted;
    }
    return accepted;
}

// ---------------------------------------------------------------------------------------------
// single-pair / single-body public entry points: a 2-body scratch world on the device
// ---------------------------------------------------------------------------------------------
struct Scratch {
    bool ready = false;
    int device = -1;
    cudaStream_t stream = nullptr;
    DevWorld dw{};
    FxLaunchCtx cx{};
    ShapeDev *d_shapes = nullptr;
    int2 *d_pairs = nullptr;
    Manifold *d_out = nullptr;
    CollisionPod *d_col = nullptr;
    int pair_cap = 0;
    // pinned staging
    float4 *h4 = nullptr;     // posrot[2] vel[2] mprop[2] force[2] aabb[2]
    float *hangle = nullptr;
    int2 *hshape = nullptr;
    ShapeDev *hshapes = nullptr;
    CollisionPod *hcol = nullptr;
    Manifold *hout = nullptr;
};
static thread_local Scratch g_scratch;

static Scratch *scratch() {
    Scratch &s = g_scratch;
    if (s.ready) { cudaSetDevice(s.device); return &s; }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        fx_set_error(1, "no CUDA device: narrow phase / solver entry points have no CPU fallback");
        return nullptr;
    }
    s.device = current_device() < n ? current_device() : 0;
    cudaSetDevice(s.device);
    if (!fx_cuda_ok(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking), "scratch stream")) return nullptr;
    DevWorld &d = s.dw;
    cudaMalloc((void **) &d.posrot, 2 * sizeof(float4)); cudaMalloc((void **) &d.vel, 2 * sizeof(float4));
    cudaMalloc((void **) &d.mprop, 2 * sizeof(float4)); cudaMalloc((void **) &d.force, 2 * sizeof(float4));
    cudaMalloc((void **) &d.aabb, 2 * sizeof(float4)); cudaMalloc((void **) &d.angle, 2 * sizeof(float));
    cudaMalloc((void **) &d.shape, 2 * sizeof(int2));
    cudaMalloc((void **) &s.d_shapes, 2 * sizeof(ShapeDev));
    cudaMalloc((void **) &s.d_pairs, sizeof(int2)); cudaMalloc((void **) &s.d_out, sizeof(Manifold));
    cudaMalloc((void **) &s.d_col, sizeof(CollisionPod));
    cudaMallocHost((void **) &s.h4, 10 * sizeof(float4)); cudaMallocHost((void **) &s.hangle, 2 * sizeof(float));
    cudaMallocHost((void **) &s.hshape, 2 * sizeof(int2)); cudaMallocHost((void **) &s.hshapes, 2 * sizeof(ShapeDev));
    cudaMallocHost((void **) &s.hcol, sizeof(CollisionPod)); cudaMallocHost((void **) &s.hout, sizeof(Manifold));
    d.shapes = s.d_shapes;
    d.n = 2;
    d.n_owned = 2;
    { const char *e = std::getenv("FEROX_NARROWPHASE"); d.narrow_mode = (e && std::strcmp(e, "group8") == 0) ? 1 : 0; }
    s.cx.stream = s.stream;
    s.cx.max_blocks = 148 * 8;
    s.ready = fx_cuda_ok(cudaGetLastError(), "scratch init");
    return s.ready ? &s : nullptr;
}

static BodyRow current_row(frBody *b) {
    BodyRow r;
    if (b->world) {
        fx_world_pull(b->world);
        const HostMirror &h = b->world->h;
        const int i = b->index;
        r.posrot = h.posrot[i]; r.angle = h.angle[i]; r.vel = h.vel[i]; r.mprop = h.mprop[i]; r.force = h.force[i]; r.aabb = h.aabb[i];
    } else {
        r = b->loc;
    }
    return r;
}
static void store_row(frBody *b, const BodyRow &r, unsigned bits) {
    if (b->world) {
        HostMirror &h = b->world->h;
        const int i = b->index;
        h.posrot[i] = r.posrot; h.angle[i] = r.angle; h.vel[i] = r.vel; h.aabb[i] = r.aabb;
        fx_world_mark(b->world, bits);
    } else {
        b->loc.posrot = r.posrot; b->loc.angle = r.angle; b->loc.vel = r.vel; b->loc.aabb = r.aabb;
    }
}

static bool scratch_upload(Scratch *s, frBody *b1, frBody *b2) {
    frBody *bs[2] = { b1, b2 };
    for (int k = 0; k < 2; ++k) {
        frBody *b = bs[k];
        if (!b) continue;
        BodyRow r = current_row(b);
        s->h4[0 + k] = r.posrot; s->h4[2 + k] = r.vel; s->h4[4 + k] = r.mprop; s->h4[6 + k] = r.force; s->h4[8 + k] = r.aabb;
        s->hangle[k] = r.angle;
        if (b->shape) { fill_shape_dev(s->hshapes[k], b->shape); s->hshape[k] = make_int2(k, (int) b->shape->type); }
        else { std::memset(&s->hshapes[k], 0, sizeof(ShapeDev)); s->hshape[k] = make_int2(-1, 0); }
    }
    cudaStream_t st = s->stream;
    DevWorld &d = s->dw;
    cudaMemcpyAsync(d.posrot, s->h4 + 0, 2 * sizeof(float4), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d.vel, s->h4 + 2, 2 * sizeof(float4), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d.mprop, s->h4 + 4, 2 * sizeof(float4), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d.force, s->h4 + 6, 2 * sizeof(float4), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d.aabb, s->h4 + 8, 2 * sizeof(float4), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d.angle, s->hangle, 2 * sizeof(float), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d.shape, s->hshape, 2 * sizeof(int2), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(s->d_shapes, s->hshapes, 2 * sizeof(ShapeDev), cudaMemcpyHostToDevice, st);
    return true;
}

// copy exactly the fields the reference's narrow phase writes (src/collision.c:264-558) into the
// caller's struct, leaving everything else (cached impulses, ids in the circle cases, timestamps,
// material) as it was
static void merge_manifold(const Manifold &m, frCollision *c) {
    c->direction.x = m.dir.x; c->direction.y = m.dir.y;
    if (m.has_ids) { c->contacts[0].id = m.id[0]; if (m.count == 2) c->contacts[1].id = m.id[1]; }
    c->contacts[0].point.x = m.p[0].x; c->contacts[0].point.y = m.p[0].y;
    c->contacts[0].depth = m.depth[0];
    if (m.count == 2) {
        c->contacts[1].point.x = m.p[1].x; c->contacts[1].point.y = m.p[1].y;
        c->contacts[1].depth = m.depth[1];
    } else {
        c->contacts[1] = c->contacts[0];
    }
    c->count = m.count;
}

extern "C" bool frComputeCollision(frBody *b1, frBody *b2, frCollision *collision) {
    if (b1 == nullptr || b2 == nullptr) return false;
    Scratch *s = scratch();
    if (!s) return false;
    scratch_upload(s, b1, b2);
    int2 pr = make_int2(0, 1);
    cudaMemcpyAsync(s->d_pairs, &pr, sizeof pr, cudaMemcpyHostToDevice, s->stream);
    fx_launch_collide_pairs(s->dw, s->cx, s->d_pairs, 1, s->d_out);
    cudaMemcpyAsync(s->hout, s->d_out, sizeof(Manifold), cudaMemcpyDeviceToHost, s->stream);
    if (!fx_cuda_ok(cudaStreamSynchronize(s->stream), "frComputeCollision")) return false;
    if (s->hout->count <= 0) return false;
    if (collision != nullptr) merge_manifold(*s->hout, collision);
    return true;
}

extern "C" int frxComputeCollisions(frWorld *w, const int *pairs, int nPairs, frCollision *out, int *hit) {
    if (!w || !w->device_ok || !pairs || nPairs <= 0) return 0;
    cudaSetDevice(w->device);
    ensure_body_capacity(w, (int) w->bodies.size());
    sync_shapes(w);
    world_push(w);
    fill_dev_params(w, 1.0f);
    int2 *d_pairs = nullptr;
    Manifold *d_out = nullptr;
    cudaMalloc((void **) &d_pairs, (size_t) nPairs * sizeof(int2));
    cudaMalloc((void **) &d_out, (size_t) nPairs * sizeof(Manifold));
    cudaMemcpyAsync(d_pairs, pairs, (size_t) nPairs * sizeof(int2), cudaMemcpyHostToDevice, w->stream);
    fx_launch_collide_pairs(w->dw, w->cx, d_pairs, nPairs, d_out);
    std::vector<Manifold> res((size_t) nPairs);
    cudaMemcpyAsync(res.data(), d_out, (size_t) nPairs * sizeof(Manifold), cudaMemcpyDeviceToHost, w->stream);
    fx_cuda_ok(cudaStreamSynchronize(w->stream), "frxComputeCollisions");
    cudaFree(d_pairs);
    cudaFree(d_out);
    int hits = 0;
    for (int i = 0; i < nPairs; ++i) {
        const bool h = res[i].count > 0;
        if (hit) hit[i] = h ? 1 : 0;
        if (h) { ++hits; if (out) merge_manifold(res[i], &out[i]); }
    }
    return hits;
}

static void pair_solve(frBody *b1, frBody *b2, frCollision *collision, int op, float inv_dt) {
    if (b1 == nullptr || b2 == nullptr || collision == nullptr) return;
    Scratch *s = scratch();
    if (!s) return;
    scratch_upload(s, b1, b2);
    std::memcpy(s->hcol, collision, sizeof(CollisionPod));
    cudaMemcpyAsync(s->d_col, s->hcol, sizeof(CollisionPod), cudaMemcpyHostToDevice, s->stream);
    fx_launch_single_pair_solve(s->dw, s->stream, s->d_col, op, inv_dt);
    cudaMemcpyAsync(s->hcol, s->d_col, sizeof(CollisionPod), cudaMemcpyDeviceToHost, s->stream);
    cudaMemcpyAsync(s->h4 + 2, s->dw.vel, 2 * sizeof(float4), cudaMemcpyDeviceToHost, s->stream);
    if (!fx_cuda_ok(cudaStreamSynchronize(s->stream), "pair solve")) return;
    std::memcpy(collision, s->hcol, sizeof(CollisionPod));
    frBody *bs[2] = { b1, b2 };
    for (int k = 0; k < 2; ++k) {
        BodyRow r = current_row(bs[k]);
        r.vel = s->h4[2 + k];
        store_row(bs[k], r, FX_DIRTY_VEL);
    }
}
extern "C" void frApplyAccumulatedImpulses(frBody *b1, frBody *b2, frCollision *collision) { pair_solve(b1, b2, collision, 0, 0.0f); }
extern "C" void frResolveCollision(frBody *b1, frBody *b2, frCollision *collision, float inverseDt) {
    pair_solve(b1, b2, collision, 1, inverseDt);
}

static void single_body(frBody *b, int op, float dt) {
    if (b == nullptr || dt <= 0.0f) return;
    Scratch *s = scratch();
    if (!s) return;
    scratch_upload(s, b, nullptr);
    fx_launch_single_body(s->dw, s->stream, 0, op, dt);
    cudaMemcpyAsync(s->h4 + 0, s->dw.posrot, sizeof(float4), cudaMemcpyDeviceToHost, s->stream);
    cudaMemcpyAsync(s->h4 + 2, s->dw.vel, sizeof(float4), cudaMemcpyDeviceToHost, s->stream);
    cudaMemcpyAsync(s->h4 + 8, s->dw.aabb, sizeof(float4), cudaMemcpyDeviceToHost, s->stream);
    cudaMemcpyAsync(s->hangle, s->dw.angle, sizeof(float), cudaMemcpyDeviceToHost, s->stream);
    if (!fx_cuda_ok(cudaStreamSynchronize(s->stream), "single body")) return;
    BodyRow r = current_row(b);
    r.posrot = s->h4[0]; r.vel = s->h4[2]; r.aabb = s->h4[8]; r.angle = s->hangle[0];
    store_row(b, r, FX_DIRTY_POSROT | FX_DIRTY_ANGLE | FX_DIRTY_VEL | FX_DIRTY_AABB);
}
extern "C" void frIntegrateForBodyVelocity(frBody *b, float dt) { single_body(b, 0, dt); }
extern "C" void frIntegrateForBodyPosition(frBody *b, float dt) { single_body(b, 1, dt); }

extern "C" int frxDeviceSinCos(const float *angles, int n, float *outSin, float *outCos) {
    if (!angles || n <= 0) return 0;
    Scratch *s = scratch();
    if (!s) return 0;
    float *d = nullptr;
    if (!fx_cuda_ok(cudaMalloc((void **) &d, 3 * (size_t) n * sizeof(float)), "cudaMalloc")) return 0;
    cudaMemcpyAsync(d, angles, (size_t) n * sizeof(float), cudaMemcpyHostToDevice, s->stream);
    fx_launch_sincos(d, n, d + n, d + 2 * (size_t) n, s->stream);
    if (outSin) cudaMemcpyAsync(outSin, d + n, (size_t) n * sizeof(float), cudaMemcpyDeviceToHost, s->stream);
    if (outCos) cudaMemcpyAsync(outCos, d + 2 * (size_t) n, (size_t) n * sizeof(float), cudaMemcpyDeviceToHost, s->stream);
    bool ok = fx_cuda_ok(cudaStreamSynchronize(s->stream), "frxDeviceSinCos");
    cudaFree(d);
    return ok ? n : 0;
}

// ---------------------------------------------------------------------------------------------
// world ray cast (reference src/world.c:291-320): query API over the synced host state
// ---------------------------------------------------------------------------------------------
extern "C" int frxRaycastBatch(frWorld *w, const frRay *rays, int nRays, frRaycastHit *hits, int *counts, int maxHitsPerRay) {
    if (!w || !rays || nRays <= 0 || !hits || !counts || maxHitsPerRay <= 0) return 0;
    if (!w->device_ok) { fx_set_error(1, "frxRaycastBatch: world has no CUDA device (no CPU fallback)"); return 0; }
    static_assert(sizeof(frRaycastHit) == 32 && sizeof(frRay) == 20, "ray POD layouts");
    for (int r = 0; r < nRays; ++r) counts[r] = 0;
    if (!(w->cell > 0.0f) || w->bodies.empty()) return 0;          // NULL hash: every hash call is a no-op (src/broad_phase.c:100,134)
    cudaSetDevice(w->device);
    ensure_body_capacity(w, (int) w->bodies.size());
    sync_shapes(w);
    world_push(w);                                                   // setters since the last step
    DevWorld &d = w->dw;
    d.n_owned = (int) w->bodies.size();
    d.n = d.n_owned;
    d.cell = w->cell;
    d.inv_cell = 1.0f / w->cell;
    d.shapes = w->d_shapes;
    cudaStream_t s = w->stream;
    void *d_rays = nullptr, *d_hits = nullptr;
    int *d_counts = nullptr;
    const size_t hit_bytes = (size_t) nRays * (size_t) maxHitsPerRay * sizeof(frRaycastHit);
    if (!fx_cuda_ok(cudaMalloc(&d_rays, (size_t) nRays * sizeof(frRay)), "cudaMalloc rays") ||
        !fx_cuda_ok(cudaMalloc(&d_hits, hit_bytes), "cudaMalloc hits") ||
        !fx_cuda_ok(cudaMalloc((void **) &d_counts, (size_t) nRays * sizeof(int)), "cudaMalloc counts")) {
        cudaFree(d_rays); cudaFree(d_hits); cudaFree(d_counts);
        return 0;
    }
    cudaMemcpyAsync(d_rays, rays, (size_t) nRays * sizeof(frRay), cudaMemcpyHostToDevice, s);
    fx_launch_raycast_batch(d, w->cx, d_rays, nRays, maxHitsPerRay, d_hits, d_counts);
    cudaMemcpyAsync(counts, d_counts, (size_t) nRays * sizeof(int), cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(hits, d_hits, hit_bytes, cudaMemcpyDeviceToHost, s);
    const bool ok = fx_cuda_ok(cudaStreamSynchronize(s), "frxRaycastBatch");
    cudaFree(d_rays); cudaFree(d_hits); cudaFree(d_counts);
    if (!ok) return 0;
    long long total = 0;
    for (int r = 0; r < nRays; ++r) {
        total += counts[r];
        const int k = counts[r] < maxHitsPerRay ? counts[r] : maxHitsPerRay;
        for (int j = 0; j < k; ++j) {
            frRaycastHit &h = hits[(size_t) r * (size_t) maxHitsPerRay + j];
            const long long idx = (long long) (intptr_t) h.body;      // the kernel left the world index in the pointer slot
            h.body = (idx >= 0 && idx < (long long) w->bodies.size()) ? w->bodies[(size_t) idx] : nullptr;
        }
    }
    return (int) (total > 0x7fffffff ? 0x7fffffff : total);
}

extern "C" int frxPointQueryBatch(frWorld *w, const frVector2 *points, int nPoints, int *bodyIndices, int *counts, int maxPerPoint) {
    if (!w || !points || nPoints <= 0 || !bodyIndices || !counts || maxPerPoint <= 0) return 0;
    if (!w->device_ok) { fx_set_error(1, "frxPointQueryBatch: world has no CUDA device (no CPU fallback)"); return 0; }
    for (int r = 0; r < nPoints; ++r) counts[r] = 0;
    if (w->bodies.empty()) return 0;
    cudaSetDevice(w->device);
    ensure_body_capacity(w, (int) w->bodies.size());
    sync_shapes(w);
    world_push(w);
    DevWorld &d = w->dw;
    d.n_owned = (int) w->bodies.size();
    d.n = d.n_owned;
    d.shapes = w->d_shapes;
    cudaStream_t s = w->stream;
    void *d_points = nullptr;
    int *d_idx = nullptr, *d_counts = nullptr;
    const size_t idx_bytes = (size_t) nPoints * (size_t) maxPerPoint * sizeof(int);
    if (!fx_cuda_ok(cudaMalloc(&d_points, (size_t) nPoints * sizeof(frVector2)), "cudaMalloc points") ||
        !fx_cuda_ok(cudaMalloc((void **) &d_idx, idx_bytes), "cudaMalloc indices") ||
        !fx_cuda_ok(cudaMalloc((void **) &d_counts, (size_t) nPoints * sizeof(int)), "cudaMalloc counts")) {
        cudaFree(d_points); cudaFree(d_idx); cudaFree(d_counts);
        return 0;
    }
    cudaMemcpyAsync(d_points, points, (size_t) nPoints * sizeof(frVector2), cudaMemcpyHostToDevice, s);
    fx_launch_point_query_batch(d, w->cx, d_points, nPoints, maxPerPoint, d_idx, d_counts);
    cudaMemcpyAsync(counts, d_counts, (size_t) nPoints * sizeof(int), cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(bodyIndices, d_idx, idx_bytes, cudaMemcpyDeviceToHost, s);
    const bool ok = fx_cuda_ok(cudaStreamSynchronize(s), "frxPointQueryBatch");
    cudaFree(d_points); cudaFree(d_idx); cudaFree(d_counts);
    if (!ok) return 0;
    long long total = 0;
    for (int r = 0; r < nPoints; ++r) total += counts[r];
    return (int) (total > 0x7fffffff ? 0x7fffffff : total);
}

// ---------------------------------------------------------------------------------------------
// world ray cast (reference src/world.c:291-320): one ray through the batched device query
// ---------------------------------------------------------------------------------------------
extern "C" void frComputeWorldRaycast(frWorld *w, frRay ray, frRaycastQueryFunc func, void *userData) {
    if (w == nullptr || func == nullptr || !(w->cell > 0.0f)) return;
    int cap = 256, count = 0;
    std::vector<frRaycastHit> hits;
    for (;;) {
        hits.assign((size_t) cap, frRaycastHit{});
        frxRaycastBatch(w, &ray, 1, hits.data(), &count, cap);
        if (count <= cap) break;
        cap = count;                                                  // rare: more than 256 bodies on one ray
    }
    for (int j = 0; j < count; ++j) func(hits[(size_t) j], userData);
}
