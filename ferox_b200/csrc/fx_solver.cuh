// fx_solver.cuh -- the per-contact arithmetic of the sequential-impulse solver, shared verbatim by
// the graph-coloured sweep, the serial (reference-order) sweep and the single-pair entry points
// frApplyAccumulatedImpulses / frResolveCollision.
//
// Restates reference src/rigid_body.c:333-411 (effective masses + warm start) and :446-551 (one
// Gauss-Seidel update of one contact), including its quirks, which parity requires (SURVEY.md
// App. A.8): the angular term of the relative velocity uses UNIT perpendiculars of the lever arms
// (:465-474) while the effective masses and the applied angular impulse use the true arms; the
// restitution term is re-applied in every iteration; the relative velocity is re-evaluated before
// the normal impulse has been applied (:501-507), i.e. it is the same value.
#pragma once
#include "fx_math.cuh"

#define FX_BAUMGARTE_FACTOR 0.2f   // FR_WORLD_BAUMGARTE_FACTOR, include/ferox.h:53-56
#define FX_BAUMGARTE_SLOP 0.01f    // FR_WORLD_BAUMGARTE_SLOP,   include/ferox.h:58-61

struct Vel3 {
    float x, y, w;
};

// position-only constants of one contact for one step
struct ContactK {
    V2 r1, r2;      // lever arms (contact point - body position)
    V2 u1, u2;      // frVector2LeftNormal(r): unit perpendiculars
    float bias, nmass, tmass;
};

FX_HD void contact_arms(ContactK &k, V2 point, V2 pa, V2 pb) {
    k.r1 = vsub(point, pa);
    k.r2 = vsub(point, pb);
    k.u1 = vperp_left(k.r1);
    k.u2 = vperp_left(k.r2);
}

// Baumgarte bias, :478-481.  fminf operand order as compiled in the reference (see pin_fminf).
FX_HD float contact_bias(float depth, float inv_dt) {
    return -(FX_BAUMGARTE_FACTOR * inv_dt) * pin_fminf(-depth + FX_BAUMGARTE_SLOP, 0.0f);
}

// effective masses, :350-385
FX_HD void contact_masses(ContactK &k, V2 n, V2 t, float inv_ma, float inv_mb, float inv_ia, float inv_ib) {
    float c1 = vcross(k.r1, n), c2 = vcross(k.r2, n);
    float nm = (inv_ma + inv_mb) + inv_ia * (c1 * c1) + inv_ib * (c2 * c2);
    k.nmass = 1.0f / nm;
    c1 = vcross(k.r1, t);
    c2 = vcross(k.r2, t);
    float tm = (inv_ma + inv_mb) + inv_ia * (c1 * c1) + inv_ib * (c2 * c2);
    k.tmass = 1.0f / tm;
}

FX_HD void apply_impulse(const ContactK &k, V2 imp, float inv_ma, float inv_mb, float inv_ia, float inv_ib,
                         Vel3 &va, Vel3 &vb) {
    va.x = va.x - imp.x * inv_ma;
    va.y = va.y - imp.y * inv_ma;
    va.w -= inv_ia * vcross(k.r1, imp);
    vb.x = vb.x + imp.x * inv_mb;
    vb.y = vb.y + imp.y * inv_mb;
    vb.w += inv_ib * vcross(k.r2, imp);
}

// warm start of one contact, :355-409
FX_HD void contact_warm_start(const ContactK &k, V2 n, V2 t, float lam_n, float lam_t, float inv_ma, float inv_mb,
                              float inv_ia, float inv_ib, Vel3 &va, Vel3 &vb) {
    V2 acc = vadd(vscale(n, lam_n), vscale(t, lam_t));
    apply_impulse(k, acc, inv_ma, inv_mb, inv_ia, inv_ib, va, vb);
}

// one Gauss-Seidel update of one contact, :457-550
FX_HD void contact_resolve(const ContactK &k, V2 n, V2 t, float friction, float restitution, float &lam_n,
                           float &lam_t, float inv_ma, float inv_mb, float inv_ia, float inv_ib, Vel3 &va,
                           Vel3 &vb) {
    V2 rv = vsub(vadd(v2(vb.x, vb.y), vscale(k.u2, vb.w)), vadd(v2(va.x, va.y), vscale(k.u1, va.w)));
    float vn = vdot(rv, n);
    float dn = ((-(1.0f + restitution) * vn) + k.bias) * k.nmass;
    {
        float old = lam_n;
        lam_n = pin_fmaxf(old + dn, 0.0f);
        dn = lam_n - old;
    }
    V2 n_imp = vscale(n, dn);
    float dt_ = -vdot(rv, t) * k.tmass;
    {
        float max_t = fabsf(friction * lam_n);
        float old = lam_t;
        lam_t = pin_fminf(pin_fmaxf(-max_t, old + dt_), max_t);
        dt_ = lam_t - old;
    }
    V2 tot = vadd(n_imp, vscale(t, dt_));
    apply_impulse(k, tot, inv_ma, inv_mb, inv_ia, inv_ib, va, vb);
}
