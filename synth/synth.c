/*
 * synth.c -- deterministic procedural generator for the scan-to-map workloads
 * (SURVEY.md section 8(d)).  Plain C, no dependencies.  It produces the byte-identical
 * inputs that both the CPU oracle and the CUDA path consume:
 *
 *   - a box-room scene (40 x 30 x 8 m) holding 12 axis-aligned pillars/crates,
 *     or a degenerate corridor (two long parallel walls + floor, nothing else in range);
 *   - raw lidar scans in firing order (column-major, rings interleaved) as
 *     x,y,z,intensity,ring,time -- the PointXYZIRT record the reference's
 *     imageProjection.cpp:8-21 consumes -- with exact ray/box intersection,
 *     N(0, sigma) range noise, random dropouts, a few duplicate rays (to exercise the
 *     "first hit wins" rule, imageProjection.cpp:623) and an optional constant-rate
 *     rotation during the sweep (to exercise deskewPoint, imageProjection.cpp:545-580);
 *   - corner / surface feature maps sampled on the scene's edges / faces with jitter;
 *   - pose helpers (ground truth and perturbed guess).
 *
 * All randomness comes from one splitmix64 stream per call, seeded by the caller
 * (seed = 20201018 + 1000*config + frame by convention), so a (seed, parameters) pair
 * names the bytes.  Geometry is evaluated in double and rounded to float once.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SYNTH_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ rng */
typedef struct { uint64_t s; int have; double spare; } rng_t;

static uint64_t rng_next(rng_t *r) {
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static double rng_u01(rng_t *r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static double rng_uniform(rng_t *r, double a, double b) { return a + (b - a) * rng_u01(r); }
static double rng_normal(rng_t *r) {
    if (r->have) { r->have = 0; return r->spare; }
    double u, v, s;
    do { u = 2.0 * rng_u01(r) - 1.0; v = 2.0 * rng_u01(r) - 1.0; s = u * u + v * v; } while (s >= 1.0 || s == 0.0);
    double m = sqrt(-2.0 * log(s) / s);
    r->spare = v * m; r->have = 1;
    return u * m;
}
static void rng_seed(rng_t *r, uint64_t seed) { r->s = seed * 0xD1342543DE82EF95ULL + 0x2545F4914F6CDD1DULL; r->have = 0; r->spare = 0; rng_next(r); }

/* ------------------------------------------------------------------ scene */
#define MAX_BOXES 16
typedef struct { double lo[3], hi[3]; } box_t;
typedef struct {
    box_t room;            /* rays start inside and hit its inner faces            */
    int open_ends;         /* corridor: the two x-end walls and the ceiling are out of range */
    int n_boxes;
    box_t boxes[MAX_BOXES];/* solid obstacles                                       */
} scene_t;

static int boxes_overlap(const box_t *a, const box_t *b, double margin) {
    for (int k = 0; k < 2; k++)
        if (a->hi[k] + margin < b->lo[k] || b->hi[k] + margin < a->lo[k]) return 0;
    return 1;
}

/* kind 0: box room with 12 crates.  kind 1: degenerate corridor (two long walls + floor). */
static void scene_build(scene_t *sc, int kind, uint64_t scene_seed) {
    rng_t r; rng_seed(&r, scene_seed ^ 0x5CE7E5EEDULL);
    memset(sc, 0, sizeof(*sc));
    if (kind == 1) {
        sc->room.lo[0] = -2000.0; sc->room.hi[0] = 2000.0;   /* end walls far out of lidar range */
        sc->room.lo[1] = 0.0;     sc->room.hi[1] = 6.0;
        sc->room.lo[2] = 0.0;     sc->room.hi[2] = 2000.0;   /* no ceiling in range */
        sc->open_ends = 1;
        sc->n_boxes = 0;
        return;
    }
    sc->room.lo[0] = 0.0; sc->room.hi[0] = 40.0;
    sc->room.lo[1] = 0.0; sc->room.hi[1] = 30.0;
    sc->room.lo[2] = 0.0; sc->room.hi[2] = 8.0;
    /* keep-out disc around the region where sensors are placed is NOT enforced: near
       returns (< 1 m) are legal input and exercise the range gate. Boxes must not overlap. */
    int placed = 0, guard = 0;
    while (placed < 12 && guard < 10000) {
        guard++;
        box_t b;
        double ex = rng_uniform(&r, 0.5, 3.0), ey = rng_uniform(&r, 0.5, 3.0), ez = rng_uniform(&r, 0.5, 3.0);
        double cx = rng_uniform(&r, 1.0 + ex * 0.5, 39.0 - ex * 0.5);
        double cy = rng_uniform(&r, 1.0 + ey * 0.5, 29.0 - ey * 0.5);
        b.lo[0] = cx - ex * 0.5; b.hi[0] = cx + ex * 0.5;
        b.lo[1] = cy - ey * 0.5; b.hi[1] = cy + ey * 0.5;
        b.lo[2] = 0.0;           b.hi[2] = ez;
        /* leave the centre strip (where sensor poses are drawn) free of obstacles */
        box_t strip = { {11.0, 9.0, 0.0}, {29.0, 21.0, 8.0} };
        if (boxes_overlap(&b, &strip, 0.0)) {
            /* allow a few boxes in the strip border only */
            if (cx > 13.0 && cx < 27.0 && cy > 11.0 && cy < 19.0) continue;
        }
        int ok = 1;
        for (int i = 0; i < placed; i++) if (boxes_overlap(&b, &sc->boxes[i], 0.3)) { ok = 0; break; }
        if (!ok) continue;
        sc->boxes[placed++] = b;
    }
    sc->n_boxes = placed;
}

/* distance along a ray (origin o inside the room, unit dir d) to the first surface */
static double scene_cast(const scene_t *sc, const double o[3], const double d[3]) {
    double t_room = 1e30;
    for (int k = 0; k < 3; k++) {
        if (d[k] > 1e-12) { double t = (sc->room.hi[k] - o[k]) / d[k]; if (t < t_room) t_room = t; }
        else if (d[k] < -1e-12) { double t = (sc->room.lo[k] - o[k]) / d[k]; if (t < t_room) t_room = t; }
    }
    double best = t_room;
    for (int b = 0; b < sc->n_boxes; b++) {
        const box_t *bx = &sc->boxes[b];
        double tn = 0.0, tf = 1e30; int miss = 0;
        for (int k = 0; k < 3; k++) {
            if (fabs(d[k]) < 1e-12) { if (o[k] < bx->lo[k] || o[k] > bx->hi[k]) { miss = 1; break; } }
            else {
                double t1 = (bx->lo[k] - o[k]) / d[k], t2 = (bx->hi[k] - o[k]) / d[k];
                if (t1 > t2) { double s = t1; t1 = t2; t2 = s; }
                if (t1 > tn) tn = t1;
                if (t2 < tf) tf = t2;
                if (tn > tf) { miss = 1; break; }
            }
        }
        if (!miss && tn > 1e-9 && tn < best) best = tn;
    }
    return best;
}

/* ------------------------------------------------------------------ rotations */
/* R = Rz(yaw) * Ry(pitch) * Rx(roll), the convention of pcl::getTransformation */
static void rot_from_rpy(double roll, double pitch, double yaw, double R[9]) {
    double A = cos(yaw), B = sin(yaw), C = cos(pitch), D = sin(pitch), E = cos(roll), F = sin(roll);
    R[0] = A * C; R[1] = A * D * F - B * E; R[2] = B * F + A * D * E;
    R[3] = B * C; R[4] = A * E + B * D * F; R[5] = B * D * E - A * F;
    R[6] = -D;    R[7] = C * F;             R[8] = C * E;
}
static void mat3_mul(const double A[9], const double B[9], double C[9]) {
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
        C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
static void mat3_vec(const double A[9], const double v[3], double o[3]) {
    for (int i = 0; i < 3; i++) o[i] = A[i * 3] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2];
}
static void rpy_from_rot(const double R[9], double *roll, double *pitch, double *yaw) {
    *roll = atan2(R[7], R[8]); *pitch = asin(-R[6]); *yaw = atan2(R[3], R[0]);
}

/* ------------------------------------------------------------------ sensor models */
/* model: 16 VLP-16, 32 HDL-32E, 64 HDL-64E-like, 128 OS1-128-like */
static double ring_elevation_deg(int n_rings, int ring) {
    switch (n_rings) {
    case 16:  return -15.0 + 2.0 * ring;
    case 32:  return -30.67 + (41.34 / 31.0) * ring;
    case 64:  return -24.8 + (26.8 / 63.0) * ring;
    case 128: return -22.5 + (45.0 / 127.0) * ring;
    default:  return -15.0 + 30.0 * ring / (double)(n_rings > 1 ? n_rings - 1 : 1);
    }
}

/*
 * synth_scan: one lidar sweep.
 *   pose[6]   = sensor pose at sweep start in the map frame (roll,pitch,yaw,x,y,z)
 *   rates[3]  = constant body rotation rates (rad/s about x,y,z, applied as RPY ramps) during the sweep
 *   out arrays sized >= n_rings*horizon + extra_cap  (extra_cap = duplicates + junk)
 * Returns the number of raw points written.
 */
SYNTH_API int synth_scan(int scene_kind, uint64_t scene_seed, uint64_t seed,
                         int n_rings, int horizon, const double pose[6], const double rates[3],
                         double range_sigma, double dropout, double dup_frac,
                         float *x, float *y, float *z, float *intensity, int32_t *ring, float *time,
                         int capacity) {
    scene_t sc; scene_build(&sc, scene_kind, scene_seed);
    rng_t r; rng_seed(&r, seed);
    double R0[9]; rot_from_rpy(pose[0], pose[1], pose[2], R0);
    double o[3] = { pose[3], pose[4], pose[5] };
    const double res = 360.0 / (double)horizon;
    const double period = 0.1;
    int n = 0;
    int n_main = n_rings * horizon;
    int n_dup = (int)(dup_frac * n_main);
    for (int it = 0; it < n_main + n_dup && n < capacity; it++) {
        int c, rg; int junk = 0;
        if (it < n_main) { c = it / n_rings; rg = it % n_rings; }
        else {
            c = (int)(rng_u01(&r) * horizon); if (c >= horizon) c = horizon - 1;
            rg = (int)(rng_u01(&r) * n_rings); if (rg >= n_rings) rg = n_rings - 1;
            if (rng_u01(&r) < 0.05) junk = 1;       /* a few records with an out-of-range ring id */
        }
        double t = period * (double)c / (double)horizon;
        /* column centre under imageProjection.cpp:605-616 : atan2(x,y) = 270 - c*res (deg) */
        double az = (270.0 - (double)c * res) * (M_PI / 180.0);
        double el = ring_elevation_deg(n_rings, rg) * (M_PI / 180.0);
        double ds[3] = { cos(el) * sin(az), cos(el) * cos(az), sin(el) };
        double Rimu[9], Rt[9], dw[3];
        rot_from_rpy(rates[0] * t, rates[1] * t, rates[2] * t, Rimu);
        mat3_mul(R0, Rimu, Rt);
        mat3_vec(Rt, ds, dw);
        double rho = scene_cast(&sc, o, dw);
        double u_drop = rng_u01(&r);
        double noise = rng_normal(&r) * range_sigma;
        double inten = rng_uniform(&r, 0.0, 255.0);
        if (it < n_main && u_drop < dropout) continue;
        if (rho > 120.0) continue;                   /* no return */
        rho += noise;
        if (rho < 0.05) rho = 0.05;
        x[n] = (float)(rho * ds[0]); y[n] = (float)(rho * ds[1]); z[n] = (float)(rho * ds[2]);
        intensity[n] = (float)inten;
        ring[n] = junk ? n_rings + 3 : rg;
        time[n] = (float)t;
        n++;
    }
    return n;
}

/* ------------------------------------------------------------------ maps */
typedef struct { double a[3], u[3], v[3]; double area; } face_t;     /* a + s*u + t*v, s,t in [0,1] */
typedef struct { double a[3], d[3]; double len; } edge_t;            /* a + s*d, s in [0,1]          */

static int add_face(face_t *F, int n, double ax, double ay, double az, double ux, double uy, double uz, double vx, double vy, double vz) {
    face_t f = { {ax, ay, az}, {ux, uy, uz}, {vx, vy, vz}, 0 };
    double lu = sqrt(ux * ux + uy * uy + uz * uz), lv = sqrt(vx * vx + vy * vy + vz * vz);
    f.area = lu * lv; F[n] = f; return n + 1;
}
static int add_edge(edge_t *E, int n, double ax, double ay, double az, double dx, double dy, double dz) {
    edge_t e = { {ax, ay, az}, {dx, dy, dz}, sqrt(dx * dx + dy * dy + dz * dz) }; E[n] = e; return n + 1;
}

static void scene_primitives(const scene_t *sc, face_t *F, int *nf, edge_t *E, int *ne, const double centre[3], double reach) {
    int f = 0, e = 0;
    box_t rm = sc->room;
    if (sc->open_ends) {   /* clip the corridor to what a sensor at `centre` can see */
        rm.lo[0] = centre[0] - reach; rm.hi[0] = centre[0] + reach; rm.hi[2] = 4.0;
    }
    double L = rm.hi[0] - rm.lo[0], W = rm.hi[1] - rm.lo[1], H = rm.hi[2] - rm.lo[2];
    /* floor (+ ceiling), 4 walls */
    f = add_face(F, f, rm.lo[0], rm.lo[1], rm.lo[2], L, 0, 0, 0, W, 0);
    if (!sc->open_ends) f = add_face(F, f, rm.lo[0], rm.lo[1], rm.hi[2], L, 0, 0, 0, W, 0);
    f = add_face(F, f, rm.lo[0], rm.lo[1], rm.lo[2], L, 0, 0, 0, 0, H);
    f = add_face(F, f, rm.lo[0], rm.hi[1], rm.lo[2], L, 0, 0, 0, 0, H);
    if (!sc->open_ends) {
        f = add_face(F, f, rm.lo[0], rm.lo[1], rm.lo[2], 0, W, 0, 0, 0, H);
        f = add_face(F, f, rm.hi[0], rm.lo[1], rm.lo[2], 0, W, 0, 0, 0, H);
        /* vertical room corners */
        e = add_edge(E, e, rm.lo[0], rm.lo[1], rm.lo[2], 0, 0, H);
        e = add_edge(E, e, rm.hi[0], rm.lo[1], rm.lo[2], 0, 0, H);
        e = add_edge(E, e, rm.lo[0], rm.hi[1], rm.lo[2], 0, 0, H);
        e = add_edge(E, e, rm.hi[0], rm.hi[1], rm.lo[2], 0, 0, H);
    }
    /* wall/floor seams along x (present in both scene kinds) */
    e = add_edge(E, e, rm.lo[0], rm.lo[1], rm.lo[2], L, 0, 0);
    e = add_edge(E, e, rm.lo[0], rm.hi[1], rm.lo[2], L, 0, 0);
    for (int b = 0; b < sc->n_boxes; b++) {
        const box_t *bx = &sc->boxes[b];
        double ex = bx->hi[0] - bx->lo[0], ey = bx->hi[1] - bx->lo[1], ez = bx->hi[2] - bx->lo[2];
        /* top + 4 sides */
        f = add_face(F, f, bx->lo[0], bx->lo[1], bx->hi[2], ex, 0, 0, 0, ey, 0);
        f = add_face(F, f, bx->lo[0], bx->lo[1], bx->lo[2], ex, 0, 0, 0, 0, ez);
        f = add_face(F, f, bx->lo[0], bx->hi[1], bx->lo[2], ex, 0, 0, 0, 0, ez);
        f = add_face(F, f, bx->lo[0], bx->lo[1], bx->lo[2], 0, ey, 0, 0, 0, ez);
        f = add_face(F, f, bx->hi[0], bx->lo[1], bx->lo[2], 0, ey, 0, 0, 0, ez);
        /* 4 vertical edges + 4 top edges */
        e = add_edge(E, e, bx->lo[0], bx->lo[1], bx->lo[2], 0, 0, ez);
        e = add_edge(E, e, bx->hi[0], bx->lo[1], bx->lo[2], 0, 0, ez);
        e = add_edge(E, e, bx->lo[0], bx->hi[1], bx->lo[2], 0, 0, ez);
        e = add_edge(E, e, bx->hi[0], bx->hi[1], bx->lo[2], 0, 0, ez);
        e = add_edge(E, e, bx->lo[0], bx->lo[1], bx->hi[2], ex, 0, 0);
        e = add_edge(E, e, bx->lo[0], bx->hi[1], bx->hi[2], ex, 0, 0);
        e = add_edge(E, e, bx->lo[0], bx->lo[1], bx->hi[2], 0, ey, 0);
        e = add_edge(E, e, bx->hi[0], bx->lo[1], bx->hi[2], 0, ey, 0);
    }
    *nf = f; *ne = e;
}

/*
 * synth_map: n_corner points on edges, n_surf points on faces, jittered, as XYZI (intensity =
 * running index, the way the reference's keyframe clouds carry an index, mapOptmization.h:927).
 * centre/reach only matter for the corridor scene.
 */
SYNTH_API void synth_map(int scene_kind, uint64_t scene_seed, uint64_t seed,
                         int n_corner, int n_surf, double jitter_corner, double jitter_surf,
                         const double centre[3], double reach,
                         float *corner_xyzi, float *surf_xyzi) {
    scene_t sc; scene_build(&sc, scene_kind, scene_seed);
    rng_t r; rng_seed(&r, seed ^ 0xA5A5A5A5ULL);
    face_t F[8 + 5 * MAX_BOXES]; edge_t E[8 + 8 * MAX_BOXES]; int nf, ne;
    scene_primitives(&sc, F, &nf, E, &ne, centre, reach);
    double tot_len = 0, tot_area = 0;
    for (int i = 0; i < ne; i++) tot_len += E[i].len;
    for (int i = 0; i < nf; i++) tot_area += F[i].area;
    for (int k = 0; k < n_corner; k++) {
        double pick = rng_u01(&r) * tot_len; int i = 0;
        while (i < ne - 1 && pick > E[i].len) { pick -= E[i].len; i++; }
        double s = rng_u01(&r);
        for (int c = 0; c < 3; c++)
            corner_xyzi[4 * k + c] = (float)(E[i].a[c] + s * E[i].d[c] + rng_uniform(&r, -jitter_corner, jitter_corner));
        corner_xyzi[4 * k + 3] = (float)k;
    }
    for (int k = 0; k < n_surf; k++) {
        double pick = rng_u01(&r) * tot_area; int i = 0;
        while (i < nf - 1 && pick > F[i].area) { pick -= F[i].area; i++; }
        double s = rng_u01(&r), t = rng_u01(&r);
        for (int c = 0; c < 3; c++)
            surf_xyzi[4 * k + c] = (float)(F[i].a[c] + s * F[i].u[c] + t * F[i].v[c] + rng_uniform(&r, -jitter_surf, jitter_surf));
        surf_xyzi[4 * k + 3] = (float)k;
    }
}

/*
 * synth_pose: ground-truth sensor pose inside the free centre strip and a perturbed guess
 * (guess = GT o delta; |dt| <= dt_max per axis, |dr| <= dr_max per axis).
 * Both as (roll,pitch,yaw,x,y,z) doubles.
 */
SYNTH_API void synth_pose(int scene_kind, uint64_t seed, double dt_max, double dr_max, double gt[6], double guess[6]) {
    rng_t r; rng_seed(&r, seed ^ 0x90530ULL);
    if (scene_kind == 1) { gt[3] = rng_uniform(&r, -5.0, 5.0); gt[4] = rng_uniform(&r, 2.0, 4.0); }
    else { gt[3] = rng_uniform(&r, 14.0, 26.0); gt[4] = rng_uniform(&r, 12.0, 18.0); }
    gt[5] = 1.8;
    gt[0] = rng_uniform(&r, -0.03, 0.03); gt[1] = rng_uniform(&r, -0.03, 0.03); gt[2] = rng_uniform(&r, -M_PI, M_PI);
    double Rg[9], Rd[9], Rq[9];
    rot_from_rpy(gt[0], gt[1], gt[2], Rg);
    double dr[3], dt[3];
    for (int k = 0; k < 3; k++) dr[k] = rng_uniform(&r, -dr_max, dr_max);
    for (int k = 0; k < 3; k++) dt[k] = rng_uniform(&r, -dt_max, dt_max);
    rot_from_rpy(dr[0], dr[1], dr[2], Rd);
    mat3_mul(Rg, Rd, Rq);
    rpy_from_rot(Rq, &guess[0], &guess[1], &guess[2]);
    double dtw[3]; mat3_vec(Rg, dt, dtw);
    for (int k = 0; k < 3; k++) guess[3 + k] = gt[3 + k] + dtw[k];
}

/*
 * synth_imu_ramp: the imuTime / imuRotX/Y/Z arrays imageProjection.cpp:323-393 would have
 * integrated for a constant-rate rotation: samples at `hz` from sweep start, rot = rate * dt.
 * Returns the number of samples (covers the 0.1 s sweep plus a margin).
 */
SYNTH_API int synth_imu_ramp(double t_scan_start, const double rates[3], double hz, int capacity,
                             double *imu_time, double *rot_x, double *rot_y, double *rot_z) {
    int n = (int)(0.12 * hz) + 1; if (n > capacity) n = capacity;
    for (int k = 0; k < n; k++) {
        double dt = (double)k / hz;
        imu_time[k] = t_scan_start + dt;
        rot_x[k] = rates[0] * dt; rot_y[k] = rates[1] * dt; rot_z[k] = rates[2] * dt;
    }
    return n;
}
