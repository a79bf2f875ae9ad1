/*
 * nbody_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the arithmetic of the reference's data-parallel systems, used only by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the
 * checker and the reported CPU baseline.  Nothing under recs_b200/ may call into this file.
 *
 * PARITY UNPINNED for floating point: the reference (martinjonsson01/recs, Rust nightly) cannot
 * be compiled in this environment and its own test-suite asserts nothing about the numerical
 * results of these systems (crates/ecs_bench_suite/tests/debug_benchmarks.rs:33-37 only runs
 * them).  The functions below follow the Rust source line by line; the cgmath 0.18.0 vector
 * semantics (Cargo.lock:1025-1026, not vendored) are restated from its published source:
 *   Point3 - Point3 -> Vector3 componentwise; magnitude2 = dot(v, v) = (x*x + y*y) + z*z;
 *   magnitude = sqrt(magnitude2); normalize_to(m) = v * (m / magnitude(v));
 *   Sum for Vector3 = left fold from zero with componentwise +.
 * rustc never contracts a*b+c into an FMA, so this file MUST be compiled with
 * -ffp-contract=off and without -ffast-math (see oracle/Makefile).
 *
 * Component layout is the reference's: Position / Velocity / Acceleration are 3 contiguous f32
 * (cgmath Point3/Vector3, crates/gfx/src/lib.rs:98-101, crates/n-body/src/lib.rs:53,70), Mass is
 * one f32 (crates/n-body/src/lib.rs:43).
 */
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define ORACLE_API __attribute__((visibility("default")))

/* crates/n-body/src/lib.rs:14 */
static const float NBODY_DT = 1.0f / 20000.0f;
/* crates/n-body/src/lib.rs:123 */
static const float GRAVITATIONAL_CONSTANT = 6.67e-11f;
/* f32::EPSILON, crates/n-body/src/lib.rs:118 */
static const float F32_EPSILON = 1.1920929e-7f;

ORACLE_API int oracle_abi_version(void) { return 1; }

/* ---- gravity: crates/n-body/src/lib.rs:104-135 ------------------------------------------- */
/* One outer body (position px,py,pz) against every (position, mass) of the nested query, in
 * query order, summed sequentially in f32 like `.sum()` (lib.rs:129-133). */
static inline void gravity_one_f32(float px, float py, float pz, const float* qpos, const float* qmass,
                                   size_t nq, float* out) {
    float sx = 0.0f, sy = 0.0f, sz = 0.0f; /* Vector3::zero() */
    for (size_t j = 0; j < nq; ++j) {
        /* let to_body = body_position - position;                               lib.rs:115 */
        const float tx = qpos[3 * j + 0] - px;
        const float ty = qpos[3 * j + 1] - py;
        const float tz = qpos[3 * j + 2] - pz;
        /* let distance_squared = to_body.magnitude2();                          lib.rs:116 */
        const float d2 = (tx * tx + ty * ty) + tz * tz;
        /* if distance_squared <= f32::EPSILON { return Vector3::zero(); }       lib.rs:118-120 */
        float cx = 0.0f, cy = 0.0f, cz = 0.0f;
        if (!(d2 <= F32_EPSILON)) {
            /* let acceleration = G * body_mass / distance_squared;              lib.rs:124 */
            const float a = (GRAVITATIONAL_CONSTANT * qmass[j]) / d2;
            /* to_body.normalize_to(acceleration) = to_body * (a / |to_body|)    lib.rs:126 */
            const float k = a / sqrtf(d2);
            cx = tx * k;
            cy = ty * k;
            cz = tz * k;
        }
        sx = sx + cx;
        sy = sy + cy;
        sz = sz + cz;
    }
    out[0] = sx; /* *acceleration = total_acceleration;                          lib.rs:134 */
    out[1] = sy;
    out[2] = sz;
}

/* Same sum in f64 ("truth" used to arbitrate between the two f32 summation orders). */
static inline void gravity_one_f64(double px, double py, double pz, const float* qpos, const float* qmass,
                                   size_t nq, double* out) {
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (size_t j = 0; j < nq; ++j) {
        const double tx = (double)qpos[3 * j + 0] - px;
        const double ty = (double)qpos[3 * j + 1] - py;
        const double tz = (double)qpos[3 * j + 2] - pz;
        const double d2 = (tx * tx + ty * ty) + tz * tz;
        if (!(d2 <= (double)F32_EPSILON)) {
            const double a = ((double)GRAVITATIONAL_CONSTANT * (double)qmass[j]) / d2;
            const double k = a / sqrt(d2);
            sx += tx * k;
            sy += ty * k;
            sz += tz * k;
        }
    }
    out[0] = sx;
    out[1] = sy;
    out[2] = sz;
}

/* Rows [row_begin, row_end) of the outer iteration; acc rows outside are untouched. */
ORACLE_API void oracle_gravity_rows(const float* pos, float* acc, size_t row_begin, size_t row_end,
                                    const float* qpos, const float* qmass, size_t nq) {
    for (size_t i = row_begin; i < row_end; ++i)
        gravity_one_f32(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], qpos, qmass, nq, acc + 3 * i);
}

ORACLE_API void oracle_gravity_rows_f64(const float* pos, double* acc, size_t row_begin, size_t row_end,
                                        const float* qpos, const float* qmass, size_t nq) {
    for (size_t i = row_begin; i < row_end; ++i)
        gravity_one_f64(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], qpos, qmass, nq, acc + 3 * i);
}

/* Arbitrary sampled rows (parity at 1 M bodies): rows[k] -> out[3k..3k+2]. */
ORACLE_API void oracle_gravity_sample(const float* pos, const uint64_t* rows, size_t n_rows, const float* qpos,
                                      const float* qmass, size_t nq, float* out_f32, double* out_f64) {
    for (size_t k = 0; k < n_rows; ++k) {
        const size_t i = (size_t)rows[k];
        if (out_f32) gravity_one_f32(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], qpos, qmass, nq, out_f32 + 3 * k);
        if (out_f64) gravity_one_f64(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], qpos, qmass, nq, out_f64 + 3 * k);
    }
}

/* The same sampled rows spread over `n_threads` host threads (per-row results do not depend on it). */
typedef struct {
    const float* pos;
    const uint64_t* rows;
    size_t begin, end;
    const float *qpos, *qmass;
    size_t nq;
    float* out_f32;
    double* out_f64;
} sample_job_t;
static void* sample_worker(void* arg) {
    const sample_job_t* j = (const sample_job_t*)arg;
    oracle_gravity_sample(j->pos, j->rows + j->begin, j->end - j->begin, j->qpos, j->qmass, j->nq,
                          j->out_f32 ? j->out_f32 + 3 * j->begin : NULL, j->out_f64 ? j->out_f64 + 3 * j->begin : NULL);
    return NULL;
}
ORACLE_API void oracle_gravity_sample_mt(const float* pos, const uint64_t* rows, size_t n_rows, const float* qpos,
                                         const float* qmass, size_t nq, float* out_f32, double* out_f64, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if ((size_t)n_threads > n_rows) n_threads = n_rows ? (int)n_rows : 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    sample_job_t* jobs = (sample_job_t*)malloc(sizeof(sample_job_t) * (size_t)n_threads);
    for (int t = 0; t < n_threads; ++t) {
        sample_job_t j = {pos, rows, n_rows * (size_t)t / (size_t)n_threads, n_rows * (size_t)(t + 1) / (size_t)n_threads,
                          qpos, qmass, nq, out_f32, out_f64};
        jobs[t] = j;
        pthread_create(&th[t], NULL, sample_worker, &jobs[t]);
    }
    for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    free(jobs);
    free(th);
}

/* ---- total energy (relative total-energy drift over 100 ticks, reported) ------------------------------
 * E = sum_i 1/2 m_i |v_i|^2  -  sum_{i<j} G m_i m_j / |p_j - p_i|, all in f64 from f64 state (f32 columns are widened by
 * the caller).  Pairs the force law masks (d2 <= f32::EPSILON, lib.rs:118-120) carry no force and so no potential
 * here either.  Not part of the reference (it never computes an energy); a diagnostic of the integration. */
typedef struct {
    const double *pos, *vel;
    const float* mass;
    size_t n;
    int t, n_threads;
    double kinetic, potential;
} energy_job_t;
static void* energy_worker(void* arg) {
    energy_job_t* j = (energy_job_t*)arg;
    double k = 0.0, u = 0.0;
    for (size_t i = (size_t)j->t; i < j->n; i += (size_t)j->n_threads) { /* interleaved rows: even pair counts */
        const double px = j->pos[3 * i], py = j->pos[3 * i + 1], pz = j->pos[3 * i + 2];
        const double vx = j->vel[3 * i], vy = j->vel[3 * i + 1], vz = j->vel[3 * i + 2];
        const double mi = (double)j->mass[i];
        k += 0.5 * mi * ((vx * vx + vy * vy) + vz * vz);
        double ui = 0.0;
        for (size_t q = i + 1; q < j->n; ++q) {
            const double tx = j->pos[3 * q] - px, ty = j->pos[3 * q + 1] - py, tz = j->pos[3 * q + 2] - pz;
            const double d2 = (tx * tx + ty * ty) + tz * tz;
            if (!(d2 <= (double)F32_EPSILON)) ui += (double)j->mass[q] / sqrt(d2);
        }
        u -= (double)GRAVITATIONAL_CONSTANT * mi * ui;
    }
    j->kinetic = k;
    j->potential = u;
    return NULL;
}
/* out[0] = kinetic, out[1] = potential (negative). */
ORACLE_API void oracle_total_energy_mt(const double* pos, const float* mass, const double* vel, size_t n, int n_threads,
                                       double* out) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    energy_job_t jobs[256];
    for (int t = 0; t < n_threads; ++t) {
        energy_job_t j = {pos, vel, mass, n, t, n_threads, 0.0, 0.0};
        jobs[t] = j;
        pthread_create(&th[t], NULL, energy_worker, &jobs[t]);
    }
    out[0] = out[1] = 0.0;
    for (int t = 0; t < n_threads; ++t) {
        pthread_join(th[t], NULL);
        out[0] += jobs[t].kinetic; /* fixed thread order: the sum does not depend on timing */
        out[1] += jobs[t].potential;
    }
}

/* ---- movement / acceleration: crates/n-body/src/lib.rs:87-102 ---------------------------- */
/* position.point += velocity * FIXED_TIME_STEP  (mul, then add; componentwise) */
ORACLE_API void oracle_axpy_rows(float* y, const float* x, float a, size_t row_begin, size_t row_end) {
    for (size_t i = 3 * row_begin; i < 3 * row_end; ++i) {
        const float t = x[i] * a;
        y[i] = y[i] + t;
    }
}
ORACLE_API float oracle_nbody_dt(void) { return NBODY_DT; }

/* simple_iter: position.0 += velocity.0 (crates/ecs_bench_suite/src/recs/simple_iter.rs:48) */
ORACLE_API void oracle_add_rows(float* y, const float* x, size_t row_begin, size_t row_end) {
    for (size_t i = 3 * row_begin; i < 3 * row_end; ++i) y[i] = y[i] + x[i];
}

/* ---- rain: crates/rain-simulation/src/lib.rs:59-96 ---------------------------------------- */
static const float RAIN_DT = 1.0f / 20.0f;          /* :15 */
static const float GRAVITY_ACCELERATION = 9.82f;    /* :59 */
static const float TERMINAL_VELOCITY = 9.0f;        /* :60 */
static const float WIND_ACCELERATION = 10.0f;       /* :72 */

/* gravity(mut velocity: Write<Velocity>)                                                :62-70 */
ORACLE_API void oracle_rain_gravity(float* vel, size_t row_begin, size_t row_end) {
    /* -GRAVITY_ACCELERATION * FIXED_TIME_STEP * Vector3::unit_y(): (f32 * f32) * Vector3 */
    const float k = (-GRAVITY_ACCELERATION) * RAIN_DT;
    const float gx = k * 0.0f, gy = k * 1.0f, gz = k * 0.0f;
    for (size_t i = row_begin; i < row_end; ++i) {
        float* v = vel + 3 * i;
        if (fabsf(v[1]) >= TERMINAL_VELOCITY) continue;
        v[0] = v[0] + gx;
        v[1] = v[1] + gy;
        v[2] = v[2] + gz;
    }
}
/* wind(mut velocity: Write<Velocity>)                                                   :86-92 */
ORACLE_API void oracle_rain_wind(float* vel, size_t row_begin, size_t row_end, const float dir[3]) {
    const float c = WIND_ACCELERATION * RAIN_DT;
    const float wx = c * dir[0], wy = c * dir[1], wz = c * dir[2];
    for (size_t i = row_begin; i < row_end; ++i) {
        float* v = vel + 3 * i;
        const float m = sqrtf((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]); /* magnitude() */
        if (m >= TERMINAL_VELOCITY) continue;
        v[0] = v[0] + wx;
        v[1] = v[1] + wy;
        v[2] = v[2] + wz;
    }
}
/* movement(mut position: Write<Position>, velocity: Read<Velocity>)                     :94-96 */
ORACLE_API void oracle_rain_movement(float* pos, const float* vel, size_t row_begin, size_t row_end) {
    oracle_axpy_rows(pos, vel, RAIN_DT, row_begin, row_end);
}
/* wind_direction()                                                                      :78-84
 * Quaternion::from_axis_angle(unit_y, Deg(10.0) * dt).rotate_vector(dir) with cgmath 0.18:
 *   Rad(deg * (PI/180 as f32)); (s, c) = sin_cos(rad * 0.5); q = (c, axis * s);
 *   q * v = (q.v x (q.v x v + v * q.s)) * 2 + v                                                */
ORACLE_API void oracle_rain_wind_direction(float dir[3]) {
    const float deg = 10.0f * RAIN_DT;
    const float rad = deg * (float)(3.14159265358979323846 / 180.0);
    const float half = rad * 0.5f;
    const float s = sinf(half), qs = cosf(half);
    const float qx = 0.0f * s, qy = 1.0f * s, qz = 0.0f * s;
    const float vx = dir[0], vy = dir[1], vz = dir[2];
    const float tx = (qy * vz - qz * vy) + vx * qs;
    const float ty = (qz * vx - qx * vz) + vy * qs;
    const float tz = (qx * vy - qy * vx) + vz * qs;
    dir[0] = (qy * tz - qz * ty) * 2.0f + vx;
    dir[1] = (qz * tx - qx * tz) * 2.0f + vy;
    dir[2] = (qx * ty - qy * tx) * 2.0f + vz;
}

/* ==============================================================================================
 * CPU baseline: the reference's executor policy restated for timing.
 *   - num_cpus worker threads (crates/scheduler/src/executor/mod.rs:129)
 *   - a system is cut into segments of max(16, N / (threads * 4)) entities
 *     (crates/ecs/src/systems/mod.rs:394-411) and every segment is one task in a shared
 *     injector queue that idle workers drain (executor/mod.rs:172-188, worker.rs:157-168)
 *   - the systems of one tick run in the order given by `order` (the precedence-graph layers,
 *     one system per layer for the n-body set), each layer a barrier (schedule/mod.rs:205-333).
 * The work-stealing deques are replaced by one atomic task counter: same task granularity, same
 * thread count, no per-task allocation, so this baseline is if anything kinder to the CPU.
 * ============================================================================================ */
enum { SYS_GRAVITY = 0, SYS_MOVEMENT = 1, SYS_ACCELERATION = 2 };

typedef struct {
    float *pos, *vel, *acc;
    const float* mass;
    size_t n;       /* bodies the nested query iterates */
    size_t n_outer; /* outer rows iterated (== n except for the bounded row sample of bench.py) */
    int n_threads;
    int n_ticks;
    int order[3];
    /* shared scheduling state */
    atomic_size_t next_task;
    size_t n_tasks, seg_size;
    int current_system;
    pthread_barrier_t barrier;
    float* pos_snapshot; /* positions the query iterates while gravity runs (== pos, read only) */
} pool_t;

static void run_segment(pool_t* p, int sys, size_t b, size_t e) {
    switch (sys) {
        case SYS_GRAVITY:
            oracle_gravity_rows(p->pos, p->acc, b, e, p->pos, p->mass, p->n);
            break;
        case SYS_MOVEMENT:
            oracle_axpy_rows(p->pos, p->vel, NBODY_DT, b, e);
            break;
        case SYS_ACCELERATION:
            oracle_axpy_rows(p->vel, p->acc, NBODY_DT, b, e);
            break;
    }
}

static void* pool_worker(void* arg) {
    pool_t* p = (pool_t*)arg;
    for (int tick = 0; tick < p->n_ticks; ++tick) {
        for (int layer = 0; layer < 3; ++layer) {
            pthread_barrier_wait(&p->barrier); /* layer published by thread 0 */
            const int sys = p->current_system;
            for (;;) {
                const size_t t = atomic_fetch_add(&p->next_task, 1);
                if (t >= p->n_tasks) break;
                const size_t b = t * p->seg_size;
                size_t e = b + p->seg_size;
                if (e > p->n_outer) e = p->n_outer;
                run_segment(p, sys, b, e);
            }
            pthread_barrier_wait(&p->barrier); /* all segments of the layer finished */
        }
    }
    return NULL;
}

/* Runs n_ticks of the n-body systems in `order` on n_threads workers; returns seconds. */
ORACLE_API double oracle_nbody_run_pool_rows(float* pos, const float* mass, float* vel, float* acc, size_t n,
                                             size_t n_outer, int n_ticks, int n_threads, const int order[3]);

ORACLE_API double oracle_nbody_run_pool(float* pos, const float* mass, float* vel, float* acc, size_t n, int n_ticks,
                                        int n_threads, const int order[3]) {
    return oracle_nbody_run_pool_rows(pos, mass, vel, acc, n, n, n_ticks, n_threads, order);
}

/* Same, with the outer iteration restricted to rows [0, n_outer) (the query still sees all n
 * bodies): the bounded sample bench.py times when a full tick would take too long on the CPU. */
ORACLE_API double oracle_nbody_run_pool_rows(float* pos, const float* mass, float* vel, float* acc, size_t n,
                                             size_t n_outer, int n_ticks, int n_threads, const int order[3]) {
    if (n_threads < 1) n_threads = 1;
    if (n_outer > n) n_outer = n;
    pool_t p;
    memset(&p, 0, sizeof p);
    p.pos = pos;
    p.vel = vel;
    p.acc = acc;
    p.mass = mass;
    p.n = n;
    p.n_outer = n_outer;
    p.n_threads = n_threads;
    p.n_ticks = n_ticks;
    memcpy(p.order, order, sizeof p.order);
    /* calculate_auto_segment_size / calculate_segment_count, systems/mod.rs:397-411 */
    size_t seg = n_outer / ((size_t)n_threads * 4);
    if (seg < 16) seg = 16;
    p.seg_size = seg;
    p.n_tasks = n_outer ? 1 + (n_outer - 1) / seg : 1;

    /* n_threads workers + this thread as the scheduling thread */
    pthread_barrier_init(&p.barrier, NULL, (unsigned)n_threads + 1);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    struct timespec t0, t1;
    for (int i = 0; i < n_threads; ++i) pthread_create(&th[i], NULL, pool_worker, &p);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int tick = 0; tick < n_ticks; ++tick) {
        for (int layer = 0; layer < 3; ++layer) {
            p.current_system = p.order[layer];
            atomic_store(&p.next_task, 0);
            pthread_barrier_wait(&p.barrier);
            pthread_barrier_wait(&p.barrier);
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    for (int i = 0; i < n_threads; ++i) pthread_join(th[i], NULL);
    free(th);
    pthread_barrier_destroy(&p.barrier);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* Single-threaded sequential tick(s) in `order` (the `Sequential` executor, ecs/src/lib.rs:439-465). */
ORACLE_API void oracle_nbody_run_sequential(float* pos, const float* mass, float* vel, float* acc, size_t n,
                                            int n_ticks, const int order[3]) {
    for (int tick = 0; tick < n_ticks; ++tick) {
        for (int layer = 0; layer < 3; ++layer) {
            switch (order[layer]) {
                case SYS_GRAVITY:
                    oracle_gravity_rows(pos, acc, 0, n, pos, mass, n);
                    break;
                case SYS_MOVEMENT:
                    oracle_axpy_rows(pos, vel, NBODY_DT, 0, n);
                    break;
                case SYS_ACCELERATION:
                    oracle_axpy_rows(vel, acc, NBODY_DT, 0, n);
                    break;
            }
        }
    }
}

/* f64 tick(s) for drift measurement (positions/velocities kept in f64 too). */
ORACLE_API void oracle_nbody_run_f64(double* pos, const float* mass, double* vel, double* acc, size_t n, int n_ticks,
                                     const int order[3]) {
    for (int tick = 0; tick < n_ticks; ++tick) {
        for (int layer = 0; layer < 3; ++layer) {
            if (order[layer] == SYS_GRAVITY) {
                for (size_t i = 0; i < n; ++i) {
                    double sx = 0, sy = 0, sz = 0;
                    for (size_t j = 0; j < n; ++j) {
                        const double tx = pos[3 * j] - pos[3 * i], ty = pos[3 * j + 1] - pos[3 * i + 1],
                                     tz = pos[3 * j + 2] - pos[3 * i + 2];
                        const double d2 = (tx * tx + ty * ty) + tz * tz;
                        if (!(d2 <= (double)F32_EPSILON)) {
                            const double k = (((double)GRAVITATIONAL_CONSTANT * (double)mass[j]) / d2) / sqrt(d2);
                            sx += tx * k;
                            sy += ty * k;
                            sz += tz * k;
                        }
                    }
                    acc[3 * i] = sx;
                    acc[3 * i + 1] = sy;
                    acc[3 * i + 2] = sz;
                }
            } else if (order[layer] == SYS_MOVEMENT) {
                for (size_t i = 0; i < 3 * n; ++i) pos[i] += vel[i] * (double)NBODY_DT;
            } else {
                for (size_t i = 0; i < 3 * n; ++i) vel[i] += acc[i] * (double)NBODY_DT;
            }
        }
    }
}
