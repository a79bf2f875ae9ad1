/* The n-body benchmark of the reference (crates/ecs_bench_suite/src/recs/n_body.rs:29-75), assembled in plain C on
 * the C ABI alone -- no Python anywhere: World + Application from include/recs_ecs.h, GPU systems from
 * include/recs_gpu.h.  This is what a host in any language with a C FFI does (INTEGRATION.md shows the Rust side).
 *
 *   cc -O2 -Iinclude examples/nbody_bench.c -o examples/nbody_bench -Lrecs_b200 -lrecs_gpu -Wl,-rpath,$PWD/recs_b200 -lm
 *   examples/nbody_bench [bodies] [ticks]
 *
 * Prints one line: bodies, ticks, ms per tick, G interactions/s, the per-tick system order the schedule produced and
 * order-independent checksums of the final columns (the GPU test compares them with the CPU oracle's run). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "recs_ecs.h"
#include "recs_gpu.h"

enum { POSITION = 1, MASS = 2, VELOCITY = 3, ACCELERATION = 4 };

#define CHECK_GPU(call)                                                                           \
    do {                                                                                          \
        int32_t st__ = (call);                                                                    \
        if (st__ != RECS_OK) {                                                                    \
            fprintf(stderr, "%s -> %d: %s\n", #call, st__, ctx ? recs_gpu_last_error(ctx) : ""); \
            return 1;                                                                             \
        }                                                                                         \
    } while (0)
#define CHECK_APP(call)                                                                      \
    do {                                                                                     \
        int32_t st__ = (call);                                                               \
        if (st__ != RECS_OK) {                                                               \
            fprintf(stderr, "%s -> %d: %s\n", #call, st__, recs_app_last_error(app));       \
            return 1;                                                                        \
        }                                                                                    \
    } while (0)

/* xorshift64*: the draws only have to be the same in this program and in the test that checks it */
static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static float uniform(float lo, float hi) {
    rng_state ^= rng_state >> 12;
    rng_state ^= rng_state << 25;
    rng_state ^= rng_state >> 27;
    const uint64_t r = rng_state * 0x2545F4914F6CDD1Dull;
    const float u = (float)(r >> 40) / 16777216.0f; /* 24 bits -> [0, 1) */
    return lo + u * (hi - lo);
}

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* `benchmark_system` of the reference's bench (crates/ecs_bench_suite/src/recs/n_body.rs:36-52): a parameterless CPU system that
 * counts the ticks.  With it in the schedule every tick passes through the host scheduler (a CPU system may record
 * commands), so recs_app_run cannot hand the remaining ticks to the context in one call. */
static void benchmark_system(void* user, uint32_t archetype, uint64_t count, void* const* columns, uint32_t n_columns) {
    (void)archetype, (void)count, (void)columns, (void)n_columns;
    ++*(uint64_t*)user;
}

int main(int argc, char** argv) {
    const uint64_t n = argc > 1 ? strtoull(argv[1], NULL, 10) : 65536;
    const uint32_t ticks = argc > 2 ? (uint32_t)atoi(argv[2]) : 10;
    const int with_benchmark_system = argc > 3 && strcmp(argv[3], "benchmark_system") == 0;
    uint64_t ticks_counted = 0;
    recs_gpu_ctx* ctx = NULL;
    recs_gpu_config cfg;
    recs_gpu_default_config(&cfg);
    CHECK_GPU(recs_gpu_create(&cfg, &ctx));
    recs_app* app = NULL;
    if (recs_app_create(ctx, &app) != RECS_OK) return 1;
    recs_world* world = recs_app_world(app);
    recs_world_register_component(world, POSITION, 12, "Position");
    recs_world_register_component(world, MASS, 4, "Mass");
    recs_world_register_component(world, VELOCITY, 12, "Velocity");
    recs_world_register_component(world, ACCELERATION, 12, "Acceleration");

    /* registration order of the benchmark: movement, acceleration, gravity (n_body.rs:54-58) */
    const uint64_t movement_ids[2] = {POSITION, VELOCITY}, acceleration_ids[2] = {VELOCITY, ACCELERATION},
                   gravity_ids[3] = {POSITION, ACCELERATION, MASS};
    uint32_t movement, acceleration, gravity;
    CHECK_APP(recs_app_add_gpu_system(app, RECS_SYS_NBODY_MOVEMENT, movement_ids, 2, &movement));
    CHECK_APP(recs_app_add_gpu_system(app, RECS_SYS_NBODY_ACCELERATION, acceleration_ids, 2, &acceleration));
    CHECK_APP(recs_app_add_gpu_system(app, RECS_SYS_NBODY_GRAVITY, gravity_ids, 3, &gravity));
    if (with_benchmark_system) {
        uint32_t index;
        CHECK_APP(recs_app_add_cpu_system(app, "benchmark_system", NULL, 0, benchmark_system, &ticks_counted, &index));
    }

    /* EVERYTHING_HEAVY (crates/n-body/src/scenes.rs:10-21), per body: position xyz, mass, velocity xyz, acceleration xyz */
    float* pos = malloc(n * 12);
    float* mass = malloc(n * 4);
    float* vel = malloc(n * 12);
    float* acc = malloc(n * 12);
    for (uint64_t i = 0; i < n; ++i) {
        for (int k = 0; k < 3; ++k) pos[3 * i + k] = uniform(-100.0f, 100.0f);
        mass[i] = uniform(5.9000e14f, 5.9724e14f);
        for (int k = 0; k < 3; ++k) vel[3 * i + k] = uniform(-10000.0f, 10000.0f);
        for (int k = 0; k < 3; ++k) acc[3 * i + k] = uniform(-1.0f, 1.0f);
    }
    const uint64_t ids[4] = {POSITION, MASS, VELOCITY, ACCELERATION};
    const void* columns[4] = {pos, mass, vel, acc};
    if (recs_world_create_entities(world, n, ids, columns, 4, NULL) != RECS_OK) {
        fprintf(stderr, "create_entities: %s\n", recs_world_last_error(world));
        return 1;
    }

    CHECK_APP(recs_app_run(app, 1, 0)); /* first tick: schedule generation, uploads, workspace sizing */
    CHECK_GPU(recs_gpu_synchronize(ctx));
    const double t0 = now_ms();
    CHECK_APP(recs_app_run(app, ticks, 0));
    CHECK_GPU(recs_gpu_synchronize(ctx));
    const double ms = (now_ms() - t0) / ticks;

    uint32_t order_all[8], batch[8], n_order = 0, order[8], n_gpu = 0;
    CHECK_APP(recs_app_last_tick_order(app, order_all, batch, 8, &n_order));
    for (uint32_t i = 0; i < n_order && i < 8; ++i)
        if (order_all[i] < 3) order[n_gpu++] = order_all[i]; /* (benchmark_system, if registered, is system 3) */
    if (with_benchmark_system && ticks_counted != (uint64_t)ticks + 1) return 1;
    const char* names[3];
    names[movement] = "movement";
    names[acceleration] = "acceleration";
    names[gravity] = "gravity";
    CHECK_APP(recs_app_sync_to_host(app));
    double sums[3] = {0, 0, 0};
    const uint64_t comps[3] = {POSITION, VELOCITY, ACCELERATION};
    for (int c = 0; c < 3; ++c) {
        void* ptr = NULL;
        uint32_t elem = 0;
        uint64_t count = 0;
        if (recs_world_column(world, 1, comps[c], &ptr, &elem, &count) != RECS_OK || count != n || elem != 12) return 1;
        const float* f = ptr;
        for (uint64_t i = 0; i < 3 * n; ++i) sums[c] += fabs((double)f[i]);
    }
    printf("NBODY_BENCH bodies=%llu ticks=%u ms_per_tick=%.4f ginteractions_per_s=%.1f order=%s>%s>%s "
           "sum_abs_pos=%.9e sum_abs_vel=%.9e sum_abs_acc=%.9e\n",
           (unsigned long long)n, ticks + 1, ms, (double)n * (double)n / ms / 1e6, names[order[0]], names[order[1]],
           names[order[2]], sums[0], sums[1], sums[2]);
    recs_app_destroy(app);
    recs_gpu_destroy(ctx);
    free(pos);
    free(mass);
    free(vel);
    free(acc);
    return 0;
}
