/*
 * oracle/synth_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
 *
 * The deterministic synthetic volume `mix` of SURVEY.md 8(d) on the host cores, so that the CPU legs of
 * bench.py (--impl reference) can create their input without touching the CUDA library.  Same formula as
 * volume-renderer_b200/python/volren_b200/workloads.py:mix_volume and the product's synth_mix_kernel
 * (double precision; libm vs CUDA exp/sin may differ in the last ulp, i.e. in a rare +-1 voxel value).
 */
#include "oracle.h"

#include <math.h>
#include <pthread.h>

typedef struct synth_job {
    void* out; int nx, ny, nz, bpv; uint32_t vmax, seed; int with_hash;
    int next_slice;     /* atomic */
} synth_job;

static void* synth_worker(void* arg)
{
    synth_job* j = (synth_job*)arg;
    const double TWO_PI = 6.283185307179586476925286766559;
    for (;;) {
        const int k = __atomic_fetch_add(&j->next_slice, 1, __ATOMIC_RELAXED);
        if (k >= j->nz) break;
        const double pz = ((double)k + 0.5) / (double)j->nz, ddz = pz - 0.5;
        const double sz = sin(TWO_PI * (5.0 * pz + 0.3));
        for (int y = 0; y < j->ny; ++y) {
            const double py = ((double)y + 0.5) / (double)j->ny, ddy = py - 0.5;
            const double sy = sin(TWO_PI * (2.0 * py + 0.2));
            for (int x = 0; x < j->nx; ++x) {
                const double px = ((double)x + 0.5) / (double)j->nx, ddx = px - 0.5;
                const double r = sqrt(ddx * ddx + ddy * ddy + ddz * ddz);
                const double s1 = (r - 0.30) / 0.04, s2 = (r - 0.15) / 0.03;
                double f = 0.70 * (0.6 * exp(-(s1 * s1)) + 0.4 * exp(-(s2 * s2)))
                         + 0.25 * (0.5 + 0.5 * sin(TWO_PI * (3.0 * px + 0.1)) * sy * sz);
                if (j->with_hash) {
                    uint32_t h = ((uint32_t)x * 73856093u) ^ ((uint32_t)y * 19349663u) ^ ((uint32_t)k * 83492791u) ^ j->seed;
                    h *= 2654435761u;
                    f += 0.05 * ((double)h / 4294967296.0);
                }
                f = fmin(fmax(f, 0.0), 1.0);
                const uint32_t v = (uint32_t)floor((double)j->vmax * f + 0.5);
                const uint64_t idx = ((uint64_t)k * (uint64_t)j->ny + (uint64_t)y) * (uint64_t)j->nx + (uint64_t)x;
                if (j->bpv == 1) ((uint8_t*)j->out)[idx] = (uint8_t)v; else ((uint16_t*)j->out)[idx] = (uint16_t)v;
            }
        }
    }
    return 0;
}

int orc_synth_mix(void* out, const int32_t dims[3], int bytes_per_voxel, uint32_t vmax, uint32_t seed, int with_hash, int nthreads)
{
    if (!out || dims[0] < 1 || dims[1] < 1 || dims[2] < 1 || (bytes_per_voxel != 1 && bytes_per_voxel != 2)) return -1;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    synth_job j = { out, dims[0], dims[1], dims[2], bytes_per_voxel, vmax, seed, with_hash, 0 };
    pthread_t th[256];
    for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], 0, synth_worker, &j);
    for (int i = 0; i < nthreads; ++i) pthread_join(th[i], 0);
    return 0;
}
