/*
 * oracle/march_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
 *
 * Scalar CPU restatement of the reference ray-march kernel, the GLSL 4.30 compute shader
 * /root/reference/VolumeRenderer.cs.  Every GLSL operator is restated as ONE correctly
 * rounded IEEE-754 binary32 operation, in source order; build with
 *     gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math
 * so that the compiler neither contracts a*b+c into an FMA nor reassociates.
 *
 * PINNED to the reference source: the shader text itself is compiled for the CPU into
 * oracle/_ref/libshader_ref.so (oracle/Makefile, oracle/shim/glsl_compat.h) and this restatement
 * equals it bit for bit (tests/test_reference_pinning.py).  This file stays because it also
 * carries the extensions the shader does not have (step override, transfer function, opacity
 * correction), the touch bitmap and the hit counter.  GLSL leaves the precision of
 * normalize/length/division implementation-defined; the IEEE definitions below (the same as
 * in glsl_compat.h) are the contract the CUDA path is checked against.
 *
 * Line references "VR.cs:n" are to /root/reference/VolumeRenderer.cs.
 */
#include "oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* GLSL min/max (spec 8.3): min(x,y) = y<x ? y : x ; max(x,y) = x<y ? y : x.             */
static inline float glsl_min(float x, float y) { return (y < x) ? y : x; }
static inline float glsl_max(float x, float y) { return (x < y) ? y : x; }

typedef struct { float x, y, z, w; } v4;

typedef struct frame_consts {
    /* VR.cs:62-83 : bounding box */
    float pmin[3], pmax[3], half_len[3];
    float denom[3];         /* bb.p_max + half_len, VR.cs:179                        */
    float step_dvr;         /* VR.cs:109 (length(vol_size.xzy))                      */
    float step_mip;         /* VR.cs:146 (length(vol_size.xyz))                      */
    float fmin, fmax;       /* float(min_val), float(max_val)                        */
    float frange;           /* float(max_val - min_val)   (int subtraction first)    */
} frame_consts;

static void make_frame_consts(const orc_params* p, frame_consts* fc)
{
    /* VR.cs:65-66 */
    int max_dim = p->dim[0] > p->dim[1] ? p->dim[0] : p->dim[1];
    max_dim = max_dim > p->dim[2] ? max_dim : p->dim[2];
    const int swz = (p->view_bottom == 1 || p->view_top == 1);

    /* VR.cs:68-78 : p_max = vec4(vol_size.xyz|xzy,1)/max_dim * vec4(voxel_size.xyz|xzy,1) */
    float n[3], vs[3];
    if (swz) {
        n[0] = (float)p->dim[0]; n[1] = (float)p->dim[2]; n[2] = (float)p->dim[1];
        vs[0] = p->voxel_size[0]; vs[1] = p->voxel_size[2]; vs[2] = p->voxel_size[1];
    } else {
        n[0] = (float)p->dim[0]; n[1] = (float)p->dim[1]; n[2] = (float)p->dim[2];
        vs[0] = p->voxel_size[0]; vs[1] = p->voxel_size[1]; vs[2] = p->voxel_size[2];
    }
    const float fmax_dim = (float)max_dim;
    for (int i = 0; i < 3; ++i) {
        float pm = n[i] / fmax_dim;          /* bb.p_max /= max_dim   */
        pm = pm * vs[i];                     /* bb.p_max *= voxel     */
        const float h = pm / 2.0f;           /* VR.cs:81              */
        fc->half_len[i] = h;
        fc->pmin[i] = 0.0f - h;              /* VR.cs:82 (p_min starts at 0) */
        fc->pmax[i] = pm - h;                /* VR.cs:83              */
        fc->denom[i] = fc->pmax[i] + h;      /* VR.cs:179             */
    }

    /* VR.cs:109 / :146 : length(p_max.xyz - p_min.xyz) / length(vec3(vol_size)) */
    const float ex = fc->pmax[0] - fc->pmin[0];
    const float ey = fc->pmax[1] - fc->pmin[1];
    const float ez = fc->pmax[2] - fc->pmin[2];
    const float diag = sqrtf(ex * ex + ey * ey + ez * ez);
    const float fx = (float)p->dim[0], fy = (float)p->dim[1], fz = (float)p->dim[2];
    const float len_xzy = sqrtf(fx * fx + fz * fz + fy * fy);
    const float len_xyz = sqrtf(fx * fx + fy * fy + fz * fz);
    fc->step_dvr = diag / len_xzy;
    fc->step_mip = diag / len_xyz;
    /* extension: step override; multiplying by 1.0f is the identity */
    fc->step_dvr = fc->step_dvr * p->step_scale;
    fc->step_mip = fc->step_mip * p->step_scale;

    fc->fmin = (float)p->min_val;
    fc->fmax = (float)p->max_val;
    fc->frange = (float)(p->max_val - p->min_val);
}

/* test accessor: the frame constants as 17 floats (pmin, pmax, half_len, denom, step_dvr, step_mip, fmin,
 * fmax, frange) so that the product's host-side evaluation can be compared with this one bit for bit */
void orc_frame_consts(const orc_params* p, float out[17])
{
    frame_consts fc;
    make_frame_consts(p, &fc);
    for (int i = 0; i < 3; ++i) {
        out[i] = fc.pmin[i]; out[3 + i] = fc.pmax[i]; out[6 + i] = fc.half_len[i]; out[9 + i] = fc.denom[i];
    }
    out[12] = fc.step_dvr; out[13] = fc.step_mip; out[14] = fc.fmin; out[15] = fc.fmax; out[16] = fc.frange;
}

/* VR.cs:175-192 */
static inline void cartesian_to_tex(const orc_params* p, const frame_consts* fc,
                                    const float pos[3], float tc[3])
{
    float px = pos[0] + fc->half_len[0];
    float py = pos[1] + fc->half_len[1];
    float pz = pos[2] + fc->half_len[2];
    px = px / fc->denom[0];
    py = py / fc->denom[1];
    pz = pz / fc->denom[2];
    pz = 1.0f - pz;
    if (p->view_top == 1) {
        tc[0] = px; tc[1] = 1.0f - pz; tc[2] = py;
    } else if (p->view_bottom == 1) {
        tc[0] = px; tc[1] = pz; tc[2] = 1.0f - py;
    } else {
        tc[0] = px; tc[1] = py; tc[2] = pz;
    }
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

typedef struct sampler {
    const uint8_t*  v8;
    const uint16_t* v16;
    int nx, ny, nz;
    uint8_t* touch;
    int atomic_touch;
} sampler;

static inline float fetch(const sampler* s, int x, int y, int z)
{
    const uint64_t idx = ((uint64_t)z * (uint64_t)s->ny + (uint64_t)y) * (uint64_t)s->nx + (uint64_t)x;
    if (s->touch) {
        const uint8_t bit = (uint8_t)(1u << (idx & 7));
        if (s->atomic_touch) __atomic_fetch_or(&s->touch[idx >> 3], bit, __ATOMIC_RELAXED);
        else s->touch[idx >> 3] |= bit;
    }
    return s->v8 ? (float)s->v8[idx] : (float)s->v16[idx];
}

/* float -> int with the GPU's F2I behaviour for NaN (0); floor for everything in range. */
static inline int floor_to_int(float f)
{
    if (!(f == f)) return 0;
    const float fl = floorf(f);
    if (fl <= -2147483648.0f) return (int)(-2147483647 - 1);
    if (fl >= 2147483648.0f) return 2147483647;
    return (int)fl;
}

/* VR.cs:121 with the texture state of RendererCore.cpp:408-419.
 * NEAREST  = de-facto behaviour of an integer texture: i = clamp(floor(u*N), 0, N-1).
 * TRILINEAR (extension): f = fma(u, N, -0.5); i0 = floor(f); w = f - i0;
 *   texels i0 and i0+1 clamped to [0,N-1] (CLAMP_TO_EDGE); lerp(a,b,w) = fma(w, b-a, a),
 *   x first, then y, then z, on float-converted voxels. */
static inline float sample_volume(const sampler* s, int filter, const float tc[3])
{
    if (filter == ORC_FILTER_NEAREST) {
        const int ix = clampi(floor_to_int(tc[0] * (float)s->nx), 0, s->nx - 1);
        const int iy = clampi(floor_to_int(tc[1] * (float)s->ny), 0, s->ny - 1);
        const int iz = clampi(floor_to_int(tc[2] * (float)s->nz), 0, s->nz - 1);
        return fetch(s, ix, iy, iz);
    }
    const float fx = fmaf(tc[0], (float)s->nx, -0.5f);
    const float fy = fmaf(tc[1], (float)s->ny, -0.5f);
    const float fz = fmaf(tc[2], (float)s->nz, -0.5f);
    const float flx = floorf(fx), fly = floorf(fy), flz = floorf(fz);
    const float wx = fx - flx, wy = fy - fly, wz = fz - flz;
    const int bx = floor_to_int(fx), by = floor_to_int(fy), bz = floor_to_int(fz);
    const int x0 = clampi(bx, 0, s->nx - 1), x1 = clampi(bx + 1, 0, s->nx - 1);
    const int y0 = clampi(by, 0, s->ny - 1), y1 = clampi(by + 1, 0, s->ny - 1);
    const int z0 = clampi(bz, 0, s->nz - 1), z1 = clampi(bz + 1, 0, s->nz - 1);
    const float v000 = fetch(s, x0, y0, z0), v100 = fetch(s, x1, y0, z0);
    const float v010 = fetch(s, x0, y1, z0), v110 = fetch(s, x1, y1, z0);
    const float v001 = fetch(s, x0, y0, z1), v101 = fetch(s, x1, y0, z1);
    const float v011 = fetch(s, x0, y1, z1), v111 = fetch(s, x1, y1, z1);
    const float c00 = fmaf(wx, v100 - v000, v000);
    const float c10 = fmaf(wx, v110 - v010, v010);
    const float c01 = fmaf(wx, v101 - v001, v001);
    const float c11 = fmaf(wx, v111 - v011, v011);
    const float c0 = fmaf(wy, c10 - c00, c00);
    const float c1 = fmaf(wy, c11 - c01, c01);
    return fmaf(wz, c1 - c0, c0);
}

/* VR.cs:121-124 : clamp + normalise.  Returns the scalar every lane of `src` holds. */
static inline float window_value(const frame_consts* fc, float s)
{
    /* clamp(x,lo,hi) = min(max(x,lo),hi) */
    float v = glsl_min(glsl_max(s, fc->fmin), fc->fmax);
    if (v <= fc->fmax && v >= fc->fmin)          /* VR.cs:123 (false only if min>max) */
        v = (v - fc->fmin) / fc->frange;         /* VR.cs:124 */
    return v;
}

/* extension: opacity from a 256-entry LUT indexed by round(v*255). */
static inline float tf_alpha(const orc_params* p, float v)
{
    const float t = v * 255.0f + 0.5f;
    int iso = floor_to_int(t);
    iso = clampi(iso, 0, 255);
    return p->tf_lut[iso];
}

static inline float opacity_correct(const orc_params* p, float a)
{
    if (p->opacity_correction && p->step_scale != 1.0f)
        return (float)(1.0 - pow(1.0 - (double)a, (double)p->step_scale));
    return a;
}

static void render_pixel(const orc_params* p, const frame_consts* fc, const sampler* smp,
                         int pix_x, int pix_y, float out[4], orc_counters* cnt)
{
    const float* cam = p->cam;
    /* ---- computeRay, VR.cs:194-216 (called with pix + 0.5, VR.cs:86) ---- */
    const float pixel_x = (float)pix_x + 0.5f;
    const float pixel_y = (float)pix_y + 0.5f;
    const float fw = (float)p->width, fh = (float)p->height;
    const float aspect = (fw * 1.0f) / fh;
    float x = aspect * (((2.0f * pixel_x) / fw) - 1.0f);
    float y = ((2.0f * pixel_y) / fh) - 1.0f;
    float z = -cam[20];
    /* normalize(vec4(x,y,z,0)) : v / sqrt(dot(v,v)) */
    float len = sqrtf(x * x + y * y + z * z + 0.0f * 0.0f);
    float dx = x / len, dy = y / len, dz = z / len, dw = 0.0f / len;
    /* view_mat * dir : column-major mat4, component = m[0][i]*v.x + m[1][i]*v.y + ... */
    float mx = cam[0] * dx + cam[4] * dy + cam[8]  * dz + cam[12] * dw;
    float my = cam[1] * dx + cam[5] * dy + cam[9]  * dz + cam[13] * dw;
    float mz = cam[2] * dx + cam[6] * dy + cam[10] * dz + cam[14] * dw;
    float mw = cam[3] * dx + cam[7] * dy + cam[11] * dz + cam[15] * dw;
    len = sqrtf(mx * mx + my * my + mz * mz + mw * mw);
    const float dir[3] = { mx / len, my / len, mz / len };
    const float org[3] = { cam[16], cam[17], cam[18] };

    cnt->rays++;

    /* ---- intersectRayAABB, VR.cs:218-238 ---- */
    float t_max = INFINITY, t_min = -INFINITY;
    const float inv[3] = { 1.0f / dir[0], 1.0f / dir[1], 1.0f / dir[2] };
    const float lo0 = (fc->pmin[0] - org[0]) * inv[0], hi0 = (fc->pmax[0] - org[0]) * inv[0];
    const float lo1 = (fc->pmin[1] - org[1]) * inv[1], hi1 = (fc->pmax[1] - org[1]) * inv[1];
    const float lo2 = (fc->pmin[2] - org[2]) * inv[2], hi2 = (fc->pmax[2] - org[2]) * inv[2];
    int hit;
    t_min = glsl_max(t_min, glsl_min(lo0, hi0));
    t_max = glsl_min(t_max, glsl_max(lo0, hi0));
    t_min = glsl_max(t_min, glsl_min(lo1, hi1));
    t_max = glsl_min(t_max, glsl_max(lo1, hi1));
    if (t_max < t_min) {
        hit = 0;
    } else {
        t_min = glsl_max(t_min, glsl_min(lo2, hi2));
        t_max = glsl_min(t_max, glsl_max(lo2, hi2));
        hit = (t_max > glsl_max(t_min, 0.0f));
    }
    if (!hit) {                                  /* VR.cs:98-101 */
        out[0] = out[1] = out[2] = out[3] = 0.0f;
        return;
    }
    cnt->rays_hit++;

    /* ---- rayMarchVolume VR.cs:104-139 / MIP VR.cs:141-173 ---- */
    const float step = p->is_mip == 1 ? fc->step_mip : fc->step_dvr;
    const float EPSILON = 0.000001f;
    float pos[3];
    for (int i = 0; i < 3; ++i) {
        const float start = org[i] + (dir[i] * t_min);        /* VR.cs:107 */
        pos[i] = start + dir[i] * EPSILON;                    /* VR.cs:114 */
    }
    const float dstep[3] = { dir[0] * step, dir[1] * step, dir[2] * step };

    float C = 0.0f, A = 0.0f;      /* dest.rgb (all equal), dest.a */
    for (int i = 0; i < 10000; ++i) {
        float tc[3];
        cartesian_to_tex(p, fc, pos, tc);
        if (tc[0] > 1.0f || tc[1] > 1.0f || tc[2] > 1.0f ||
            tc[0] < 0.0f || tc[1] < 0.0f || tc[2] < 0.0f || A >= 0.95f)     /* VR.cs:118 */
            break;

        const float s = sample_volume(smp, p->filter, tc);                 /* VR.cs:121 */
        cnt->samples++;
        const float v = window_value(fc, s);                               /* VR.cs:122-124 */
        float src_rgb = v, src_a = v;
        if (p->use_tf) src_a = tf_alpha(p, v);

        if (p->is_mip == 1) {
            src_rgb = src_rgb * p->alpha_scale;                            /* VR.cs:163 */
            src_a = src_a * p->alpha_scale;
            if (A < src_a) { C = src_rgb; A = src_a; }                     /* VR.cs:164-167 */
        } else {
            src_a = src_a * p->alpha_scale;                                /* VR.cs:130 */
            src_a = opacity_correct(p, src_a);
            src_rgb = src_rgb * src_a;                                     /* VR.cs:131 */
            const float t = 1.0f - A;                                      /* VR.cs:132 */
            C = C + src_rgb * t;
            A = A + src_a * t;
            if (A > 0.99f) break;                                          /* VR.cs:134 */
        }
        pos[0] = pos[0] + dstep[0];                                        /* VR.cs:136 */
        pos[1] = pos[1] + dstep[1];
        pos[2] = pos[2] + dstep[2];
    }
    out[0] = out[1] = out[2] = C;
    out[3] = A;
}

typedef struct job {
    const orc_params* p;
    const frame_consts* fc;
    sampler smp;
    float* rgba;
    int next_row;            /* atomic */
    orc_counters cnt;
} job;

typedef struct worker { job* j; orc_counters cnt; } worker;

static void* worker_main(void* arg)
{
    worker* w = (worker*)arg;
    job* j = w->j;
    const orc_params* p = j->p;
    for (;;) {
        const int k = __atomic_fetch_add(&j->next_row, 1, __ATOMIC_RELAXED);
        const long row = (long)p->row_begin + (long)k * (long)p->row_stride;
        if (row >= p->row_end || row >= p->height) break;
        float* out = j->rgba + (size_t)row * (size_t)p->width * 4;
        for (int x = 0; x < p->width; ++x)
            render_pixel(p, j->fc, &j->smp, x, (int)row, out + (size_t)x * 4, &w->cnt);
    }
    return NULL;
}

int orc_render(const orc_params* p, const void* voxels, float* rgba,
               uint8_t* touch, orc_counters* counters, int nthreads)
{
    if (!p || !voxels || !rgba) return -1;
    if (p->width <= 0 || p->height <= 0) return -1;
    if (p->dim[0] <= 0 || p->dim[1] <= 0 || p->dim[2] <= 0) return -1;
    if (p->bytes_per_voxel != 1 && p->bytes_per_voxel != 2) return -1;
    if (p->row_stride <= 0 || p->row_begin < 0) return -1;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;

    frame_consts fc;
    make_frame_consts(p, &fc);

    job j;
    memset(&j, 0, sizeof j);
    j.p = p; j.fc = &fc; j.rgba = rgba;
    j.smp.v8 = p->bytes_per_voxel == 1 ? (const uint8_t*)voxels : NULL;
    j.smp.v16 = p->bytes_per_voxel == 2 ? (const uint16_t*)voxels : NULL;
    j.smp.nx = p->dim[0]; j.smp.ny = p->dim[1]; j.smp.nz = p->dim[2];
    j.smp.touch = touch; j.smp.atomic_touch = nthreads > 1;

    worker w[256];
    pthread_t th[256];
    for (int i = 0; i < nthreads; ++i) { w[i].j = &j; memset(&w[i].cnt, 0, sizeof(orc_counters)); }
    if (nthreads == 1) {
        worker_main(&w[0]);
    } else {
        for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, worker_main, &w[i]);
        for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
    }
    if (counters) {
        memset(counters, 0, sizeof *counters);
        for (int i = 0; i < nthreads; ++i) {
            counters->rays += w[i].cnt.rays;
            counters->rays_hit += w[i].cnt.rays_hit;
            counters->samples += w[i].cnt.samples;
        }
    }
    return 0;
}

uint64_t orc_popcount(const uint8_t* touch, uint64_t nvoxels)
{
    uint64_t n = 0;
    const uint64_t nbytes = nvoxels >> 3;
    for (uint64_t i = 0; i < nbytes; ++i) n += (uint64_t)__builtin_popcount(touch[i]);
    const unsigned rem = (unsigned)(nvoxels & 7);
    if (rem) n += (uint64_t)__builtin_popcount(touch[nbytes] & ((1u << rem) - 1u));
    return n;
}
