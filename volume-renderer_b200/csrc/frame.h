// frame.h -- per-frame constants of the ray march, computed once on the host.
//
// Everything in VolumeRenderer.cs that is uniform over the dispatch (bounding box :62-83,
// step size :109/:146, the tex-coord denominator :179, the float forms of the window
// uniforms :122-124) is evaluated here with one IEEE binary32 operation per GLSL operator,
// in source order.  This translation unit must be compiled WITHOUT floating-point
// contraction (-ffp-contract=off); the device code only consumes the results.
#pragma once

#include <cmath>
#include <cstdint>

#include "volren_b200.h"

namespace vr {

enum DivMode : int {
    DIV_RECIP_EXACT = 0,   // every divisor is a power of two: q * (1/D) is exact
    DIV_MARKSTEIN = 1,     // q*y ; r = fma(-D,q0,q) ; q1 = fma(r,y,q0); verified on device
    DIV_IEEE = 2           // div.rn.f32
};

struct FrameConsts {
    // image / partition
    int32_t W, H;
    int32_t rank, world, tile_rows;
    int32_t compact;              // 1: output holds owned rows only, packed
    int32_t row0;                 // first local (owned) row of this launch: a band of the partition
    // camera block (Camera::setUBO, Camera.cpp:59-80)
    float cam[21];
    // bounding box, VolumeRenderer.cs:62-83
    float pmin[3], pmax[3], half_len[3];
    float denom[3];               // bb.p_max + half_len, VolumeRenderer.cs:179
    float inv_denom[3];           // RN(1/denom)
    float step;                   // VolumeRenderer.cs:109 (DVR) or :146 (MIP), * step_scale
    // volume
    int32_t dim[3];
    float dimf[3];
    // window, VolumeRenderer.cs:122-124
    float fmin, fmax, frange, inv_frange;
    int32_t window_ordered;       // min_val <= max_val (else :123 is false for every sample)
    // uniforms / extensions
    float alpha_scale;
    float step_scale;
    int32_t is_mip, view_top, view_bottom;
    int32_t filter, use_tf, opacity_correction;
    int32_t tc_div_mode, win_div_mode;
};

inline bool is_pow2_float(float v)
{
    if (!(v > 0.0f) || std::isinf(v)) return false;
    int e = 0;
    return std::frexp(v, &e) == 0.5f;
}

// host-side evaluation; mirrors oracle/march_oracle.c:make_frame_consts operation by operation
inline void compute_frame_consts(FrameConsts& fc, int W, int H, const int32_t dim[3],
                                 const float voxel_size[3], const float cam[21], const vr_params& p)
{
    fc.W = W; fc.H = H;
    for (int i = 0; i < 21; ++i) fc.cam[i] = cam[i];

    int max_dim = dim[0] > dim[1] ? dim[0] : dim[1];
    max_dim = max_dim > dim[2] ? max_dim : dim[2];
    const bool swz = (p.view_bottom == 1 || p.view_top == 1);
    float n[3], vs[3];
    if (swz) {
        n[0] = (float)dim[0]; n[1] = (float)dim[2]; n[2] = (float)dim[1];
        vs[0] = voxel_size[0]; vs[1] = voxel_size[2]; vs[2] = voxel_size[1];
    } else {
        n[0] = (float)dim[0]; n[1] = (float)dim[1]; n[2] = (float)dim[2];
        vs[0] = voxel_size[0]; vs[1] = voxel_size[1]; vs[2] = voxel_size[2];
    }
    const float fmax_dim = (float)max_dim;
    bool all_pow2 = true;
    for (int i = 0; i < 3; ++i) {
        float pm = n[i] / fmax_dim;
        pm = pm * vs[i];
        float h = pm / 2.0f;
        fc.half_len[i] = h;
        fc.pmin[i] = 0.0f - h;
        float pmx = pm - h;
        fc.pmax[i] = pmx;
        float d = pmx + h;
        fc.denom[i] = d;
        fc.inv_denom[i] = 1.0f / d;
        all_pow2 = all_pow2 && is_pow2_float(d);
    }
    float ex = fc.pmax[0] - fc.pmin[0];
    float ey = fc.pmax[1] - fc.pmin[1];
    float ez = fc.pmax[2] - fc.pmin[2];
    float exx = ex * ex, eyy = ey * ey, ezz = ez * ez;
    float dsum = exx + eyy;
    dsum = dsum + ezz;
    const float diag = std::sqrt(dsum);
    const float fx = (float)dim[0], fy = (float)dim[1], fz = (float)dim[2];
    float xx = fx * fx, yy = fy * fy, zz = fz * fz;
    float l_xzy = xx + zz; l_xzy = l_xzy + yy;      // length(vol_size.xzy), :109
    float l_xyz = xx + yy; l_xyz = l_xyz + zz;      // length(vol_size.xyz), :146
    const float len = std::sqrt((float)(p.is_mip == 1 ? l_xyz : l_xzy));
    float step = diag / len;
    step = step * p.step_scale;
    fc.step = step;

    for (int i = 0; i < 3; ++i) { fc.dim[i] = dim[i]; fc.dimf[i] = (float)dim[i]; }
    fc.fmin = (float)p.min_val;
    fc.fmax = (float)p.max_val;
    fc.frange = (float)(p.max_val - p.min_val);
    fc.inv_frange = 1.0f / fc.frange;
    fc.window_ordered = p.min_val <= p.max_val;
    fc.alpha_scale = p.alpha_scale;
    fc.step_scale = p.step_scale;
    fc.is_mip = p.is_mip == 1; fc.view_top = p.view_top == 1; fc.view_bottom = p.view_bottom == 1;
    fc.filter = p.filter; fc.use_tf = p.use_tf != 0;
    fc.opacity_correction = (p.opacity_correction != 0) && (p.step_scale != 1.0f);
    fc.tc_div_mode = all_pow2 ? DIV_RECIP_EXACT : DIV_MARKSTEIN;   // MARKSTEIN is verified later
    fc.win_div_mode = (p.max_val > p.min_val) ? DIV_MARKSTEIN : DIV_IEEE;
}

}  // namespace vr
