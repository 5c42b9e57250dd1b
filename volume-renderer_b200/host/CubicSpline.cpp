// CubicSpline.cpp -- see CubicSpline.h.  Reference: src/CubicSpline.cpp ("CS:n").
#include "CubicSpline.h"

using vr::vec4;

CubicSpline::CubicSpline() {}
CubicSpline::~CubicSpline() {}

vec4 CubicSpline::getPointOnSpline(float t, float segment_idx)                       // CS:13-18
{
    const CubicCoefficiants& k = spline[(size_t)segment_idx];
    return k.a + t * (k.b + t * (k.c + t * k.d));
}

vec4 CubicSpline::getPointOnSpline(int iso_val)                                      // CS:20-40
{
    float t = 0;
    int segment_idx = 0;
    const int n = (int)control_points.size();
    for (int i = 0; i < n; i++) {
        if (control_points[i].iso_value == iso_val) return control_points[i].color;
        if (control_points[i].iso_value > iso_val) {
            if (i == 0) break;   // below the first knot: the reference reads control_points[-1]
            segment_idx = i - 1;
            t = (iso_val - control_points[i - 1].iso_value) /
                (float)(control_points[i].iso_value - control_points[i - 1].iso_value);
            break;
        }
    }
    return getPointOnSpline(t, (float)segment_idx);
}

void CubicSpline::recomputeCoefficients(int inserted_idx, const std::vector<TransferFuncControlPoint>& cps)   // CS:42-48
{
    const vec4 one(1.0f);
    for (int i = inserted_idx; i < (int)cps.size() - 1; i++)
        coeffs[i] = one / ((4.0f * one) - coeffs[i - 1]);
    coeffs.push_back(one / ((2.0f * one) - coeffs.back()));
}

void CubicSpline::calcCubicSpline(const std::vector<TransferFuncControlPoint>& cp_list)   // CS:50-115
{
    control_points = cp_list;
    const int n = (int)control_points.size() - 1;     // n segments
    spline.clear();
    coeffs.clear();
    if (n < 1) return;
    std::vector<vec4> deltas(n + 1), derivs(n + 1);

    // forward elimination of the tridiagonal system [2 1; 1 4 1; ...; 1 2] D = 3 * dy
    const vec4 one(1.0f);
    coeffs.push_back(vec4(0.5f));
    for (int i = 1; i < n; i++) coeffs.push_back(one / ((4.0f * one) - coeffs[i - 1]));
    coeffs.push_back(one / ((2.0f * one) - coeffs[n - 1]));

    deltas[0] = 3.0f * (control_points[1].color - control_points[0].color) * coeffs[0];
    for (int i = 1; i < n; i++)
        deltas[i] = (3.0f * (control_points[i + 1].color - control_points[i - 1].color) - deltas[i - 1]) * coeffs[i];
    deltas[n] = (3.0f * (control_points[n].color - control_points[n - 1].color) - deltas[n - 1]) * coeffs[n];

    // back substitution
    derivs[n] = deltas[n];
    for (int i = n - 1; i >= 0; i--) derivs[i] = deltas[i] - coeffs[i] * derivs[i + 1];

    for (int i = 0; i < n; i++) {
        const vec4& y0 = control_points[i].color;
        const vec4& y1 = control_points[i + 1].color;
        CubicCoefficiants k;
        k.a = y0;
        k.b = derivs[i];
        k.c = 3.0f * (y1 - y0) - 2.0f * derivs[i] - derivs[i + 1];
        k.d = 2.0f * (y0 - y1) + derivs[i] + derivs[i + 1];
        spline.push_back(k);
    }
}

void CubicSpline::bakeAlphaLUT(float lut[256])
{
    if (control_points.size() < 2) { for (int i = 0; i < 256; ++i) lut[i] = 0.0f; return; }
    const int lo = control_points.front().iso_value, hi = control_points.back().iso_value;
    for (int i = 0; i < 256; ++i) {
        const int iso = i < lo ? lo : (i > hi ? hi : i);
        float a = getPointOnSpline(iso).w;
        a = a < 0.0f ? 0.0f : (a > 1.0f ? 1.0f : a);     // AlphaControlSplineWidget.cpp:247
        lut[i] = a;
    }
}
