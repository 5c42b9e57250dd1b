// CubicSpline.cpp -- see CubicSpline.h.  Reference: src/CubicSpline.cpp ("CS:n").
#include "CubicSpline.h"

using vr::vec4;

CubicSpline::CubicSpline() {}
CubicSpline::~CubicSpline() {}

vec4 CubicSpline::getPointOnSpline(float t, float segment_idx)                       // CS:13-18
{
    const SegmentPoly& k = segments_[(size_t)segment_idx];
    return k.p0 + t * (k.p1 + t * (k.p2 + t * k.p3));
}

vec4 CubicSpline::getPointOnSpline(int iso_val)                                      // CS:20-40
{
    float t = 0;
    int segment_idx = 0;
    const int n = (int)knots_.size();
    for (int i = 0; i < n; i++) {
        if (knots_[i].iso_value == iso_val) return knots_[i].color;
        if (knots_[i].iso_value > iso_val) {
            if (i == 0) break;   // below the first knot: the reference reads knots_[-1]
            segment_idx = i - 1;
            t = (iso_val - knots_[i - 1].iso_value) /
                (float)(knots_[i].iso_value - knots_[i - 1].iso_value);
            break;
        }
    }
    return getPointOnSpline(t, (float)segment_idx);
}

void CubicSpline::recomputeCoefficients(int inserted_idx, const KnotList& cps)   // CS:42-48
{
    const vec4 one(1.0f);
    for (int i = inserted_idx; i < (int)cps.size() - 1; i++)
        elim_[i] = one / ((4.0f * one) - elim_[i - 1]);
    elim_.push_back(one / ((2.0f * one) - elim_.back()));
}

void CubicSpline::calcCubicSpline(const KnotList& cp_list)   // CS:50-115
{
    knots_ = cp_list;
    const int n = (int)knots_.size() - 1;     // n segments
    segments_.clear();
    elim_.clear();
    if (n < 1) return;
    std::vector<vec4> deltas(n + 1), derivs(n + 1);

    // forward elimination of the tridiagonal system [2 1; 1 4 1; ...; 1 2] D = 3 * dy
    const vec4 one(1.0f);
    elim_.push_back(vec4(0.5f));
    for (int i = 1; i < n; i++) elim_.push_back(one / ((4.0f * one) - elim_[i - 1]));
    elim_.push_back(one / ((2.0f * one) - elim_[n - 1]));

    deltas[0] = 3.0f * (knots_[1].color - knots_[0].color) * elim_[0];
    for (int i = 1; i < n; i++)
        deltas[i] = (3.0f * (knots_[i + 1].color - knots_[i - 1].color) - deltas[i - 1]) * elim_[i];
    deltas[n] = (3.0f * (knots_[n].color - knots_[n - 1].color) - deltas[n - 1]) * elim_[n];

    // back substitution
    derivs[n] = deltas[n];
    for (int i = n - 1; i >= 0; i--) derivs[i] = deltas[i] - elim_[i] * derivs[i + 1];

    for (int i = 0; i < n; i++) {
        const vec4& y0 = knots_[i].color;
        const vec4& y1 = knots_[i + 1].color;
        SegmentPoly k;
        k.p0 = y0;
        k.p1 = derivs[i];
        k.p2 = 3.0f * (y1 - y0) - 2.0f * derivs[i] - derivs[i + 1];
        k.p3 = 2.0f * (y0 - y1) + derivs[i] + derivs[i + 1];
        segments_.push_back(k);
    }
}

void CubicSpline::bakeAlphaLUT(float lut[256])
{
    if (knots_.size() < 2) { for (int i = 0; i < 256; ++i) lut[i] = 0.0f; return; }
    const int lo = knots_.front().iso_value, hi = knots_.back().iso_value;
    for (int i = 0; i < 256; ++i) {
        const int iso = i < lo ? lo : (i > hi ? hi : i);
        float a = getPointOnSpline(iso).w;
        a = a < 0.0f ? 0.0f : (a > 1.0f ? 1.0f : a);     // AlphaControlSplineWidget.cpp:247
        lut[i] = a;
    }
}
