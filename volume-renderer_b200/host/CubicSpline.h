// CubicSpline.h -- natural cubic spline through transfer-function control points.
//
// Public surface = what the reference's TF editor calls on its spline object
// (include/CubicSpline.h:7-31; call sites src/UI/elements/AlphaControlSplineWidget.cpp:60,146,153,
// 188,247): the nested control-point type, calcCubicSpline, recomputeCoefficients and the two
// getPointOnSpline overloads -- same names and argument meaning, vr::vec4 instead of glm::vec4.
// bakeAlphaLUT is the extension that turns the opacity curve into the 256-entry table the CUDA
// march reads (vr_params::tf_lut).  Numerics: every vec4 operation is one IEEE binary32 operation
// per lane, in the order of src/CubicSpline.cpp:50-115 (pinned by SURVEY.md 8c's known answers).
#pragma once

#include <string>
#include <vector>

#include "vecmath.h"

class CubicSpline
{
public:
    // one knot of the editor: iso value on the x axis, RGBA (only .w = opacity is used by the march)
    struct TransferFuncControlPoint
    {
        std::string label;
        int iso_value;
        vr::vec4 color;
    };
    typedef std::vector<TransferFuncControlPoint> KnotList;

    CubicSpline();
    ~CubicSpline();

    // fit: forward elimination + back substitution of the natural-spline tridiagonal system
    void calcCubicSpline(const KnotList& control_points);
    // continue the forward-elimination factors from knot `inserted_idx` on (editor inserted a knot)
    void recomputeCoefficients(int inserted_idx, const KnotList& control_points);

    // evaluate at an iso value (exact knots return the knot colour) / at parameter t of one segment
    vr::vec4 getPointOnSpline(int iso_value);
    vr::vec4 getPointOnSpline(float t, float segment_idx);

    // extension: lut[i] = clamp(getPointOnSpline(clamp(i, first knot, last knot)).w, 0, 1)
    void bakeAlphaLUT(float lut[256]);

private:
    // per-segment polynomial y(t) = p0 + t (p1 + t (p2 + t p3)), t in [0,1]
    struct SegmentPoly { vr::vec4 p0, p1, p2, p3; };

    std::vector<TransferFuncControlPoint> knots_;   // copy of the last fitted knot list
    std::vector<vr::vec4> elim_;                     // forward-elimination factors, one per knot
    std::vector<SegmentPoly> segments_;              // one per knot interval
};
