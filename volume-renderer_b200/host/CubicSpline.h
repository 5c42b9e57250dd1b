// CubicSpline.h -- natural cubic spline through transfer-function control points, the
// reference's include/CubicSpline.h:7-31 surface (calcCubicSpline / getPointOnSpline), plus
// bakeAlphaLUT which turns the opacity curve into the 256-entry table the CUDA march reads.
#pragma once

#include <string>
#include <vector>

#include "vecmath.h"

class CubicSpline
{
    public:
        CubicSpline();
        ~CubicSpline();

        struct TransferFuncControlPoint
        {
            std::string label;
            int iso_value;
            vr::vec4 color;
        };
        void calcCubicSpline(const std::vector<TransferFuncControlPoint>& control_points);
        void recomputeCoefficients(int inserted_idx, const std::vector<TransferFuncControlPoint>& control_points);
        vr::vec4 getPointOnSpline(int iso_value);
        vr::vec4 getPointOnSpline(float t, float segment_idx);

        // extension: lut[i] = clamp(getPointOnSpline(clamp(i, first knot, last knot)).w, 0, 1)
        void bakeAlphaLUT(float lut[256]);

    private:
        struct CubicCoefficiants { vr::vec4 a, b, c, d; };
        std::vector<vr::vec4> coeffs;
        std::vector<CubicCoefficiants> spline;
        std::vector<TransferFuncControlPoint> control_points;
};
