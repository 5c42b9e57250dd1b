/*
 * oracle/spline_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
 *
 * Restatement of /root/reference/src/CubicSpline.cpp ("CS:n"): natural cubic spline through
 * N+1 control points in vec4, one rounded fp32 operation per glm operator, source order.
 * Default knots used by the tests: AlphaControlSplineWidget.cpp:56-59.
 * PINNED to the reference source: src/CubicSpline.cpp compiled unmodified into oracle/_ref/libhost_ref.so
 * (tests/test_reference_pinning.py compares every iso value of random knot sets bit for bit).
 */
#include "oracle.h"

#include <string.h>

int orc_spline_calc(orc_spline* s, int n_points, const int32_t* iso, const float* color4)   /* CS:50-115 */
{
    if (n_points < 2 || n_points > ORC_SPLINE_MAX_KNOTS) return -1;
    memset(s, 0, sizeof *s);
    s->n_points = n_points;
    for (int i = 0; i < n_points; ++i) {
        s->iso[i] = iso[i];
        for (int k = 0; k < 4; ++k) s->color[i][k] = color4[i * 4 + k];
    }
    const int n = n_points - 1;
    float deltas[ORC_SPLINE_MAX_KNOTS][4];

    for (int k = 0; k < 4; ++k) {
        /* CS:88-95 */
        s->coeffs[0][k] = 0.5f;
        for (int i = 1; i < n; ++i)
            s->coeffs[i][k] = 1.0f / ((4.0f * 1.0f) - s->coeffs[i - 1][k]);
        s->coeffs[n][k] = 1.0f / ((2.0f * 1.0f) - s->coeffs[n - 1][k]);

        /* CS:97-101 */
        deltas[0][k] = 3.0f * (s->color[1][k] - s->color[0][k]) * s->coeffs[0][k];
        for (int i = 1; i < n; ++i)
            deltas[i][k] = (3.0f * (s->color[i + 1][k] - s->color[i - 1][k]) - deltas[i - 1][k]) * s->coeffs[i][k];
        deltas[n][k] = (3.0f * (s->color[n][k] - s->color[n - 1][k]) - deltas[n - 1][k]) * s->coeffs[n][k];

        /* CS:103-105 */
        s->deriv[n][k] = deltas[n][k];
        for (int i = n - 1; i >= 0; --i)
            s->deriv[i][k] = deltas[i][k] - s->coeffs[i][k] * s->deriv[i + 1][k];

        /* CS:110-114 */
        for (int i = 0; i < n; ++i) {
            s->a[i][k] = s->color[i][k];
            s->b[i][k] = s->deriv[i][k];
            s->c[i][k] = 3.0f * (s->color[i + 1][k] - s->color[i][k]) - 2.0f * s->deriv[i][k] - s->deriv[i + 1][k];
            s->d[i][k] = 2.0f * (s->color[i][k] - s->color[i + 1][k]) + s->deriv[i][k] + s->deriv[i + 1][k];
        }
    }
    return 0;
}

void orc_spline_eval_t(const orc_spline* s, float t, int seg, float out[4])     /* CS:13-18 */
{
    for (int k = 0; k < 4; ++k)
        out[k] = s->a[seg][k] + t * (s->b[seg][k] + t * (s->c[seg][k] + t * s->d[seg][k]));
}

void orc_spline_eval_iso(const orc_spline* s, int iso_val, float out[4])        /* CS:20-40 */
{
    float t = 0;
    int seg = 0;
    for (int i = 0; i < s->n_points; ++i) {
        if (s->iso[i] == iso_val) {
            memcpy(out, s->color[i], sizeof(float) * 4);
            return;
        } else if (s->iso[i] > iso_val) {
            if (i == 0) break;   /* reference indexes control_points[-1] (UB); callers clamp */
            seg = i - 1;
            t = (float)(iso_val - s->iso[i - 1]) / (float)(s->iso[i] - s->iso[i - 1]);
            break;
        }
    }
    orc_spline_eval_t(s, t, seg, out);
}

void orc_spline_bake_alpha_lut(const orc_spline* s, float lut[256])
{
    const int lo = s->iso[0], hi = s->iso[s->n_points - 1];
    for (int i = 0; i < 256; ++i) {
        int iso = i < lo ? lo : (i > hi ? hi : i);
        float c[4];
        orc_spline_eval_iso(s, iso, c);
        float a = c[3];
        a = a < 0.0f ? 0.0f : (a > 1.0f ? 1.0f : a);      /* AlphaControlSplineWidget.cpp:247 */
        lut[i] = a;
    }
}
