"""Second, independent restatement of VolumeRenderer.cs for single pixels in numpy float32
scalars (slow Python loops; small cases only).  Written from the shader text, not from
oracle/march_oracle.c, to catch transcription errors in the C oracle."""
import math

import numpy as np

f32 = np.float32


def _gmin(x, y):
    return y if y < x else x


def _gmax(x, y):
    return y if x < y else x


def shade_pixel(px, py, W, H, dims, vox3d, cam, *, voxel_size=(1, 1, 1), alpha_scale=1.0, min_val=0, max_val=255,
                is_mip=0, view_top=0, view_bottom=0, trilinear=False, step_scale=1.0, tf_lut=None):
    """vox3d indexed [z][y][x].  Returns (r, g, b, a) as float32."""
    with np.errstate(all="ignore"):
        N = [int(d) for d in dims]
        cam = [f32(c) for c in cam]
        vs = [f32(v) for v in voxel_size]
        max_dim = max(N)
        swz = view_top == 1 or view_bottom == 1
        order = (0, 2, 1) if swz else (0, 1, 2)
        p_max = [f32(N[o]) / f32(max_dim) * vs[o] for o in order]
        half = [p / f32(2.0) for p in p_max]
        p_min = [f32(0) - h for h in half]
        p_max = [p - h for p, h in zip(p_max, half)]

        # computeRay
        pixel_x, pixel_y = f32(px) + f32(0.5), f32(py) + f32(0.5)
        aspect = (f32(W) * f32(1.0)) / f32(H)
        x = aspect * ((f32(2.0) * pixel_x / f32(W)) - f32(1))
        y = (f32(2.0) * pixel_y / f32(H)) - f32(1)
        z = -cam[20]
        ln = f32(math.sqrt(f32(f32(f32(x * x + y * y) + z * z) + f32(0) * f32(0)))) if False else np.sqrt(f32(x * x + y * y + z * z + f32(0) * f32(0)))
        d = [x / ln, y / ln, z / ln, f32(0) / ln]
        m = []
        for r in range(4):
            m.append(f32(f32(f32(cam[0 + r] * d[0] + cam[4 + r] * d[1]) + cam[8 + r] * d[2]) + cam[12 + r] * d[3]))
        ln = np.sqrt(f32(f32(f32(m[0] * m[0] + m[1] * m[1]) + m[2] * m[2]) + m[3] * m[3]))
        direc = [m[0] / ln, m[1] / ln, m[2] / ln]
        org = [cam[16], cam[17], cam[18]]

        # intersectRayAABB
        t_max, t_min = f32(np.inf), f32(-np.inf)
        inv = [f32(1) / c for c in direc]
        lo = [(p_min[i] - org[i]) * inv[i] for i in range(3)]
        hi = [(p_max[i] - org[i]) * inv[i] for i in range(3)]
        t_min = _gmax(t_min, _gmin(lo[0], hi[0])); t_max = _gmin(t_max, _gmax(lo[0], hi[0]))
        t_min = _gmax(t_min, _gmin(lo[1], hi[1])); t_max = _gmin(t_max, _gmax(lo[1], hi[1]))
        if t_max < t_min:
            return (f32(0),) * 4
        t_min = _gmax(t_min, _gmin(lo[2], hi[2])); t_max = _gmin(t_max, _gmax(lo[2], hi[2]))
        if not (t_max > _gmax(t_min, f32(0))):
            return (f32(0),) * 4

        diag = np.sqrt(f32(f32((p_max[0] - p_min[0]) ** 2 + (p_max[1] - p_min[1]) ** 2) + (p_max[2] - p_min[2]) ** 2))
        fx, fy, fz = f32(N[0]), f32(N[1]), f32(N[2])
        if is_mip == 1:
            vlen = np.sqrt(f32(f32(fx * fx + fy * fy) + fz * fz))
        else:
            vlen = np.sqrt(f32(f32(fx * fx + fz * fz) + fy * fy))
        step = f32(diag / vlen) * f32(step_scale)
        EPS = f32(0.000001)
        pos = [f32(f32(org[i] + direc[i] * t_min) + direc[i] * EPS) for i in range(3)]
        dstep = [direc[i] * step for i in range(3)]
        denom = [p_max[i] + half[i] for i in range(3)]

        def fetch(ix, iy, iz):
            ix = min(max(ix, 0), N[0] - 1); iy = min(max(iy, 0), N[1] - 1); iz = min(max(iz, 0), N[2] - 1)
            return f32(vox3d[iz][iy][ix])

        C, A = f32(0), f32(0)
        fmin, fmax = f32(min_val), f32(max_val)
        for _ in range(10000):
            p = [f32(pos[i] + half[i]) / denom[i] for i in range(3)]
            p[2] = f32(1) - p[2]
            if view_top == 1:
                tc = [p[0], f32(1) - p[2], p[1]]
            elif view_bottom == 1:
                tc = [p[0], p[2], f32(1) - p[1]]
            else:
                tc = p
            if any(c > f32(1) for c in tc) or any(c < f32(0) for c in tc) or A >= f32(0.95):
                break
            if not trilinear:
                idx = [int(math.floor(tc[i] * f32(N[i]))) for i in range(3)]
                s = fetch(*idx)
            else:
                f = [f32(np.float64(tc[i]) * np.float64(N[i]) - 0.5) for i in range(3)]      # fma = one rounding
                b = [int(math.floor(v)) for v in f]
                w = [f32(f[i] - f32(math.floor(f[i]))) for i in range(3)]
                def lerp(a, bb, ww):
                    return f32(np.float64(ww) * np.float64(f32(bb - a)) + np.float64(a))
                c = [[lerp(fetch(b[0], b[1] + j, b[2] + k), fetch(b[0] + 1, b[1] + j, b[2] + k), w[0]) for j in (0, 1)] for k in (0, 1)]
                c0 = lerp(c[0][0], c[0][1], w[1]); c1 = lerp(c[1][0], c[1][1], w[1])
                s = lerp(c0, c1, w[2])
            v = _gmin(_gmax(s, fmin), fmax)
            if v <= fmax and v >= fmin:
                v = f32(v - fmin) / f32(max_val - min_val)
            rgb, a = v, v
            if tf_lut is not None:
                iso = int(math.floor(f32(f32(v * f32(255.0)) + f32(0.5))))
                a = f32(tf_lut[min(max(iso, 0), 255)])
            if is_mip == 1:
                rgb = rgb * f32(alpha_scale); a = a * f32(alpha_scale)
                if A < a:
                    C, A = rgb, a
            else:
                a = a * f32(alpha_scale)
                rgb = rgb * a
                t = f32(1) - A
                C = f32(C + rgb * t)
                A = f32(A + a * t)
                if A > f32(0.99):
                    break
            pos = [f32(pos[i] + dstep[i]) for i in range(3)]
        return (C, C, C, A)
