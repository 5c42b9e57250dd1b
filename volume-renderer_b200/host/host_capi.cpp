// host_capi.cpp -- C-linkage test/tool hooks over the C++ host mirror (Camera, CubicSpline,
// VolumeIO, RendererCore) so that Python tests and bench.py can drive it the way the
// reference's RendererGUI drives RendererCore.  Not part of the drop-in boundary
// (that is include/volren_b200.h); see INTEGRATION.md.
#include <cstring>
#include <string>
#include <vector>

#include "Camera.h"
#include "CubicSpline.h"
#include "ImageIO.h"
#include "RendererCore.h"
#include "VolumeIO.h"
#include "../csrc/frame.h"      // host-side evaluation of the per-frame constants (header only)

#define VRH_API extern "C" __attribute__((visibility("default")))

// ------------------------------------------------------------------ Camera
VRH_API Camera* vrh_camera_new(float y_fov, float rot_speed, float mov_speed) { return new Camera(y_fov, rot_speed, mov_speed); }
VRH_API void vrh_camera_free(Camera* c) { delete c; }
VRH_API void vrh_camera_reset(Camera* c) { c->resetCamera(); }
VRH_API void vrh_camera_set_orientation(Camera* c, float zoom, float zenith, float azimuth) { c->setOrientation(zoom, zenith, azimuth); }
VRH_API void vrh_camera_set_spherical(Camera* c, float radius, float zenith, float azimuth) { c->setSpherical(radius, zenith, azimuth); }
VRH_API int vrh_camera_is_changed(Camera* c) { return c->is_changed ? 1 : 0; }
VRH_API void vrh_camera_ubo(Camera* c, float out21[21])
{
    std::vector<float> d;
    c->setUBO(d);
    std::memcpy(out21, d.data(), sizeof(float) * 21);
}

// ------------------------------------------------------------------ CubicSpline
VRH_API CubicSpline* vrh_spline_new(int n, const int* iso, const float* color4)
{
    std::vector<CubicSpline::TransferFuncControlPoint> cps;
    for (int i = 0; i < n; ++i)
        cps.push_back({"K" + std::to_string(i), iso[i], vr::vec4(color4[i * 4], color4[i * 4 + 1], color4[i * 4 + 2], color4[i * 4 + 3])});
    CubicSpline* s = new CubicSpline();
    s->calcCubicSpline(cps);
    return s;
}
VRH_API void vrh_spline_free(CubicSpline* s) { delete s; }
VRH_API void vrh_spline_eval_iso(CubicSpline* s, int iso, float out4[4])
{
    const vr::vec4 v = s->getPointOnSpline(iso);
    out4[0] = v.x; out4[1] = v.y; out4[2] = v.z; out4[3] = v.w;
}
VRH_API void vrh_spline_eval_t(CubicSpline* s, float t, int seg, float out4[4])
{
    const vr::vec4 v = s->getPointOnSpline(t, (float)seg);
    out4[0] = v.x; out4[1] = v.y; out4[2] = v.z; out4[3] = v.w;
}
VRH_API void vrh_spline_bake_alpha_lut(CubicSpline* s, float lut[256]) { s->bakeAlphaLUT(lut); }

// ------------------------------------------------------------------ VolumeIO
struct vrh_pvm { vr::PvmVolume v; std::string error; };
VRH_API vrh_pvm* vrh_pvm_decode(const uint8_t* file, uint64_t bytes)
{
    vrh_pvm* p = new vrh_pvm();
    if (!vr::pvmDecode(file, bytes, p->v, p->error)) p->v.payload.clear();
    return p;
}
VRH_API vrh_pvm* vrh_pvm_read(const char* fn)
{
    vrh_pvm* p = new vrh_pvm();
    if (!vr::pvmReadFile(fn, p->v, p->error)) p->v.payload.clear();
    return p;
}
VRH_API int vrh_pvm_ok(vrh_pvm* p) { return p->error.empty() && p->v.width > 0 ? 1 : 0; }
VRH_API const char* vrh_pvm_error(vrh_pvm* p) { return p->error.c_str(); }
VRH_API void vrh_pvm_header(vrh_pvm* p, uint32_t out_dims_comp_ver[5], float scale[3])
{
    out_dims_comp_ver[0] = p->v.width; out_dims_comp_ver[1] = p->v.height; out_dims_comp_ver[2] = p->v.depth;
    out_dims_comp_ver[3] = p->v.components; out_dims_comp_ver[4] = (uint32_t)p->v.version;
    for (int i = 0; i < 3; ++i) scale[i] = p->v.scale[i];
}
VRH_API uint64_t vrh_pvm_payload_bytes(vrh_pvm* p) { return p->v.payload.size(); }
VRH_API const uint8_t* vrh_pvm_payload(vrh_pvm* p) { return p->v.payload.data(); }
VRH_API const char* vrh_pvm_string(vrh_pvm* p, int which)
{
    switch (which) { case 0: return p->v.description.c_str(); case 1: return p->v.courtesy.c_str();
                     case 2: return p->v.parameter.c_str(); default: return p->v.comment.c_str(); }
}
VRH_API void vrh_pvm_free(vrh_pvm* p) { delete p; }
VRH_API uint32_t vrh_dds_checksum(const uint8_t* data, uint64_t bytes) { return vr::ddsChecksum(data, bytes); }

// read-only memory map of a payload file (the .raw path): size and single bytes at 64-bit offsets
VRH_API vr::MappedFile* vrh_raw_map(const char* fn)
{
    vr::MappedFile* m = new vr::MappedFile();
    std::string err;
    if (!m->open(fn, err)) { delete m; return nullptr; }
    return m;
}
VRH_API uint64_t vrh_raw_size(vr::MappedFile* m) { return m->size(); }
VRH_API int vrh_raw_byte(vr::MappedFile* m, uint64_t offset) { return offset < m->size() ? (int)m->data()[offset] : -1; }
VRH_API void vrh_raw_free(vr::MappedFile* m) { delete m; }
VRH_API int vrh_checked_volume_bytes(const uint64_t dims[3], uint64_t bytes_per_voxel, uint64_t* bytes)
{
    return vr::checkedVolumeBytes(dims, bytes_per_voxel, *bytes) ? 1 : 0;
}

VRH_API int vrh_rawinf_write(const char* raw_fn, const int dims[3], const float spacing[3])
{
    vr::RawInf inf;
    for (int i = 0; i < 3; ++i) { inf.dims[i] = dims[i]; inf.spacing[i] = spacing[i]; }
    return vr::writeRawInf(raw_fn, inf) ? 1 : 0;
}
// returns 1 ok, 0 parse error (title/msg filled), -1 no sidecar
VRH_API int vrh_rawinf_read(const char* raw_fn, int dims[3], float spacing[3], char* title, char* msg, int cap)
{
    vr::RawInf inf; bool exists = false; std::string t, m;
    const bool ok = vr::readRawInf(raw_fn, inf, exists, t, m);
    if (!exists) return -1;
    for (int i = 0; i < 3; ++i) { dims[i] = inf.dims[i]; spacing[i] = inf.spacing[i]; }
    std::strncpy(title, t.c_str(), cap - 1); title[cap - 1] = 0;
    std::strncpy(msg, m.c_str(), cap - 1); msg[cap - 1] = 0;
    return ok ? 1 : 0;
}

VRH_API int vrh_write_image(const char* fn, const char* ext, int w, int h, const uint8_t* rgb)
{
    const std::string e(ext);
    if (e == ".png") return vr::writePNG(fn, w, h, rgb) ? 1 : 0;
    if (e == ".jpg") return vr::writeJPG(fn, w, h, rgb) ? 1 : 0;
    if (e == ".bmp") return vr::writeBMP(fn, w, h, rgb) ? 1 : 0;
    if (e == ".ppm") return vr::writePPM(fn, w, h, rgb) ? 1 : 0;
    return 0;
}

// ------------------------------------------------------------------ RendererCore
VRH_API RendererCore* vrh_core_new(int device, int width, int height)
{
    RendererCore* r = new RendererCore();
    r->cuda_device = device;
    r->window_size = vr::ivec2{width, height};          // RendererGUI.cpp:38-39
    r->framebuffer_size = vr::ivec2{width, height};
    return r;
}
VRH_API void vrh_core_free(RendererCore* r) { delete r; }
VRH_API int vrh_core_setup(RendererCore* r, char* err, int cap)
{
    try { r->setup(); } catch (const std::exception& e) { std::strncpy(err, e.what(), cap - 1); err[cap - 1] = 0; return 0; }
    return 1;
}
VRH_API int vrh_core_load_shader(RendererCore* r, const char* fn) { return r->loadShader(fn, false) ? 1 : 0; }
VRH_API void vrh_core_set_datasize(RendererCore* r, int bytes) { r->datasize_bytes = bytes; }
VRH_API void vrh_core_set_raw_info(RendererCore* r, const int dims[3], const float spacing[3])
{
    r->tex3D_dim = vr::ivec3{dims[0], dims[1], dims[2]};
    r->voxel_size = vr::vec3{spacing[0], spacing[1], spacing[2]};
}
VRH_API int vrh_core_check_raw_inf(RendererCore* r, const char* fn) { return r->checkRawInfFile(fn) ? 1 : 0; }
VRH_API void vrh_core_read_volume(RendererCore* r, const char* fn) { r->readVolumeData(fn); }
VRH_API void vrh_core_render(RendererCore* r) { r->render(); }
VRH_API int vrh_core_read_frame(RendererCore* r, float* rgba)
{
    std::vector<float> f;
    if (!r->readFrame(f)) return 0;
    std::memcpy(rgba, f.data(), f.size() * sizeof(float));
    return 1;
}
VRH_API int vrh_core_save_image(RendererCore* r, const char* fn, const char* ext) { return r->saveImage(fn, ext) ? 1 : 0; }
VRH_API void vrh_core_camera_orient(RendererCore* r, float zoom, float zenith, float azimuth) { r->main_cam.setOrientation(zoom, zenith, azimuth); }
VRH_API void vrh_core_camera_reset(RendererCore* r) { r->main_cam.resetCamera(); }
VRH_API void vrh_core_camera_ubo(RendererCore* r, float out21[21]) { vrh_camera_ubo(&r->main_cam, out21); }
// field mutated first, setter called after -- exactly what RendererGUI.cpp:336-358,382-386 does
VRH_API void vrh_core_gui_alpha(RendererCore* r, float a) { r->alpha_scale = a; r->setAlpha(); }
VRH_API void vrh_core_gui_mip(RendererCore* r, int on) { r->use_mip = on != 0; r->setMIP(); }
VRH_API void vrh_core_gui_min(RendererCore* r, int v) { r->min_val = v; r->setMinVal(); }
VRH_API void vrh_core_gui_max(RendererCore* r, int v) { r->max_val = v; r->setMaxVal(); }
VRH_API void vrh_core_gui_view(RendererCore* r, int top, int bottom)
{
    r->rotate_to_top = top != 0; r->rotate_to_bottom = bottom != 0; r->setInitialCameraRotation();
}
VRH_API void vrh_core_ext_filter(RendererCore* r, int f) { r->setFilter(f); }
VRH_API void vrh_core_ext_step(RendererCore* r, float s, int oc) { r->setStepScale(s, oc != 0); }
VRH_API void vrh_core_ext_tf(RendererCore* r, const float* lut) { r->setTransferFunction(lut); }
VRH_API void vrh_core_ext_kernel(RendererCore* r, int k) { r->setKernel(k); }
VRH_API void vrh_core_get_params(RendererCore* r, vr_params* out) { *out = r->params; }

struct vrh_core_state {
    int tex3D_dim[3]; float voxel_size[3];
    int datasize_bytes, min_val, max_val, min_dataset_val, max_dataset_val;
    int workgroups_x, workgroups_y;
    float alpha_scale, kerneltime_sum;
    float last_kernel_ms; unsigned last_kernel_used;
};
VRH_API void vrh_core_state_get(RendererCore* r, vrh_core_state* s)
{
    s->tex3D_dim[0] = r->tex3D_dim.x; s->tex3D_dim[1] = r->tex3D_dim.y; s->tex3D_dim[2] = r->tex3D_dim.z;
    s->voxel_size[0] = r->voxel_size.x; s->voxel_size[1] = r->voxel_size.y; s->voxel_size[2] = r->voxel_size.z;
    s->datasize_bytes = r->datasize_bytes; s->min_val = r->min_val; s->max_val = r->max_val;
    s->min_dataset_val = r->min_dataset_val; s->max_dataset_val = r->max_dataset_val;
    s->workgroups_x = r->workgroups_x; s->workgroups_y = r->workgroups_y;
    s->alpha_scale = r->alpha_scale; s->kerneltime_sum = r->kerneltime_sum;
    s->last_kernel_ms = r->last_stats.kernel_ms; s->last_kernel_used = r->last_stats.kernel_used;
}
VRH_API void vrh_core_reset_kerneltime(RendererCore* r) { r->kerneltime_sum = 0; }     // RendererGUI.cpp:58-60
VRH_API void vrh_core_strings(RendererCore* r, char* title, char* msg, char* dataset, char* shader, int cap)
{
    auto cp = [&](char* d, const std::string& s) { std::strncpy(d, s.c_str(), cap - 1); d[cap - 1] = 0; };
    cp(title, r->title); cp(msg, r->msg); cp(dataset, r->loaded_dataset); cp(shader, r->loaded_shader);
}
VRH_API void vrh_core_clear_popup(RendererCore* r) { r->title.clear(); r->msg.clear(); }   // RendererGUI.cpp:90-96
VRH_API void vrh_core_histogram(RendererCore* r, float out[256]) { std::memcpy(out, r->histogram.data(), 256 * sizeof(float)); }

// ---- per-frame constants as the product evaluates them on the host (csrc/frame.h), for the CPU tests:
//      pmin[3], pmax[3], half_len[3], denom[3], step, fmin, fmax, frange, inv_denom[3], inv_frange, tc_div_mode
VRH_API void vrh_frame_consts(int W, int H, const int32_t dim[3], const float voxel_size[3], const float cam21[21],
                              const vr_params* p, float out[22])
{
    vr::FrameConsts fc;
    std::memset(&fc, 0, sizeof fc);
    vr::compute_frame_consts(fc, W, H, dim, voxel_size, cam21, *p);
    for (int i = 0; i < 3; ++i) {
        out[i] = fc.pmin[i]; out[3 + i] = fc.pmax[i]; out[6 + i] = fc.half_len[i]; out[9 + i] = fc.denom[i];
        out[16 + i] = fc.inv_denom[i];
    }
    out[12] = fc.step; out[13] = fc.fmin; out[14] = fc.fmax; out[15] = fc.frange;
    out[19] = fc.inv_frange; out[20] = (float)fc.tc_div_mode; out[21] = (float)fc.win_div_mode;
}
