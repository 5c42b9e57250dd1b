// RendererCore.cpp -- see RendererCore.h.  Reference: src/RendererCore.cpp ("RC:n").
#include "RendererCore.h"

#include <exception>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <stdexcept>

#include "ImageIO.h"
#include "VolumeIO.h"

RendererCore::RendererCore() : main_cam(30), histogram(256, 0.0f)                     // RC:13-26
{
    voxel_size = vr::vec3{1.0f, 1.0f, 1.0f};
    tex3D_dim = vr::ivec3{0, 0, 0};
    alpha_scale = 1;
    min_val = 0;
    max_val = 0;
    max_dataset_val = min_dataset_val = 0;
    datasize_bytes = -1;
    kerneltime_sum = 0.0;
    workgroups_x = workgroups_y = 0;
    use_mip = rotate_to_bottom = rotate_to_top = false;
    cuda_device = 0;
    ctx = nullptr;
    vr_params_default(&params);
    std::memset(&last_stats, 0, sizeof last_stats);
}

RendererCore::~RendererCore()                                                         // RC:28-32
{
    if (ctx) vr_destroy(ctx);
}

void RendererCore::reportAbiError(const char* title_text)
{
    title = title_text;
    msg = vr_last_error();
}

void RendererCore::setup()                                                            // RC:34-44
{
    setupFBO();
}

void RendererCore::setupFBO()                                                         // RC:184-219
{
    if (ctx) { vr_destroy(ctx); ctx = nullptr; }
    if (vr_create(cuda_device, framebuffer_size.x, framebuffer_size.y, &ctx) != VR_OK)
        throw std::runtime_error(std::string("Framebuffer not complete. ") + vr_last_error());
}

bool RendererCore::checkRawInfFile(std::string fn)                                    // RC:46-54
{
    std::ifstream inf_file(fn + ".inf");
    return (bool)inf_file;
}

void RendererCore::pushParams()
{
    if (ctx && vr_set_params(ctx, &params) != VR_OK) reportAbiError("Error!");
}

void RendererCore::setAlpha()                                                         // RC:56-60
{
    if (!loaded_shader.empty()) { params.alpha_scale = alpha_scale; pushParams(); }
}

void RendererCore::setMinVal()                                                        // RC:62-71
{
    if (!loaded_shader.empty()) {
        params.min_val = (datasize_bytes == 2) ? min_val + 1000 : min_val;
        pushParams();
    }
}

void RendererCore::setMaxVal()                                                        // RC:73-82
{
    if (!loaded_shader.empty()) {
        params.max_val = (datasize_bytes == 2) ? max_val + 1000 : max_val;
        pushParams();
    }
}

void RendererCore::setMIP()                                                           // RC:84-88
{
    if (!loaded_shader.empty()) { params.is_mip = use_mip ? 1 : 0; pushParams(); }
}

void RendererCore::setInitialCameraRotation()                                         // RC:90-98
{
    if (!loaded_shader.empty()) {
        main_cam.resetCamera();
        params.view_top = rotate_to_top ? 1 : 0;
        params.view_bottom = rotate_to_bottom ? 1 : 0;
        pushParams();
    }
}

void RendererCore::setUniforms()                                                      // RC:100-110
{
    if (!loaded_shader.empty() && ctx) {
        const float vs[3] = {voxel_size.x, voxel_size.y, voxel_size.z};
        if (vr_set_voxel_size(ctx, vs) != VR_OK) reportAbiError("Error!");
    }
    setAlpha();
    setMinVal();
    setMaxVal();
    setMIP();
    setInitialCameraRotation();
}

bool RendererCore::loadShader(std::string fn, bool /*reload*/)                        // RC:112-136
{
    // The march kernel is compiled into libvolren_b200.so; the name is only recorded so that
    // the GUI's "shader loaded" gating (RendererGUI.cpp:150,162,167) keeps working.
    if (fn.empty()) { loaded_shader.clear(); return false; }
    loaded_shader = fn;
    const int workgroup_size[2] = {16, 16};                     // VolumeRenderer.cs:3
    workgroups_x = window_size.x / workgroup_size[0];           // RC:121-122 (display only)
    workgroups_y = window_size.y / workgroup_size[1];
    params.alpha_scale = alpha_scale;                           // RC:125
    pushParams();
    if (!loaded_dataset.empty()) {
        setUniforms();
        main_cam.resetCamera();
    }
    return true;
}

void RendererCore::setupUBO(bool /*is_update*/)                                       // RC:221-240
{
    std::vector<float> cam_data;
    main_cam.setUBO(cam_data);
    if (ctx && vr_set_camera(ctx, cam_data.data()) != VR_OK) reportAbiError("Error!");
}

void RendererCore::render()                                                           // RC:138-163
{
    if (!ctx) return;
    if (main_cam.is_changed) setupUBO(true);
    vr_render_stats st;
    std::memset(&st, 0, sizeof st);
    if (vr_render_device(ctx, nullptr, 0, nullptr, &st) != VR_OK) { reportAbiError("Render failed"); return; }
    last_stats = st;
    kerneltime_sum += st.kernel_ms;                             // RC:153 (milliseconds)
}

bool RendererCore::readFrame(std::vector<float>& rgba)
{
    if (!ctx) return false;
    rgba.resize((size_t)framebuffer_size.x * framebuffer_size.y * 4);
    if (vr_read_frame(ctx, rgba.data()) != VR_OK) { reportAbiError("Error!"); return false; }
    return true;
}

bool RendererCore::saveImage(std::string fn, std::string ext)                         // RC:165-182
{
    if (!ctx) return false;
    std::vector<uint8_t> rgb((size_t)framebuffer_size.x * framebuffer_size.y * 3);
    if (vr_read_rgb8(ctx, rgb.data(), /*flip_vertical=*/1) != VR_OK) { reportAbiError("Error!"); return false; }
    if (ext == ".png") return vr::writePNG(fn, framebuffer_size.x, framebuffer_size.y, rgb.data());
    if (ext == ".bmp") return vr::writeBMP(fn, framebuffer_size.x, framebuffer_size.y, rgb.data());
    if (ext == ".jpg") return vr::writeJPG(fn, framebuffer_size.x, framebuffer_size.y, rgb.data());   // quality 100, RC:176
    if (ext == ".ppm") return vr::writePPM(fn, framebuffer_size.x, framebuffer_size.y, rgb.data());
    return false;
}

void RendererCore::readVolumeData(std::string fn)                                     // RC:242-447
{
    // nothing may escape into the GUI's frame loop: the reference reports through title/msg only (RC:264-301)
    try {
        readVolumeDataImpl(fn);
    } catch (const std::exception& e) {
        msg = std::string("Error reading volume: ") + e.what(); title = "Error!";
    }
}

void RendererCore::readVolumeDataImpl(const std::string& fn)
{
    const std::string ext = fn.length() >= 3 ? fn.substr(fn.length() - 3, 3) : std::string();
    std::vector<uint8_t> volume;
    vr::MappedFile mapped;
    vr::PvmVolume pvm;
    const uint8_t* voxels = nullptr;

    if (ext == "raw") {
        vr::RawInf inf;
        bool exists = false;
        std::string t, m;
        if (vr::readRawInf(fn, inf, exists, t, m)) {
            tex3D_dim = vr::ivec3{inf.dims[0], inf.dims[1], inf.dims[2]};
            voxel_size = vr::vec3{inf.spacing[0], inf.spacing[1], inf.spacing[2]};
        } else if (exists) {
            title = t; msg = m;
            return;
        } else {
            // no sidecar: write one from the user-provided parameters (RC:304-317)
            vr::RawInf w;
            w.dims[0] = tex3D_dim.x; w.dims[1] = tex3D_dim.y; w.dims[2] = tex3D_dim.z;
            w.spacing[0] = voxel_size.x; w.spacing[1] = voxel_size.y; w.spacing[2] = voxel_size.z;
            vr::writeRawInf(fn, w);
        }
        std::string merr;
        if (!mapped.open(fn, merr)) { msg = "Failed to Open RAW file..."; title = "Error!"; return; }
        if (tex3D_dim.x <= 0 || tex3D_dim.y <= 0 || tex3D_dim.z <= 0) {
            msg = "Texture Dimensions shouldn't contain any zeroes. Please provide a valid .raw.inf file.";
            title = "Invalid Data Size!";
            return;
        }
        if (datasize_bytes != 1 && datasize_bytes != 2) { msg = "Choose UINT8 or UINT16 first."; title = "Error!"; return; }
        // 64-bit, range checked before anything is allocated (RC:327 multiplies into an `int`)
        const uint64_t d64[3] = {(uint64_t)tex3D_dim.x, (uint64_t)tex3D_dim.y, (uint64_t)tex3D_dim.z};
        uint64_t need = 0;
        if (!vr::checkedVolumeBytes(d64, (uint64_t)datasize_bytes, need)) {
            msg = "Texture Dimensions must be between 1 and 16384. Please provide a valid .raw.inf file.";
            title = "Invalid Data Size!";
            return;
        }
        if (mapped.size() >= need) {
            voxels = mapped.data();                     // zero-copy: the upload reads the mapped pages
        } else {
            // short file: the reference's value-initialised buffer leaves the tail at zero (RC:329-337)
            volume.assign(need, 0);
            if (mapped.size()) std::memcpy(volume.data(), mapped.data(), (size_t)mapped.size());
            voxels = volume.data();
        }
    } else {
        std::string err;
        if (!vr::pvmReadFile(fn, pvm, err)) { msg = "Error reading PVM file"; title = "Error!"; return; }
        const uint64_t pd[3] = {pvm.width, pvm.height, pvm.depth};
        uint64_t pbytes = 0;
        if (!vr::checkedVolumeBytes(pd, 1, pbytes)) { msg = "Error reading PVM file"; title = "Error!"; return; }
        tex3D_dim = vr::ivec3{(int)pvm.width, (int)pvm.height, (int)pvm.depth};
        voxel_size = vr::vec3{pvm.scale[0], pvm.scale[1], pvm.scale[2]};
        if (datasize_bytes != 1 && datasize_bytes != 2) datasize_bytes = (int)pvm.components == 2 ? 2 : 1;
        // like the reference, the menu's 1/2-byte choice wins over `components` (RC:345-348),
        // but never read past the payload
        const uint64_t need = (uint64_t)pvm.width * pvm.height * pvm.depth * (uint64_t)datasize_bytes;
        if (pvm.payload.size() < need) { msg = "Error reading PVM file"; title = "Error!"; return; }
        voxels = pvm.payload.data();
    }

    std::cout << "Dataset dimensions: " << tex3D_dim.x << ", " << tex3D_dim.y << ", " << tex3D_dim.z << std::endl;
    std::cout << "Dataset Aspect ratio: " << voxel_size.x << ", " << voxel_size.y << ", " << voxel_size.z << std::endl;

    if (!ctx) { msg = "Renderer not set up"; title = "Error!"; return; }
    // upload (RC:408-419); the min/max scan and the histogram (RC:360-405) run on the GPU
    const uint64_t dims[3] = {(uint64_t)tex3D_dim.x, (uint64_t)tex3D_dim.y, (uint64_t)tex3D_dim.z};
    const float vs[3] = {voxel_size.x, voxel_size.y, voxel_size.z};
    if (vr_upload_volume(ctx, voxels, dims, datasize_bytes, vs) != VR_OK) { reportAbiError("Error!"); return; }
    vr_volume_stats st;
    if (vr_volume_stats_get(ctx, &st) != VR_OK) { reportAbiError("Error!"); return; }
    min_val = min_dataset_val = st.min_value;                   // RC:375-384
    max_val = max_dataset_val = st.max_value;
    for (int i = 0; i < 256; ++i) histogram[i] = st.histogram[i];

    title = "File Loaded!";
    msg = "File Loaded Successfully!";

    if (!loaded_shader.empty()) {
        setUniforms();
        main_cam.resetCamera();
    }
    const size_t idx = fn.find_last_of("/");
    loaded_dataset = (idx == std::string::npos) ? fn : fn.substr(idx + 1);
}

void RendererCore::setFilter(int f) { params.filter = f; pushParams(); }
void RendererCore::setStepScale(float s, bool oc) { params.step_scale = s; params.opacity_correction = oc ? 1 : 0; pushParams(); }
void RendererCore::setKernel(int k) { params.kernel = k; pushParams(); }
void RendererCore::setTransferFunction(const float* lut256)
{
    if (lut256) { params.use_tf = 1; std::memcpy(params.tf_lut, lut256, sizeof(float) * 256); }
    else params.use_tf = 0;
    pushParams();
}
