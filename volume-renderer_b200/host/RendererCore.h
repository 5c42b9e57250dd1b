// RendererCore.h -- the reference's renderer host class (include/RendererCore.h:9-44) with the
// OpenGL calls replaced by calls into the C-ABI of libvolren_b200.so.  Member names, method
// names and their order of effects are kept so that the reference's RendererGUI (a friend, as
// in the reference) compiles against this header unchanged apart from the glm -> vr:: vector
// typedefs.  Nothing in this class touches CUDA directly.
#pragma once

#include <string>
#include <vector>

#include "Camera.h"
#include "vecmath.h"
#include "volren_b200.h"

class RendererGUI;

class RendererCore
{
public:
    RendererCore();
    ~RendererCore();
    void setup();       // allocates the frame (vr_create); throws std::runtime_error like setupFBO
    void render();      // one frame; adds the kernel milliseconds to kerneltime_sum (RendererCore.cpp:153)

    // The reference declares everything below private and lets `friend class RendererGUI` in
    // (include/RendererCore.h:17-18); the headless tools and tests of this repo are the "GUI"
    // here, so it is public.  Names are the drop-in contract.
    friend class RendererGUI;

    // ---- uniform setters: the GUI mutates the field, then calls the setter (RendererGUI.cpp:336-358,382-386)
    void setAlpha();
    void setMinVal();                   // 16-bit data: uniform = GUI value + 1000 (RendererCore.cpp:66-69)
    void setMaxVal();
    void setMIP();
    void setUniforms();
    void setInitialCameraRotation();    // also resets the camera (RendererCore.cpp:94)
    void setupFBO();
    void setupUBO(bool is_update = false);
    // ---- loading / saving (RendererGUI.cpp:128-138,191-222)
    void readVolumeData(std::string fn);            // .raw + .raw.inf, .pvm; errors -> title/msg, no throw
    bool checkRawInfFile(std::string fn);
    bool saveImage(std::string fn, std::string ext);
    bool loadShader(std::string fn, bool reload);   // the kernel is built in: records the name, returns true

    // ---- extensions of the CUDA backend
    void setFilter(int vr_filter);
    void setStepScale(float step_scale, bool opacity_correction);
    void setTransferFunction(const float* lut256);  // nullptr disables
    void setKernel(int vr_kernel);
    bool readFrame(std::vector<float>& rgba);       // frame of the last render(): W*H*4 floats, bottom row first

    // ---- state the GUI reads and writes directly
    Camera main_cam;
    std::vector<float> histogram;                   // 256 bins, 0..100 (RendererCore.cpp:386-405)
    std::string loaded_dataset, loaded_shader;
    std::string msg, title;                         // both non-empty => the GUI shows a popup (RendererGUI.cpp:90-96)
    float alpha_scale;
    float kerneltime_sum;                           // milliseconds; read and zeroed by the GUI once per second
    int workgroups_x, workgroups_y;                 // shown in the profiler window; = image size / 16
    int datasize_bytes;                             // 1 | 2, chosen in the GUI's load menu
    int min_val, max_val, max_dataset_val, min_dataset_val;
    bool use_mip, rotate_to_bottom, rotate_to_top;
    vr::vec3 voxel_size;
    vr::ivec3 tex3D_dim;
    vr::ivec2 window_size, framebuffer_size;

    // ---- what replaces the GL object names (vol_tex3D, camera_ubo_ID, fbo_ID, cs_programID ...)
    int cuda_device;              // which GPU renders
    vr_context* ctx;              // the C-ABI context
    vr_params params;             // shadow of the uniform block
    vr_render_stats last_stats;

private:
    void readVolumeDataImpl(const std::string& fn);
    void pushParams();
    void reportAbiError(const char* title_text);
};
