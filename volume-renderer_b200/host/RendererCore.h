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
        void setup();
        void render();

    // The reference declares everything below private and lets `friend class RendererGUI` in;
    // the headless tools and tests of this repo are the "GUI" here, so it is public.
    public:
        friend class RendererGUI;
        void setAlpha();
        void setMinVal();
        void setMaxVal();
        void setMIP();
        void setUniforms();
        void setInitialCameraRotation();
        void setupFBO();
        void setupUBO(bool is_update = false);
        void readVolumeData(std::string fn);
        bool checkRawInfFile(std::string fn);
        bool saveImage(std::string fn, std::string ext);
        bool loadShader(std::string fn, bool reload);

        // extensions of the CUDA backend: sampling filter, step override,
        // opacity correction, transfer-function LUT, device choice, kernel choice
        void setFilter(int vr_filter);
        void setStepScale(float step_scale, bool opacity_correction);
        void setTransferFunction(const float* lut256);       // nullptr disables
        void setKernel(int vr_kernel);
        // the frame of the last render(), W*H*4 floats, bottom row first
        bool readFrame(std::vector<float>& rgba);

        Camera main_cam;
        std::vector<float> histogram;
        std::string loaded_dataset, loaded_shader, msg, title;
        float alpha_scale, kerneltime_sum;
        int workgroups_x, workgroups_y, datasize_bytes, min_val, max_val, max_dataset_val, min_dataset_val;
        bool use_mip, rotate_to_bottom, rotate_to_top;
        vr::vec3 voxel_size;
        vr::ivec3 tex3D_dim;
        vr::ivec2 window_size, framebuffer_size;

        int cuda_device;              // replaces the GL context: which GPU renders
        vr_context* ctx;              // replaces vol_tex3D / camera_ubo_ID / fbo_ID / cs_programID
        vr_params params;             // shadow of the uniform block
        vr_render_stats last_stats;

    private:
        void pushParams();
        void reportAbiError(const char* title_text);
};
