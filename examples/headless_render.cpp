// headless_render.cpp -- the reference's RendererGUI::run(), minus the GUI: the same sequence of
// RendererCore / Camera calls the ImGui front end makes (src/RendererGUI.cpp:36-51,100-101,124-138,
// 191-222 of the reference), written against volume-renderer_b200/host/.  Shows that the host mirror
// is driven exactly like the reference's class: set the public fields, call the setters, render(),
// saveImage().  Build (after volume-renderer_b200/build.sh), with the host sources compiled into the
// application exactly as INTEGRATION.md section 2 tells a maintainer of the reference to do:
//   H=volume-renderer_b200/host; g++ -std=c++17 -O2 -ffp-contract=off -Iinclude -I$H examples/headless_render.cpp
//       $H/RendererCore.cpp $H/Camera.cpp $H/CubicSpline.cpp $H/VolumeIO.cpp $H/ImageIO.cpp
//       -Lvolume-renderer_b200/lib -lvolren_b200 -Wl,-rpath,$PWD/volume-renderer_b200/lib -o examples/headless_render
// Run (needs a B200; there is no CPU path):
//   examples/headless_render volume.raw|volume.pvm <1|2 bytes per voxel> out.png [zenith azimuth] [mip]
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <string>

#include "RendererCore.h"

int main(int argc, char** argv)
{
    if (argc < 4) {
        std::fprintf(stderr, "usage: %s volume.raw|volume.pvm bytes_per_voxel out.png [zenith azimuth] [mip]\n", argv[0]);
        return 2;
    }
    try {
        RendererCore volren;
        // RendererGUI.cpp:38-40 -- the window size is fixed before setup()
        volren.window_size = vr::ivec2{1280, 720};
        volren.framebuffer_size = volren.window_size;
        volren.setup();
        // RendererGUI.cpp:51 -- "load the shader": the kernel is built in, the call records the name
        if (!volren.loadShader("VolumeRenderer.cs", false)) { std::fprintf(stderr, "%s\n", volren.msg.c_str()); return 1; }
        // RendererGUI.cpp:124-138,191-203 -- the load menu sets datasize_bytes, then readVolumeData
        volren.datasize_bytes = std::atoi(argv[2]);
        volren.readVolumeData(argv[1]);
        if (!volren.title.empty()) {      // errors arrive as popup strings, as in the reference (RendererGUI.cpp:90-96)
            std::fprintf(stderr, "%s: %s\n", volren.title.c_str(), volren.msg.c_str());
            return 1;
        }
        std::printf("%s: %d x %d x %d, spacing %g %g %g, values %d..%d\n", volren.loaded_dataset.c_str(),
                    volren.tex3D_dim.x, volren.tex3D_dim.y, volren.tex3D_dim.z,
                    volren.voxel_size.x, volren.voxel_size.y, volren.voxel_size.z,
                    volren.min_dataset_val, volren.max_dataset_val);
        // RendererGUI.cpp:336-358,382-386 -- sliders: mutate the field, then call the setter
        volren.alpha_scale = 0.05f;
        volren.setAlpha();
        volren.min_val = volren.min_dataset_val;
        volren.setMinVal();
        volren.max_val = volren.max_dataset_val;
        volren.setMaxVal();
        volren.use_mip = argc > 6 && std::string(argv[6]) == "mip";
        volren.setMIP();
        volren.setFilter(VR_FILTER_TRILINEAR);          // extension of this backend
        // GlfwManager.cpp:179,213 -- mouse drags arrive as setOrientation(zoom, zenith, azimuth)
        if (argc > 5) volren.main_cam.setOrientation(0.0f, (float)std::atof(argv[4]), (float)std::atof(argv[5]));
        // RendererGUI.cpp:100-101 -- one frame; the GUI reads kerneltime_sum (milliseconds) once per second
        volren.render();
        std::printf("kernel %.3f ms\n", volren.kerneltime_sum);
        // RendererGUI.cpp:221-222
        const std::string out = argv[3];
        const std::string ext = out.substr(out.find_last_of('.') + 1);
        if (!volren.saveImage(out, ext)) { std::fprintf(stderr, "%s: %s\n", volren.title.c_str(), volren.msg.c_str()); return 1; }
        return 0;
    } catch (const std::exception& e) {   // main.cpp:10-17 of the reference
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
}
