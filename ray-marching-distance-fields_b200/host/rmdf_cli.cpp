// rmdf_cli — headless stand-in for the reference's Main.hs/App.hs loop: load the reflection map,
// build/verify the pre-convolved cache, render one frame of a scene (optionally tile by tile, like
// the `T` key, App.hs:107) and save the screenshot (`S` key → saveFrameBufferToPNG).
//   rmdf_cli [--env file.hdr] [--cache dir] [--scene 0..3|julia] [--w N] [--h N] [--time T]
//            [--tiled] [--fast] [--smooth] [--device D] --out frame.png
#include "rmdf_host.hpp"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

using namespace rmdf_host;

int main(int argc, char **argv)
{
    std::string env, cache, out, scene = "2";
    int w = 512, h = 512, device = 0;                              // Main.hs:50-51
    double time = 0.0;
    bool tiled = false, fast = false, smooth = false;
    for (int i = 1; i < argc; i++) {
        auto next = [&](const char *flag) -> const char * { if (i + 1 >= argc) { std::fprintf(stderr, "%s needs a value\n", flag); std::exit(1); } return argv[++i]; };
        if (!std::strcmp(argv[i], "--env")) env = next("--env");
        else if (!std::strcmp(argv[i], "--cache")) cache = next("--cache");
        else if (!std::strcmp(argv[i], "--scene")) scene = next("--scene");
        else if (!std::strcmp(argv[i], "--w")) w = std::atoi(next("--w"));
        else if (!std::strcmp(argv[i], "--h")) h = std::atoi(next("--h"));
        else if (!std::strcmp(argv[i], "--time")) time = std::atof(next("--time"));
        else if (!std::strcmp(argv[i], "--device")) device = std::atoi(next("--device"));
        else if (!std::strcmp(argv[i], "--out")) out = next("--out");
        else if (!std::strcmp(argv[i], "--tiled")) tiled = true;
        else if (!std::strcmp(argv[i], "--fast")) fast = true;
        else if (!std::strcmp(argv[i], "--smooth")) smooth = true;
        else if (!std::strcmp(argv[i], "--help")) { std::puts("usage: rmdf_cli [--env f.hdr] [--cache dir] [--scene 0..3|julia] [--w N] [--h N] [--time T] [--tiled] [--fast] [--smooth] [--device D] --out frame.png"); return 0; }
        else { std::fprintf(stderr, "unknown argument %s\n", argv[i]); return 1; }
    }
    if (out.empty()) { std::fprintf(stderr, "--out is required\n"); return 1; }
    try {
        ShaderRenderer sr(env, device, cache);
        FrameBuffer fb(w, h);
        if (scene == "julia") {
            fb.fillFrameBuffer([&](int fw, int fh, uint32_t *v) { sr.juliaAnimated(fw, fh, v, smooth, time); return 0; });
        } else {
            const FragmentShader shd = (FragmentShader)std::atoi(scene.c_str());
            const int mode = fast ? RMDF_MODE_FAST : RMDF_MODE_PARITY;
            fb.fillFrameBuffer([&](int fw, int fh, uint32_t *v) {
                if (!tiled) sr.drawShaderTile(shd, -1, fw, fh, time, v, mode);
                else for (int idx = 0; !(idx > 0 && isTileIdxFirstTile(idx)); idx++) sr.drawShaderTile(shd, idx, fw, fh, time, v, mode);
                return 0;
            });
        }
        fb.saveFrameBufferToPNG(out);
        const rmdf_counters c = sr.counters();
        std::printf("Saved screenshot of framebuffer to %s (%dx%d, %llu samples, %llu DE evaluations, last kernel %.3f ms)\n", out.c_str(), w, h,
                    (unsigned long long)c.samples, (unsigned long long)(c.march_de + c.shade_de), c.last_kernel_ms);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "rmdf_cli: %s\n", e.what());
        return 2;
    }
    return 0;
}
