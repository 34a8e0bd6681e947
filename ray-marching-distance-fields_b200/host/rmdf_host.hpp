// rmdf_host.hpp — C++ host-side mirror of the reference's interface for the hot path, on top
// of the C ABI (include/rmdf.h).  The reference's host is compiled Haskell (ShaderRendering.hs,
// FrameBuffer.hs, HDREnvMap.hs); GHC is absent from this image, so this is the compiled caller
// that exercises the boundary the way the patched reference would.
// Same names, argument meaning and error behaviour (errors → std::runtime_error carrying
// rmdf_last_error(), the reference's `Left String` / traceAndThrow).
#pragma once
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "../../include/rmdf.h"

namespace rmdf_host {

// JuicyPixels `Image PixelRGBF`: interleaved RGB float, row 0 = top
struct HDRImage {
    int w = 0, h = 0;
    std::vector<float> rgb;
};

// ShaderRendering.hs:46-47
enum class FragmentShader : int { FSDECornellBoxShader = 0, FSDETestShader, FSMBPower8Shader, FSMBGeneralShader };

constexpr int tilesX = 8, tilesY = 8, nTiles = tilesX * tilesY;                 // ShaderRendering.hs:49-52
inline bool isTileIdxLastTile(int idx) { return idx % nTiles == nTiles - 1; }   // :54-55
inline bool isTileIdxFirstTile(int idx) { return idx % nTiles == 0; }           // :57-58

// HDREnvMap.hs:31-52 / JP.saveRadianceImage (RGBE convention: DESIGN.md C8)
HDRImage loadHDRImage(const std::string &fn);
void saveRadianceImage(const std::string &fn, const HDRImage &img);

// FrameBuffer.hs:39-228 without GL: w·h Word32 0xAABBGGRR, idx = px + py·w, row 0 = bottom
class FrameBuffer {
public:
    FrameBuffer(int w, int h) { resizeFrameBuffer(w, h); }
    void resizeFrameBuffer(int w, int h);                                        // :79-111 (cleared to opaque black)
    std::pair<int, int> getFrameBufferDim() const { return { w_, h_ }; }
    // :117-158 — f(w, h, vec) writes the pixels
    template <class F> auto fillFrameBuffer(F &&f) { return f(w_, h_, data_.data()); }
    void saveFrameBufferToPNG(const std::string &fn) const;                      // :215-228 (flipped, A = 0xFF)
    const std::vector<uint32_t> &data() const { return data_; }
private:
    int w_ = 0, h_ = 0;
    std::vector<uint32_t> data_;
};

// ShaderRendering.hs:36-44,60-110 — RAII instead of the bracket
class ShaderRenderer {
public:
    // withShaderRenderer shdFn reflMapFn: the shader file has no counterpart (scenes are compiled
    // ahead of time); reflMapFn may be empty = no environment (all maps sample as 1.0)
    explicit ShaderRenderer(const std::string &reflMapFn = "", int device = 0, const std::string &cacheDir = "");
    ~ShaderRenderer();
    ShaderRenderer(const ShaderRenderer &) = delete;
    ShaderRenderer &operator=(const ShaderRenderer &) = delete;

    // :151-196 — tileIdx < 0 = Nothing
    void drawShaderTile(FragmentShader shd, int tileIdx, int w, int h, double time, uint32_t *fb,
                        int mode = RMDF_MODE_PARITY, int resolveLog2 = 0);
    // :131-149 — returns the number of cache files written
    int buildPreConvolvedHDREnvMapCache(const HDRImage &reflMap, const std::vector<std::pair<float, std::string>> &powfn);
    HDRImage resizeHDRImage(const HDRImage &src, int dstw);                      // HDREnvMap.hs:169-195
    std::vector<HDRImage> cosineConvolveHDREnvMap(const HDRImage &src, const std::vector<float> &powers);   // :217-254
    void latLongHDREnvMapToCubeMap(int slot, const HDRImage &latlong);           // :118-163
    void juliaAnimated(int w, int h, uint32_t *fb, bool smooth, double tick);    // Fractal2D.hs:63-98
    rmdf_counters counters();
    static std::string mkCacheFn(const std::string &reflMapFn, float power);     // ShaderRendering.hs:72-74
    rmdf_ctx *ctx() { return ctx_; }
private:
    rmdf_ctx *ctx_ = nullptr;
};

}  // namespace rmdf_host
