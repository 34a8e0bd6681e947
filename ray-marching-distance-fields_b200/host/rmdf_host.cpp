// rmdf_host.cpp — see rmdf_host.hpp.  File formats only (Radiance RGBE, PNG); every
// computation goes through librmdf_b200.so.
#include "rmdf_host.hpp"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <sys/stat.h>

namespace rmdf_host {

static void check(int rc, const char *what)
{
    if (rc != RMDF_OK) throw std::runtime_error(std::string(what) + ": " + rmdf_last_error());
}

// ---- Radiance .hdr ----------------------------------------------------------------------------
HDRImage loadHDRImage(const std::string &fn)
{
    std::ifstream f(fn, std::ios::binary);
    if (!f) throw std::runtime_error("Not an HDR RGBF image: " + fn);
    std::string data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    size_t pos = 0;
    for (;;) {                                                   // header lines up to the empty one
        size_t e = data.find('\n', pos);
        if (e == std::string::npos) throw std::runtime_error("Not an HDR RGBF image: " + fn);
        bool empty = (e == pos);
        pos = e + 1;
        if (empty) break;
    }
    size_t e = data.find('\n', pos);
    HDRImage img;
    if (e == std::string::npos || std::sscanf(data.substr(pos, e - pos).c_str(), "-Y %d +X %d", &img.h, &img.w) != 2)
        throw std::runtime_error("Not an HDR RGBF image: " + fn);
    pos = e + 1;
    const int w = img.w, h = img.h;
    img.rgb.resize((size_t)w * h * 3);
    std::vector<unsigned char> line((size_t)w * 4);
    const unsigned char *p = reinterpret_cast<const unsigned char *>(data.data());
    const size_t n = data.size();
    for (int y = 0; y < h; y++) {
        if (pos + 4 <= n && w >= 8 && w < 32768 && p[pos] == 2 && p[pos + 1] == 2 && ((p[pos + 2] << 8) | p[pos + 3]) == w) {
            pos += 4;
            for (int c = 0; c < 4; c++)
                for (int x = 0; x < w;) {
                    if (pos >= n) throw std::runtime_error("truncated HDR image: " + fn);
                    int cnt = p[pos++];
                    if (cnt > 128) { cnt -= 128; if (pos >= n || x + cnt > w) throw std::runtime_error("bad RLE: " + fn); while (cnt--) line[(size_t)(x++) * 4 + c] = p[pos]; pos++; }
                    else { if (pos + cnt > n || x + cnt > w) throw std::runtime_error("bad RLE: " + fn); while (cnt--) line[(size_t)(x++) * 4 + c] = p[pos++]; }
                }
        } else {
            if (pos + (size_t)w * 4 > n) throw std::runtime_error("truncated HDR image: " + fn);
            std::memcpy(line.data(), p + pos, (size_t)w * 4);
            pos += (size_t)w * 4;
        }
        for (int x = 0; x < w; x++) {                              // JuicyPixels toFloat
            const float s = std::ldexp(1.0f, (int)line[(size_t)x * 4 + 3] - 136);
            for (int c = 0; c < 3; c++) img.rgb[((size_t)y * w + x) * 3 + c] = ((float)line[(size_t)x * 4 + c] + 0.0f) * s;
        }
    }
    return img;
}

void saveRadianceImage(const std::string &fn, const HDRImage &img)
{
    std::ofstream f(fn, std::ios::binary);
    if (!f) throw std::runtime_error("cannot write " + fn);
    f << "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y " << img.h << " +X " << img.w << "\n";
    std::vector<unsigned char> out((size_t)img.w * img.h * 4);
    for (size_t i = 0; i < (size_t)img.w * img.h; i++) {          // JuicyPixels toRGBE
        const float *v = &img.rgb[i * 3];
        float d = std::fmax(v[0], std::fmax(v[1], v[2]));
        unsigned char *o = &out[i * 4];
        if (d <= 1e-32f) { o[0] = o[1] = o[2] = o[3] = 0; continue; }
        int e;
        const float coeff = std::frexp(d, &e) * 255.9999f / d;
        for (int c = 0; c < 3; c++) o[c] = (unsigned char)(int)std::trunc(v[c] * coeff);
        o[3] = (unsigned char)(e + 128);
    }
    f.write(reinterpret_cast<const char *>(out.data()), (std::streamsize)out.size());
    if (!f) { f.close(); std::remove(fn.c_str()); throw std::runtime_error("write failed: " + fn); }   // ShaderRendering.hs:146-148
}

// ---- FrameBuffer ---------------------------------------------------------------------------------
void FrameBuffer::resizeFrameBuffer(int w, int h)
{
    w_ = w; h_ = h;
    data_.assign((size_t)w * h, 0xFF000000u);
}

static uint32_t crc32_of(const unsigned char *p, size_t n, uint32_t crc = 0)
{
    static uint32_t table[256];
    static bool init = false;
    if (!init) { for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; } init = true; }
    crc = ~crc;
    for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
    return ~crc;
}

void FrameBuffer::saveFrameBufferToPNG(const std::string &fn) const
{
    // RGBA8 PNG with stored (uncompressed) deflate blocks — no zlib dependency
    std::vector<unsigned char> raw;
    raw.reserve((size_t)h_ * ((size_t)w_ * 4 + 1));
    for (int y = 0; y < h_; y++) {                                 // row y of the PNG = texture row h-1-y
        raw.push_back(0);
        for (int x = 0; x < w_; x++) {
            const uint32_t px = data_[(size_t)(h_ - 1 - y) * w_ + x];
            raw.push_back(px & 0xFF); raw.push_back((px >> 8) & 0xFF); raw.push_back((px >> 16) & 0xFF); raw.push_back(0xFF);
        }
    }
    std::vector<unsigned char> z = { 0x78, 0x01 };
    uint32_t a = 1, b = 0;
    for (unsigned char c : raw) { a = (a + c) % 65521; b = (b + a) % 65521; }
    for (size_t off = 0; off < raw.size() || off == 0; off += 65535) {
        const size_t len = std::min<size_t>(65535, raw.size() - off);
        z.push_back(off + len >= raw.size() ? 1 : 0);
        z.push_back(len & 0xFF); z.push_back(len >> 8); z.push_back(~len & 0xFF); z.push_back((~len >> 8) & 0xFF);
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + len);
        if (raw.empty()) break;
    }
    const uint32_t adler = (b << 16) | a;
    for (int s = 24; s >= 0; s -= 8) z.push_back((adler >> s) & 0xFF);
    std::ofstream f(fn, std::ios::binary);
    if (!f) throw std::runtime_error("cannot write " + fn);
    auto chunk = [&](const char *type, const std::vector<unsigned char> &payload) {
        std::vector<unsigned char> buf(type, type + 4);
        buf.insert(buf.end(), payload.begin(), payload.end());
        const uint32_t len = (uint32_t)payload.size(), crc = crc32_of(buf.data(), buf.size());
        const unsigned char l[4] = { (unsigned char)(len >> 24), (unsigned char)(len >> 16), (unsigned char)(len >> 8), (unsigned char)len };
        const unsigned char c[4] = { (unsigned char)(crc >> 24), (unsigned char)(crc >> 16), (unsigned char)(crc >> 8), (unsigned char)crc };
        f.write((const char *)l, 4); f.write((const char *)buf.data(), (std::streamsize)buf.size()); f.write((const char *)c, 4);
    };
    f.write("\x89PNG\r\n\x1a\n", 8);
    std::vector<unsigned char> ihdr(13);
    for (int i = 0; i < 4; i++) { ihdr[i] = (w_ >> (24 - 8 * i)) & 0xFF; ihdr[4 + i] = (h_ >> (24 - 8 * i)) & 0xFF; }
    ihdr[8] = 8; ihdr[9] = 6; ihdr[10] = ihdr[11] = ihdr[12] = 0;
    chunk("IHDR", ihdr); chunk("IDAT", z); chunk("IEND", {});
}

// ---- ShaderRenderer ------------------------------------------------------------------------------
std::string ShaderRenderer::mkCacheFn(const std::string &reflMapFn, float power)
{
    const size_t dot = reflMapFn.find_last_of('.'), slash = reflMapFn.find_last_of('/');
    const std::string stem = (dot != std::string::npos && (slash == std::string::npos || dot > slash)) ? reflMapFn.substr(0, dot) : reflMapFn;
    std::ostringstream s;                                          // Haskell `show (8 :: Float)` = "8.0"
    s << stem << "_cache_pow_" << power;
    if (std::floor(power) == power) s << ".0";
    s << ".hdr";
    return s.str();
}

ShaderRenderer::ShaderRenderer(const std::string &reflMapFn, int device, const std::string &cacheDir)
{
    check(rmdf_create(&ctx_, device), "withShaderRenderer - Init failed");
    try {
        float xyz[288];
        check(rmdf_cornell_box_vertices(xyz), "mkCornellBoxVerticesTex");
        check(rmdf_set_cornell(ctx_, xyz, 96), "mkCornellBoxVerticesTex");
        if (!reflMapFn.empty()) {
            const HDRImage reflMap = loadHDRImage(reflMapFn);
            std::string base = reflMapFn;
            if (!cacheDir.empty()) { const size_t s = reflMapFn.find_last_of('/'); base = cacheDir + "/" + (s == std::string::npos ? reflMapFn : reflMapFn.substr(s + 1)); }
            std::vector<std::pair<float, std::string>> powfn;
            for (float p : { 1.0f, 8.0f, 64.0f, 512.0f }) powfn.push_back({ p, mkCacheFn(base, p) });      // :71-75
            buildPreConvolvedHDREnvMapCache(reflMap, powfn);
            latLongHDREnvMapToCubeMap(RMDF_ENV_REFLECTION, reflMap);
            for (size_t i = 0; i < powfn.size(); i++) latLongHDREnvMapToCubeMap(RMDF_ENV_COS_1 + (int)i, loadHDRImage(powfn[i].second));
        }
    } catch (...) {
        rmdf_destroy(ctx_);
        throw;
    }
}
ShaderRenderer::~ShaderRenderer() { rmdf_destroy(ctx_); }

int ShaderRenderer::buildPreConvolvedHDREnvMapCache(const HDRImage &reflMap, const std::vector<std::pair<float, std::string>> &powfn)
{
    std::vector<std::pair<float, std::string>> missing;
    for (auto &pf : powfn) { struct stat st; if (stat(pf.second.c_str(), &st) != 0) missing.push_back(pf); }
    if (missing.empty()) return 0;
    const HDRImage resized = resizeHDRImage(reflMap, 256);         // :139
    std::vector<float> powers;
    for (auto &pf : missing) powers.push_back(pf.first);
    const std::vector<HDRImage> conv = cosineConvolveHDREnvMap(resized, powers);
    for (size_t i = 0; i < missing.size(); i++) saveRadianceImage(missing[i].second, conv[i]);
    return (int)missing.size();
}

HDRImage ShaderRenderer::resizeHDRImage(const HDRImage &src, int dstw)
{
    HDRImage out;
    out.w = dstw; out.h = rmdf_envmap_resize_height(src.w, src.h, dstw);
    out.rgb.resize((size_t)out.w * out.h * 3);
    check(rmdf_envmap_resize(ctx_, src.rgb.data(), src.w, src.h, dstw, out.rgb.data()), "resizeHDRImage");
    return out;
}

std::vector<HDRImage> ShaderRenderer::cosineConvolveHDREnvMap(const HDRImage &src, const std::vector<float> &powers)
{
    const size_t n = (size_t)src.w * src.h * 3;
    std::vector<float> out(n * powers.size());
    check(rmdf_envmap_convolve(ctx_, src.rgb.data(), src.w, src.h, powers.data(), (int)powers.size(), out.data()), "cosineConvolveHDREnvMap");
    std::vector<HDRImage> imgs(powers.size());
    for (size_t i = 0; i < powers.size(); i++) { imgs[i].w = src.w; imgs[i].h = src.h; imgs[i].rgb.assign(out.begin() + i * n, out.begin() + (i + 1) * n); }
    return imgs;
}

void ShaderRenderer::latLongHDREnvMapToCubeMap(int slot, const HDRImage &latlong)
{
    check(rmdf_envmap_bind(ctx_, slot, latlong.rgb.data(), latlong.w, latlong.h), "latLongHDREnvMapToCubeMap");
}

void ShaderRenderer::drawShaderTile(FragmentShader shd, int tileIdx, int w, int h, double time, uint32_t *fb, int mode, int resolveLog2)
{
    rmdf_params p;
    std::memset(&p, 0, sizeof p);
    p.scene = (int)shd; p.fb_w = w; p.fb_h = h; p.time = (float)time;
    p.tile_idx = tileIdx < 0 ? -1 : tileIdx;
    p.resolve_log2 = resolveLog2; p.mode = mode; p.shard_count = 1;
    check(rmdf_render(ctx_, &p, fb), "drawShaderTile");
}

void ShaderRenderer::juliaAnimated(int w, int h, uint32_t *fb, bool smooth, double tick)
{
    check(rmdf_julia(ctx_, w, h, smooth ? 1 : 0, tick, 0, h, 0, fb), "juliaAnimated");
}

rmdf_counters ShaderRenderer::counters()
{
    rmdf_counters c;
    check(rmdf_counters_get(ctx_, &c), "rmdf_counters_get");
    return c;
}

}  // namespace rmdf_host
