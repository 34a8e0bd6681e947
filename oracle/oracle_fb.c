/* oracle_fb.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * Fixed-function frame-buffer behaviour the reference relies on
 * (/root/reference/FrameBuffer.hs):
 *   :59-62,153-154,187-195  HighQualityDownscaling = glGenerateMipmap + trilinear
 *        minification; showing an s× frame buffer at 1× samples mip level log2 s.
 *   :215-228 saveFrameBufferToPNG = level 0, rows flipped, alpha forced to 0xFF.
 * glGenerateMipmap's filter is implementation-defined (GL 3.3 §3.8.11); the oracle
 * uses the 2×2 box with round-to-nearest per level on the 8-bit gamma-encoded
 * values, (a+b+c+d+2)>>2, every level re-quantised to RGBA8.  Odd sizes: the last
 * row/column is dropped (floor(w/2)), which never happens for the levels the
 * configs resolve to.  PARITY STATUS: "parity unpinned".
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

static uint32_t box4(uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    uint32_t out = 0;
    int s;
    for (s = 0; s < 32; s += 8) {
        uint32_t v = ((a >> s) & 255u) + ((b >> s) & 255u) + ((c >> s) & 255u) + ((d >> s) & 255u);
        out |= ((v + 2u) >> 2) << s;
    }
    return out;
}
/* out: (w>>levels)·(h>>levels) pixels */
void oracle_mip_resolve(const uint32_t *fb, int w, int h, int levels, uint32_t *out)
{
    uint32_t *cur = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)w * (size_t)h);
    int l, x, y;
    memcpy(cur, fb, sizeof(uint32_t) * (size_t)w * (size_t)h);
    for (l = 0; l < levels; l++) {
        int nw = w >> 1, nh = h >> 1;
        for (y = 0; y < nh; y++)
            for (x = 0; x < nw; x++)
                cur[(size_t)y * nw + x] = box4(cur[(size_t)(2 * y) * w + 2 * x], cur[(size_t)(2 * y) * w + 2 * x + 1],
                                               cur[(size_t)(2 * y + 1) * w + 2 * x], cur[(size_t)(2 * y + 1) * w + 2 * x + 1]);
        w = nw; h = nh;
    }
    memcpy(out, cur, sizeof(uint32_t) * (size_t)w * (size_t)h);
    free(cur);
}
/* FrameBuffer.hs:222-226: PNG row y = texture row h-1-y, A = 0xFF */
void oracle_screenshot_rows(const uint32_t *fb, int w, int h, uint8_t *rgba)
{
    int x, y;
    for (y = 0; y < h; y++)
        for (x = 0; x < w; x++) {
            uint32_t p = fb[(size_t)(h - 1 - y) * w + x];
            uint8_t *o = rgba + ((size_t)y * w + x) * 4;
            o[0] = (uint8_t)(p & 255u); o[1] = (uint8_t)((p >> 8) & 255u); o[2] = (uint8_t)((p >> 16) & 255u); o[3] = 0xFF;
        }
}
