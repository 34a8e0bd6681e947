/* oracle_julia.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of /root/reference/Fractal2D.hs:19-25 (magnitudeSq,
 * fractionalIterCnt) and :63-98 (juliaAnimated).  Haskell Float = binary32;
 * sin/cos/log on Float call libm sinf/cosf/logf, `Complex Float` multiplication is
 * (a·c − b·d) :+ (a·d + b·c)  (Data.Complex).  Compile with -ffp-contract=off.
 * PARITY STATUS: "parity unpinned" (no reference tests; see oracle_march.c).
 */
#include <stdint.h>
#include <stddef.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "oracle.h"

static float magnitude_sq(float re, float im) { return re * re + im * im; }

/* Fractal2D.hs:24-25 */
static float fractional_iter_cnt(int iter, float re, float im)
{
    float v = (float)iter - logf(logf(magnitude_sq(re, im))) / logf(2.0f);
    return v > 0.0f ? v : 0.0f;            /* max 0 v */
}

/* rows [y0,y1) of the w×h frame; fb holds all h rows, idx = px + py·w (row 0 is the
 * bottom row of the GL texture, FrameBuffer.hs:147-152) */
void oracle_julia(int w, int h, uint32_t *fb, int smooth, double tick, int y0, int y1, int threads)
{
    const float PI_F = 3.14159265358979323846f;
    float f_tick = (float)tick;
    float q1 = f_tick / 17.0f, q2 = f_tick / 61.0f, q3 = f_tick / 71.0f;
    float scaled_tick = q1 - truncf(q1);      /* snd . properFraction */
    float scaled_tick2 = q2 - truncf(q2);
    float scaled_tick3 = q3 - truncf(q3);
    float two_pi = scaled_tick * 2.0f * PI_F;
    float julia_r = sinf(two_pi) * (0.7f > scaled_tick2 ? 0.7f : scaled_tick2);
    float julia_i = cosf(two_pi) * (0.7f > scaled_tick3 ? 0.7f : scaled_tick3);
    float fw = (float)w, fh = (float)h;
    float ratio = fw / fh;
    float xshift = 1.45f * ratio;
    const int max_iter = 40;
    int py;
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads > 0 ? threads : omp_get_max_threads()) if (threads != 1)
#endif
    for (py = y0; py < y1; py++) {
        int px;
        for (px = 0; px < w; px++) {
            float fpx = (float)px, fpy = (float)py;
            float y = (fpy / fh) * 2.9f - 1.45f;
            float x = (fpx / fw) * 2.9f * ratio - xshift;
            float zr = x, zi = y;
            int iter = 0;
            float val;
            for (;;) {
                float nr, ni;
                if (iter == max_iter || magnitude_sq(zr, zi) > 4.0f * 4.0f) break;
                nr = (zr * zr - zi * zi) + julia_r;
                ni = (zr * zi + zi * zr) + julia_i;
                if (nr == zr && ni == zi) { iter = max_iter; break; }    /* 1-cycle */
                zr = nr; zi = ni;
                iter++;
            }
            if (smooth) {
                float ic = (iter == max_iter) ? (float)max_iter : fractional_iter_cnt(iter, zr, zi);
                val = ic / (float)max_iter * 255.0f;
            } else {
                val = (float)iter / (float)max_iter * 255.0f;
            }
            fb[(size_t)px + (size_t)py * (size_t)w] = ((uint32_t)(int32_t)truncf(val)) << 8;
        }
    }
}

/* Fractal2D.hs:98 `forSegmentsConcurrently Nothing 0 h`: rows [y0,y1) cut by makeNSegments
 * (ConcurrentSegments.hs:14-23) into nseg contiguous segments, one thread per segment — the
 * reference's own thread decomposition (nseg defaults to the processor count there). */
void oracle_julia_segments(int w, int h, uint32_t *fb, int smooth, double tick, int y0, int y1, int nseg)
{
    int pairs[2 * 1024], n, i;
    if (nseg < 1) nseg = 1;
    if (nseg > 1024) nseg = 1024;
    n = oracle_make_segments(nseg, y0, y1, pairs, 1024);
#ifdef _OPENMP
#pragma omp parallel for schedule(static, 1) num_threads(n > 0 ? n : 1)
#endif
    for (i = 0; i < n; i++)
        oracle_julia(w, h, fb, smooth, tick, pairs[2 * i], pairs[2 * i + 1], 1);
}
