/* CPU oracle (TEST INFRASTRUCTURE ONLY): plain-C restatement of the reference's integer tile-grid arithmetic and
 * blend-ramp values, independent of both the NumPy oracle (oracle/tiling.py) and the product (csrc/hostutil.cpp),
 * so the three can be cross-checked.  Parity unpinned by reference tests (it has none); pinned by the
 * hand-derived golden numbers in tests/golden/tile_grid.json.
 *
 * Follows /root/reference/src/tensorrt/img2img_render.cpp:7-66 (calculateTiles) and
 * /root/reference/src/tensorrt/img2img_load.cpp:29-52 (createTileWeights).
 * Build: gcc -O2 -shared -fPIC oracle/tiling_c.c -o oracle/_build/liboracle_tiling.so -lm   (see __graft_entry__.build) */
#include <math.h>

typedef struct { int x, y, w, h; } orc_rect;

/* returns tile count; info[8] = nx, ny, sIn.w, sIn.h, iov.x, iov.y, oov.x, oov.y */
int orc_calculate_tiles(int in_w, int in_h, int out_w, int out_h, int tw, int th, int otw, int oth, int scaling,
                        double ovx, double ovy, orc_rect* in_rects, orc_rect* out_rects, int cap, int* info) {
    const int sot_w = tw * scaling, sot_h = tw * scaling; /* render.cpp:11-14 uses .width twice */
    const int sin_w = (int)lround((double)otw / sot_w * tw);        /* :17 */
    const int sin_h = (int)lround((double)oth / sot_h * th);        /* :18 */
    const int iov_x = (int)lround(tw * ovx), iov_y = (int)lround(th * ovy);          /* :21-24 */
    const int oov_x = (int)lround(sot_w * ovx), oov_y = (int)lround(sot_h * ovy);    /* :26-29 */
    const int nx = (int)lround(ceil((double)(in_w - iov_x) / (sin_w - iov_x)));      /* :32 */
    const int ny = (int)lround(ceil((double)(in_h - iov_y) / (sin_h - iov_y)));      /* :33 */
    int k = 0;
    for (int i = 0; i < nx; ++i)                                                     /* :43 */
        for (int j = 0; j < ny; ++j, ++k) {                                          /* :44 */
            if (k >= cap) continue;
            in_rects[k].x = -((tw - sin_w) / 2) + i * sin_w - i * iov_x;             /* :47 */
            in_rects[k].y = -((th - sin_h) / 2) + j * sin_h - j * iov_y;             /* :48 */
            in_rects[k].w = tw; in_rects[k].h = th;
            const int x = i * otw - i * oov_x, y = j * oth - j * oov_y;              /* :54-55 */
            out_rects[k].x = x; out_rects[k].y = y;
            out_rects[k].w = x + otw > out_w ? out_w - x : otw;                      /* :59 */
            out_rects[k].h = y + oth > out_h ? out_h - y : oth;                      /* :60 */
        }
    if (info) { info[0] = nx; info[1] = ny; info[2] = sin_w; info[3] = sin_h; info[4] = iov_x; info[5] = iov_y; info[6] = oov_x; info[7] = oov_y; }
    return nx * ny;
}

/* top-edge ramp of createTileWeights: rows r < overlap get (float)((double)(r+1)/(overlap+1)) (load.cpp:34-38) */
int orc_blend_ramp(int overlap, float* ramp) {
    const int height = overlap + 1;
    for (int i = 1; i < height; ++i) ramp[i - 1] = (float)((double)i / height);
    return overlap;
}
