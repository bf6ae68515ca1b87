/*
 * nr_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, never the product path).
 *
 * Plain-C restatement of the two live CUDA kernels of the reference
 * (Rebirth-Alex/neural_renderer_v2_pytorch):
 *
 *   nro_face_index_map*   <- neural_renderer_torch/cuda/rasterize_cuda_kernel.cu:52-153
 *                            (face_index_map_forward_safe_cuda_kernel)
 *   nro_weight_map        <- neural_renderer_torch/cuda/rasterize_cuda_kernel.cu:246-308
 *                            (compute_weight_map_cuda_kernel)
 *
 * The reference has NO CPU implementation of these two stages
 * (cuda/rasterize_cuda.cpp:60-61 rejects CPU tensors), so this file restates the
 * *device* arithmetic, including the FMA contraction nvcc 12.9 applies to the
 * reference source for sm_100a (read from `cuobjdump -sass` of the reference
 * .cu; see DESIGN.md "Exact arithmetic"):
 *
 *   c1 = fma(yp-y0, x1-x0, -rn((y1-y0)*(xp-x0)))          (c2, c3 same shape)
 *   det = fma(x1, y2-y0, fma(x2, y0-y1, rn(x0*(y1-y2))))
 *   w0 = fma(yp, x2-x1, rn(xp*(y1-y2))) + fma(x1, y2, -rn(x2*y1))   (w1, w2 cyclic)
 *   back-face test: two rounded products, not contracted
 *   all divisions IEEE-754 round-to-nearest float divisions;
 *   "computed in double then narrowed" expressions (xp, yp, 1./sum) are innocuous
 *   double roundings (53 >= 2*24+2) and equal the float operation.
 *
 * Must be compiled with -ffp-contract=off so that gcc never fuses anything that
 * is not spelled fmaf() here.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Built with -mfma so fmaf() is one instruction; oracle/__init__.py refuses to load
 * the library on a host CPU without FMA instead of taking SIGILL. */

/* pixel centre, rasterize_cuda_kernel.cu:76-77 */
static inline float nro_pix(int i, int is) {
    return (float)((2. * i + 1 - is) / is);
}

/*
 * One face against one pixel, with the running z-buffer state of that pixel.
 * Statement order follows rasterize_cuda_kernel.cu:82-149.
 */
static inline void nro_face_pixel(const float *face, float xp, float yp, float near_, float far_,
                                  int draw_backside, float delta, int fn, float *depth_min,
                                  int32_t *face_index_min) {
    const float x0 = face[0], y0 = face[1], z0 = face[2];
    const float x1 = face[3], y1 = face[4], z1 = face[5];
    const float x2 = face[6], y2 = face[7], z2 = face[8];

    /* :94-97 bounding box on the pixel centre, strict inequalities */
    if (xp < x0 && xp < x1 && xp < x2) return;
    if (x0 < xp && x1 < xp && x2 < xp) return;
    if (yp < y0 && yp < y1 && yp < y2) return;
    if (y0 < yp && y1 < yp && y2 < yp) return;

    /* :100-104 back-face cull: two rounded products (not contracted by nvcc) */
    if (!draw_backside) {
        const float a = (y2 - y0) * (x1 - x0);
        const float b = (y1 - y0) * (x2 - x0);
        if (a > b) return;
    }

    /* :107-116 edge functions */
    const float c1 = fmaf(yp - y0, x1 - x0, -((y1 - y0) * (xp - x0)));
    const float c2 = fmaf(yp - y1, x2 - x1, -((y2 - y1) * (xp - x1)));
    if (c1 * c2 < 0) return;
    const float c3 = fmaf(yp - y2, x0 - x2, -((y0 - y2) * (xp - x2)));
    if (c2 * c3 < 0) return;

    /* :118-121 degenerate faces; compared in double against 1e-8 */
    const float det = fmaf(x1, y2 - y0, fmaf(x2, y0 - y1, x0 * (y1 - y2)));
    if (fabs((double)det) < 0.00000001) return;

    /* :124-126 early-out on the running minimum */
    if (*depth_min < z0 && *depth_min < z1 && *depth_min < z2) return;

    /* :129-136 barycentric weights */
    float w0 = fmaf(yp, x2 - x1, xp * (y1 - y2)) + fmaf(x1, y2, -(x2 * y1));
    float w1 = fmaf(yp, x0 - x2, xp * (y2 - y0)) + fmaf(x2, y0, -(x0 * y2));
    float w2 = fmaf(yp, x1 - x0, xp * (y0 - y1)) + fmaf(x0, y1, -(x1 * y0));
    const float w_sum = (w0 + w1) + w2;
    w0 = w0 / w_sum;
    w1 = w1 / w_sum;
    w2 = w2 / w_sum;

    /* :139-142 perspective-correct depth and near/far */
    const float s = (w0 / z0 + w1 / z1) + w2 / z2;
    const float zp = 1.0f / s;
    if (zp <= near_ || far_ <= zp) return;

    /* :145-148 z-test with hysteresis, float subtraction */
    if (zp <= *depth_min - delta) {
        *depth_min = zp;
        *face_index_min = fn;
    }
}

/*
 * Literal restatement: every pixel scans every face in index order.
 * faces [B, nf, 9] f32, fim [B, is, is] i32.
 */
void nro_face_index_map(const float *faces, int32_t *fim, int B, int nf, int is, float near_,
                        float far_, int draw_backside, float delta) {
    const long rows = (long)B * is;
#pragma omp parallel for schedule(dynamic, 4)
    for (long r = 0; r < rows; ++r) {
        const int bn = (int)(r / is), yi = (int)(r % is);
        const float yp = nro_pix(yi, is);
        const float *fb = faces + (size_t)bn * nf * 9;
        for (int xi = 0; xi < is; ++xi) {
            const float xp = nro_pix(xi, is);
            float depth_min = far_;
            int32_t best = -1;
            for (int fn = 0; fn < nf; ++fn)
                nro_face_pixel(fb + (size_t)fn * 9, xp, yp, near_, far_, draw_backside, delta, fn,
                               &depth_min, &best);
            fim[(size_t)r * is + xi] = best;
        }
    }
}

/*
 * Same result, less work: the y bounding-box rejection (:96-97) does not depend on
 * xi and has no side effect, so it is hoisted out of the pixel loop (one candidate
 * list per image row, in ascending face order). Used for the larger test cases and
 * as the CPU baseline; checked against nro_face_index_map in tests/.
 */
void nro_face_index_map_rows(const float *faces, int32_t *fim, int B, int nf, int is, float near_,
                             float far_, int draw_backside, float delta) {
    const long rows = (long)B * is;
#pragma omp parallel
    {
        int32_t *cand = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nf > 0 ? nf : 1));
#pragma omp for schedule(dynamic, 4)
        for (long r = 0; r < rows; ++r) {
            const int bn = (int)(r / is), yi = (int)(r % is);
            const float yp = nro_pix(yi, is);
            const float *fb = faces + (size_t)bn * nf * 9;
            int nc = 0;
            for (int fn = 0; fn < nf; ++fn) {
                const float *f = fb + (size_t)fn * 9;
                const float y0 = f[1], y1 = f[4], y2 = f[7];
                if (yp < y0 && yp < y1 && yp < y2) continue;
                if (y0 < yp && y1 < yp && y2 < yp) continue;
                cand[nc++] = fn;
            }
            for (int xi = 0; xi < is; ++xi) {
                const float xp = nro_pix(xi, is);
                float depth_min = far_;
                int32_t best = -1;
                for (int c = 0; c < nc; ++c)
                    nro_face_pixel(fb + (size_t)cand[c] * 9, xp, yp, near_, far_, draw_backside,
                                   delta, cand[c], &depth_min, &best);
                fim[(size_t)r * is + xi] = best;
            }
        }
        free(cand);
    }
}

/*
 * compute_weight_map, rasterize_cuda_kernel.cu:246-308.
 * wmap [B*is*is, 3] must be zero-filled by the caller (background pixels are not
 * written, exactly like the reference: rasterize.py:71).
 */
void nro_weight_map(const float *faces, const int32_t *fim, float *wmap, int B, int nf, int is) {
    const long rows = (long)B * is;
#pragma omp parallel for schedule(static)
    for (long r = 0; r < rows; ++r) {
        const int bn = (int)(r / is), yi = (int)(r % is);
        const float yp = nro_pix(yi, is);
        for (int xi = 0; xi < is; ++xi) {
            const size_t i = (size_t)r * is + xi;
            const int fi = fim[i];
            if (fi < 0) continue;
            const float xp = nro_pix(xi, is);
            const float *f = faces + ((size_t)bn * nf + fi) * 9;
            const float x0 = f[0], y0 = f[1], x1 = f[3], y1 = f[4], x2 = f[6], y2 = f[7];
            float w[3];
            w[0] = fmaf(yp, x2 - x1, xp * (y1 - y2)) + fmaf(x1, y2, -(x2 * y1));
            w[1] = fmaf(yp, x0 - x2, xp * (y2 - y0)) + fmaf(x2, y0, -(x0 * y2));
            w[2] = fmaf(yp, x1 - x0, xp * (y0 - y1)) + fmaf(x0, y1, -(x1 * y0));
            float w_sum = (w[0] + w[1]) + w[2];
            if (w_sum < 0) {
                w[0] = -w[0];
                w[1] = -w[1];
                w[2] = -w[2];
            }
            /* max(w, 0.) is evaluated in double; fmax drops a NaN operand like CUDA's */
            w[0] = (float)fmax((double)w[0], 0.);
            w[1] = (float)fmax((double)w[1], 0.);
            w[2] = (float)fmax((double)w[2], 0.);
            w_sum = (w[0] + w[1]) + w[2];
            for (int j = 0; j < 3; ++j) {
                float q = w[j] / w_sum;
                q = (float)fmax(fmin((double)q, 1.), 0.);
                wmap[i * 3 + j] = q;
            }
        }
    }
}

int nro_version(void) { return 1; }
