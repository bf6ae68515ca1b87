// nr_backward.cu -- fused backward of the rasterize path.
//
// One thread per internal pixel (same 16x16 tiling / 8x4 warp blocks as the forward):
//   1. upstream gradient through the 2x2 anti-aliasing mean, flip and permute
//      (autograd of rasterize.py:315-328), read straight from grad_images [B,C,S,S];
//   2. the Differentiation stencil (differentiation.py:13-36, utils.py:75-101) on the
//      internal-resolution image -> d loss / d (x, y) of the pixel;
//   3. coordinate_map backward (rasterize.py:91-97): w_k * grad_xy onto the three vertices of
//      the pixel's face, straight into grad_vertices (the reference goes through a
//      [B,nf,3,3] intermediate and two index_put scatters: rasterize.py:232, utils.py:104-114);
//   4. sample_textures backward (rasterize.py:100-153): bilinear taps into grad_textures, and the
//      perspective-correct uv path into face z and vertices_textures;
//   5. depth-map backward (rasterize.py:80-88) into face z.
// The weight map is a constant for autograd in the reference (it comes out of a CUDA kernel with
// no autograd edge, rasterize.py:75) and is recomputed here from the face instead of being stored.
#include "nr_kernels.h"

namespace nr {

// d loss / d (x, y) at one pixel. I(c, dy, dx) / G(c, dy, dx) load the image / upstream gradient
// at internal pixel (yi + dy, xi + dx); they are only called for in-range neighbours.
template <class LoadI, class LoadG>
__device__ __forceinline__ void diff_stencil(int yi, int xi, int R, int C, float inv_step, LoadI I,
                                             LoadG G, float &gx, float &gy) {
    const bool ym = yi > 0, yp = yi + 1 < R, xm = xi > 0, xp = xi + 1 < R;
    // r[i] = -sum_c (I[i]-I[i+1]) g[i+1] / step ; l[i] = -sum_c (I[i+1]-I[i]) g[i] / step
    float ry_i = 0.f, ry_m = 0.f, ly_m = 0.f, ly_i = 0.f;
    float rx_i = 0.f, rx_m = 0.f, lx_m = 0.f, lx_i = 0.f;
    for (int c = 0; c < C; ++c) {
        const float ic = I(c, 0, 0), gc = G(c, 0, 0);
        if (yp) {
            const float in = I(c, 1, 0), gn = G(c, 1, 0);
            ry_i = __fadd_rn(ry_i, __fmul_rn(__fsub_rn(ic, in), gn));
            ly_i = __fadd_rn(ly_i, __fmul_rn(__fsub_rn(in, ic), gc));
        }
        if (ym) {
            const float ip = I(c, -1, 0), gp = G(c, -1, 0);
            ry_m = __fadd_rn(ry_m, __fmul_rn(__fsub_rn(ip, ic), gc));
            ly_m = __fadd_rn(ly_m, __fmul_rn(__fsub_rn(ic, ip), gp));
        }
        if (xp) {
            const float in = I(c, 0, 1), gn = G(c, 0, 1);
            rx_i = __fadd_rn(rx_i, __fmul_rn(__fsub_rn(ic, in), gn));
            lx_i = __fadd_rn(lx_i, __fmul_rn(__fsub_rn(in, ic), gc));
        }
        if (xm) {
            const float ip = I(c, 0, -1), gp = G(c, 0, -1);
            rx_m = __fadd_rn(rx_m, __fmul_rn(__fsub_rn(ip, ic), gc));
            lx_m = __fadd_rn(lx_m, __fmul_rn(__fsub_rn(ic, ip), gp));
        }
    }
    // torch divides by the python scalar `step` as a multiplication by 1/step on CUDA
    const float gyr = __fadd_rn(__fmul_rn(-ry_i, inv_step), __fmul_rn(-ry_m, inv_step));
    const float gyl = __fadd_rn(__fmul_rn(-ly_m, inv_step), __fmul_rn(-ly_i, inv_step));
    const float gxr = __fadd_rn(__fmul_rn(-rx_i, inv_step), __fmul_rn(-rx_m, inv_step));
    const float gxl = __fadd_rn(__fmul_rn(-lx_m, inv_step), __fmul_rn(-lx_i, inv_step));
    gy = nr_maximum(gyr, gyl);
    gx = nr_maximum(gxr, gxl);
}

__device__ __forceinline__ int first_argmin3(const float a[3]) {
    int k = 0;
    if (a[1] < a[k]) k = 1;
    if (a[2] < a[k]) k = 2;
    return k;
}
__device__ __forceinline__ int first_argmax3(const float a[3]) {
    int k = 0;
    if (a[1] > a[k]) k = 1;
    if (a[2] > a[k]) k = 2;
    return k;
}

// backward of min(max(x0, lo), hi): gradient shares for x0, lo, hi (ties split evenly, like
// torch.max / torch.min on two tensors)
__device__ __forceinline__ void clamp_shares(float x0, float lo, float hi, float &sx, float &slo,
                                             float &shi) {
    const float x1 = fmaxf(x0, lo);
    float s1 = (x1 < hi) ? 1.f : ((x1 > hi) ? 0.f : 0.5f);
    shi = 1.f - s1;
    const float sa = (x0 > lo) ? 1.f : ((x0 < lo) ? 0.f : 0.5f);
    sx = s1 * sa;
    slo = s1 * (1.f - sa);
}

// Backward of sample_texture(): scatters the four bilinear taps into grad_tex (planar
// [3, H, W] of this view) and returns the gradients w.r.t. face depths and corner uv's.
__device__ __forceinline__ void sample_texture_backward(const float *__restrict__ tex_b,
                                                        float *__restrict__ gtex_b, int H, int W,
                                                        float eps, const float q[3], const float z[3],
                                                        const float u[3], const float v[3],
                                                        const float g[3], float gz[3], float gu[3],
                                                        float gv[3]) {
    const TexCoord tc = texel_coord(q, z, u, v, eps);
    const float depth = tc.depth, nx = tc.nx, ny = tc.ny, x0 = tc.x0, y0 = tc.y0, xf = tc.xf, yf = tc.yf;
    const float *zz = tc.zz;
    const float ulo = fminf(u[0], fminf(u[1], u[2])), uhi = __fsub_rn(fmaxf(u[0], fmaxf(u[1], u[2])), eps);
    const float vlo = fminf(v[0], fminf(v[1], v[2])), vhi = __fsub_rn(fmaxf(v[0], fmaxf(v[1], v[2])), eps);
    const float xff = floorf(xf), yff = floorf(yf), xcf = xff + 1.f, ycf = yff + 1.f;
    const int xfi = (int)xff, yfi = (int)yff, xci = (int)xcf, yci = (int)ycf;
    const float ax = xcf - xf, bx = xf - xff, ay = ycf - yf, by = yf - yff;
    const float w1 = ay * ax, w2 = ay * bx, w3 = by * ax, w4 = by * bx;
    const int T = H * W;
    const int i1 = yfi * W + xfi, i2 = yfi * W + xci, i3 = yci * W + xfi, i4 = yci * W + xci;
    const bool ok1 = (unsigned)i1 < (unsigned)T, ok2 = (unsigned)i2 < (unsigned)T;
    const bool ok3 = (unsigned)i3 < (unsigned)T, ok4 = (unsigned)i4 < (unsigned)T;
    float d1 = 0.f, d2 = 0.f, d3 = 0.f, d4 = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float *p = tex_b + (size_t)c * T;
        float *gp = gtex_b ? gtex_b + (size_t)c * T : nullptr;
        const float gc = g[c];
        if (ok1) { d1 += gc * __ldg(p + i1); if (gp && gc != 0.f) atomicAdd(gp + i1, w1 * gc); }
        if (ok2) { d2 += gc * __ldg(p + i2); if (gp && gc != 0.f) atomicAdd(gp + i2, w2 * gc); }
        if (ok3) { d3 += gc * __ldg(p + i3); if (gp && gc != 0.f) atomicAdd(gp + i3, w3 * gc); }
        if (ok4) { d4 += gc * __ldg(p + i4); if (gp && gc != 0.f) atomicAdd(gp + i4, w4 * gc); }
    }
    const float gxf = ay * (d2 - d1) + by * (d4 - d3);
    const float gyf = ax * (d3 - d1) + bx * (d4 - d2);
    float sx, sxlo, sxhi, sy, sylo, syhi;
    clamp_shares(x0, ulo, uhi, sx, sxlo, sxhi);
    clamp_shares(y0, vlo, vhi, sy, sylo, syhi);
    const float gx0 = gxf * sx, gy0 = gyf * sy;
    const float gnx = gx0 * depth, gny = gy0 * depth;
    const float gdepth = gx0 * nx + gy0 * ny;
    const float gD = -gdepth * depth * depth;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float iz = 1.f / zz[k];
        gu[k] = gnx * q[k] * iz;
        gv[k] = gny * q[k] * iz;
        gz[k] = -(gnx * q[k] * u[k] + gny * q[k] * v[k] + gD * q[k]) * iz * iz;
    }
    // clamp bounds depend on the corner uv's themselves (min / max over the corners)
    gu[first_argmin3(u)] += gxf * sxlo;
    gu[first_argmax3(u)] += gxf * sxhi;
    gv[first_argmin3(v)] += gyf * sylo;
    gv[first_argmax3(v)] += gyf * syhi;
}

__global__ void __launch_bounds__(TILE_THREADS)
k_backward(const BackwardArgs a) {
    const int tid = threadIdx.x;
    const int tile = blockIdx.x, b = blockIdx.y;
    const int tx = tile % a.ntx, ty = tile / a.ntx;
    const int R = a.R, S = a.S, C = a.C;
    int px, py;
    tile_pixel(tid, px, py);
    const int xi = tx * TILE + px, yi = ty * TILE + py;
    if (xi >= R || yi >= R) return;
    const int u_ = R - 1 - yi, v_ = R - 1 - xi;
    const bool aa = (a.flags & FLAG_AA) != 0;

    const float *Ib = a.internal + (size_t)b * C * R * R;
    const float *Gb = a.grad_images + (size_t)b * C * S * S;
    auto LI = [&](int c, int dy, int dx) -> float {
        return __ldg(Ib + ((size_t)c * R + (u_ - dy)) * R + (v_ - dx));
    };
    auto LG = [&](int c, int dy, int dx) -> float {
        const int uu = u_ - dy, vv = v_ - dx;
        return aa ? __fmul_rn(__ldg(Gb + ((size_t)c * S + (uu >> 1)) * S + (vv >> 1)), 0.25f)
                  : __ldg(Gb + ((size_t)c * S + uu) * S + vv);
    };
    const float stepf = (float)(2. / R);
    const float inv_step = __frcp_rn(stepf);
    float gx, gy;
    diff_stencil(yi, xi, R, C, inv_step, LI, LG, gx, gy);

    const int f = a.fim[((size_t)b * R + yi) * R + xi];
    if (f < 0) return;   // gradients only reach the mesh through foreground pixels

    int vid[3];
    if (a.faces) {
        vid[0] = __ldg(a.faces + 3 * (size_t)f);
        vid[1] = __ldg(a.faces + 3 * (size_t)f + 1);
        vid[2] = __ldg(a.faces + 3 * (size_t)f + 2);
    } else {
        vid[0] = 3 * f; vid[1] = 3 * f + 1; vid[2] = 3 * f + 2;
    }
    const float *vb = a.verts + (size_t)b * a.nv * 3;
    float X[3], Y[3], Z[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        X[k] = __ldg(vb + 3 * (size_t)vid[k]);
        Y[k] = __ldg(vb + 3 * (size_t)vid[k] + 1);
        Z[k] = __ldg(vb + 3 * (size_t)vid[k] + 2);
    }
    const float xp = pix_center(xi, R), yp = pix_center(yi, R);
    float q[3];
    raw_weights(xp, yp, X[0], Y[0], X[1], Y[1], X[2], Y[2], q[0], q[1], q[2]);
    normalize_weights(q[0], q[1], q[2]);

    float gz[3] = {0.f, 0.f, 0.f};
    int c0 = 0;
    if (a.flags & FLAG_RGB) {
        const float g[3] = {LG(0, 0, 0), LG(1, 0, 0), LG(2, 0, 0)};
        if (g[0] != 0.f || g[1] != 0.f || g[2] != 0.f) {
            const int32_t *fti = a.ft + 3 * (size_t)f;
            const float *vtb = a.vt + (size_t)b * a.nvt * 2;
            int tvid[3];
            float u[3], v[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                tvid[k] = __ldg(fti + k);
                const float2 uv = __ldg(reinterpret_cast<const float2 *>(vtb) + tvid[k]);
                u[k] = uv.x;
                v[k] = uv.y;
            }
            float gu[3], gv[3];
            sample_texture_backward(a.tex + (size_t)b * 3 * a.H * a.W,
                                    a.grad_tex ? a.grad_tex + (size_t)b * 3 * a.H * a.W : nullptr, a.H,
                                    a.W, a.eps, q, Z, u, v, g, gz, gu, gv);
            if (a.grad_vt) {
                float *gvt = a.grad_vt + (size_t)b * a.nvt * 2;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (gu[k] != 0.f) atomicAdd(gvt + 2 * (size_t)tvid[k], gu[k]);
                    if (gv[k] != 0.f) atomicAdd(gvt + 2 * (size_t)tvid[k] + 1, gv[k]);
                }
            }
        }
        c0 = 3;
    }
    if (a.flags & FLAG_SIL) ++c0;
    if (a.flags & FLAG_DEPTH) {
        const float gd = LG(c0, 0, 0);
        if (gd != 0.f) {
            const float s = (q[0] / Z[0] + q[1] / Z[1]) + q[2] / Z[2];
            const float dm = 1.f / s;
            const float t = gd * dm * dm;
#pragma unroll
            for (int k = 0; k < 3; ++k) gz[k] += t * q[k] / (Z[k] * Z[k]);
        }
    }

    float *gvb = a.grad_verts + (size_t)b * a.nv * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float *dst = gvb + 3 * (size_t)vid[k];
        const float cx = q[k] * gx, cy = q[k] * gy;
        if (cx != 0.f) atomicAdd(dst, cx);
        if (cy != 0.f) atomicAdd(dst + 1, cy);
        if (gz[k] != 0.f) atomicAdd(dst + 2, gz[k]);
    }
}

// Differentiation.backward on channels-last tensors (the public differentiation() op).
__global__ void __launch_bounds__(256)
k_differentiation_backward(const float *__restrict__ images, const float *__restrict__ grad_out,
                           float *__restrict__ grad_xy, int B, int R, int C) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)B * R * R;
    if (i >= total) return;
    const int xi = (int)(i % R), yi = (int)((i / R) % R);
    const float *Ib = images + (size_t)i * C;
    const float *Gb = grad_out + (size_t)i * C;
    auto LI = [&](int c, int dy, int dx) -> float { return __ldg(Ib + ((long long)dy * R + dx) * C + c); };
    auto LG = [&](int c, int dy, int dx) -> float { return __ldg(Gb + ((long long)dy * R + dx) * C + c); };
    const float inv_step = __frcp_rn((float)(2. / R));
    float gx, gy;
    diff_stencil(yi, xi, R, C, inv_step, LI, LG, gx, gy);
    grad_xy[i * 2] = gx;
    grad_xy[i * 2 + 1] = gy;
}

cudaError_t launch_backward(const BackwardArgs &a, cudaStream_t stream) {
    if (a.B <= 0 || a.R <= 0) return cudaSuccess;
    dim3 grid(a.ntx * a.ntx, a.B);
    ProfScope p(PROF_BACKWARD, stream);
    k_backward<<<grid, TILE_THREADS, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_differentiation_backward(const float *images, const float *grad_output,
                                            float *grad_coordinates, int B, int R, int C,
                                            cudaStream_t stream) {
    const long long total = (long long)B * R * R;
    if (total <= 0) return cudaSuccess;
    ProfScope p(PROF_DIFF_BACKWARD, stream);
    k_differentiation_backward<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
        images, grad_output, grad_coordinates, B, R, C);
    return cudaGetLastError();
}

}  // namespace nr
