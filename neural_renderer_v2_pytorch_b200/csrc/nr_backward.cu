// nr_backward.cu -- fused backward of the rasterize path.
//
// Persistent over the forward's non-empty 16x16 tiles; one thread per internal pixel, a warp owns two
// 16-pixel row segments of the tile (so the pixels of one face form runs in lane order):
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

// Backward of sample_texture(): returns the four tap indices (-1 when the tap reads as zero),
// the tap weights, and the gradients w.r.t. face depths and corner uv's.
template <bool NEED_UV>
__device__ __forceinline__ void sample_texture_backward(const float *__restrict__ tex_b, int H, int W,
                                                        float eps, const float q[3], const float z[3],
                                                        const float u[3], const float v[3], const float *aux,
                                                        const float g_lit[3], const float cw[3], float g[3],
                                                        float rgb_tex[3], int tap[4], float tw[4],
                                                        int &cell, float gz[3], float gu[3], float gv[3]) {
    // exact divisions like the forward: texel coordinates reach ~1e3, where the fast division's 2 ulp
    // would move the bilinear weights by more than the 1e-5 gradient tolerance.  With the forward's aux map
    // the three quotients are read back instead (aux = depth, nx, ny of this pixel).
    const TexCoord tc = aux ? texel_coord_stored(aux[0], aux[1], aux[2], z, u, v, eps) : texel_coord<true>(q, z, u, v, eps);
    const float depth = tc.depth, nx = tc.nx, ny = tc.ny, x0 = tc.x0, y0 = tc.y0, xf = tc.xf, yf = tc.yf;
    const float *zz = tc.zz;
    const float ulo = fminf(u[0], fminf(u[1], u[2])), uhi = __fsub_rn(fmaxf(u[0], fmaxf(u[1], u[2])), eps);
    const float vlo = fminf(v[0], fminf(v[1], v[2])), vhi = __fsub_rn(fmaxf(v[0], fmaxf(v[1], v[2])), eps);
    const float xff = floorf(xf), yff = floorf(yf), xcf = xff + 1.f, ycf = yff + 1.f;
    const int xfi = (int)xff, yfi = (int)yff, xci = (int)xcf, yci = (int)ycf;
    const float ax = xcf - xf, bx = xf - xff, ay = ycf - yf, by = yf - yff;
    tw[0] = ay * ax; tw[1] = ay * bx; tw[2] = by * ax; tw[3] = by * bx;
    const int T = H * W;
    tap[0] = yfi * W + xfi; tap[1] = yfi * W + xci; tap[2] = yci * W + xfi; tap[3] = yci * W + xci;
    cell = (yfi << 16) ^ (xfi & 0xffff);   // identifies the bilinear cell: equal cell => equal four taps
    // g = upstream gradient of the UNLIT sample (lit = unlit * cw, rasterize.py:283); rgb_tex = unlit sample
    float d[4];
    g[0] = g_lit[0] * cw[0]; g[1] = g_lit[1] * cw[1]; g[2] = g_lit[2] * cw[2];
    rgb_tex[0] = rgb_tex[1] = rgb_tex[2] = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        if ((unsigned)tap[t] >= (unsigned)T) tap[t] = -1;
        d[t] = 0.f;
        if (tap[t] >= 0) {
            const float *pt = tex_b + tap[t];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float tv = __ldg(pt + c * T);
                d[t] += g[c] * tv;
                rgb_tex[c] += tw[t] * tv;
            }
        }
    }
    const float gxf = ay * (d[1] - d[0]) + by * (d[3] - d[2]);
    const float gyf = ax * (d[2] - d[0]) + bx * (d[3] - d[1]);
    float sx, sxlo = 0.f, sxhi = 0.f, sy, sylo = 0.f, syhi = 0.f;
    clamp_shares(x0, ulo, uhi, sx, sxlo, sxhi);
    clamp_shares(y0, vlo, vhi, sy, sylo, syhi);
    const float gx0 = gxf * sx, gy0 = gyf * sy;
    const float gnx = gx0 * depth, gny = gy0 * depth;
    const float gdepth = gx0 * nx + gy0 * ny;
    const float gD = -gdepth * depth * depth;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float iz = 1.f / zz[k];
        if (NEED_UV) {
            gu[k] = gnx * q[k] * iz;
            gv[k] = gny * q[k] * iz;
        }
        gz[k] = -(gnx * q[k] * u[k] + gny * q[k] * v[k] + gD * q[k]) * iz * iz;
    }
    if (NEED_UV) {
        // clamp bounds depend on the corner uv's themselves (min / max over the corners)
        gu[first_argmin3(u)] += gxf * sxlo;
        gu[first_argmax3(u)] += gxf * sxhi;
        gv[first_argmin3(v)] += gyf * sylo;
        gv[first_argmax3(v)] += gyf * syhi;
    }
}

// Warp-level segmented sum over runs of equal `key` in lane order (a warp is one image row
// segment, so the pixels of one face / one texel cell form runs).  Afterwards the LAST lane of
// every run holds the run total in v[]; returns true on those lanes.  Fixed summation order.
// Lanes with valid == false never join a run.
template <int N>
__device__ __forceinline__ bool run_reduce(int key, bool valid, float (&v)[N], int lane) {
    const int prev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool prev_valid = __shfl_up_sync(0xffffffffu, (int)valid, 1) != 0;
    // an invalid lane is a run of its own and separates the runs around it
    // runs never cross lane 16 (a warp holds two 16-pixel rows), so four doubling steps suffice
    const bool head = ((lane & 15) == 0) || !valid || !prev_valid || (key != prev);
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    if (heads != 0xffffffffu) {
        const int seg_start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) {
            const bool take = (lane - d >= seg_start);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const float t = __shfl_up_sync(0xffffffffu, v[i], d);
                if (take) v[i] += t;
            }
        }
    }
    return (lane == 31) || ((heads >> (lane + 1)) & 1u);
}

// C = channel count at compile time (0: read it from the arguments).
// Persistent over the forward's list of non-empty 16x16 tiles (grid-stride); a warp owns two
// 16-pixel row segments of the tile, so the pixels of one face form runs in lane order.
// Accumulate one contribution.  Float atomics (default) commute only up to rounding, so the result
// depends on arrival order; in deterministic mode the value is rounded ONCE to 64-bit fixed point
// (value * 2^k) and added with integer atomics, which are exactly associative: the sum is bit-identical
// from run to run whatever the order.  The warp-level run sums in front of it have a fixed order.
template <bool DET, class Index>
__device__ __forceinline__ void accumulate(float *dst_f, long long *dst_i, Index idx, float v, float scale) {
    if (DET) {
        atomicAdd(reinterpret_cast<unsigned long long *>(dst_i + idx),
                  (unsigned long long)__double2ll_rn((double)v * (double)scale));
    } else {
        atomicAdd(dst_f + idx, v);
    }
}

// AUX: the forward left the normalised weights (and, with colour, the texel coordinate) of every foreground pixel
// in the aux map: nothing of that is re-derived here (9 scattered vertex loads and 12 IEEE divisions less per pixel).
template <int CT, bool NEED_UV, bool DET, bool AUX>
__global__ void __launch_bounds__(TILE_THREADS, 4)
k_backward(const BackwardArgs a) {
    pdl_wait();         // the forward's maps, tile list and zero-filled accumulators (nr_kernels.h)
    const int lane = threadIdx.x & 31, wrow = threadIdx.x >> 5;
    const int R = a.R, S = a.S, C = CT ? CT : a.C;
    const int nt = a.ntx * a.ntx;
    TileList tl;
    if (a.tile_list) tl = open_tile_list(a.tile_list, a.B * nt);
    const int count = a.tile_list ? tl.total : a.B * nt;
    for (int work = blockIdx.x; work < count; work += gridDim.x) {
    int b, tx, ty;
    if (a.tile_list) {
        const int4 e = tile_entry(tl, work);
        b = e.x; tx = e.y & 0xffff; ty = e.y >> 16;
    } else {
        b = work / nt;
        const int tile = work - b * nt;
        ty = tile / a.ntx; tx = tile - ty * a.ntx;
    }
    const int xi = tx * TILE + (lane & 15), yi = ty * TILE + wrow * 2 + (lane >> 4);
    const bool valid = (xi < R) && (yi < R);
    const int f = valid ? __ldg(a.fim + ((size_t)b * R + yi) * R + xi) : -1;
    const bool fg = f >= 0;
    // gradients only reach the mesh through foreground pixels (to_map links nothing else,
    // utils.py:104-114): a warp without any is done with this tile
    if (__ballot_sync(0xffffffffu, fg) == 0u) continue;

    const int u_ = R - 1 - yi, v_ = R - 1 - xi;
    const bool aa = (a.flags & FLAG_AA) != 0;
    float gx = 0.f, gy = 0.f;
    float gcen[5] = {0.f, 0.f, 0.f, 0.f, 0.f};   // upstream gradient at this pixel, per channel
    if (fg) {
        // neighbour coordinates, clamped so every load is in range; out-of-range pairs are masked
        const bool has_yp = yi + 1 < R, has_ym = yi > 0, has_xp = xi + 1 < R, has_xm = xi > 0;
        const int u_yp = has_yp ? u_ - 1 : u_, u_ym = has_ym ? u_ + 1 : u_;
        const int v_xp = has_xp ? v_ - 1 : v_, v_xm = has_xm ? v_ + 1 : v_;
        const float *Ib = a.internal + (size_t)b * C * R * R;
        const float *Gb = a.grad_images + (size_t)b * C * S * S;
        const float gs = aa ? 0.25f : 1.f;   // 2x2 mean backward (rasterize.py:323-328), exact
        const int sh = aa ? 1 : 0;
        // one 64-bit pointer per array (this pixel, channel 0); neighbours and channels are 32-bit element
        // offsets from it (R <= 32768, so a plane has < 2^31 elements)
        const float *pI = Ib + (size_t)u_ * R + v_;
        const float *pG = Gb + (size_t)(u_ >> sh) * S + (v_ >> sh);
        const int dI_yp = (u_yp - u_) * R, dI_ym = (u_ym - u_) * R, dI_xp = v_xp - v_, dI_xm = v_xm - v_;
        const int dG_yp = ((u_yp >> sh) - (u_ >> sh)) * S, dG_ym = ((u_ym >> sh) - (u_ >> sh)) * S;
        const int dG_xp = (v_xp >> sh) - (v_ >> sh), dG_xm = (v_xm >> sh) - (v_ >> sh);
        const int planeI = R * R, planeG = S * S;
        float ry_i = 0.f, ry_m = 0.f, ly_m = 0.f, ly_i = 0.f, rx_i = 0.f, rx_m = 0.f, lx_m = 0.f, lx_i = 0.f;
#pragma unroll
        for (int c = 0; c < (CT ? CT : 5); ++c) {
            if (c < C) {
                const float *Ic = pI + c * planeI, *Gc = pG + c * planeG;
                const float ic = __ldg(Ic), iyp = __ldg(Ic + dI_yp), iym = __ldg(Ic + dI_ym);
                const float ixp = __ldg(Ic + dI_xp), ixm = __ldg(Ic + dI_xm);
                const float gc = __fmul_rn(__ldg(Gc), gs), gyp = __fmul_rn(__ldg(Gc + dG_yp), gs);
                const float gym = __fmul_rn(__ldg(Gc + dG_ym), gs), gxp = __fmul_rn(__ldg(Gc + dG_xp), gs);
                const float gxm = __fmul_rn(__ldg(Gc + dG_xm), gs);
                gcen[c] = gc;
                // differentiation.py:19-29; a clamped (out-of-range) neighbour equals the centre, so its
                // difference is exactly zero and the term vanishes like the zero padding does
                ry_i = __fadd_rn(ry_i, __fmul_rn(__fsub_rn(ic, iyp), gyp));
                ly_i = __fadd_rn(ly_i, __fmul_rn(__fsub_rn(iyp, ic), gc));
                ry_m = __fadd_rn(ry_m, __fmul_rn(__fsub_rn(iym, ic), gc));
                ly_m = __fadd_rn(ly_m, __fmul_rn(__fsub_rn(ic, iym), gym));
                rx_i = __fadd_rn(rx_i, __fmul_rn(__fsub_rn(ic, ixp), gxp));
                lx_i = __fadd_rn(lx_i, __fmul_rn(__fsub_rn(ixp, ic), gc));
                rx_m = __fadd_rn(rx_m, __fmul_rn(__fsub_rn(ixm, ic), gc));
                lx_m = __fadd_rn(lx_m, __fmul_rn(__fsub_rn(ic, ixm), gxm));
            }
        }
        const float inv_step = __frcp_rn((float)(2. / R));
        const float gyr = __fadd_rn(__fmul_rn(-ry_i, inv_step), __fmul_rn(-ry_m, inv_step));
        const float gyl = __fadd_rn(__fmul_rn(-ly_m, inv_step), __fmul_rn(-ly_i, inv_step));
        const float gxr = __fadd_rn(__fmul_rn(-rx_i, inv_step), __fmul_rn(-rx_m, inv_step));
        const float gxl = __fadd_rn(__fmul_rn(-lx_m, inv_step), __fmul_rn(-lx_i, inv_step));
        gy = nr_maximum(gyr, gyl);
        gx = nr_maximum(gxr, gxl);
    }

    // ---- per-pixel contributions
    int vid[3] = {0, 0, 0};
    float vg[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // (x, y, z) gradient of the 3 corners
    float tg[12] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // 4 taps x rgb
    int tap[4] = {-1, -1, -1, -1};
    int cell = 0;
    bool has_tex = false;
    // CT = 1 is a lone silhouette or depth channel, CT >= 3 always holds the three colour channels
    const bool rgb = (CT == 1) ? false : ((CT >= 3) ? true : (a.flags & FLAG_RGB) != 0);
    constexpr bool LIT = (CT == 0);      // lights only in the generic variants (keeps the common ones small)
    const bool has_z = rgb || (a.flags & FLAG_DEPTH) != 0;
    if (fg) {
        if (a.faces) {
            vid[0] = __ldg(a.faces + 3 * (size_t)f);
            vid[1] = __ldg(a.faces + 3 * (size_t)f + 1);
            vid[2] = __ldg(a.faces + 3 * (size_t)f + 2);
        } else {
            vid[0] = 3 * f; vid[1] = 3 * f + 1; vid[2] = 3 * f + 2;
        }
        const float *vb = a.verts + (size_t)b * a.nv * 3;
        float Z[3] = {1.f, 1.f, 1.f}, q[3], tcs[3] = {0.f, 0.f, 0.f};
        if (AUX) {
            const float *ap = a.aux + (((size_t)b * R + yi) * R + xi) * (rgb ? 6 : 3);
            if (rgb) {
                const float2 a0 = __ldg(reinterpret_cast<const float2 *>(ap)), a1 = __ldg(reinterpret_cast<const float2 *>(ap) + 1),
                             a2 = __ldg(reinterpret_cast<const float2 *>(ap) + 2);
                q[0] = a0.x; q[1] = a0.y; q[2] = a1.x;
                tcs[0] = a1.y; tcs[1] = a2.x; tcs[2] = a2.y;
            } else {
                q[0] = __ldg(ap); q[1] = __ldg(ap + 1); q[2] = __ldg(ap + 2);
            }
            if (has_z) {
#pragma unroll
                for (int k = 0; k < 3; ++k) Z[k] = __ldg(vb + 3 * (size_t)vid[k] + 2);
            }
        } else {
            float X[3], Y[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                X[k] = __ldg(vb + 3 * (size_t)vid[k]);
                Y[k] = __ldg(vb + 3 * (size_t)vid[k] + 1);
                Z[k] = __ldg(vb + 3 * (size_t)vid[k] + 2);
            }
            const float xp = pix_center(xi, R), yp = pix_center(yi, R);
            raw_weights(xp, yp, X[0], Y[0], X[1], Y[1], X[2], Y[2], q[0], q[1], q[2]);
            normalize_weights<true>(q[0], q[1], q[2]);   // exact: 2 ulp on q move texel coordinates of ~1e3 by 1e-4
        }
        float gz[3] = {0.f, 0.f, 0.f};
        int c0 = 0;
        if (rgb) {
            const float g_lit[3] = {gcen[0], gcen[1], gcen[2]};
            if (g_lit[0] != 0.f || g_lit[1] != 0.f || g_lit[2] != 0.f) {
                const int32_t *fti = a.ft + 3 * (size_t)f;
                const float *vtb = a.vt + (size_t)b * a.nvt * 2;
                int tvid[3];
                float u[3], v[3];
                bool tex_ok = true;      // an out-of-range texture-vertex index contributes nothing (the forward drew black)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    tvid[k] = __ldg(fti + k);
                    const bool ok = (unsigned)tvid[k] < (unsigned)a.nvt;
                    tex_ok &= ok;
                    if (!ok) tvid[k] = 0;
                    const float2 uv = __ldg(reinterpret_cast<const float2 *>(vtb) + tvid[k]);
                    u[k] = uv.x;
                    v[k] = uv.y;
                }
                if (tex_ok) {
                float gu[3] = {0.f, 0.f, 0.f}, gv[3] = {0.f, 0.f, 0.f}, tw[4], g[3], rgb_tex[3];
                float cw[3] = {1.f, 1.f, 1.f}, nrm[3] = {0.f, 0.f, 0.f};
                const bool lit = LIT && a.lights.num > 0;
                if (lit) {
                    const float *vnb = a.lights.vnormals + (size_t)b * a.nv * 3;
#pragma unroll
                    for (int k = 0; k < 3; ++k)
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            nrm[c] = __fadd_rn(nrm[c], __fmul_rn(q[k], __ldg(vnb + 3 * (size_t)vid[k] + c)));
                    light_weights(a.lights, b, a.B, nrm, cw, nullptr, nullptr);
                }
                sample_texture_backward<NEED_UV>(a.tex + (size_t)b * 3 * a.H * a.W, a.H, a.W, a.eps, q, Z, u, v, AUX ? tcs : nullptr,
                                        g_lit, cw, g, rgb_tex, tap, tw, cell, gz, gu, gv);
                if (lit && a.lights.grad_vnormals) {
                    // d loss / d colour weight = upstream * unlit sample; then through the lights to the normal
                    const float gcw[3] = {g_lit[0] * rgb_tex[0], g_lit[1] * rgb_tex[1], g_lit[2] * rgb_tex[2]};
                    float cw2[3], gn[3];
                    light_weights(a.lights, b, a.B, nrm, cw2, gcw, gn);
                    const size_t ovn = (size_t)b * a.nv * 3;
#pragma unroll
                    for (int k = 0; k < 3; ++k)
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            if (gn[c] != 0.f)
                                accumulate<DET>(a.lights.grad_vnormals, a.det_vn, ovn + 3 * (size_t)vid[k] + c, q[k] * gn[c], a.det_scale);
                }
                has_tex = true;
#pragma unroll
                for (int t = 0; t < 4; ++t)
#pragma unroll
                    for (int c = 0; c < 3; ++c) tg[t * 3 + c] = tw[t] * g[c];
                if (NEED_UV) {
                    const size_t o = (size_t)b * a.nvt * 2;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        if (gu[k] != 0.f) accumulate<DET>(a.grad_vt, a.det_vt, o + 2 * (size_t)tvid[k], gu[k], a.det_scale);
                        if (gv[k] != 0.f) accumulate<DET>(a.grad_vt, a.det_vt, o + 2 * (size_t)tvid[k] + 1, gv[k], a.det_scale);
                    }
                }
                }   // tex_ok
            }
            c0 = 3;
        }
        if (a.flags & FLAG_SIL) ++c0;
        if (a.flags & FLAG_DEPTH) {
            const float gd = gcen[c0 < 5 ? c0 : 4];
            if (gd != 0.f) {
                const float s = (q[0] / Z[0] + q[1] / Z[1]) + q[2] / Z[2];
                const float dm = 1.f / s;
                const float t = gd * dm * dm;
#pragma unroll
                for (int k = 0; k < 3; ++k) gz[k] += t * q[k] / (Z[k] * Z[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            vg[3 * k] = q[k] * gx;
            vg[3 * k + 1] = q[k] * gy;
            vg[3 * k + 2] = gz[k];
        }
    }

    // ---- vertices: one atomic per (run of equal face, corner, coordinate)
    {
        bool nz = false;
#pragma unroll
        for (int i = 0; i < 9; ++i) nz |= (vg[i] != 0.f);
        if (__ballot_sync(0xffffffffu, nz) != 0u) {
            const bool tail = run_reduce<9>(f, fg, vg, lane);
            if (tail && fg) {
                const size_t o = (size_t)b * a.nv * 3;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const size_t ok = o + 3 * (size_t)vid[k];
                    float *pf = a.grad_verts + ok;
                    long long *pi = DET ? a.det_verts + ok : nullptr;
                    // one test per corner instead of one per component (every tested reduction costs a branch
                    // with a reconvergence barrier); z only moves where there is colour or depth to explain
                    if ((vg[3 * k] != 0.f) | (vg[3 * k + 1] != 0.f) | (vg[3 * k + 2] != 0.f)) {
                        accumulate<DET>(pf, pi, 0, vg[3 * k], a.det_scale);
                        accumulate<DET>(pf, pi, 1, vg[3 * k + 1], a.det_scale);
                        if (has_z) accumulate<DET>(pf, pi, 2, vg[3 * k + 2], a.det_scale);
                    }
                }
            }
        }
    }
    // ---- textures: one atomic per (run of equal base texel, tap, channel)
    if (rgb && a.grad_tex) {
        const bool has = has_tex;
        if (__ballot_sync(0xffffffffu, has) != 0u) {
            const bool tail = run_reduce<12>(cell, has, tg, lane);
            if (tail && has) {
                const size_t T = (size_t)a.H * a.W, o = (size_t)b * 3 * T;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    if (tap[t] < 0) continue;
                    float *pf = a.grad_tex + o + tap[t];
                    long long *pi = DET ? a.det_tex + o + tap[t] : nullptr;
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        accumulate<DET>(pf, pi, c * (int)T, tg[t * 3 + c], a.det_scale);
                }
            }
        }
    }
    }   // tiles
}

__global__ void __launch_bounds__(256)
k_fixed_to_float(const long long *__restrict__ src, float *__restrict__ dst, size_t n, float inv_scale) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += (float)((double)src[i] * (double)inv_scale);
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
    const long long tiles = (long long)a.ntx * a.ntx * a.B;
    // exactly the resident CTAs (4 per SM): a second wave of the static tile stride only adds a tail
    const int grid = (int)(tiles < (long long)a.sm_count * 4 ? tiles : (long long)a.sm_count * 4);
    dim3 block(TILE_THREADS);
    ProfScope p(PROF_BACKWARD, stream);
    if (a.det_verts) {
        if (a.grad_vt) launch_after(2, (long long)a.B * a.R * a.R, k_backward<0, true, true, false>, dim3(grid), block, 0, stream, a);
        else launch_after(2, (long long)a.B * a.R * a.R, k_backward<0, false, true, false>, dim3(grid), block, 0, stream, a);
    } else if (a.grad_vt) {
        if (a.aux) launch_after(2, (long long)a.B * a.R * a.R, k_backward<0, true, false, true>, dim3(grid), block, 0, stream, a);
        else launch_after(2, (long long)a.B * a.R * a.R, k_backward<0, true, false, false>, dim3(grid), block, 0, stream, a);
    } else if (a.aux) {
        switch (a.lights.num > 0 ? 0 : a.C) {
            case 1: launch_after(2, (long long)a.B * a.R * a.R, k_backward<1, false, false, true>, dim3(grid), block, 0, stream, a); break;
            case 3: launch_after(2, (long long)a.B * a.R * a.R, k_backward<3, false, false, true>, dim3(grid), block, 0, stream, a); break;
            case 4: launch_after(2, (long long)a.B * a.R * a.R, k_backward<4, false, false, true>, dim3(grid), block, 0, stream, a); break;
            default: launch_after(2, (long long)a.B * a.R * a.R, k_backward<0, false, false, true>, dim3(grid), block, 0, stream, a); break;
        }
    } else {
        switch (a.lights.num > 0 ? 0 : a.C) {
            case 1: launch_after(2, (long long)a.B * a.R * a.R, k_backward<1, false, false, false>, dim3(grid), block, 0, stream, a); break;
            case 3: launch_after(2, (long long)a.B * a.R * a.R, k_backward<3, false, false, false>, dim3(grid), block, 0, stream, a); break;
            case 4: launch_after(2, (long long)a.B * a.R * a.R, k_backward<4, false, false, false>, dim3(grid), block, 0, stream, a); break;
            default: launch_after(2, (long long)a.B * a.R * a.R, k_backward<0, false, false, false>, dim3(grid), block, 0, stream, a); break;
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || !a.det_verts) return e;
    // fixed point -> float, added onto the caller's (zero-filled) gradient tensors
    const float inv = 1.f / a.det_scale;
    auto conv = [&](const long long *src, float *dst, size_t n) {
        if (src && dst && n) k_fixed_to_float<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(src, dst, n, inv);
    };
    conv(a.det_verts, a.grad_verts, (size_t)a.B * a.nv * 3);
    conv(a.det_tex, a.grad_tex, (size_t)a.B * 3 * a.H * a.W);
    conv(a.det_vt, a.grad_vt, (size_t)a.B * a.nvt * 2);
    conv(a.det_vn, a.lights.grad_vnormals, (size_t)a.B * a.nv * 3);
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
