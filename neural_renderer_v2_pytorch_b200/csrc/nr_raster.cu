// nr_raster.cu -- tile rasterizer with fused shading epilogue.
//
// One CTA per (view, 16x16 tile), one thread per pixel.  The tile's face list (ascending face
// index) is staged through shared memory in chunks of 256 records; each warp owns an 8x4 pixel
// block, culls the chunk against that block 32 faces at a time (one lane per face, ballot),
// and only the surviving faces are evaluated per pixel with the reference's arithmetic:
//   rasterize_cuda_kernel.cu:94-149  (z-buffer, sequential 1e-4 hysteresis)
// The epilogue fuses what the reference does in ~60 torch ops and 9*B host round trips:
//   rasterize_cuda_kernel.cu:246-308 weight map, rasterize.py:100-153 texture sampling,
//   :240-242 silhouettes, :80-88 depth, :295-310 channel merge, :315-316 permute + flip,
//   :321-328 2x2 anti-aliasing mean.
#include "nr_kernels.h"

namespace nr {

// Ascending sort of a[0..n) in shared memory, n <= SMEM_SORT_CAP. Face ids inside one tile list
// are unique.  Already-sorted lists (the common case: the scatter kernel mostly claims slots in
// face order) only pay the check.
__device__ void sort_ids_smem(int *a, int *scratch, int n) {
    const int tid = threadIdx.x;
    int unsorted = 0;
    for (int i = tid; i + 1 < n; i += TILE_THREADS) unsorted |= (a[i] > a[i + 1]);
    if (!__syncthreads_or(unsorted)) return;
    if (n <= TILE_THREADS) {
        // rank sort: position = number of smaller ids
        int v = 0, rank = 0;
        if (tid < n) {
            v = a[tid];
            for (int j = 0; j < n; ++j) rank += (a[j] < v);
        }
        __syncthreads();
        if (tid < n) a[rank] = v;
        __syncthreads();
        return;
    }
    (void)scratch;
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int k = 2; k <= np2; k <<= 1) {
        for (int i = tid; i < np2 / 2; i += TILE_THREADS) {
            const int blk = i / (k / 2), off = i % (k / 2);
            const int lo = blk * k + off, hi = blk * k + k - 1 - off;
            if (hi < n) {
                const int x = a[lo], y = a[hi];
                if (x > y) {
                    a[lo] = y;
                    a[hi] = x;
                }
            }
        }
        __syncthreads();
        for (int j = k / 4; j >= 1; j >>= 1) {
            for (int i = tid; i < np2 / 2; i += TILE_THREADS) {
                const int lo = (i / j) * 2 * j + (i % j), hi = lo + j;
                if (hi < n) {
                    const int x = a[lo], y = a[hi];
                    if (x > y) {
                        a[lo] = y;
                        a[hi] = x;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// Perspective-correct bilinear texture sample of one foreground pixel, rasterize.py:100-153.
// q = weight map, z = face depths, uv = texel coordinates of the 3 face corners.
__device__ __forceinline__ void sample_texture(const float *__restrict__ tex_b, int H, int W,
                                               float eps, const float q[3], const float z[3],
                                               const float u[3], const float v[3], float rgb[3]) {
    const TexCoord tc = texel_coord(q, z, u, v, eps);
    const float xf = tc.xf, yf = tc.yf;
    const float xff = floorf(xf), yff = floorf(yf);
    const float xcf = __fadd_rn(xff, 1.f), ycf = __fadd_rn(yff, 1.f);
    const int xfi = (int)xff, yfi = (int)yff, xci = (int)xcf, yci = (int)ycf;
    const float w1 = __fmul_rn(__fsub_rn(ycf, yf), __fsub_rn(xcf, xf));
    const float w2 = __fmul_rn(__fsub_rn(ycf, yf), __fsub_rn(xf, xff));
    const float w3 = __fmul_rn(__fsub_rn(yf, yff), __fsub_rn(xcf, xf));
    const float w4 = __fmul_rn(__fsub_rn(yf, yff), __fsub_rn(xf, xff));
    const int T = H * W;
    const int i1 = yfi * W + xfi, i2 = yfi * W + xci, i3 = yci * W + xfi, i4 = yci * W + xci;
    // to_map (utils.py:104-114) yields zero for a negative index; an index >= H*W is an
    // IndexError in the reference and reads as zero here.
    const bool ok1 = (unsigned)i1 < (unsigned)T, ok2 = (unsigned)i2 < (unsigned)T;
    const bool ok3 = (unsigned)i3 < (unsigned)T, ok4 = (unsigned)i4 < (unsigned)T;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float *p = tex_b + (size_t)c * T;
        const float t1 = ok1 ? __ldg(p + i1) : 0.f, t2 = ok2 ? __ldg(p + i2) : 0.f;
        const float t3 = ok3 ? __ldg(p + i3) : 0.f, t4 = ok4 ? __ldg(p + i4) : 0.f;
        rgb[c] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, t1), __fmul_rn(w2, t2)), __fmul_rn(w3, t3)),
                           __fmul_rn(w4, t4));
    }
}

__global__ void __launch_bounds__(TILE_THREADS)
k_raster(const RasterArgs a) {
    __shared__ int s_ids[SMEM_SORT_CAP];
    __shared__ float4 s_rec[TILE_THREADS][4];
    __shared__ uint2 s_bb[TILE_THREADS];

    if (a.hdr->overflow) return;
    const int tid = threadIdx.x, lane = tid & 31;
    const int tile = blockIdx.x, b = blockIdx.y;
    const int tx = tile % a.ntx, ty = tile / a.ntx;
    const int R = a.R;
    int px, py;
    tile_pixel(tid, px, py);
    const int xi = tx * TILE + px, yi = ty * TILE + py;
    const bool valid = (xi < R) && (yi < R);
    const float xp = pix_center(xi, R), yp = pix_center(yi, R);
    // pixel block owned by this warp, for the warp-level cull
    const int wx0 = tx * TILE + ((tid >> 5) & 1) * WARP_BW, wx1 = wx0 + WARP_BW - 1;
    const int wy0 = ty * TILE + (tid >> 6) * WARP_BH, wy1 = wy0 + WARP_BH - 1;

    const int tflat = b * a.ntx * a.ntx + tile;
    const int n = a.tile_count[tflat];
    const int32_t *list = a.pairs + a.tile_offset[tflat];
    const bool in_smem = (n <= SMEM_SORT_CAP);
    if (in_smem && n > 0) {
        for (int i = tid; i < n; i += TILE_THREADS) s_ids[i] = list[i];
        __syncthreads();
        sort_ids_smem(s_ids, nullptr, n);
    }

    float depth_min = a.far_plane;
    int best = -1;
    float bw0 = 0.f, bw1 = 0.f, bw2 = 0.f, bz0 = 0.f, bz1 = 0.f, bz2 = 0.f;
    const FaceRec *rec_b = a.rec + (size_t)b * a.nf;

    for (int c0 = 0; c0 < n; c0 += TILE_THREADS) {
        const int cn = min(TILE_THREADS, n - c0);
        if (tid < cn) {
            const int fid = in_smem ? s_ids[c0 + tid] : list[c0 + tid];
            const float4 *rp = reinterpret_cast<const float4 *>(rec_b + fid);
            const float4 q0 = __ldg(rp), q1 = __ldg(rp + 1), q2 = __ldg(rp + 2);
            const float x0 = q0.x, y0 = q0.y, z0 = q0.z, x1 = q0.w;
            const float y1 = q1.x, z1 = q1.y, x2 = q1.z, y2 = q1.w, z2 = q2.x;
            s_rec[tid][0] = make_float4(x0, y0, x1, y1);
            s_rec[tid][1] = make_float4(x2, y2, __fsub_rn(x1, x0), __fsub_rn(y1, y0));
            s_rec[tid][2] = make_float4(__fsub_rn(x2, x1), __fsub_rn(y2, y1), __fsub_rn(x0, x2), __fsub_rn(y0, y2));
            s_rec[tid][3] = make_float4(z0, z1, z2, __int_as_float(fid));
            s_bb[tid] = make_uint2(__float_as_uint(q2.y), __float_as_uint(q2.z));
        }
        __syncthreads();
        for (int g = 0; g < cn; g += 32) {
            bool hit = false;
            if (g + lane < cn) {
                const uint2 bb = s_bb[g + lane];
                const int xlo = bb.x & 0xffff, xhi = bb.x >> 16, ylo = bb.y & 0xffff, yhi = bb.y >> 16;
                hit = (xlo <= wx1) && (xhi >= wx0) && (ylo <= wy1) && (yhi >= wy0);
            }
            unsigned m = __ballot_sync(0xffffffffu, hit);
            while (m) {
                const int j = g + __ffs(m) - 1;
                m &= m - 1;
                const uint2 bb = s_bb[j];
                // :94-97, exact by construction of the pixel box
                if (xi < (int)(bb.x & 0xffff) || xi > (int)(bb.x >> 16) || yi < (int)(bb.y & 0xffff) ||
                    yi > (int)(bb.y >> 16))
                    continue;
                const float4 A = s_rec[j][0], Bq = s_rec[j][1], Cq = s_rec[j][2];
                // :107-116
                const float c1 = __fmaf_rn(__fsub_rn(yp, A.y), Bq.z, -__fmul_rn(Bq.w, __fsub_rn(xp, A.x)));
                const float c2 = __fmaf_rn(__fsub_rn(yp, A.w), Cq.x, -__fmul_rn(Cq.y, __fsub_rn(xp, A.z)));
                if (__fmul_rn(c1, c2) < 0.f) continue;
                const float c3 = __fmaf_rn(__fsub_rn(yp, Bq.y), Cq.z, -__fmul_rn(Cq.w, __fsub_rn(xp, Bq.x)));
                if (__fmul_rn(c2, c3) < 0.f) continue;
                const float4 D = s_rec[j][3];
                // :124-126
                if (depth_min < D.x && depth_min < D.y && depth_min < D.z) continue;
                // :129-136
                float w0, w1, w2;
                raw_weights(xp, yp, A.x, A.y, A.z, A.w, Bq.x, Bq.y, w0, w1, w2);
                const float ws = __fadd_rn(__fadd_rn(w0, w1), w2);
                const float n0 = __fdiv_rn(w0, ws), n1 = __fdiv_rn(w1, ws), n2 = __fdiv_rn(w2, ws);
                // :139-142
                const float s = __fadd_rn(__fadd_rn(__fdiv_rn(n0, D.x), __fdiv_rn(n1, D.y)), __fdiv_rn(n2, D.z));
                const float zp = __frcp_rn(s);
                if (zp <= a.near_plane || a.far_plane <= zp) continue;
                // :145-148
                if (zp <= __fsub_rn(depth_min, a.delta)) {
                    depth_min = zp;
                    best = __float_as_int(D.w);
                    bw0 = w0; bw1 = w1; bw2 = w2;
                    bz0 = D.x; bz1 = D.y; bz2 = D.z;
                }
            }
        }
        __syncthreads();
    }

    // ------------------------------------------------------------------ epilogue
    const bool fg = best >= 0;
    float q[3] = {0.f, 0.f, 0.f};
    if (fg) {
        q[0] = bw0; q[1] = bw1; q[2] = bw2;
        normalize_weights(q[0], q[1], q[2]);
    }
    const size_t pix = ((size_t)b * R + yi) * R + xi;
    if (valid) {
        a.fim[pix] = best;
        if (a.wmap) {
            float *w = a.wmap + pix * 3;
            w[0] = q[0]; w[1] = q[1]; w[2] = q[2];
        }
    }
    float dm = 0.f;
    if (fg && ((a.flags & FLAG_DEPTH) || a.dmap))
        dm = __fdiv_rn(1.f, __fadd_rn(__fadd_rn(__fdiv_rn(q[0], bz0), __fdiv_rn(q[1], bz1)), __fdiv_rn(q[2], bz2)));
    if (valid && a.dmap) a.dmap[pix] = dm;
    if (!a.images) return;

    float ch[5];
    int C = 0;
    if (a.flags & FLAG_RGB) {
        float rgb[3] = {0.f, 0.f, 0.f};
        if (fg) {
            const int32_t *fti = a.ft + 3 * (size_t)best;
            const float *vtb = a.vt + (size_t)b * a.nvt * 2;
            float u[3], v[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int t = __ldg(fti + k);
                const float2 uv = __ldg(reinterpret_cast<const float2 *>(vtb) + t);
                u[k] = uv.x;
                v[k] = uv.y;
            }
            const float z[3] = {bz0, bz1, bz2};
            sample_texture(a.tex + (size_t)b * 3 * a.H * a.W, a.H, a.W, a.eps, q, z, u, v, rgb);
        }
        ch[0] = rgb[0]; ch[1] = rgb[1]; ch[2] = rgb[2];
        C = 3;
    }
    if (a.flags & FLAG_SIL) ch[C++] = fg ? 1.f : 0.f;
    if (a.flags & FLAG_DEPTH) ch[C++] = dm;

    const int u_ = R - 1 - yi, v_ = R - 1 - xi;   // flipped coordinates, rasterize.py:316
    if (!(a.flags & FLAG_AA)) {
        if (valid) {
            for (int c = 0; c < C; ++c) a.images[(((size_t)b * C + c) * R + u_) * R + v_] = ch[c];
        }
    } else {
        const int S = a.S;
        const bool writer = valid && !(xi & 1) && !(yi & 1);
        for (int c = 0; c < C; ++c) {
            const float me = ch[c];
            if (valid && a.internal) a.internal[(((size_t)b * C + c) * R + u_) * R + v_] = me;
            // quad in flipped coordinates: F[2Y][2X] is (yi odd, xi odd); rasterize.py:323-328
            const float px_ = __shfl_xor_sync(0xffffffffu, me, 1);   // same row, other column
            const float py_ = __shfl_xor_sync(0xffffffffu, me, 8);   // other row, same column
            const float pd_ = __shfl_xor_sync(0xffffffffu, me, 9);
            if (writer) {
                // me = (even, even) -> F[2Y+1][2X+1]; py_ = (odd row, even col) -> F[2Y][2X+1]
                // px_ = (even row, odd col) -> F[2Y+1][2X]; pd_ = (odd, odd) -> F[2Y][2X]
                const float sum = __fadd_rn(__fadd_rn(__fadd_rn(pd_, px_), py_), me);
                a.images[(((size_t)b * C + c) * S + (u_ >> 1)) * S + (v_ >> 1)] = __fmul_rn(sum, 0.25f);
            }
        }
    }
}

// compute_weight_map_c compatibility kernel (rasterize_cuda_kernel.cu:246-308): one thread per
// pixel, faces given as the gathered [B, nf, 3, 3] tensor, background pixels untouched.
__global__ void __launch_bounds__(256)
k_weight_map_compat(const float *__restrict__ faces, const int32_t *__restrict__ fim,
                    float *__restrict__ wmap, long long total, int nf, int R) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int fi = fim[i];
    if (fi < 0) return;
    const long long pp = (long long)R * R;
    const int bn = (int)(i / pp), pn = (int)(i % pp);
    const int yi = pn / R, xi = pn % R;
    const float xp = pix_center(xi, R), yp = pix_center(yi, R);
    const float *f = faces + ((size_t)bn * nf + fi) * 9;
    float w0, w1, w2;
    raw_weights(xp, yp, f[0], f[1], f[3], f[4], f[6], f[7], w0, w1, w2);
    normalize_weights(w0, w1, w2);
    wmap[i * 3 + 0] = w0;
    wmap[i * 3 + 1] = w1;
    wmap[i * 3 + 2] = w2;
}

cudaError_t launch_raster(const RasterArgs &a, cudaStream_t stream) {
    if (a.B <= 0 || a.R <= 0) return cudaSuccess;
    dim3 grid(a.ntx * a.ntx, a.B);
    ProfScope p(PROF_RASTER, stream);
    k_raster<<<grid, TILE_THREADS, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_weight_map_compat(const float *faces, const int32_t *fim, float *wmap, int B,
                                     int nf, int R, cudaStream_t stream) {
    const long long total = (long long)B * R * R;
    if (total <= 0) return cudaSuccess;
    ProfScope p(PROF_WEIGHT_MAP, stream);
    k_weight_map_compat<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(faces, fim, wmap, total, nf, R);
    return cudaGetLastError();
}

}  // namespace nr
