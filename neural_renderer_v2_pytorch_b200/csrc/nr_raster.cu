// nr_raster.cu -- tile rasterizer with fused shading epilogue.
//
// Persistent kernel: one WARP owns an 8x4 pixel block of a non-empty tile (16x16, or 8x8 for dense
// meshes), claimed dynamically from the heavy-first tile list.  It walks the tile's face list
// (ascending face index) 32 entries at a time: one lane per face culls against the block (pixel box,
// then a conservative edge-interval test, ballot), the survivors are staged in the warp's slice of
// shared memory, and only those are evaluated per pixel with the reference's arithmetic:
//   rasterize_cuda_kernel.cu:94-149  (z-buffer, sequential 1e-4 hysteresis)
// The epilogue fuses what the reference does in ~60 torch ops and 9*B host round trips:
//   rasterize_cuda_kernel.cu:246-308 weight map, rasterize.py:100-153 texture sampling,
//   :240-242 silhouettes, :80-88 depth, :295-310 channel merge, :315-316 permute + flip,
//   :321-328 2x2 anti-aliasing mean.
#include "nr_shade.cuh"

namespace nr {

// Conservative test "no pixel of the block [xa, xb] x [ya, yb] (pixel-centre coordinates) can pass
// the reference's inside test for this face" (rasterize_cuda_kernel.cu:107-116).  The reference accepts
// a pixel iff c1*c2 >= 0 and c2*c3 >= 0 in float arithmetic, which also lets every pixel with c2 == 0
// (or an underflowing product) through.  So a block is only rejected when, over the whole block, c2
// is bounded away from zero AND c1 or c3 is bounded away from zero with the opposite sign.  The edge
// functions are affine, so their range over the block is spanned by the four corners; the bounds
// carry a margin 40x above the rounding error of the reference's float evaluation.  NaN never rejects.
__device__ __forceinline__ bool block_outside_face(float x0, float y0, float x1, float y1, float x2,
                                                   float y2, float xa, float xb, float ya, float yb) {
    const float dx10 = x1 - x0, dy10 = y1 - y0, dx21 = x2 - x1, dy21 = y2 - y1, dx02 = x0 - x2, dy02 = y0 - y2;
    float lo1, hi1, lo2, hi2, lo3, hi3;
    {
        const float a = (ya - y0) * dx10, b = (yb - y0) * dx10, c = dy10 * (xa - x0), d = dy10 * (xb - x0);
        lo1 = fminf(a, b) - fmaxf(c, d);
        hi1 = fmaxf(a, b) - fminf(c, d);
        const float m = 1e-5f * (fmaxf(fabsf(a), fabsf(b)) + fmaxf(fabsf(c), fabsf(d))) + 1e-30f;
        lo1 -= m;
        hi1 += m;
    }
    {
        const float a = (ya - y1) * dx21, b = (yb - y1) * dx21, c = dy21 * (xa - x1), d = dy21 * (xb - x1);
        lo2 = fminf(a, b) - fmaxf(c, d);
        hi2 = fmaxf(a, b) - fminf(c, d);
        const float m = 1e-5f * (fmaxf(fabsf(a), fabsf(b)) + fmaxf(fabsf(c), fabsf(d))) + 1e-30f;
        lo2 -= m;
        hi2 += m;
    }
    {
        const float a = (ya - y2) * dx02, b = (yb - y2) * dx02, c = dy02 * (xa - x2), d = dy02 * (xb - x2);
        lo3 = fminf(a, b) - fmaxf(c, d);
        hi3 = fmaxf(a, b) - fminf(c, d);
        const float m = 1e-5f * (fmaxf(fabsf(a), fabsf(b)) + fmaxf(fabsf(c), fabsf(d))) + 1e-30f;
        lo3 -= m;
        hi3 += m;
    }
    const bool c2pos = lo2 > 0.f, c2neg = hi2 < 0.f;
    return (c2pos && (hi1 < 0.f || hi3 < 0.f)) || (c2neg && (lo1 > 0.f || lo3 > 0.f));
}


// Persistent kernel over the non-empty tiles.  The unit of work is one WARP = one 8x4 pixel block
// of a tile, claimed with one atomicAdd (the next claim is in flight while the current block is
// processed); the blocks of a tile are neighbours in the claim order, so its records are shared through
// L1 / L2.  Warps never wait for each other: a block without candidate faces costs one cull pass.
// Per 32 faces of the tile list: every lane loads one face record, tests its pixel box against the
// block, the survivors are compacted (ballot) into this warp's slice of shared memory with their edge
// deltas and their pixel box as a bit mask over the block, then evaluated per pixel in list order.
// Every pixel of the block is written (background included), see "output initialisation" above.
constexpr int RASTER_WARPS = TILE_THREADS / 32;

// NR_LAZY_Z: the z-test of a REGULAR candidate (face_z_regular, weights_one_sign: nothing cancels) is decided
// with fast_zp() whenever the outcome is clear of its error bound, which it is for everything but near-ties:
// the seven IEEE divisions of the reference's depth (~65 instructions) are then never executed.  The running
// minimum may therefore be the cheap depth of the current winner (dexact == false); the first comparison that
// is not clear recomputes it exactly from the winner's weights, so every DECISION equals the reference's.
#ifndef NR_LAZY_Z
#define NR_LAZY_Z 1
#endif
// NR_CPASYNC_STAGE (experiment, profiles/r2_cpasync_ab.txt): the 48-byte record of the NEXT group's face goes
// global -> shared with three cp.async (LDGSTS) per lane into a per-warp double buffer instead of three LDG.128
// into registers.  The records a warp needs are a gather (tile list -> record), so a bulk-tensor TMA copy has
// nothing contiguous to move; cp.async is the asynchronous copy that fits, and it was measured.
#ifndef NR_CPASYNC_STAGE
#define NR_CPASYNC_STAGE 0
#endif
constexpr int REC_Q = NR_LAZY_Z ? 5 : 4;       // float4 per staged face

__device__ __noinline__ float exact_zp_call(float w0, float w1, float w2, float z0, float z1, float z2) {
    return exact_zp(w0, w1, w2, z0, z1, z2);
}

// Variants: RGB (texture sampling), AA (2x2 mean epilogue), FULL (everything optional: lights,
// backgrounds, the weight / depth maps of rasterize_maps, resolutions that are no multiple of 16).
// The plain variants leave that code out, which halves their size: at 75 KB the one-size kernel spent
// as many issue slots waiting for instructions as for memory.
template <bool RGB, bool AA, bool FULL, bool FINE>
__global__ void __launch_bounds__(TILE_THREADS, 4)
k_raster(const RasterArgs a) {
    pdl_wait();         // the binning kernel's records, lists and header (nr_kernels.h)
    pdl_trigger();      // the backward may be scheduled as this kernel's CTAs leave
    // a tile is 16x16 pixels = 8 warp blocks (8x4 each), or in the FINE variants 8x8 = 2 blocks
    constexpr int TSZ = FINE ? FINE_TILE : TILE, BLK_SHIFT = FINE ? 1 : 3, BLKS = 1 << BLK_SHIFT;
    __shared__ float4 s_rec[RASTER_WARPS][32][REC_Q];
#if NR_CPASYNC_STAGE
    __shared__ float4 s_stage[RASTER_WARPS][2][32][3];
#endif
    __shared__ uint32_t s_bb[RASTER_WARPS][32];

    // If the pair list did not fit the workspace, the lists are unusable: every block then scans ALL
    // faces of its view (the records are complete regardless), which is the reference's own loop with
    // the block cull in front.  Slow but correct, and the host grows the workspace for the next call.
    const bool overflow = a.hdr->overflow != 0;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int R = a.R;
    const TileList tl = open_tile_list(a.tile_list, a.B * a.ntx * a.ntx);
    const int items = tl.total * BLKS;
    constexpr bool aa = AA;
    const bool has_bg = FULL && RGB && a.lights.backgrounds != nullptr;
    const bool pow2 = (R & (R - 1)) == 0;
    const float invR = 1.f / (float)R;          // exact for power-of-two R
    const unsigned lt_mask = (1u << lane) - 1u;
    float4 (*my_rec)[REC_Q] = s_rec[wid];
    uint32_t *my_bb = s_bb[wid];

    // ---- fill items (stores only, nr_shade.cuh), interleaved with the raster items: claim i also performs
    // fill item i (a static share per warp was slower: the warps holding the heavy tiles kept their fills
    // for the end)
    const FillPlan plan = make_fill_plan(a);
    const int fill_items = plan.fill_items;

    // Scheduling: the CTA claims EIGHT consecutive items (the blocks of one 16x16 tile) with one global
    // atomicAdd and its warps take them one by one from a shared cursor, so the blocks of a tile run on the
    // same SM at about the same time (face records and lists shared through L1) and no warp waits for
    // another except while a batch is being fetched.  One shared word holds both the batch and the number
    // of items taken from it: s_cur = (first item of the batch / 8) << 5 | taken; the warp that draws
    // taken == 8 fetches the next batch, later ones (at most 7: 5 bits are plenty) wait for it.
    constexpr int GRAB = 1, BATCH = 8;
    __shared__ unsigned s_cur;
    const int all_items = max(items, fill_items);
    if (threadIdx.x == 0) s_cur = ((unsigned)atomicAdd(&a.hdr->work_counter, BATCH) >> 3) << 5;
    __syncthreads();
    while (true) {
        int first = 0;
        if (lane == 0) {
            while (true) {
                const unsigned v = atomicAdd(&s_cur, 1u);
                const unsigned taken = v & 31u;
                if (taken < BATCH) {
                    first = (int)((v >> 5) << 3) + (int)taken;
                    break;
                }
                if (taken == BATCH) {
                    const unsigned base = (unsigned)atomicAdd(&a.hdr->work_counter, BATCH);
                    atomicExch(&s_cur, (base >> 3) << 5);
                } else {
                    while ((*reinterpret_cast<volatile unsigned *>(&s_cur) >> 5) == (v >> 5)) { }
                }
            }
        }
        first = __shfl_sync(0xffffffffu, first, 0);
        if (first >= all_items) break;
        if (first < fill_items) do_fill_item<AA, FULL, FINE>(a, plan, first, lane);
    for (int item = first; item < min(first + GRAB, items); ++item) {
        const int4 e0 = tile_entry(tl, item >> BLK_SHIFT);
        const int sub = item & (BLKS - 1);
        const int b = e0.x, n = overflow ? a.nf : e0.w;
        const int32_t *list = a.pairs + e0.z;
        const int wx0 = (e0.y & 0xffff) * TSZ + (FINE ? 0 : (sub & 1) * WARP_BW), wx1 = wx0 + WARP_BW - 1;
        const int wy0 = (e0.y >> 16) * TSZ + (FINE ? sub : (sub >> 1)) * WARP_BH, wy1 = wy0 + WARP_BH - 1;
        const int xi = wx0 + (lane & 7), yi = wy0 + (lane >> 3);
        const bool valid = (xi < R) && (yi < R);
        const float xp = pow2 ? __fmul_rn((float)(2 * xi + 1 - R), invR) : pix_center(xi, R);
        const float yp = pow2 ? __fmul_rn((float)(2 * yi + 1 - R), invR) : pix_center(yi, R);
        const FaceRec *rec_b = a.rec + (size_t)b * a.nf;
        // pixel centres of the block's corner pixels, for the conservative edge cull
        const float xa = pow2 ? __fmul_rn((float)(2 * wx0 + 1 - R), invR) : pix_center(wx0, R);
        const float xb = pow2 ? __fmul_rn((float)(2 * wx1 + 1 - R), invR) : pix_center(wx1, R);
        const float ya = pow2 ? __fmul_rn((float)(2 * wy0 + 1 - R), invR) : pix_center(wy0, R);
        const float yb = pow2 ? __fmul_rn((float)(2 * wy1 + 1 - R), invR) : pix_center(wy1, R);

        float depth_min = a.far_plane;
        bool dexact = true;       // depth_min is the reference's value (not the cheap depth of the current winner)
        int best = -1;
        float bw0 = 0.f, bw1 = 0.f, bw2 = 0.f, bz0 = 0.f, bz1 = 0.f, bz2 = 0.f;

        // software pipeline over the list: while group g is culled and evaluated, the records of
        // group g+1 and the ids of group g+2 are already in flight
        auto load_id = [&](int i) -> int { return i < n ? (overflow ? i : __ldg(list + i)) : -1; };
        int fid_next = load_id(lane);
        int fid_next2 = load_id(32 + lane);
#if NR_CPASYNC_STAGE
        int stage = 0;
        auto stage_record = [&](int st, int f_) {
            if (f_ >= 0) {
                const float4 *rp = reinterpret_cast<const float4 *>(rec_b + f_);
                const unsigned dst = (unsigned)__cvta_generic_to_shared(&s_stage[wid][st][lane][0]);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(rp));
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 16), "l"(rp + 1));
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 32), "l"(rp + 2));
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        stage_record(0, fid_next);
#else
        float4 n0 = make_float4(0.f, 0.f, 0.f, 0.f), n1 = n0, n2 = n0;
        if (fid_next >= 0) {
            const float4 *rp = reinterpret_cast<const float4 *>(rec_b + fid_next);
            n0 = __ldg(rp); n1 = __ldg(rp + 1); n2 = __ldg(rp + 2);
        }
#endif
        for (int g = 0; g < n; g += 32) {
            // ---- one face per lane: cull against this warp's block, compact the survivors
            const int fid = fid_next;
#if NR_CPASYNC_STAGE
            fid_next = fid_next2;
            stage_record(stage ^ 1, fid_next);                    // next group's records in flight ...
            asm volatile("cp.async.wait_group 1;" ::: "memory");   // ... while this group's have landed
            float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0, q2 = q0;
            if (fid >= 0) {
                q0 = s_stage[wid][stage][lane][0]; q1 = s_stage[wid][stage][lane][1]; q2 = s_stage[wid][stage][lane][2];
            }
            stage ^= 1;
#else
            const float4 q0 = n0, q1 = n1, q2 = n2;
            fid_next = fid_next2;
            if (fid_next >= 0) {
                const float4 *rp = reinterpret_cast<const float4 *>(rec_b + fid_next);
                n0 = __ldg(rp); n1 = __ldg(rp + 1); n2 = __ldg(rp + 2);
            }
#endif
            fid_next2 = load_id(g + 64 + lane);
            bool hit = false;
            if (fid >= 0) {
                const uint32_t bx = __float_as_uint(q2.y), by = __float_as_uint(q2.z);
                hit = ((int)(bx & 0xffff) <= wx1) && ((int)(bx >> 16) >= wx0) && ((int)(by & 0xffff) <= wy1) &&
                      ((int)(by >> 16) >= wy0);
                if (hit) hit = !block_outside_face(q0.x, q0.y, q0.w, q1.x, q1.z, q1.w, xa, xb, ya, yb);
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m == 0u) continue;
            if (hit) {
                const int slot = __popc(m & lt_mask);      // keeps list order
                const float x0 = q0.x, y0 = q0.y, z0 = q0.z, x1 = q0.w;
                const float y1 = q1.x, z1 = q1.y, x2 = q1.z, y2 = q1.w, z2 = q2.x;
                my_rec[slot][0] = make_float4(x0, y0, x1, y1);
                my_rec[slot][1] = make_float4(x2, y2, __fsub_rn(x1, x0), __fsub_rn(y1, y0));
                my_rec[slot][2] = make_float4(__fsub_rn(x2, x1), __fsub_rn(y2, y1), __fsub_rn(x0, x2), __fsub_rn(y0, y2));
                my_rec[slot][3] = make_float4(z0, z1, z2, __int_as_float(fid));
#if NR_LAZY_Z
                // reciprocal depths and the smallest corner depth of a regular face (NaN marks an irregular one)
                const bool zreg = face_z_regular(z0, z1, z2);
                my_rec[slot][4] = make_float4(fast_rcp(z0), fast_rcp(z1), fast_rcp(z2),
                                              zreg ? fminf(z0, fminf(z1, z2)) : __int_as_float(0x7fc00000));
#endif
                // the face's pixel box (:94-97, exact by construction) as a bit mask over this block's 32
                // pixels (bit = lane): the columns / rows it covers, relative to the block origin
                const uint32_t bx = __float_as_uint(q2.y), by = __float_as_uint(q2.z);
                const int c0 = max((int)(bx & 0xffff) - wx0, 0), c1 = min((int)(bx >> 16) - wx0, WARP_BW - 1);
                const int r0 = max((int)(by & 0xffff) - wy0, 0), r1 = min((int)(by >> 16) - wy0, WARP_BH - 1);
                const uint32_t cols = ((2u << c1) - 1u) & ~((1u << c0) - 1u);          // 8 bits
                const uint32_t rows = ((2u << r1) - 1u) & ~((1u << r0) - 1u);          // 4 bits
                my_bb[slot] = cols * ((rows * 0x00204081u) & 0x01010101u);              // row r -> bits 8r .. 8r+7
            }
            __syncwarp();
            const int nh = __popc(m);
            // ---- phase 1: coverage of every surviving face, one bit per face (cheap, all lanes)
            unsigned inside = 0u;
            // (branch-free: at least one lane of the block is inside every surviving face's pixel box, so the
            // warp executes the edge functions anyway)
            for (int j = 0; j < nh; ++j) {
                const bool in_box = (my_bb[j] >> lane) & 1u;
                const float4 A = my_rec[j][0], Bq = my_rec[j][1], Cq = my_rec[j][2];
                // :107-116
                const float c1 = __fmaf_rn(__fsub_rn(yp, A.y), Bq.z, -__fmul_rn(Bq.w, __fsub_rn(xp, A.x)));
                const float c2 = __fmaf_rn(__fsub_rn(yp, A.w), Cq.x, -__fmul_rn(Cq.y, __fsub_rn(xp, A.z)));
                const float c3 = __fmaf_rn(__fsub_rn(yp, Bq.y), Cq.z, -__fmul_rn(Cq.w, __fsub_rn(xp, Bq.x)));
                const bool rejected = (__fmul_rn(c1, c2) < 0.f) | (__fmul_rn(c2, c3) < 0.f);
                if (in_box && !rejected) inside |= 1u << j;
            }
            // ---- phase 2: every lane walks ITS OWN covering faces in list order (the z-test is
            // sequential per pixel only), so lanes evaluate different faces at the same time and the
            // loop runs max-depth-complexity times instead of once per surviving face
            while (inside) {
                const int j = __ffs(inside) - 1;
                inside &= inside - 1;
                const float4 D = my_rec[j][3];
#if NR_LAZY_Z
                const float4 E = my_rec[j][4];
                // surely depth_min < every corner depth (:124-126 skips the face); false for an irregular face
                if (depth_min * (1.f + FAST_Z_REL) < E.w) continue;
                const float4 A = my_rec[j][0], Bq = my_rec[j][1];
                float w0, w1, w2;
                raw_weights(xp, yp, A.x, A.y, A.z, A.w, Bq.x, Bq.y, w0, w1, w2);
                if (E.w == E.w && weights_one_sign(w0, w1, w2)) {
                    const float zf = fast_zp(w0, w1, w2, E.x, E.y, E.z);
                    const float m = 2.f * FAST_Z_REL * fmaxf(zf, depth_min);     // error of zf plus that of depth_min
                    const float t = depth_min - a.delta;
                    // (every comparison is false for a NaN: such a candidate takes the exact path)
                    if (zf < a.near_plane - m || zf > a.far_plane + m || zf > t + m) continue;      // :140-148 reject it
                    if (zf > a.near_plane + m && zf < a.far_plane - m && zf < t - m) {              // ... accept it
                        depth_min = zf;
                        dexact = false;
                        best = __float_as_int(D.w);
                        bw0 = w0; bw1 = w1; bw2 = w2;
                        bz0 = D.x; bz1 = D.y; bz2 = D.z;
                        continue;
                    }
                }
                // not clear: the reference's own arithmetic, against the exact running minimum
                if (!dexact) {
                    depth_min = exact_zp_call(bw0, bw1, bw2, bz0, bz1, bz2);
                    dexact = true;
                }
                if (depth_min < D.x && depth_min < D.y && depth_min < D.z) continue;      // :124-126
                const float zp = exact_zp_call(w0, w1, w2, D.x, D.y, D.z);                 // :129-139
#else
                // :124-126
                if (depth_min < D.x && depth_min < D.y && depth_min < D.z) continue;
                const float4 A = my_rec[j][0], Bq = my_rec[j][1];
                // :129-136
                float w0, w1, w2;
                raw_weights(xp, yp, A.x, A.y, A.z, A.w, Bq.x, Bq.y, w0, w1, w2);
                const float zp = exact_zp(w0, w1, w2, D.x, D.y, D.z);
#endif
                // :139-142
                if (zp <= a.near_plane || a.far_plane <= zp) continue;
                // :145-148
                if (zp <= __fsub_rn(depth_min, a.delta)) {
                    depth_min = zp;
                    best = __float_as_int(D.w);
                    bw0 = w0; bw1 = w1; bw2 = w2;
                    bz0 = D.x; bz1 = D.y; bz2 = D.z;
                }
            }
            __syncwarp();
        }
        // ---------------------------------------------- epilogue: every pixel of the block is written
        shade_block<RGB, AA, FULL>(a, b, xi, yi, valid, best, bw0, bw1, bw2, bz0, bz1, bz2, has_bg);
    }   // items of this claim
    }   // claims
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

// The optional per-pixel maps (weight_map, depth_map: rasterize_maps / the compat operators, not
// on the hot path) are written for foreground pixels only, so they start as zeros.
cudaError_t launch_background_fill(const RasterArgs &a, cudaStream_t stream) {
    if (a.B <= 0 || a.R <= 0 || (!a.wmap && !a.dmap)) return cudaSuccess;
    ProfScope p(PROF_MEMSET, stream);
    const size_t P = (size_t)a.B * a.R * a.R;
    cudaError_t e = cudaSuccess;
    if (a.wmap) e = cudaMemsetAsync(a.wmap, 0, P * 3 * sizeof(float), stream);
    if (e == cudaSuccess && a.dmap) e = cudaMemsetAsync(a.dmap, 0, P * sizeof(float), stream);
    return e;
}

cudaError_t launch_raster(const RasterArgs &a, cudaStream_t stream) {
    if (a.B <= 0 || a.R <= 0) return cudaSuccess;
    const long long tiles = (long long)a.ntx * a.ntx * a.B;
    const int grid = (int)(tiles < (long long)a.sm_count * 8 ? tiles : (long long)a.sm_count * 8);
    const bool rgb = (a.flags & FLAG_RGB) != 0, aa = (a.flags & FLAG_AA) != 0;
    // everything optional goes to the FULL variants
    const bool full = a.lights.num > 0 || a.lights.backgrounds || a.wmap || a.dmap || !a.images || (a.R & 15);
    ProfScope p(PROF_RASTER, stream);
    if (a.fine) {       // dense meshes: the full-featured kernels over 8x8 tiles
        switch ((rgb ? 1 : 0) | (aa ? 2 : 0)) {
            case 0: launch_after(1, (long long)a.B * a.R * a.R, k_raster<false, false, true, true>, dim3(grid), dim3(TILE_THREADS), 0, stream, a); break;
            case 1: launch_after(1, (long long)a.B * a.R * a.R, k_raster<true, false, true, true>, dim3(grid), dim3(TILE_THREADS), 0, stream, a); break;
            case 2: launch_after(1, (long long)a.B * a.R * a.R, k_raster<false, true, true, true>, dim3(grid), dim3(TILE_THREADS), 0, stream, a); break;
            default: launch_after(1, (long long)a.B * a.R * a.R, k_raster<true, true, true, true>, dim3(grid), dim3(TILE_THREADS), 0, stream, a); break;
        }
        return cudaGetLastError();
    }
    const int variant = (rgb ? 1 : 0) | (aa ? 2 : 0) | (full ? 4 : 0);
    switch (variant) {
        case 0: launch_after(1, (long long)a.B * a.R * a.R, k_raster<false, false, false, false>, dim3(grid), dim3(TILE_THREADS), 0, stream, a); break;
        case 1: launch_after(1, (long long)a.B * a.R * a.R, k_raster<true, false, false, false>, dim3(grid), dim3(TILE_THREADS), 0, stream, a); break;
        case 2: launch_after(1, (long long)a.B * a.R * a.R, k_raster<false, true, false, false>, dim3(grid), dim3(TILE_THREADS), 0, stream, a); break;
        case 3: launch_after(1, (long long)a.B * a.R * a.R, k_raster<true, true, false, false>, dim3(grid), dim3(TILE_THREADS), 0, stream, a); break;
        case 4: launch_after(1, (long long)a.B * a.R * a.R, k_raster<false, false, true, false>, dim3(grid), dim3(TILE_THREADS), 0, stream, a); break;
        case 5: launch_after(1, (long long)a.B * a.R * a.R, k_raster<true, false, true, false>, dim3(grid), dim3(TILE_THREADS), 0, stream, a); break;
        case 6: launch_after(1, (long long)a.B * a.R * a.R, k_raster<false, true, true, false>, dim3(grid), dim3(TILE_THREADS), 0, stream, a); break;
        default: launch_after(1, (long long)a.B * a.R * a.R, k_raster<true, true, true, false>, dim3(grid), dim3(TILE_THREADS), 0, stream, a); break;
    }
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
