// nr_raster_dense.cu -- rasterizer for meshes of SMALL triangles (hundreds of faces per 16x16 tile, a few
// pixels per face: a 100 k-face sphere at 512^2, a million random triangles at 1024^2).
//
// There the block kernel (nr_raster.cu) wastes its lanes: it evaluates every face of a tile list for all 32
// pixels of a warp block although the face covers a handful, and it needs the lists in ascending face order,
// which costs a sort as long as the rasterization itself.  This kernel turns the loops around:
//
//   one CTA per non-empty tile, ONE THREAD PER FACE of the tile's (unsorted) list.  The thread walks the
//   pixels of the face's exact pixel box inside the tile (:94-97), evaluates the reference's edge functions
//   (:107-116) and, where the pixel is covered, the cheap depth fast_zp() (nr_shade.cuh), and merges
//   (depth, face) into the tile's z-buffer in shared memory with a 64-bit compare-and-swap.
//
// The reference's z-test is a SEQUENTIAL scan with a 1e-4 hysteresis (:145-148), not a minimum.  But for a
// pixel whose closest candidate is closer than every other candidate by more than the hysteresis (plus the
// error bound of the cheap depth), the scan provably ends on that candidate whatever the order: a candidate
// is only ever replaced by one at least delta closer, and the closest one is accepted whenever it arrives.
// So the kernel keeps, per pixel, the minimum and a "contested" bit, set whenever a candidate comes within
// the band of the minimum of its moment (the later of two near candidates always sees the earlier one or a
// closer one, so no near-tie with the final minimum is missed), or is irregular in any way (weights of mixed
// sign: the c2 == 0 quirk of :109,114; depths that are not ordinary positive numbers; a depth within the
// error bound of near or far).  Contested pixels - a fraction of a percent - are then resolved EXACTLY by one
// warp each: it collects every face of the list that covers the pixel with the reference's own depth
// (exact_zp), orders the candidates by face index and replays the reference's scan (:124-148) on them.
// face_index_map is bit-identical to the block kernel's and the reference's (tests: every parity case runs
// through both kernels).
#include "nr_shade.cuh"

namespace nr {

constexpr unsigned long long DENSE_EMPTY = ~0ull;
constexpr int DENSE_BIG_AREA = 48;          // faces covering more pixels of the tile are rasterized by a whole warp
constexpr int DENSE_BIG_QUEUE = 256;
constexpr int DENSE_CAND = 64;              // candidates per contested pixel held in shared memory (more: slow path)

struct DenseShared {
    unsigned long long key[TILE * TILE];    // (bits of the cheap depth) << 32 | face
    unsigned flag[TILE * TILE / 32];        // contested pixels
    int big[DENSE_BIG_QUEUE];               // list positions of large faces
    int nbig;
    int item;
    // per warp: candidates of the pixel being resolved, unsorted then sorted by face index
    int cand_f[TILE_THREADS / 32][2][DENSE_CAND];
    float cand_z[TILE_THREADS / 32][2][DENSE_CAND];
    float cand_c[TILE_THREADS / 32][2][DENSE_CAND];
};

struct DenseFace {
    float x0, y0, z0, x1, y1, z1, x2, y2, z2;
    float iz0, iz1, iz2;
    bool zreg;
    int fid;
};

// One (face, pixel) candidate: coverage by the reference's edge functions, then the cheap depth into the
// tile's z-buffer.  p = pixel index inside the tile (row * 16 + column).
__device__ __forceinline__ void dense_candidate(DenseShared &sh, const DenseFace &f, float xp, float yp, int p,
                                                float near_plane, float far_plane, float delta) {
    // :107-116
    const float c1 = __fmaf_rn(__fsub_rn(yp, f.y0), __fsub_rn(f.x1, f.x0), -__fmul_rn(__fsub_rn(f.y1, f.y0), __fsub_rn(xp, f.x0)));
    const float c2 = __fmaf_rn(__fsub_rn(yp, f.y1), __fsub_rn(f.x2, f.x1), -__fmul_rn(__fsub_rn(f.y2, f.y1), __fsub_rn(xp, f.x1)));
    const float c3 = __fmaf_rn(__fsub_rn(yp, f.y2), __fsub_rn(f.x0, f.x2), -__fmul_rn(__fsub_rn(f.y0, f.y2), __fsub_rn(xp, f.x2)));
    if ((__fmul_rn(c1, c2) < 0.f) | (__fmul_rn(c2, c3) < 0.f)) return;
    float w0, w1, w2;
    raw_weights(xp, yp, f.x0, f.y0, f.x1, f.y1, f.x2, f.y2, w0, w1, w2);
    const unsigned bit = 1u << (p & 31);
    unsigned *flag = &sh.flag[p >> 5];
    if (!(f.zreg && weights_one_sign(w0, w1, w2))) {
        atomicOr(flag, bit);            // irregular: the exact scan decides this pixel
        return;
    }
    const float zf = fast_zp(w0, w1, w2, f.iz0, f.iz1, f.iz2);
    const float m = FAST_Z_REL * zf;
    // :140-142 rejects zp <= near and zp >= far; a depth in (far - delta, far) is valid but can never pass the
    // z-test against the initial minimum `far`, nor against a smaller one: it is irrelevant as well
    const float top = far_plane - delta;
    if (zf < near_plane - m || zf > top + m) return;
    // (a NaN fails every comparison above and the one below: contested)
    if (!(zf > near_plane + m && zf < top - m && zf * 1e-6f < delta)) {
        atomicOr(flag, bit);
        return;
    }
    const unsigned long long mine = ((unsigned long long)__float_as_uint(zf) << 32) | (unsigned)f.fid;
    unsigned long long old = *reinterpret_cast<volatile unsigned long long *>(&sh.key[p]);
    while (true) {
        if (old != DENSE_EMPTY) {
            const float zo = __uint_as_float((unsigned)(old >> 32));
            // within the hysteresis (plus both error bounds and the rounding of depth_min - delta) of the
            // minimum of this moment: order may matter
            if (fabsf(zf - zo) < delta + 2.5f * FAST_Z_REL * fmaxf(zf, zo)) atomicOr(flag, bit);
        }
        if (mine >= old) break;
        const unsigned long long prev = atomicCAS(&sh.key[p], old, mine);
        if (prev == old) break;
        old = prev;
    }
}

// Exact resolution of one contested pixel by a whole warp (all lanes call it with the same arguments).
// Returns the winning face of the reference's scan (:82-149) restricted to this pixel, or -1.
__device__ __noinline__ int dense_resolve_pixel(DenseShared &sh, const RasterArgs &a, const FaceRec *rec_b,
                                                const int32_t *list, int n, bool overflow, int xi, int yi, float xp,
                                                float yp, int lane, int wid) {
    const unsigned lt_mask = (1u << lane) - 1u;
    int *cf = sh.cand_f[wid][0];
    float *cz = sh.cand_z[wid][0], *cc = sh.cand_c[wid][0];
    int count = 0;
    // ---- every face of the list that covers the pixel with a depth the reference would look at
    auto probe = [&](int i, int &fid, float &zp, float &zc) -> bool {
        fid = overflow ? i : __ldg(list + i);
        const float4 *rp = reinterpret_cast<const float4 *>(rec_b + fid);
        const float4 q2 = __ldg(rp + 2);
        const uint32_t bx = __float_as_uint(q2.y), by = __float_as_uint(q2.z);
        if (xi < (int)(bx & 0xffff) || xi > (int)(bx >> 16) || yi < (int)(by & 0xffff) || yi > (int)(by >> 16)) return false;
        const float4 q0 = __ldg(rp), q1 = __ldg(rp + 1);
        const float x0 = q0.x, y0 = q0.y, z0 = q0.z, x1 = q0.w, y1 = q1.x, z1 = q1.y, x2 = q1.z, y2 = q1.w, z2 = q2.x;
        const float c1 = __fmaf_rn(__fsub_rn(yp, y0), __fsub_rn(x1, x0), -__fmul_rn(__fsub_rn(y1, y0), __fsub_rn(xp, x0)));
        const float c2 = __fmaf_rn(__fsub_rn(yp, y1), __fsub_rn(x2, x1), -__fmul_rn(__fsub_rn(y2, y1), __fsub_rn(xp, x1)));
        const float c3 = __fmaf_rn(__fsub_rn(yp, y2), __fsub_rn(x0, x2), -__fmul_rn(__fsub_rn(y0, y2), __fsub_rn(xp, x2)));
        if ((__fmul_rn(c1, c2) < 0.f) | (__fmul_rn(c2, c3) < 0.f)) return false;
        float w0, w1, w2;
        raw_weights(xp, yp, x0, y0, x1, y1, x2, y2, w0, w1, w2);
        zp = exact_zp(w0, w1, w2, z0, z1, z2);
        // the scan skips a face when depth_min < z0 && depth_min < z1 && depth_min < z2 (:124-126): with a NaN
        // corner depth that is never true, otherwise it is depth_min < min(z)
        zc = (z0 != z0 || z1 != z1 || z2 != z2) ? -__int_as_float(0x7f800000) : fminf(z0, fminf(z1, z2));
        // :140-142; a NaN depth passes this test but then fails zp <= depth_min - delta: it never matters
        return !(zp <= a.near_plane || a.far_plane <= zp) && zp == zp;
    };
    for (int g = 0; g < n; g += 32) {
        const int i = g + lane;
        int fid = 0;
        float zp = 0.f, zc = 0.f;
        const bool keep = i < n && probe(i, fid, zp, zc);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        const int slot = count + __popc(m & lt_mask);
        if (keep && slot < DENSE_CAND) {
            cf[slot] = fid;
            cz[slot] = zp;
            cc[slot] = zc;
        }
        count += __popc(m);
    }
    __syncwarp();
    float depth_min = a.far_plane;
    int best = -1;
    if (count <= DENSE_CAND) {
        // ---- order by face index (rank = number of smaller ids), then the reference's scan
        int *sf = sh.cand_f[wid][1];
        float *sz = sh.cand_z[wid][1], *sc = sh.cand_c[wid][1];
        const int f0 = lane < count ? cf[lane] : 0x7fffffff, f1 = lane + 32 < count ? cf[lane + 32] : 0x7fffffff;
        int r0 = 0, r1 = 0;
        for (int j = 0; j < count; ++j) {
            const int x = cf[j];
            r0 += x < f0;
            r1 += x < f1;
        }
        if (lane < count) { sf[r0] = f0; sz[r0] = cz[lane]; sc[r0] = cc[lane]; }
        if (lane + 32 < count) { sf[r1] = f1; sz[r1] = cz[lane + 32]; sc[r1] = cc[lane + 32]; }
        __syncwarp();
        for (int j = 0; j < count; ++j) {
            const float zp = sz[j];
            if (depth_min < sc[j]) continue;                         // :124-126
            if (zp <= __fsub_rn(depth_min, a.delta)) {                // :145-148
                depth_min = zp;
                best = sf[j];
            }
        }
        __syncwarp();
        return best;
    }
    // ---- more candidates than the buffer holds (dozens of layers over one pixel): walk the faces in
    // ascending index by repeated selection of the smallest id above the last one
    int last = -1;
    while (true) {
        int mn = 0x7fffffff;
        float mzp = 0.f, mzc = 0.f;
        for (int g = 0; g < n; g += 32) {
            const int i = g + lane;
            if (i >= n) continue;
            const int id = overflow ? i : __ldg(list + i);
            if (id <= last || id >= mn) continue;
            int fid;
            float zp, zc;
            if (probe(i, fid, zp, zc)) { mn = fid; mzp = zp; mzc = zc; }
        }
        int wmn = mn;
        for (int o = 16; o; o >>= 1) wmn = min(wmn, __shfl_xor_sync(0xffffffffu, wmn, o));
        if (wmn == 0x7fffffff) break;
        const unsigned owner = __ballot_sync(0xffffffffu, mn == wmn);
        const int src = __ffs(owner) - 1;
        const float zp = __shfl_sync(0xffffffffu, mzp, src), zc = __shfl_sync(0xffffffffu, mzc, src);
        last = wmn;
        if (depth_min < zc) continue;
        if (zp <= __fsub_rn(depth_min, a.delta)) {
            depth_min = zp;
            best = wmn;
        }
    }
    return best;
}

template <bool RGB, bool AA, bool FULL>
__global__ void __launch_bounds__(TILE_THREADS, 4)
k_raster_dense(const RasterArgs a) {
    __shared__ DenseShared sh;
    const bool overflow = a.hdr->overflow != 0;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int R = a.R;
    const PixGrid grid(R);
    const TileList tl = open_tile_list(a.tile_list, a.B * a.ntx * a.ntx);
    const bool has_bg = FULL && RGB && a.lights.backgrounds != nullptr;
    const FillPlan plan = make_fill_plan(a);
    // claim i = tile i of the heavy-first list, and its eight warps perform fill items 8 i .. 8 i + 7
    const int all_items = max(tl.total, (plan.fill_items + 7) >> 3);
    int px, py;
    tile_pixel(tid, px, py);                 // the pixel this thread resolves and shades
    const int p_own = py * TILE + px;

    if (tid == 0) sh.item = atomicAdd(&a.hdr->work_counter, 1);
    __syncthreads();
    while (true) {
        const int item = sh.item;
        if (item >= all_items) break;
        const int fi = item * 8 + wid;
        if (fi < plan.fill_items) do_fill_item<AA, FULL, false>(a, plan, fi, lane);
        if (item < tl.total) {
            const int4 e0 = tile_entry(tl, item);
            const int b = e0.x, n = overflow ? a.nf : e0.w;
            const int32_t *list = a.pairs + e0.z;
            const int tx0 = (e0.y & 0xffff) * TILE, ty0 = (e0.y >> 16) * TILE;
            const FaceRec *rec_b = a.rec + (size_t)b * a.nf;
            sh.key[tid] = DENSE_EMPTY;
            if (tid < TILE * TILE / 32) sh.flag[tid] = 0u;
            if (tid == 0) sh.nbig = 0;
            __syncthreads();

            // ---- one thread per face of the list
            auto load_face = [&](int i, DenseFace &f, int &bx0, int &bx1, int &by0, int &by1) -> bool {
                f.fid = overflow ? i : __ldg(list + i);
                const float4 *rp = reinterpret_cast<const float4 *>(rec_b + f.fid);
                const float4 q2 = __ldg(rp + 2);
                const uint32_t bx = __float_as_uint(q2.y), by = __float_as_uint(q2.z);
                // the exact pixel box of the face (:94-97), clipped to this tile
                bx0 = max((int)(bx & 0xffff), tx0); bx1 = min((int)(bx >> 16), tx0 + TILE - 1);
                by0 = max((int)(by & 0xffff), ty0); by1 = min((int)(by >> 16), ty0 + TILE - 1);
                if (bx0 > bx1 || by0 > by1) return false;
                const float4 q0 = __ldg(rp), q1 = __ldg(rp + 1);
                f.x0 = q0.x; f.y0 = q0.y; f.z0 = q0.z; f.x1 = q0.w;
                f.y1 = q1.x; f.z1 = q1.y; f.x2 = q1.z; f.y2 = q1.w; f.z2 = q2.x;
                f.zreg = face_z_regular(f.z0, f.z1, f.z2);
                f.iz0 = fast_rcp(f.z0); f.iz1 = fast_rcp(f.z1); f.iz2 = fast_rcp(f.z2);
                return true;
            };
            for (int i = tid; i < n; i += TILE_THREADS) {
                DenseFace f;
                int bx0, bx1, by0, by1;
                if (!load_face(i, f, bx0, bx1, by0, by1)) continue;
                if ((bx1 - bx0 + 1) * (by1 - by0 + 1) > DENSE_BIG_AREA) {
                    const int q = atomicAdd(&sh.nbig, 1);
                    if (q < DENSE_BIG_QUEUE) {
                        sh.big[q] = i;
                        continue;
                    }
                }
                for (int y = by0; y <= by1; ++y) {
                    const float yp = grid.center(y);
                    for (int x = bx0; x <= bx1; ++x)
                        dense_candidate(sh, f, grid.center(x), yp, (y - ty0) * TILE + (x - tx0), a.near_plane, a.far_plane, a.delta);
                }
            }
            __syncthreads();
            // ---- large faces: a warp per face, a lane per pixel of its box
            const int nbig = min(sh.nbig, DENSE_BIG_QUEUE);
            for (int q = wid; q < nbig; q += TILE_THREADS / 32) {
                DenseFace f;
                int bx0, bx1, by0, by1;
                if (!load_face(sh.big[q], f, bx0, bx1, by0, by1)) continue;
                const int w = bx1 - bx0 + 1, area = w * (by1 - by0 + 1);
                for (int k = lane; k < area; k += 32) {
                    const int y = by0 + k / w, x = bx0 + k % w;
                    dense_candidate(sh, f, grid.center(x), grid.center(y), (y - ty0) * TILE + (x - tx0), a.near_plane, a.far_plane, a.delta);
                }
            }
            __syncthreads();

            // ---- resolve: the minimum where it is uncontested, the reference's scan elsewhere
            const int xi = tx0 + px, yi = ty0 + py;
            const bool valid = xi < R && yi < R;
            const unsigned long long key = sh.key[p_own];
            int best = key == DENSE_EMPTY ? -1 : (int)(unsigned)(key & 0xffffffffu);
            const bool contested = (sh.flag[p_own >> 5] >> (p_own & 31)) & 1u;
            const float xp = grid.center(xi), yp = grid.center(yi);
            unsigned todo = __ballot_sync(0xffffffffu, contested && valid);
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const int rx = __shfl_sync(0xffffffffu, xi, src), ry = __shfl_sync(0xffffffffu, yi, src);
                const float rxp = __shfl_sync(0xffffffffu, xp, src), ryp = __shfl_sync(0xffffffffu, yp, src);
                const int r = dense_resolve_pixel(sh, a, rec_b, list, n, overflow, rx, ry, rxp, ryp, lane, wid);
                if (lane == src) best = r;
            }
            // ---- the winner's raw weights and corner depths for the shading epilogue
            float bw0 = 0.f, bw1 = 0.f, bw2 = 0.f, bz0 = 0.f, bz1 = 0.f, bz2 = 0.f;
            if (best >= 0) {
                const float4 *rp = reinterpret_cast<const float4 *>(rec_b + best);
                const float4 q0 = __ldg(rp), q1 = __ldg(rp + 1), q2 = __ldg(rp + 2);
                raw_weights(xp, yp, q0.x, q0.y, q0.w, q1.x, q1.z, q1.w, bw0, bw1, bw2);
                bz0 = q0.z; bz1 = q1.y; bz2 = q2.x;
            }
            shade_block<RGB, AA, FULL>(a, b, xi, yi, valid, best, bw0, bw1, bw2, bz0, bz1, bz2, has_bg);
        }
        __syncthreads();
        if (tid == 0) sh.item = atomicAdd(&a.hdr->work_counter, 1);
        __syncthreads();
    }
}

cudaError_t launch_raster_dense(const RasterArgs &a, cudaStream_t stream) {
    if (a.B <= 0 || a.R <= 0) return cudaSuccess;
    const long long tiles = (long long)a.ntx * a.ntx * a.B;
    const int grid = (int)(tiles < (long long)a.sm_count * 4 ? tiles : (long long)a.sm_count * 4);
    const bool rgb = (a.flags & FLAG_RGB) != 0, aa = (a.flags & FLAG_AA) != 0;
    const bool full = a.lights.num > 0 || a.lights.backgrounds || a.wmap || a.dmap || !a.images || (a.R & 15);
    ProfScope p(PROF_RASTER_DENSE, stream);
    // silhouettes of dense meshes are the measured case; everything else takes the full-featured variants
    if (!rgb && !aa && !full) k_raster_dense<false, false, false><<<grid, TILE_THREADS, 0, stream>>>(a);
    else if (rgb && aa) k_raster_dense<true, true, true><<<grid, TILE_THREADS, 0, stream>>>(a);
    else if (rgb) k_raster_dense<true, false, true><<<grid, TILE_THREADS, 0, stream>>>(a);
    else if (aa) k_raster_dense<false, true, true><<<grid, TILE_THREADS, 0, stream>>>(a);
    else k_raster_dense<false, false, true><<<grid, TILE_THREADS, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace nr
