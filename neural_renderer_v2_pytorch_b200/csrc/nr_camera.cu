// nr_camera.cu -- fused camera transform.
//
// World-space vertices + per-view camera -> screen-space vertices in one kernel, and its backward in
// one kernel, instead of the ~10 + ~15 elementwise / bmm launches of
//   look_at   (neural_renderer_torch/look_at.py:28-42):  (v - eye) @ R^T
//   perspective (neural_renderer_torch/perspective.py:9-17): x / z / width, y / z / width, z
// The 3x3 rotation R (rows = camera x, y, z axes) is built by the caller from the eye / at / up
// vectors with torch ops on [B,3] tensors, so its own backward (normalize, cross) stays in autograd.
#include "../../include/nr_b200.h"
#include "nr_kernels.h"

namespace nr {

constexpr int CAM_THREADS = 256;

// shared != 0: `verts` is ONE mesh [1, nv, 3] seen by every camera (blockIdx.y still walks the views)
__global__ void __launch_bounds__(CAM_THREADS)
k_camera_forward(const float *__restrict__ verts, const float *__restrict__ rot, const float *__restrict__ eye,
                 float *__restrict__ out, int nv, int perspective, float width, int shared) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv) return;
    const float *R = rot + (size_t)b * 9, *E = eye + (size_t)b * 3;
    const float *v = verts + ((size_t)(shared ? 0 : b) * nv + i) * 3;
    const float d0 = __fsub_rn(v[0], E[0]), d1 = __fsub_rn(v[1], E[1]), d2 = __fsub_rn(v[2], E[2]);
    float c[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) c[r] = __fmaf_rn(R[3 * r + 2], d2, __fmaf_rn(R[3 * r + 1], d1, __fmul_rn(R[3 * r], d0)));
    float *o = out + ((size_t)b * nv + i) * 3;
    if (perspective) {
        o[0] = __fdiv_rn(__fdiv_rn(c[0], c[2]), width);
        o[1] = __fdiv_rn(__fdiv_rn(c[1], c[2]), width);
    } else {
        o[0] = c[0];
        o[1] = c[1];
    }
    o[2] = c[2];
}

// d loss / d (vertex, rotation, eye) of one (view, vertex): gd = gradient of the world-space vertex,
// acc[0..8] += d rotation (row-major), acc[9..11] += d eye
__device__ __forceinline__ void camera_backward_one(const float *R, const float *E, const float *v, const float *g,
                                                    int perspective, float width, float gd[3], float acc[12]) {
    const float d[3] = {v[0] - E[0], v[1] - E[1], v[2] - E[2]};
    float gc[3] = {g[0], g[1], g[2]};
    if (perspective) {
        float c[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) c[r] = R[3 * r] * d[0] + R[3 * r + 1] * d[1] + R[3 * r + 2] * d[2];
        const float iz = 1.f / c[2], izw = iz / width;
        gc[2] = g[2] - (g[0] * c[0] + g[1] * c[1]) * iz * izw;
        gc[0] = g[0] * izw;
        gc[1] = g[1] * izw;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) gd[j] = R[j] * gc[0] + R[3 + j] * gc[1] + R[6 + j] * gc[2];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[3 * r + j] = gc[r] * d[j];
    acc[9] = -gd[0];
    acc[10] = -gd[1];
    acc[11] = -gd[2];
}

// CTA sum of acc[12] -> partial[12] (fixed order)
__device__ __forceinline__ void camera_block_sum(float acc[12], float (*s_red)[12], float *partial) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        float x = acc[k];
        for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) s_red[wid][k] = x;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        float x = 0.f;
        for (int w = 0; w < CAM_THREADS / 32; ++w) x += s_red[w][threadIdx.x];
        partial[threadIdx.x] = x;
    }
    __syncthreads();
}

// grad_verts [B,nv,3] written; per-CTA partial sums of d loss / d R (9) and d loss / d eye (3) written
// to partial [B, gridDim.x, 12] (summed by the caller in a fixed order: deterministic), unless it is null.
__global__ void __launch_bounds__(CAM_THREADS)
k_camera_backward(const float *__restrict__ verts, const float *__restrict__ rot, const float *__restrict__ eye,
                  const float *__restrict__ gout, float *__restrict__ gverts, float *__restrict__ partial, int nv,
                  int perspective, float width) {
    __shared__ float s_red[CAM_THREADS / 32][12];
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float acc[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[k] = 0.f;
    if (i < nv) {
        float gd[3];
        camera_backward_one(rot + (size_t)b * 9, eye + (size_t)b * 3, verts + ((size_t)b * nv + i) * 3,
                            gout + ((size_t)b * nv + i) * 3, perspective, width, gd, acc);
        float *gv = gverts + ((size_t)b * nv + i) * 3;
        gv[0] = gd[0];
        gv[1] = gd[1];
        gv[2] = gd[2];
    }
    if (partial) camera_block_sum(acc, s_red, partial + ((size_t)b * gridDim.x + blockIdx.x) * 12);
}

// One mesh [1,nv,3] seen by B cameras: a thread owns a vertex and walks the views in order, so the gradient of
// the shared mesh is summed in registers (fixed order, no [B,nv,3] intermediate, no atomics).
__global__ void __launch_bounds__(CAM_THREADS)
k_camera_backward_shared(const float *__restrict__ verts, const float *__restrict__ rot, const float *__restrict__ eye,
                         const float *__restrict__ gout, float *__restrict__ gverts, float *__restrict__ partial, int B,
                         int nv, int perspective, float width) {
    __shared__ float s_red[CAM_THREADS / 32][12];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < nv;
    float v[3] = {0.f, 0.f, 0.f}, sum[3] = {0.f, 0.f, 0.f};
    if (in) {
        v[0] = verts[3 * (size_t)i];
        v[1] = verts[3 * (size_t)i + 1];
        v[2] = verts[3 * (size_t)i + 2];
    }
    for (int b = 0; b < B; ++b) {
        float acc[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) acc[k] = 0.f;
        if (in) {
            float gd[3];
            camera_backward_one(rot + (size_t)b * 9, eye + (size_t)b * 3, v, gout + ((size_t)b * nv + i) * 3, perspective,
                                width, gd, acc);
            sum[0] += gd[0];
            sum[1] += gd[1];
            sum[2] += gd[2];
        }
        if (partial) camera_block_sum(acc, s_red, partial + ((size_t)b * gridDim.x + blockIdx.x) * 12);
    }
    if (in) {
        gverts[3 * (size_t)i] = sum[0];
        gverts[3 * (size_t)i + 1] = sum[1];
        gverts[3 * (size_t)i + 2] = sum[2];
    }
}

// ---- shared mesh on several GPUs: camera backward FUSED with the all-reduce of its result ------------------
// Multi-view optimisation of one mesh on N GPUs (BASELINE config 3): every rank holds B of the views; the
// gradient of the shared [1,nv,3] mesh is the sum over all views of all ranks.  Instead of this kernel followed
// by an NCCL all-reduce of 12*nv bytes (a 0.16 ms step grows by 50 us on 8 GPUs), the exchange happens INSIDE
// the kernel over NVLink peer memory (every rank maps every rank's buffer: symmetric memory):
//
//   A thread owns a vertex.  It sums the gradient over the local views in registers (as above) and PUSHES the
//   three sums into its slots of every peer's receive buffer, each as one 8-byte word (value, epoch): the word is
//   written atomically, so a reader that finds this step's epoch in it has the value too - no flag behind a
//   fence (a release / acquire pair at system scope cost 20 us per step; NCCL's LL protocol makes the same
//   trade: twice the bytes, no fence).  Then it polls its own buffer for the three words of every peer, adds
//   them IN RANK ORDER (every rank gets the same bits, run to run) and writes the result.  Nothing else
//   synchronises: no flags, no CTA or grid barrier, no kernel boundary; words travel while other CTAs compute.
//
// Receive slots of two consecutive steps alternate (epoch parity): a rank cannot be more than one step ahead of a
// peer (it needs that peer's words of the step before), so nobody overwrites a word that is still to be read.
// Epochs only grow, so the buffers are never reset.  A thread pushes BEFORE it polls, so the wait cannot deadlock
// whatever part of the grid is resident.
constexpr int CAM_MAX_RANKS = 16;
struct PeerExchange {
    int rank, world;
    unsigned long long *recv[CAM_MAX_RANKS];    // every rank's receive buffer: [2][world][nv * 3] words (value | epoch << 32)
    int *epoch;                                  // this rank's {step counter, CTAs done, peer timed out, -} (device memory, zeroed once)
};

__device__ __forceinline__ void st_word_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_word_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(CAM_THREADS)
k_camera_backward_shared_exchange(const float *__restrict__ verts, const float *__restrict__ rot, const float *__restrict__ eye,
                                  const float *__restrict__ gout, float *__restrict__ gverts, float *__restrict__ partial,
                                  int B, int nv, int perspective, float width, const PeerExchange px) {
    __shared__ float s_red[CAM_THREADS / 32][12];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < nv;
    // this step (bumped by the last CTA to finish); unsigned: it wraps after 2^32 steps, and only equality and parity are used
    const unsigned e = (unsigned)*reinterpret_cast<volatile int *>(px.epoch) + 1u;
    float v[3] = {0.f, 0.f, 0.f}, sum[3] = {0.f, 0.f, 0.f};
    if (in) {
        v[0] = verts[3 * (size_t)i];
        v[1] = verts[3 * (size_t)i + 1];
        v[2] = verts[3 * (size_t)i + 2];
    }
    for (int b = 0; b < B; ++b) {
        float acc[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) acc[k] = 0.f;
        if (in) {
            float gd[3];
            camera_backward_one(rot + (size_t)b * 9, eye + (size_t)b * 3, v, gout + ((size_t)b * nv + i) * 3, perspective,
                                width, gd, acc);
            sum[0] += gd[0];
            sum[1] += gd[1];
            sum[2] += gd[2];
        }
        if (partial) camera_block_sum(acc, s_red, partial + ((size_t)b * gridDim.x + blockIdx.x) * 12);
    }
    if (in) {
        const size_t half = (size_t)nv * 3;
        const unsigned long long tag = (unsigned long long)e << 32;
        // ---- push my words into every peer's slots [parity][my rank]
        const size_t mine = ((size_t)(e & 1) * px.world + px.rank) * half + 3 * (size_t)i;
        for (int r = 0; r < px.world; ++r) {
            if (r == px.rank) continue;
            unsigned long long *dst = px.recv[r] + mine;
#pragma unroll
            for (int k = 0; k < 3; ++k) st_word_sys(dst + k, tag | __float_as_uint(sum[k]));
        }
        // ---- the peers' words in my buffer, added in rank order: the same sum, bit for bit, on every rank
        float t[3] = {0.f, 0.f, 0.f};
        for (int r = 0; r < px.world; ++r) {
            if (r == px.rank) {
                t[0] += sum[0]; t[1] += sum[1]; t[2] += sum[2];
                continue;
            }
            const unsigned long long *src = px.recv[px.rank] + ((size_t)(e & 1) * px.world + r) * half + 3 * (size_t)i;
            unsigned long long w0 = ld_word_sys(src), w1 = ld_word_sys(src + 1), w2 = ld_word_sys(src + 2);
            // (a peer that never arrives - a crashed process - must not hang the GPU: give up after a few seconds
            // and say so in epoch[2]; the gradient of this step is then garbage)
            long long spins = 0;
            while ((unsigned)(w0 >> 32) != e || (unsigned)(w1 >> 32) != e || (unsigned)(w2 >> 32) != e) {
                __nanosleep(32);
                if (++spins > (1ll << 25)) {
                    px.epoch[2] = 1;
                    break;
                }
                w0 = ld_word_sys(src); w1 = ld_word_sys(src + 1); w2 = ld_word_sys(src + 2);
            }
            t[0] += __uint_as_float((unsigned)w0); t[1] += __uint_as_float((unsigned)w1); t[2] += __uint_as_float((unsigned)w2);
        }
        gverts[3 * (size_t)i] = t[0];
        gverts[3 * (size_t)i + 1] = t[1];
        gverts[3 * (size_t)i + 2] = t[2];
    }
    // ---- the last CTA to finish opens the next step
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(px.epoch + 1, 1) == (int)gridDim.x - 1) {
            px.epoch[1] = 0;
            __threadfence();
            *reinterpret_cast<volatile int *>(px.epoch) = (int)e;
        }
    }
}

}  // namespace nr

extern "C" {

int nr_camera_partial_blocks(int32_t num_vertices) { return (num_vertices + nr::CAM_THREADS - 1) / nr::CAM_THREADS; }

int nr_camera_forward(const float *vertices, const float *rotation, const float *eye, float *out, int32_t batch,
                      int32_t num_vertices, int32_t perspective, float width, int32_t shared_mesh, void *stream) {
    if (!vertices || !rotation || !eye || !out || batch < 0 || num_vertices < 0 || batch > 65535) return NR_ERR_INVALID_ARGUMENT;
    if (batch == 0 || num_vertices == 0) return NR_OK;
    dim3 grid(nr_camera_partial_blocks(num_vertices), batch);
    nr::ProfScope p(nr::PROF_CAMERA_FORWARD, (cudaStream_t)stream);
    nr::k_camera_forward<<<grid, nr::CAM_THREADS, 0, (cudaStream_t)stream>>>(vertices, rotation, eye, out, num_vertices,
                                                                            perspective, width, shared_mesh);
    return cudaGetLastError() == cudaSuccess ? NR_OK : NR_ERR_CUDA;
}

int nr_camera_backward(const float *vertices, const float *rotation, const float *eye, const float *grad_out,
                       float *grad_vertices, float *partial, int32_t batch, int32_t num_vertices,
                       int32_t perspective, float width, int32_t shared_mesh, void *stream) {
    if (!vertices || !rotation || !eye || !grad_out || !grad_vertices || batch < 0 || num_vertices < 0 || batch > 65535)
        return NR_ERR_INVALID_ARGUMENT;
    if (batch == 0 || num_vertices == 0) return NR_OK;
    nr::ProfScope p(nr::PROF_CAMERA_BACKWARD, (cudaStream_t)stream);
    if (shared_mesh) {
        nr::k_camera_backward_shared<<<nr_camera_partial_blocks(num_vertices), nr::CAM_THREADS, 0, (cudaStream_t)stream>>>(
            vertices, rotation, eye, grad_out, grad_vertices, partial, batch, num_vertices, perspective, width);
    } else {
        dim3 grid(nr_camera_partial_blocks(num_vertices), batch);
        nr::k_camera_backward<<<grid, nr::CAM_THREADS, 0, (cudaStream_t)stream>>>(vertices, rotation, eye, grad_out,
                                                                                 grad_vertices, partial, num_vertices,
                                                                                 perspective, width);
    }
    return cudaGetLastError() == cudaSuccess ? NR_OK : NR_ERR_CUDA;
}

int nr_camera_exchange_bytes(int32_t num_vertices, int32_t world) {
    // receive slots [2 (epoch parity)][world][nv * 3] of 8-byte words
    const size_t total = 2 * (size_t)world * (size_t)num_vertices * 3 * sizeof(unsigned long long);
    return (num_vertices <= 0 || world < 1 || total > 0x7fffffff) ? -1 : (int)total;
}

int nr_camera_backward_shared_allreduce(const float *vertices, const float *rotation, const float *eye, const float *grad_out,
                                        float *grad_vertices, float *partial, int32_t batch, int32_t num_vertices,
                                        int32_t perspective, float width, int32_t rank, int32_t world,
                                        void *const *peer_buffers, int32_t *epoch, void *stream) {
    if (!vertices || !rotation || !eye || !grad_out || !grad_vertices || !peer_buffers || !epoch || batch < 0 ||
        num_vertices <= 0 || world < 1 || world > nr::CAM_MAX_RANKS || rank < 0 || rank >= world)
        return NR_ERR_INVALID_ARGUMENT;
    const int slices = nr_camera_partial_blocks(num_vertices);
    // the wait inside the kernel needs every CTA of the grid resident at once
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nr::k_camera_backward_shared_exchange, nr::CAM_THREADS, 0);
    if ((long long)slices > (long long)sms * per_sm) return NR_ERR_INVALID_ARGUMENT;
    nr::PeerExchange px;
    px.rank = rank;
    px.world = world;
    for (int r = 0; r < world; ++r) {
        if (!peer_buffers[r]) return NR_ERR_INVALID_ARGUMENT;
        px.recv[r] = (unsigned long long *)peer_buffers[r];
    }
    px.epoch = epoch;
    nr::ProfScope p(nr::PROF_CAMERA_BACKWARD, (cudaStream_t)stream);
    nr::k_camera_backward_shared_exchange<<<slices, nr::CAM_THREADS, 0, (cudaStream_t)stream>>>(
        vertices, rotation, eye, grad_out, grad_vertices, partial, batch, num_vertices, perspective, width, px);
    return cudaGetLastError() == cudaSuccess ? NR_OK : NR_ERR_CUDA;
}

}  // extern "C"
