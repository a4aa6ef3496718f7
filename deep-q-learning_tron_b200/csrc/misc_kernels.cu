// misc_kernels.cu -- export/import of env state, action policies, replay ring push/gather/sample.
#include <cuda_bf16.h>

#include "common.cuh"
#include "launch.h"

namespace tron {

// ------------------------------------------------------------------------------------------------
// state export / import (history, drop-in shims, tests)
// ------------------------------------------------------------------------------------------------
__global__ void export_meta_kernel(const uint2* __restrict__ meta, int n, int8_t* heads, uint8_t* alive, uint8_t* done,
                                   uint8_t* winner, int32_t* ep_len) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const uint2 m = meta[e];
    const uint32_t f = m.y & 0xFFu;
    if (heads) ((uint32_t*)heads)[e] = m.x;
    if (alive) { alive[2 * e] = f & 1u; alive[2 * e + 1] = (f >> 1) & 1u; }
    if (done) done[e] = (f >> 2) & 1u;
    if (winner) winner[e] = (f >> TRON_FLAG_WINNER_SHIFT) & 3u;
    if (ep_len) ep_len[e] = (int32_t)(m.y >> 16);
}

__global__ void import_meta_kernel(uint2* __restrict__ meta, int n, const int8_t* heads, const uint8_t* alive,
                                   const uint8_t* done, const uint8_t* winner, const int32_t* ep_len) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    uint2 m = meta[e];
    uint32_t f = m.y & 0xFFu;
    uint32_t k = m.y >> 16;
    if (heads) m.x = ((const uint32_t*)heads)[e];
    if (alive) f = (f & ~3u) | (alive[2 * e] ? 1u : 0u) | (alive[2 * e + 1] ? 2u : 0u);
    if (done) f = (f & ~TRON_FLAG_DONE) | (done[e] ? TRON_FLAG_DONE : 0u);
    if (winner) f = (f & ~(3u << TRON_FLAG_WINNER_SHIFT)) | ((winner[e] & 3u) << TRON_FLAG_WINNER_SHIFT);
    if (ep_len) k = (uint32_t)ep_len[e] & 0xFFFFu;
    f &= ~0x20u;  // TRON_FLAG_BOXES_VALID: an imported grid has unknown dirty boxes
    m.y = f | (k << 16);
    meta[e] = m;
}

int launch_export_meta(const void* meta, int n, int8_t* heads, uint8_t* alive, uint8_t* done, uint8_t* winner,
                       int32_t* ep_len, cudaStream_t s) {
    export_meta_kernel<<<(n + 255) / 256, 256, 0, s>>>((const uint2*)meta, n, heads, alive, done, winner, ep_len);
    return cudaGetLastError() == cudaSuccess ? TRON_OK : TRON_ERR_CUDA;
}
int launch_import_meta(void* meta, int n, const int8_t* heads, const uint8_t* alive, const uint8_t* done,
                       const uint8_t* winner, const int32_t* ep_len, cudaStream_t s) {
    import_meta_kernel<<<(n + 255) / 256, 256, 0, s>>>((uint2*)meta, n, heads, alive, done, winner, ep_len);
    return cudaGetLastError() == cudaSuccess ? TRON_OK : TRON_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------
// policies
// ------------------------------------------------------------------------------------------------
__global__ void advance_counter_kernel(unsigned long long* c, unsigned long long delta) { *c += delta; }
int launch_advance_counter(uint64_t* c, uint64_t delta, cudaStream_t s) {
    advance_counter_kernel<<<1, 1, 0, s>>>((unsigned long long*)c, delta);
    return cudaGetLastError() == cudaSuccess ? TRON_OK : TRON_ERR_CUDA;
}

__global__ void random_actions_kernel(uint8_t* actions, int n, unsigned long long seed, unsigned long long counter,
                                      const unsigned long long* counter_dev, unsigned long long base) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    if (counter_dev) counter += *counter_dev;
    const uint4 r = philox(seed, counter, base + (unsigned long long)e, TAG_ACTION, 0);
    ((uchar2*)actions)[e] = make_uchar2((unsigned char)(r.x >> 30), (unsigned char)(r.y >> 30));
}

// epsilon-greedy (reference DDQN.py:90-110): explore iff u <= eps, else first argmax of the 4 q-values
template <typename QT>
__global__ void select_actions_kernel(const QT* __restrict__ q, int n, float eps, uint8_t* actions, unsigned long long seed,
                                      unsigned long long counter, const unsigned long long* counter_dev, unsigned long long base) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (counter_dev) counter += *counter_dev;
    float v[4];
    if constexpr (sizeof(QT) == 4) {
        const float4 f = ((const float4*)q)[i];
        v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    } else {
        const uint2 w = ((const uint2*)q)[i];
        v[0] = __uint_as_float(w.x << 16); v[1] = __uint_as_float(w.x & 0xFFFF0000u);
        v[2] = __uint_as_float(w.y << 16); v[3] = __uint_as_float(w.y & 0xFFFF0000u);
    }
    const uint4 r = philox(seed, counter, base + (unsigned long long)i, TAG_EPS, 0);
    const float u = (float)(r.x >> 8) * (1.0f / 16777216.0f);
    int best = 0;
#pragma unroll
    for (int j = 1; j < 4; ++j) if (v[j] > v[best]) best = j;
    actions[i] = (uint8_t)(u <= eps ? (r.y >> 30) : (uint32_t)best);
}

int launch_random_actions(uint8_t* actions, int n, uint64_t seed, uint64_t counter, const uint64_t* cdev, uint64_t base, cudaStream_t s) {
    random_actions_kernel<<<(n + 255) / 256, 256, 0, s>>>(actions, n, seed, counter, (const unsigned long long*)cdev, base);
    return cudaGetLastError() == cudaSuccess ? TRON_OK : TRON_ERR_CUDA;
}
int launch_select_actions(const void* q, int q_dtype, int n, float eps, uint8_t* actions, uint64_t seed, uint64_t counter,
                          const uint64_t* cdev, uint64_t base, cudaStream_t s) {
    const unsigned long long* cd = (const unsigned long long*)cdev;
    if (q_dtype == TRON_F32) select_actions_kernel<float><<<(n + 255) / 256, 256, 0, s>>>((const float*)q, n, eps, actions, seed, counter, cd, base);
    else if (q_dtype == TRON_BF16) select_actions_kernel<uint16_t><<<(n + 255) / 256, 256, 0, s>>>((const uint16_t*)q, n, eps, actions, seed, counter, cd, base);
    else return TRON_ERR_INVALID;
    return cudaGetLastError() == cudaSuccess ? TRON_OK : TRON_ERR_CUDA;
}

// pop_up (reference tron/util.py:11-37) on already-encoded observations: [n, cells] -> [n, 3, cells] planes
// wall = (o == -1); my = 1 if o == -2, 10 if o == 10; enemy = 1 if o == -3, 10 if o == -10.
__global__ void pop_up_kernel(const void* __restrict__ obs, int in_dtype, long long n_maps, int cells, void* planes, int out_dtype) {
    const long long total = n_maps * cells;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        float o;
        if (in_dtype == TRON_F32) o = ((const float*)obs)[i];
        else if (in_dtype == TRON_BF16) o = __uint_as_float((uint32_t)((const uint16_t*)obs)[i] << 16);
        else if (in_dtype == TRON_I8) o = (float)((const int8_t*)obs)[i];
        else if (in_dtype == TRON_I32) o = (float)((const int32_t*)obs)[i];
        else o = (float)((const long long*)obs)[i];
        const float w = o == -1.f ? 1.f : 0.f;
        const float m = o == -2.f ? 1.f : o == 10.f ? 10.f : 0.f;
        const float e = o == -3.f ? 1.f : o == -10.f ? 10.f : 0.f;
        const long long map = i / cells, c = i - map * cells;
        const long long b = map * 3 * cells + c;
        if (out_dtype == TRON_F32) { ((float*)planes)[b] = w; ((float*)planes)[b + cells] = m; ((float*)planes)[b + 2 * cells] = e; }
        else if (out_dtype == TRON_BF16) {
            ((__nv_bfloat16*)planes)[b] = __float2bfloat16_rn(w); ((__nv_bfloat16*)planes)[b + cells] = __float2bfloat16_rn(m);
            ((__nv_bfloat16*)planes)[b + 2 * cells] = __float2bfloat16_rn(e);
        } else { ((int8_t*)planes)[b] = (int8_t)w; ((int8_t*)planes)[b + cells] = (int8_t)m; ((int8_t*)planes)[b + 2 * cells] = (int8_t)e; }
    }
}
int launch_pop_up(const void* obs, int in_dtype, int64_t n_maps, int cells, void* planes, int out_dtype, cudaStream_t s) {
    const long long total = n_maps * cells;
    const int blocks = (int)min((long long)148 * 16, (total + 255) / 256);
    pop_up_kernel<<<blocks, 256, 0, s>>>(obs, in_dtype, n_maps, cells, planes, out_dtype);
    return cudaGetLastError() == cudaSuccess ? TRON_OK : TRON_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------
// replay ring.  A batched step yields 2N transitions whose frames are already contiguous
// ([N,2,P,C] == [2N, frame]), so push is a wrapped streaming copy and gather a row gather.
// ------------------------------------------------------------------------------------------------
// copy n frames of `fb` bytes (fb % 16 == 0) into ring slots (cursor+i) % capacity; one warp per frame chunk
__global__ void replay_push_frames_kernel(uint4* __restrict__ ring_s, uint4* __restrict__ ring_s2, const uint4* __restrict__ s,
                                          const uint4* __restrict__ s2, long long n, int vec_per_frame,
                                          unsigned long long cursor, long long capacity) {
    const long long total = n * vec_per_frame;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long f = i / vec_per_frame;
        const int v = (int)(i - f * vec_per_frame);
        const long long slot = (long long)((cursor + (unsigned long long)f) % (unsigned long long)capacity);
        const uint4 a = __ldcs(s + i), b = __ldcs(s2 + i);
        ring_s[slot * vec_per_frame + v] = a;
        ring_s2[slot * vec_per_frame + v] = b;
    }
}
__global__ void replay_push_frames_bytes_kernel(uint8_t* ring_s, uint8_t* ring_s2, const uint8_t* s, const uint8_t* s2, long long n,
                                                long long fb, unsigned long long cursor, long long capacity) {
    const long long total = n * fb;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long f = i / fb, v = i - f * fb;
        const long long slot = (long long)((cursor + (unsigned long long)f) % (unsigned long long)capacity);
        ring_s[slot * fb + v] = s[i];
        ring_s2[slot * fb + v] = s2[i];
    }
}
__global__ void replay_push_scalars_kernel(uint8_t* ring_a, float* ring_r, uint8_t* ring_d, const uint8_t* a, const float* r,
                                           const uint8_t* d, int done_stride, long long n, unsigned long long cursor,
                                           long long capacity) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long slot = (long long)((cursor + (unsigned long long)i) % (unsigned long long)capacity);
    ring_a[slot] = a[i];
    ring_r[slot] = r[i];
    ring_d[slot] = d[done_stride == 2 ? i / 2 : i];
}

int launch_replay_push(const replay_ring* ring, uint64_t cursor, const void* s, const void* s2, const uint8_t* action,
                       const float* reward, const uint8_t* done, int done_stride, int64_t n, cudaStream_t st) {
    const long long es = tron_dtype_size(ring->frame_dtype);
    const long long fb = (long long)ring->frame_elems * es;
    const bool vec = (fb % 16 == 0) && ((((uintptr_t)s | (uintptr_t)s2 | (uintptr_t)ring->state | (uintptr_t)ring->next_state) & 15u) == 0);
    if (vec) {
        const long long total = n * (fb / 16);
        const int blocks = (int)min((long long)148 * 16, (total + 255) / 256);
        replay_push_frames_kernel<<<blocks, 256, 0, st>>>((uint4*)ring->state, (uint4*)ring->next_state, (const uint4*)s, (const uint4*)s2, n,
                                                          (int)(fb / 16), cursor, ring->capacity);
    } else {
        const long long total = n * fb;
        const int blocks = (int)min((long long)148 * 16, (total + 255) / 256);
        replay_push_frames_bytes_kernel<<<blocks, 256, 0, st>>>((uint8_t*)ring->state, (uint8_t*)ring->next_state, (const uint8_t*)s,
                                                                (const uint8_t*)s2, n, fb, cursor, ring->capacity);
    }
    replay_push_scalars_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(ring->action, ring->reward, ring->done, action, reward, done,
                                                                      done_stride, n, cursor, ring->capacity);
    return cudaGetLastError() == cudaSuccess ? TRON_OK : TRON_ERR_CUDA;
}

// gather: one CTA per sampled transition; converts frame dtype -> out dtype (f32 | bf16)
__device__ __forceinline__ float load_f32(const void* p, int dt, long long i) {
    if (dt == TRON_F32) return ((const float*)p)[i];
    if (dt == TRON_BF16) return __uint_as_float((uint32_t)((const uint16_t*)p)[i] << 16);
    return (float)((const int8_t*)p)[i];
}
__device__ __forceinline__ void store_out(void* p, int dt, long long i, float v) {
    if (dt == TRON_F32) ((float*)p)[i] = v;
    else ((__nv_bfloat16*)p)[i] = __float2bfloat16_rn(v);
}
__global__ void replay_gather_kernel(const replay_ring ring, const long long* __restrict__ idx, long long k, void* out_s, void* out_s2,
                                     int out_dtype, long long* out_a, float* out_r, float* out_d) {
    const long long row = blockIdx.x;
    if (row >= k) return;
    const long long slot = min(max(idx[row], 0ll), (long long)ring.capacity - 1);  // a bad index must not read outside the ring
    const int F = ring.frame_elems;
    const int fd = ring.frame_dtype;
    if (fd == out_dtype && ((size_t)F * tron_elem(fd)) % 16 == 0) {  // same dtype: 16-byte row copy
        const int nv = (int)((size_t)F * tron_elem(fd) / 16);
        const uint4* a = (const uint4*)((const char*)ring.state + (size_t)slot * F * tron_elem(fd));
        const uint4* b = (const uint4*)((const char*)ring.next_state + (size_t)slot * F * tron_elem(fd));
        uint4* oa = (uint4*)((char*)out_s + (size_t)row * F * tron_elem(fd));
        uint4* ob = (uint4*)((char*)out_s2 + (size_t)row * F * tron_elem(fd));
        for (int v = threadIdx.x; v < nv; v += blockDim.x) { oa[v] = a[v]; ob[v] = b[v]; }
    } else if (fd == TRON_BF16 && out_dtype == TRON_F32 && F % 8 == 0) {  // bf16 ring -> f32 batch: 16-byte loads, 2 x 16-byte stores
        const int nv = F / 8;
        const uint4* a = (const uint4*)((const uint16_t*)ring.state + (size_t)slot * F);
        const uint4* b = (const uint4*)((const uint16_t*)ring.next_state + (size_t)slot * F);
        uint4* oa = (uint4*)((float*)out_s + (size_t)row * F);
        uint4* ob = (uint4*)((float*)out_s2 + (size_t)row * F);
        for (int v = threadIdx.x; v < nv; v += blockDim.x) {
            const uint4 x = a[v], y = b[v];
            oa[2 * v] = make_uint4(x.x << 16, x.x & 0xFFFF0000u, x.y << 16, x.y & 0xFFFF0000u);
            oa[2 * v + 1] = make_uint4(x.z << 16, x.z & 0xFFFF0000u, x.w << 16, x.w & 0xFFFF0000u);
            ob[2 * v] = make_uint4(y.x << 16, y.x & 0xFFFF0000u, y.y << 16, y.y & 0xFFFF0000u);
            ob[2 * v + 1] = make_uint4(y.z << 16, y.z & 0xFFFF0000u, y.w << 16, y.w & 0xFFFF0000u);
        }
    } else {
        for (int j = threadIdx.x; j < F; j += blockDim.x) {
            store_out(out_s, out_dtype, row * F + j, load_f32(ring.state, fd, slot * F + j));
            store_out(out_s2, out_dtype, row * F + j, load_f32(ring.next_state, fd, slot * F + j));
        }
    }
    if (threadIdx.x == 0) {
        out_a[row] = (long long)ring.action[slot];
        out_r[row] = ring.reward[slot];
        out_d[row] = (float)ring.done[slot];
    }
}
int launch_replay_gather(const replay_ring* ring, const int64_t* idx, int64_t k, void* out_s, void* out_s2, int out_dtype,
                         int64_t* out_a, float* out_r, float* out_d, cudaStream_t st) {
    replay_gather_kernel<<<(unsigned)k, 128, 0, st>>>(*ring, (const long long*)idx, k, out_s, out_s2, out_dtype, (long long*)out_a, out_r, out_d);
    return cudaGetLastError() == cudaSuccess ? TRON_OK : TRON_ERR_CUDA;
}

// Floyd's sampling without replacement, one warp: for j = size-k..size-1: t = U[0,j]; take t unless taken, else j.
__global__ void replay_sample_kernel(long long size, int k, unsigned long long seed, unsigned long long counter, long long* idx) {
    extern __shared__ long long chosen[];
    const int lane = threadIdx.x;
    for (int i = 0; i < k; ++i) {
        const unsigned long long j = (unsigned long long)(size - k + i);
        const uint4 r = philox(seed, counter, (unsigned long long)i, TAG_SAMPLE, 0);
        const unsigned long long x = ((unsigned long long)r.x << 32) | r.y;
        const unsigned long long t = __umul64hi(x, j + 1);
        bool dup = false;
        for (int q = lane; q < i; q += 32) dup |= (unsigned long long)chosen[q] == t;
        dup = __any_sync(0xFFFFFFFFu, dup);
        __syncwarp();
        if (lane == 0) chosen[i] = (long long)(dup ? j : t);
        __syncwarp();
    }
    for (int i = lane; i < k; i += 32) idx[i] = chosen[i];
}
int launch_replay_sample(int64_t size, int k, uint64_t seed, uint64_t counter, int64_t* idx, cudaStream_t st) {
    replay_sample_kernel<<<1, 32, (size_t)k * sizeof(long long), st>>>(size, k, seed, counter, (long long*)idx);
    return cudaGetLastError() == cudaSuccess ? TRON_OK : TRON_ERR_CUDA;
}

}  // namespace tron
