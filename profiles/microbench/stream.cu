// stream.cu — how fast can one pass over several SoA columns go on B200?  (the fixed cost of a blocked read-phase pass)
// Per element: read u32 + f64 + f64 + u32 (24 B), write f64 + u32 (12 B).  Variants differ in load width / hints / layout.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
template <int V>
__global__ void __launch_bounds__(256) k_scalar(const uint32_t* __restrict__ off, const double* __restrict__ own, double* __restrict__ sum, uint32_t* __restrict__ cnt, uint64_t n) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    if (V == 0) { const uint32_t o = off[t]; const double w = own[t]; sum[t] = sum[t] + w; cnt[t] = cnt[t] + o; }
    else { const uint32_t o = __ldcs(off + t); const double w = __ldcs(own + t); __stcs(sum + t, __ldcs(sum + t) + w); __stcs(cnt + t, __ldcs(cnt + t) + o); }
}
// 4 consecutive elements per thread, 128-bit accesses
template <int V>
__global__ void __launch_bounds__(256) k_vec4(const uint32_t* __restrict__ off, const double* __restrict__ own, double* __restrict__ sum, uint32_t* __restrict__ cnt, uint64_t n) {
    const uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (t >= n) return;
    uint4 o, c; double2 w0, w1, s0, s1;
    if (V == 0) {
        o = *reinterpret_cast<const uint4*>(off + t); c = *reinterpret_cast<const uint4*>(cnt + t);
        w0 = *reinterpret_cast<const double2*>(own + t); w1 = *reinterpret_cast<const double2*>(own + t + 2);
        s0 = *reinterpret_cast<const double2*>(sum + t); s1 = *reinterpret_cast<const double2*>(sum + t + 2);
    } else {
        o = __ldcs(reinterpret_cast<const uint4*>(off + t)); c = __ldcs(reinterpret_cast<const uint4*>(cnt + t));
        w0 = __ldcs(reinterpret_cast<const double2*>(own + t)); w1 = __ldcs(reinterpret_cast<const double2*>(own + t + 2));
        s0 = __ldcs(reinterpret_cast<const double2*>(sum + t)); s1 = __ldcs(reinterpret_cast<const double2*>(sum + t + 2));
    }
    s0.x += w0.x; s0.y += w0.y; s1.x += w1.x; s1.y += w1.y; c.x += o.x; c.y += o.y; c.z += o.z; c.w += o.w;
    if (V == 0) {
        *reinterpret_cast<double2*>(sum + t) = s0; *reinterpret_cast<double2*>(sum + t + 2) = s1; *reinterpret_cast<uint4*>(cnt + t) = c;
    } else {
        __stcs(reinterpret_cast<double2*>(sum + t), s0); __stcs(reinterpret_cast<double2*>(sum + t + 2), s1); __stcs(reinterpret_cast<uint4*>(cnt + t), c);
    }
}
// out-of-place accumulators (read sum/cnt, write sum2/cnt2): no read-modify-write of the same lines
__global__ void __launch_bounds__(256) k_oop(const uint32_t* __restrict__ off, const double* __restrict__ own, const double* __restrict__ sum, const uint32_t* __restrict__ cnt,
                                               double* __restrict__ sum2, uint32_t* __restrict__ cnt2, uint64_t n) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint32_t o = __ldcs(off + t); const double w = __ldcs(own + t); __stcs(sum2 + t, __ldcs(sum + t) + w); __stcs(cnt2 + t, __ldcs(cnt + t) + o);
}
// read-only (reduction into nothing) and copy references
__global__ void __launch_bounds__(256) k_read(const uint32_t* __restrict__ off, const double* __restrict__ own, const double* __restrict__ sum, const uint32_t* __restrict__ cnt, double* out, uint64_t n) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double v = (double)__ldcs(off + t) + __ldcs(own + t) + __ldcs(sum + t) + (double)__ldcs(cnt + t);
    if (v == 12345.678) out[0] = v;
}
__global__ void __launch_bounds__(256) k_copy(const double2* __restrict__ a, double2* __restrict__ b, uint64_t n2) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n2) b[t] = a[t];
}
int main() {
    const uint64_t n = 100000000;
    uint32_t *off, *cnt, *cnt2; double *own, *sum, *sum2;
    CK(cudaMalloc(&off, n * 4)); CK(cudaMalloc(&cnt, n * 4)); CK(cudaMalloc(&cnt2, n * 4)); CK(cudaMalloc(&own, n * 8)); CK(cudaMalloc(&sum, n * 8)); CK(cudaMalloc(&sum2, n * 8));
    CK(cudaMemset(off, 0, n * 4)); CK(cudaMemset(cnt, 0, n * 4)); CK(cudaMemset(own, 0, n * 8)); CK(cudaMemset(sum, 0, n * 8));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto time = [&](const char* name, double bytes, auto&& launch) {
        float best = 1e9f, ms;
        for (int r = 0; r < 5; ++r) { CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
        CK(cudaGetLastError());
        printf("%-44s %7.3f ms  %6.2f TB/s\n", name, best, bytes / best / 1e9);
    };
    const unsigned g = (unsigned)((n + 255) / 256), g4 = (unsigned)((n / 4 + 255) / 256);
    time("copy 800 MB -> 800 MB (double2)", 1.6e9, [&] { k_copy<<<(unsigned)((n / 2 + 255) / 256), 256>>>((const double2*)own, (double2*)sum2, n / 2); });
    time("read-only 4 columns (24 B/elem)", 24.0 * n, [&] { k_read<<<g, 256>>>(off, own, sum, cnt, sum2, n); });
    time("scalar, plain ld/st (36 B/elem)", 36.0 * n, [&] { k_scalar<0><<<g, 256>>>(off, own, sum, cnt, n); });
    time("scalar, .cs hints", 36.0 * n, [&] { k_scalar<1><<<g, 256>>>(off, own, sum, cnt, n); });
    time("4 elems/thread 128-bit, plain", 36.0 * n, [&] { k_vec4<0><<<g4, 256>>>(off, own, sum, cnt, n); });
    time("4 elems/thread 128-bit, .cs", 36.0 * n, [&] { k_vec4<1><<<g4, 256>>>(off, own, sum, cnt, n); });
    time("scalar .cs, out-of-place accumulators", 36.0 * n, [&] { k_oop<<<g, 256>>>(off, own, sum, cnt, sum2, cnt2, n); });
    return 0;
}
