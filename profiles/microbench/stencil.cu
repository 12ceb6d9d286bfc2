// stencil.cu — microbenchmark for the grid-stencil read phase (reduce transition over the implicit Moore stencil of
// connect_raster_neighbors!, BASELINE config 2: Game of Life on 4096 x 4096, periodic).
// The engine's reduce_stencil_kernel runs a thread per cell: decode the position, 8 one-byte neighbour loads (L1 hits), 8 folds,
// finish, store — 0.147 ms per generation, issue-bound at ~100 instructions per cell.  Shapes compared here, all with
// the same generic fold / finish functor (nothing Life-specific such as bit packing):
//   (a) thread per cell, 8 loads                              — the engine's shape
//   (b) warp marches down a 30-cell wide strip: one load per lane and row, left / right neighbours by shuffle, three rows kept in
//       registers (sliding window); lanes 0 and 31 only carry the halo columns
//   (c) as (b) with two rows finished per iteration (more independent work per thread)
// Every shape is checked against a CPU evaluation of the same generation.
//   usage: stencil [nx] [ny]          Written in round 1 after the GPU budget was spent: NOT RUN YET.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

struct Cell { uint8_t active; };
struct Life {                                                   // the reduce transition of transitions/gol.h
    struct Acc { uint32_t n; };
    __host__ __device__ static void init(Acc& a) { a.n = 0; }
    __host__ __device__ static void fold(const Cell&, const Cell& nb, Acc& a) { a.n += nb.active; }
    __host__ __device__ static void finish(Cell& self, const Acc& a) { self.active = (a.n == 3 || (self.active && a.n == 2)) ? 1 : 0; }
};

// (a) thread per cell
template <class F>
__global__ void __launch_bounds__(256) per_cell(const Cell* __restrict__ in, Cell* __restrict__ out, uint32_t nx, uint32_t ny) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nx * ny) return;
    const uint32_t x = i % nx, y = i / nx;
    Cell self = in[i];
    typename F::Acc a; F::init(a);
    const uint32_t xm = x ? x - 1 : nx - 1, xp = x + 1 < nx ? x + 1 : 0, ym = y ? y - 1 : ny - 1, yp = y + 1 < ny ? y + 1 : 0;
    F::fold(self, in[ym * nx + xm], a); F::fold(self, in[ym * nx + x], a); F::fold(self, in[ym * nx + xp], a);
    F::fold(self, in[y * nx + xm], a);                                      F::fold(self, in[y * nx + xp], a);
    F::fold(self, in[yp * nx + xm], a); F::fold(self, in[yp * nx + x], a); F::fold(self, in[yp * nx + xp], a);
    F::finish(self, a);
    out[i] = self;
}

// (b) / (c) warp per strip of 30 columns, RY rows per warp, ROWS rows finished per iteration
template <class F, int RY, int ROWS>
__global__ void __launch_bounds__(256) strip(const Cell* __restrict__ in, Cell* __restrict__ out, uint32_t nx, uint32_t ny) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t strips = (nx + 29) / 30, bands = (ny + RY - 1) / RY;
    if (warp >= strips * bands) return;
    const uint32_t sx = warp % strips, by = warp / strips;
    const int32_t xs = (int32_t)(sx * 30) - 1 + (int32_t)lane;              // column this lane carries (lane 0 / 31: halo)
    const uint32_t x = xs < 0 ? nx - 1 : ((uint32_t)xs >= nx ? (uint32_t)xs - nx : (uint32_t)xs);
    const bool owner = lane >= 1 && lane <= 30 && (uint32_t)xs < nx && xs >= 0;
    const uint32_t y0 = by * RY, y1 = y0 + RY < ny ? y0 + RY : ny;
    auto row = [&](uint32_t y) { return in[(size_t)y * nx + x]; };
    Cell up = row(y0 ? y0 - 1 : ny - 1), mid = row(y0);
    Cell upl, upr, midl, midr;
    upl.active = __shfl_up_sync(0xffffffffu, up.active, 1); upr.active = __shfl_down_sync(0xffffffffu, up.active, 1);
    midl.active = __shfl_up_sync(0xffffffffu, mid.active, 1); midr.active = __shfl_down_sync(0xffffffffu, mid.active, 1);
    for (uint32_t y = y0; y < y1; y += ROWS) {
#pragma unroll
        for (int k = 0; k < ROWS; ++k) {
            const uint32_t yy = y + k;
            if (yy >= y1) break;
            const Cell dn = row(yy + 1 < ny ? yy + 1 : 0);
            Cell dnl, dnr;
            dnl.active = __shfl_up_sync(0xffffffffu, dn.active, 1); dnr.active = __shfl_down_sync(0xffffffffu, dn.active, 1);
            Cell self = mid;
            typename F::Acc a; F::init(a);
            F::fold(self, upl, a); F::fold(self, up, a); F::fold(self, upr, a);
            F::fold(self, midl, a);                      F::fold(self, midr, a);
            F::fold(self, dnl, a); F::fold(self, dn, a); F::fold(self, dnr, a);
            F::finish(self, a);
            if (owner) out[(size_t)yy * nx + x] = self;
            up = mid; upl = midl; upr = midr; mid = dn; midl = dnl; midr = dnr;
        }
    }
}

static void cpu_step(const std::vector<uint8_t>& in, std::vector<uint8_t>& out, uint32_t nx, uint32_t ny) {
    for (uint32_t y = 0; y < ny; ++y)
        for (uint32_t x = 0; x < nx; ++x) {
            Cell self{in[(size_t)y * nx + x]};
            Life::Acc a; Life::init(a);
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    if (!dx && !dy) continue;
                    const uint32_t xx = (x + nx + dx) % nx, yy = (y + ny + dy) % ny;
                    Life::fold(self, Cell{in[(size_t)yy * nx + xx]}, a);
                }
            Life::finish(self, a);
            out[(size_t)y * nx + x] = self.active;
        }
}

int main(int argc, char** argv) {
    setvbuf(stdout, nullptr, _IONBF, 0);
    const uint32_t nx = argc > 1 ? atoi(argv[1]) : 4096, ny = argc > 2 ? atoi(argv[2]) : 4096;
    const size_t n = (size_t)nx * ny;
    std::vector<uint8_t> h(n), ref(n), got(n);
    uint64_t s = 88172645463325252ull;
    for (size_t i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (s % 100) < 35; }
    cpu_step(h, ref, nx, ny);
    Cell *in, *out; CK(cudaMalloc(&in, n)); CK(cudaMalloc(&out, n));
    CK(cudaMemcpy(in, h.data(), n, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto run = [&](const char* name, auto&& launch) {
        CK(cudaMemset(out, 2, n));
        float best = 1e9f, ms;
        for (int r = 0; r < 5; ++r) { CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
        CK(cudaGetLastError());
        CK(cudaMemcpy(got.data(), out, n, cudaMemcpyDeviceToHost));
        size_t bad = 0; for (size_t i = 0; i < n; ++i) bad += got[i] != ref[i];
        printf("%-52s %8.4f ms  %7.1f G cells/s  %6.1f GB/s of 2 B/cell   mismatches %zu\n", name, best, n / best / 1e6, 2.0 * n / best / 1e6, bad);
    };
    printf("Game of Life generation on %u x %u (periodic Moore stencil), generic fold/finish functor\n", nx, ny);
    run("(a) thread per cell, 8 loads", [&] { per_cell<Life><<<(unsigned)((n + 255) / 256), 256>>>(in, out, nx, ny); });
    auto strips = [&](int ry) { return (unsigned)((((size_t)((nx + 29) / 30) * ((ny + ry - 1) / ry)) * 32 + 255) / 256); };
    run("(b) warp per 30-column strip, 32 rows, 1 row/iter", [&] { strip<Life, 32, 1><<<strips(32), 256>>>(in, out, nx, ny); });
    run("(b) warp per 30-column strip, 64 rows, 1 row/iter", [&] { strip<Life, 64, 1><<<strips(64), 256>>>(in, out, nx, ny); });
    run("(b) warp per 30-column strip, 128 rows, 1 row/iter", [&] { strip<Life, 128, 1><<<strips(128), 256>>>(in, out, nx, ny); });
    run("(c) warp per 30-column strip, 64 rows, 2 rows/iter", [&] { strip<Life, 64, 2><<<strips(64), 256>>>(in, out, nx, ny); });
    run("(c) warp per 30-column strip, 64 rows, 4 rows/iter", [&] { strip<Life, 64, 4><<<strips(64), 256>>>(in, out, nx, ny); });
    return 0;
}
