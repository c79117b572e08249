// Micro-benchmark of tcgen05.ld (TMEM -> registers) on one SM: cycles per 32x32b.x16 / .x32 load for 1..8 warps, with and
// without the tensor pipe accumulating into other TMEM columns at the same time.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_bench.bin tools/tmem_bench.cu
#include <cstdio>
#include <cstdlib>
#include "../hello_b200/csrc/tc_ptx.cuh"
using namespace hello;

template <int X32, int BATCH>
__global__ void __launch_bounds__(320, 1) bench(int warps, int mma_on, int reps, long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 64 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); }
    if (warp == 9) { ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 512); ptx::tmem_relinquish(); }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *(volatile uint32_t*)&tmem_slot;
    if (warp < warps) {
        const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;   // columns [0,128): loads
        float acc = 0.f;
        __syncwarp();
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            float v[BATCH][X32 ? 32 : 16];
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {
                if (X32) ptx::tmem_ld32(tl + (b & 1) * 32, v[b]); else ptx::tmem_ld16(tl + (b & 3) * 16, v[b]);
            }
            ptx::tmem_wait_ld();
#pragma unroll
            for (int b = 0; b < BATCH; ++b) acc += v[b][0] + v[b][(X32 ? 31 : 15)];
        }
        long long t1 = clock64();
        if (lane == 0) { out[warp] = t1 - t0; out[16 + warp] = (long long)acc; }
    } else if (warp == 8 && mma_on) {
        // keep the tensor pipe busy: 128 x 128 x 16 MMAs into columns [256, 384)
        const uint32_t idesc = ptx::idesc_bf16_m128(128);
        const uint32_t a0 = ptx::desc_lo(ptx::smem_u32(smem), 2048), b0 = ptx::desc_lo(ptx::smem_u32(smem) + 16384, 2048);
        for (int r = 0; r < reps * 4; ++r) {
#pragma unroll
            for (int k = 0; k < 8; ++k) ptx::mma_bf16_ss(tmem + 256, a0, b0, idesc, 1u);
        }
        ptx::tc_commit(ptx::smem_u32(&bar));
        ptx::mbar_wait(ptx::smem_u32(&bar), 0);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 9) ptx::tmem_dealloc(tmem, 512);
}

template <int X32, int BATCH>
void run(long long* d) {
    cudaFuncSetAttribute(bench<X32, BATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int mma : {0, 1})
        for (int warps : {1, 4, 8}) {
            const int reps = 200;
            cudaMemset(d, 0, 256);
            bench<X32, BATCH><<<1, 320, 64 * 1024>>>(warps, mma, reps, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
            long long h[32]; cudaMemcpy(h, d, 256, cudaMemcpyDeviceToHost);
            long long mx = 0; for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
            const double per_round = (double)mx / reps;
            const double bytes = (double)warps * BATCH * (X32 ? 32 : 16) * 32 * 4;
            printf("x%-2d batch %d warps %d mma %d | %7.1f cycles per round (%d loads + wait) | %6.1f B/clk SM-wide\n", X32 ? 32 : 16, BATCH,
                   warps, mma, per_round, BATCH, bytes / per_round);
        }
}

int main() {
    long long* d; cudaMalloc(&d, 256);
    run<0, 1>(d); run<0, 2>(d); run<0, 4>(d); run<1, 1>(d); run<1, 2>(d);
    return 0;
}
