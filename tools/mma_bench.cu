// Micro-benchmark of tcgen05.mma (kind::f16, bf16, M=128, K=16, cta_group::1) issue/execute rates on one SM:
// cycles per MMA as a function of N, of how many independent accumulators the stream rotates over, of the A
// operand's row offset (16-byte shifted taps vs 128-byte aligned), and of how many warps issue concurrently.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_bench.bin tools/mma_bench.cu
#include <cstdio>
#include <cstdlib>
#include "../hello_b200/csrc/tc_ptx.cuh"
using namespace hello;

struct Cfg { int warps, reps; };

template <int N, int NACC, int SHIFT, int SAME>
__global__ void __launch_bounds__(160, 1) bench(Cfg c, long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[4];
    __shared__ uint32_t tmem_slot;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) ptx::mbar_init(ptx::smem_u32(&bars[i]), 1); ptx::fence_mbar_init(); }
    if (warp == 4) { ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 512); ptx::tmem_relinquish(); }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *(volatile uint32_t*)&tmem_slot;
    constexpr uint32_t LBO = 1952;
    if (warp < c.warps) {
        constexpr uint32_t idesc = ptx::idesc_bf16_m128(N);
        const uint32_t a_base = ptx::smem_u32(smem) + warp * 32768 + SHIFT;   // A: chunk arrays, this warp's region
        const uint32_t b_base = ptx::smem_u32(smem) + 131072 + warp * 16384;  // B: 4 units of N x 16
        const uint32_t a0 = ptx::desc_lo(a_base, LBO);
        const uint32_t b0 = ptx::desc_lo(b_base, N * 16);
        const uint32_t d0 = tmem + warp * 128;
        __syncwarp();
        long long t0 = clock64();
#pragma unroll 1
        for (int r = 0; r < c.reps; ++r) {
#pragma unroll
            for (int k = 0; k < 36; ++k) {
                constexpr int dummy = 0; (void)dummy;
                const uint32_t acc = (uint32_t)(k % NACC) * N;
                const uint32_t ao = SAME ? 0u : (uint32_t)(k % 3) + (uint32_t)(k / 3 % 4) * 2 * (LBO >> 4);
                const uint32_t bo = SAME ? 0u : (uint32_t)(k % 4) * ((N * 32) >> 4);
                ptx::mma_bf16_ss(d0 + acc, a0 + ao, b0 + bo, idesc, 1u);
            }
        }
        long long t1 = clock64();
        ptx::tc_commit(ptx::smem_u32(&bars[warp]));
        ptx::mbar_wait(ptx::smem_u32(&bars[warp]), 0);
        long long t2 = clock64();
        if (lane == 0) { out[warp * 2] = t1 - t0; out[warp * 2 + 1] = t2 - t0; }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 4) ptx::tmem_dealloc(tmem, 512);
}

template <int N, int NACC, int SHIFT, int SAME>
void run(long long* d) {
    cudaFuncSetAttribute(bench<N, NACC, SHIFT, SAME>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int warps : {1, 2, 4}) {
        if (N * NACC > 128 && warps > 512 / (N * NACC)) continue;
        Cfg c{warps, 20};
        cudaMemset(d, 0, 64);
        bench<N, NACC, SHIFT, SAME><<<1, 160, 200 * 1024>>>(c, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
        long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
        long long iss = 0, tot = 0;
        for (int w = 0; w < warps; ++w) { iss = h[2 * w] > iss ? h[2 * w] : iss; tot = h[2 * w + 1] > tot ? h[2 * w + 1] : tot; }
        const double nm = 36.0 * c.reps * warps;
        printf("%4d %5d %7d %5d %6d | %9.1f %9.1f\n", N, NACC, SHIFT, warps, SAME, iss / nm, tot / nm);
    }
}

template <int N>
void run_n(long long* d) {
    run<N, 1, 0, 1>(d);
    run<N, 1, 0, 0>(d);
    run<N, 1, 16, 0>(d);
    if (N * 2 <= 512) { run<N, 2, 0, 0>(d); run<N, 2, 16, 0>(d); }
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    printf("%4s %5s %7s %5s %6s | %9s %9s\n", "N", "n_acc", "a_shift", "warps", "sameAB", "issue/MMA", "total/MMA");
    run_n<16>(d); run_n<32>(d); run_n<64>(d); run_n<96>(d); run_n<128>(d); run_n<192>(d); run_n<256>(d);
    return 0;
}
