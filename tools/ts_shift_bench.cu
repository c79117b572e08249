// Micro-benchmark of the tcgen05.mma form with the A operand in tensor memory ("TS"), its .ashift qualifier and
// tcgen05.shift: (1) what exactly gets shifted and when, on a known pattern; (2) cycles per MMA against the
// shared-memory-operand ("SS") form, alone and while other warps keep the shared-memory pipe busy with stores.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ts_shift_bench.bin tools/ts_shift_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_bf16.h>
#include "../hello_b200/csrc/tc_ptx.cuh"
using namespace hello;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 db;\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}" ::"r"(d),
        "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(acc), "r"(ptx::DESC_HI_SBO128)
        : "memory");
}
__device__ __forceinline__ void mma_ts_ashift(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 db;\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16.ashift [%0], [%1], db, %3, p;\n\t}" ::"r"(d),
        "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(acc), "r"(ptx::DESC_HI_SBO128)
        : "memory");
}
__device__ __forceinline__ void tc_shift_down(uint32_t a_tmem) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.shift.cta_group::1.down [%0];\n\t}" ::"r"(a_tmem)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}

// ------------------------------------------------------------------------------------------------ semantics
// A[m][k]: k=0 -> m+1, k=1 -> 1, k=2 -> 100+m (only rows < 128 exact in bf16 up to 256: 100+m rounds, fine as a tag),
// second A tile (columns 8..15 of the A region): k=0 -> 2*(m+1) capped.  B = selector: D[m][n] = A[m][n] for n < 16.
// out[test][m][0..3]
__global__ void __launch_bounds__(160, 1) semantics(float* out, int variant) {
    __shared__ __align__(128) uint8_t bsm[16 * 32];          // B: [2 chunks][16 rows][8 bf16]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 16 * 32 / 2; i += blockDim.x) reinterpret_cast<uint16_t*>(bsm)[i] = 0;
    __syncthreads();
    if (threadIdx.x < 16) {
        // B[n][k] = (n == k): element (chunk c = k/8, row n, e = k%8)
        const int n = threadIdx.x, k = n;
        reinterpret_cast<__nv_bfloat16*>(bsm)[(k / 8) * 16 * 8 + n * 8 + (k % 8)] = __float2bfloat16(1.f);
    }
    if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); }
    if (warp == 4) { ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 512); ptx::tmem_relinquish(); }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *(volatile uint32_t*)&tmem_slot;
    const uint32_t A0 = tmem + 256;                             // A region: columns 256..271 (two K=16 tiles)
    if (warp < 4) {
        const int m = threadIdx.x;
        const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
        uint32_t v[8];
        for (int c = 0; c < 8; ++c) v[c] = 0;
        auto pk = [](float lo, float hi) { return ptx::pack_bf16x2(lo, hi); };
        v[0] = pk((float)(m + 1), 1.f);
        v[1] = pk((float)(m % 64 + 100), 0.f);
        tmem_st8(tl + 256, v);
        v[0] = pk((float)(2 * (m % 100) + 2), 1.f);
        v[1] = 0;
        tmem_st8(tl + 264, v);
        // poison D
        float z[16];
        for (int c = 0; c < 16; ++c) z[c] = -7.f;
        for (int t = 0; t < 8; ++t) ptx::tmem_st16(tl + t * 16, z);
        ptx::tmem_wait_st();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (warp == 0) {
        const uint32_t b0 = ptx::desc_lo(ptx::smem_u32(bsm), 16 * 16);
        constexpr uint32_t idesc = ptx::idesc_bf16_m128(16);
        if (variant == 0) {
            mma_ts(tmem + 0, A0, b0, idesc, 0);            // D0: plain
            mma_ts_ashift(tmem + 16, A0, b0, idesc, 0);    // D1: with .ashift
            mma_ts(tmem + 32, A0, b0, idesc, 0);           // D2: plain again (did the shift persist?)
            mma_ts(tmem + 48, A0 + 8, b0, idesc, 0);       // D3: the neighbouring A tile (was it touched?)
            tc_shift_down(A0);
            mma_ts(tmem + 64, A0, b0, idesc, 0);           // D4: after an explicit tcgen05.shift
            mma_ts(tmem + 80, A0 + 8, b0, idesc, 0);       // D5: neighbouring tile after the explicit shift
            mma_ts_ashift(tmem + 96, A0, b0, idesc, 0);    // D6: .ashift again
            mma_ts(tmem + 112, A0, b0, idesc, 0);          // D7: plain
        } else {
            // .ashift with accumulation: D0 = A + shift(A)?  and two shifts in a row
            mma_ts(tmem + 0, A0, b0, idesc, 0);
            mma_ts_ashift(tmem + 0, A0, b0, idesc, 1);
            mma_ts_ashift(tmem + 0, A0, b0, idesc, 1);     // D0 = sum of three "taps"
            mma_ts(tmem + 16, A0, b0, idesc, 0);           // D1: state of A afterwards
            tc_shift_down(A0 + 8);
            tc_shift_down(A0 + 8);
            mma_ts(tmem + 32, A0 + 8, b0, idesc, 0);       // D2: tile 1 after two explicit shifts
            mma_ts(tmem + 48, A0, b0, idesc, 0);           // D3: tile 0 untouched by them?
            for (int t = 4; t < 8; ++t) mma_ts(tmem + t * 16, A0, b0, idesc, 0);
        }
        ptx::tc_commit(ptx::smem_u32(&bar));
    }
    ptx::mbar_wait(ptx::smem_u32(&bar), 0);
    ptx::tc_fence_after();
    if (warp < 4) {
        const int m = threadIdx.x;
        const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
        for (int t = 0; t < 8; ++t) {
            float x[16];
            ptx::tmem_ld16(tl + t * 16, x);
            ptx::tmem_wait_ld();
            for (int c = 0; c < 4; ++c) out[(t * 128 + m) * 4 + c] = x[c];
        }
        // raw A region afterwards
        uint32_t a[8];
        tmem_ld8(tl + 256, a);
        ptx::tmem_wait_ld();
        out[(8 * 128 + m) * 4 + 0] = __uint_as_float(a[0] << 16);
        out[(8 * 128 + m) * 4 + 1] = __uint_as_float(a[0] & 0xffff0000u);
        out[(8 * 128 + m) * 4 + 2] = __uint_as_float(a[1] << 16);
        tmem_ld8(tl + 264, a);
        ptx::tmem_wait_ld();
        out[(8 * 128 + m) * 4 + 3] = __uint_as_float(a[0] << 16);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 4) ptx::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ throughput
// MODE 0: SS (A from shared memory), 1: TS, 2: TS with .ashift on every MMA, 3: TS + one tcgen05.shift per MMA
// `noise` warps (5..) store 16 bytes per thread back to back into shared memory while the MMAs run.
struct Cfg { int reps, noise; };
template <int N, int MODE>
__global__ void __launch_bounds__(32 * 14, 1) bench(Cfg c, long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __shared__ volatile int stop;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); stop = 0; }
    if (warp == 4) { ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 512); ptx::tmem_relinquish(); }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *(volatile uint32_t*)&tmem_slot;
    constexpr uint32_t LBO = 1952;
    if (warp == 0) {
        constexpr uint32_t idesc = ptx::idesc_bf16_m128(N);
        const uint32_t a0 = ptx::desc_lo(ptx::smem_u32(smem), LBO);
        const uint32_t b0 = ptx::desc_lo(ptx::smem_u32(smem) + 65536, N * 16);
        const uint32_t d0 = tmem, at = tmem + 256;
        __syncwarp();
        long long t0 = clock64();
#pragma unroll 1
        for (int r = 0; r < c.reps; ++r) {
#pragma unroll
            for (int k = 0; k < 24; ++k) {
                const uint32_t bo = (uint32_t)(k % 4) * ((N * 32) >> 4);
                const uint32_t ao = (uint32_t)(k % 3) + (uint32_t)(k / 3 % 4) * 2 * (LBO >> 4);
                const uint32_t ta = at + (uint32_t)(k / 3 % 4) * 8;
                if (MODE == 0) ptx::mma_bf16_ss(d0, a0 + ao, b0 + bo, idesc, 1u);
                else if (MODE == 1) mma_ts(d0, ta, b0 + bo, idesc, 1u);
                else if (MODE == 2) mma_ts_ashift(d0, ta, b0 + bo, idesc, 1u);
                else { tc_shift_down(ta); mma_ts(d0, ta, b0 + bo, idesc, 1u); }
            }
        }
        long long t1 = clock64();
        ptx::tc_commit(ptx::smem_u32(&bar));
        ptx::mbar_wait(ptx::smem_u32(&bar), 0);
        long long t2 = clock64();
        if (lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; stop = 1; }
    } else if (warp >= 5 && warp < 5 + c.noise) {
        uint8_t* p = smem + 100 * 1024 + (warp - 5) * 4096 + lane * 16;
        uint4 v = make_uint4(lane, warp, 1, 2);
        long long n = 0;
        while (!stop) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(ptx::smem_u32(p + i * 512)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
            ++n;
        }
        if (lane == 0) out[2 + warp - 5] = n * 8;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 4) ptx::tmem_dealloc(tmem, 512);
}

template <int N, int MODE>
void run(long long* d) {
    cudaFuncSetAttribute(bench<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    for (int noise : {0, 4, 8}) {
        Cfg c{40, noise};
        cudaMemset(d, 0, 128);
        bench<N, MODE><<<1, 32 * 14, 160 * 1024>>>(c, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s (N=%d mode=%d)\n", cudaGetErrorString(e), N, MODE); exit(1); }
        long long h[16]; cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
        const double nm = 24.0 * c.reps;
        long long st = 0;
        for (int w = 0; w < noise; ++w) st += h[2 + w];
        printf("%4d %5d %6d | %9.1f %9.1f | %8.1f\n", N, MODE, noise, h[0] / nm, h[1] / nm,
               st ? (double)st * 512.0 / (double)h[1] : 0.0);
    }
}
template <int N>
void run_n(long long* d) { run<N, 0>(d); run<N, 1>(d); run<N, 2>(d); run<N, 3>(d); }

int main() {
    float* o; cudaMalloc(&o, 9 * 128 * 4 * 4);
    static float h[9 * 128 * 4];
    const int rows[] = {0, 1, 2, 3, 4, 30, 31, 32, 33, 34, 62, 63, 64, 65, 66, 94, 95, 96, 97, 124, 125, 126, 127};
    for (int variant = 0; variant < 2; ++variant) {
        cudaMemset(o, 0, sizeof(h));
        semantics<<<1, 160>>>(o, variant);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("semantics error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, o, sizeof(h), cudaMemcpyDeviceToHost);
        printf("semantics variant %d: per lane m, D_t[m][0] (A[m][0]=m+1), D_t[m][1] (=1), D_t[m][2] (=100+m%%64) for t=0..7; then raw A\n", variant);
        for (int m : rows) {
            printf("m=%3d |", m);
            for (int t = 0; t < 8; ++t) printf(" %5.0f/%1.0f/%3.0f", h[(t * 128 + m) * 4], h[(t * 128 + m) * 4 + 1], h[(t * 128 + m) * 4 + 2]);
            printf(" | A: %4.0f %2.0f %4.0f  A1: %4.0f\n", h[(8 * 128 + m) * 4], h[(8 * 128 + m) * 4 + 1], h[(8 * 128 + m) * 4 + 2], h[(8 * 128 + m) * 4 + 3]);
        }
    }
    long long* d; cudaMalloc(&d, 128);
    printf("%4s %5s %6s | %9s %9s | %8s\n", "N", "mode", "noise", "issue/MMA", "total/MMA", "noiseB/clk");
    run_n<16>(d); run_n<32>(d); run_n<64>(d); run_n<128>(d); run_n<256>(d);
    return 0;
}
