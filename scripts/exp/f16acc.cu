// Experiment: tcgen05.mma kind::f16 with an F16 accumulator + tcgen05.ld ... pack::16b.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o f16acc f16acc.cu ; run on a B200.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t a) {
    uint64_t d = 0;
    d |= (uint64_t)((a >> 4) & 0x3fff);
    d |= (uint64_t)((128 >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((512 >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__global__ void __launch_bounds__(128, 1) k(const __half *A, const __half *B, float *out, uint32_t *raw, int cfmt_f16) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint4 *sa = reinterpret_cast<uint4 *>(smem), *sb = reinterpret_cast<uint4 *>(smem + 8192);
    for (int i = threadIdx.x; i < 512; i += 128) { sa[i] = reinterpret_cast<const uint4 *>(A)[i]; sb[i] = reinterpret_cast<const uint4 *>(B)[i]; }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (MMA)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_slot;
    if (threadIdx.x == 0) {
        // idesc: c_format (bits 4-5): 0=F16 1=F32; a/b format (bits 7-9,10-12): 0=F16
        const uint32_t idesc = ((cfmt_f16 ? 0u : 1u) << 4) | (0u << 7) | (0u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
        for (int ks = 0; ks < 2; ++ks) {
            uint64_t ad = umma_desc(smem_u32(smem) + ks * 256), bd = umma_desc(smem_u32(smem + 8192) + ks * 256);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(ks) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // wait
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
    const int row = warp * 32 + lane;
    // (1) raw 32-bit view of the first 32 columns
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) raw[row * 32 + j] = r[j];
    if (cfmt_f16) {
        // (2) packed view: x32 registers <- 64 columns, two f16 per register
        for (int half = 0; half < 2; ++half) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                           "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                         : "r"(taddr + half * 64) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 32; ++j) {
                __half2 h = *reinterpret_cast<__half2 *>(&r[j]);
                out[row * 128 + half * 64 + 2 * j] = __low2float(h);
                out[row * 128 + half * 64 + 2 * j + 1] = __high2float(h);
            }
        }
    } else {
        for (int c = 0; c < 4; ++c) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                           "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                         : "r"(taddr + c * 32) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 32; ++j) out[row * 128 + c * 32 + j] = __uint_as_float(r[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(128u) : "memory");
}

static size_t tile_idx(int r, int c) { return (size_t)(((r >> 3) * 4 + (c >> 3)) * 64 + (r & 7) * 8 + (c & 7)); }

int main() {
    std::vector<__half> A(4096), B(4096);
    std::vector<double> a(128 * 32), b(128 * 32);
    srand(1);
    for (int r = 0; r < 128; ++r)
        for (int c = 0; c < 32; ++c) {
            float x = (rand() / (float)RAND_MAX - 0.5f) * 6.f, y = (rand() / (float)RAND_MAX - 0.5f) * 6.f;
            __half hx = __float2half(x), hy = __float2half(y);
            A[tile_idx(r, c)] = hx; B[tile_idx(r, c)] = hy;
            a[r * 32 + c] = __half2float(hx); b[r * 32 + c] = __half2float(hy);
        }
    __half *dA, *dB; float *dO; uint32_t *dR;
    cudaMalloc(&dA, 8192); cudaMalloc(&dB, 8192); cudaMalloc(&dO, 128 * 128 * 4); cudaMalloc(&dR, 128 * 32 * 4);
    cudaMemcpy(dA, A.data(), 8192, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), 8192, cudaMemcpyHostToDevice);
    for (int f16 = 0; f16 < 2; ++f16) {
        cudaMemset(dO, 0xff, 128 * 128 * 4);
        k<<<1, 128, 16384>>>(dA, dB, dO, dR, f16);
        cudaError_t e = cudaDeviceSynchronize();
        printf("cfmt_f16=%d: %s\n", f16, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<float> o(128 * 128); std::vector<uint32_t> raw(128 * 32);
        cudaMemcpy(o.data(), dO, o.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(raw.data(), dR, raw.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxrel = 0;
        for (int r = 0; r < 128; ++r)
            for (int c = 0; c < 128; ++c) {
                double s = 0; for (int kk = 0; kk < 32; ++kk) s += a[r * 32 + kk] * b[c * 32 + kk];
                double err = fabs(o[r * 128 + c] - s); if (err > maxerr) maxerr = err;
                if (fabs(s) > 1) { double rel = err / fabs(s); if (rel > maxrel) maxrel = rel; }
            }
        printf("  max abs err %.5g  max rel err(|s|>1) %.5g\n", maxerr, maxrel);
        double s00 = 0, s01 = 0; for (int kk = 0; kk < 32; ++kk) { s00 += a[kk] * b[kk]; s01 += a[kk] * b[32 + kk]; }
        printf("  expected D[0][0]=%.4f D[0][1]=%.4f ; raw col0=0x%08x col1=0x%08x  out[0][0..3]=%.4f %.4f %.4f %.4f\n", s00, s01, raw[0], raw[1], o[0], o[1], o[2], o[3]);
    }
    return 0;
}
