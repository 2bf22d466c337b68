// Shared helpers for the sm_100a kernels of the MEPOL k-NN entropy hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mepol_b200.h"

namespace mepol {

void set_error(const char *fmt, ...);

#define MEPOL_CUDA_CHECK(expr)                                                              \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            ::mepol::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,         \
                               cudaGetErrorString(_e));                                     \
            return MEPOL_E_CUDA;                                                            \
        }                                                                                   \
    } while (0)

#define MEPOL_REQUIRE(cond, code, ...)                                                      \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            ::mepol::set_error(__VA_ARGS__);                                                \
            return (code);                                                                  \
        }                                                                                   \
    } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Bump allocator over a caller-owned workspace (256-byte aligned slices).
struct Carver {
    char *base;
    size_t off;
    explicit Carver(void *p) : base(static_cast<char *>(p)), off(0) {}
    template <typename T>
    T *take(size_t count) {
        off = align_up(off, 256);
        T *r = reinterpret_cast<T *>(base ? base + off : nullptr);
        off += count * sizeof(T);
        return r;
    }
    size_t bytes() const { return align_up(off, 256); }
};

int device_sm_count();

// ------------------------------------------------------------------ device-side helpers
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Next representable float above a non-negative finite value (conservative rounding of bounds).
__device__ __forceinline__ float next_up(float v) { return __uint_as_float(__float_as_uint(v) + 1u); }
__device__ __forceinline__ float double_to_float_up(double v) {
    float f = __double2float_ru(v);
    return f;
}

}  // namespace mepol
