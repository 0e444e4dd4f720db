// Bounds-asserting debug build of the CG kernels (-DMSGPU_CHECKED, scripts/checked_build.py ->
// dealiiscale_b200/_build/libmsgpu_checked.so; never the product library).
//
// compute-sanitizer is not available on the GPU pool this was developed on, and the kernels that keep their state in
// dynamic shared memory, distributed shared memory and tensor memory do their own index arithmetic (padded columns,
// halos, guard zones, per-warp column blocks).  In a checked build every shared-memory access of those kernels goes
// through CheckedPtr, which knows the window it may touch, and every tensor-memory address is compared with the
// column block the warp owns; a violation reports its source line through the device-side assert and fails the launch.  In the product build
// CheckedPtr<T> is T* and the checks expand to nothing (tests/test_checked_build.py: no assert message in libmsgpu.so,
// register budgets of the kernels as the launch code expects them).
#pragma once
#include <cassert>
#include <cstdint>

namespace msgpu {

#ifdef MSGPU_CHECKED
// (the device-side assert machinery: message, source line, block and thread on stderr, then cudaErrorAssert)
#define MSGPU_CHECK(cond)                                                                              \
  do {                                                                                                 \
    if (!(cond)) __assert_fail("[msgpu checked build] " #cond, __FILE__, __LINE__, __PRETTY_FUNCTION__); \
  } while (0)

template <class T>
struct CheckedPtr {
  T* p;
  const T* lo;
  const T* hi;  // may touch [lo, hi)
  __device__ __forceinline__ T& operator[](long long i) const {
    MSGPU_CHECK(p + i >= lo && p + i < hi);
    return p[i];
  }
  __device__ __forceinline__ T& operator*() const { return (*this)[0]; }
  __device__ __forceinline__ CheckedPtr operator+(long long i) const { return CheckedPtr{p + i, lo, hi}; }
  __device__ __forceinline__ CheckedPtr operator-(long long i) const { return CheckedPtr{p - i, lo, hi}; }
  __device__ __forceinline__ explicit operator bool() const { return p != nullptr; }
  __device__ __forceinline__ operator CheckedPtr<const T>() const { return CheckedPtr<const T>{p, lo, hi}; }
};
template <class T>
__device__ __forceinline__ CheckedPtr<T> checked(T* p, const T* lo, const T* hi) {
  return CheckedPtr<T>{p, lo, hi};
}
// n consecutive elements starting at q must lie inside the window (vector loads, hand-offs to helpers)
template <class T>
__device__ __forceinline__ T* raw(const CheckedPtr<T>& q, long long n = 1) {
  MSGPU_CHECK(q.p >= q.lo && q.p + n <= q.hi);
  return q.p;
}
#else
#define MSGPU_CHECK(cond) ((void)0)
template <class T>
using CheckedPtr = T*;
template <class T>
__device__ __forceinline__ T* checked(T* p, const T*, const T*) {
  return p;
}
template <class T>
__device__ __forceinline__ T* raw(T* q, long long = 1) {
  return q;
}
#endif

// Tensor-memory address a (lane quadrant in bits 31:16, column in 15:0): columns [col, col + ncols) must lie inside
// [lo, hi), the block of columns this warp owns.
#define MSGPU_CHECK_TMEM(a, ncols, lo, hi) \
  MSGPU_CHECK((((a) & 0xffffu) >= (lo)) && (((a) & 0xffffu) + (ncols) <= (hi)))

}  // namespace msgpu
