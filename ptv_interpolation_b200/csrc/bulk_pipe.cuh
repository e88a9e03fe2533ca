// Shared by the stencil kernels: mbarrier / 1-D bulk-copy (TMA engine) primitives for shared-memory row rings,
// and the exact division by a grid spacing.  sm_100a only.
#pragma once
#include <math.h>
#include <stdint.h>

namespace ptv {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(done) : "r"(bar), "r"(parity), "r"(10000000u) : "memory");  // suspend-time hint: sleep, do not poll
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// t / h, correctly rounded, from rh = RN(1 / h): two Newton corrections with exact FMA residuals.  After the
// first one q is a faithful quotient, and for a faithful q the second one returns RN(t / h) (Markstein 1990).
// Needs t, h and t / h well inside the normal range (checked: by the caller for h, here for t).  Zero numerators
// (masked faces: the common case) stay on the straight-line path -- the sequence yields a zero, only its sign is
// set afterwards -- and the rest (subnormal, huge, inf, nan) takes the IEEE division out of line, so the loop
// bodies that divide stay small.  ptv_selftest_division compares it with __ddiv_rn on the device.
static __device__ __noinline__ double div_by_spacing_slow(double t, double h) { return __ddiv_rn(t, h); }

__device__ __forceinline__ double div_by_spacing(double t, double h, double rh) {
  double q = __dmul_rn(t, rh);
  double r = __fma_rn(-h, q, t);
  q = __fma_rn(r, rh, q);
  r = __fma_rn(-h, q, t);
  q = __fma_rn(r, rh, q);
  const unsigned e = ((unsigned)__double2hiint(t) >> 20) & 0x7ffu;
  if (t == 0.0) {
    q = h > 0.0 ? t : -t;  // 0 / h is a signed zero
  } else if (e - 323u >= 1400u) {  // |t| outside [2^-700, 2^700)
    q = div_by_spacing_slow(t, h);
  }
  return q;
}

// div_by_spacing's range for the divisor (host check)
static inline bool spacing_ok(double h) {
  return h == h && fabs(h) > 0x1p-300 && fabs(h) < 0x1p300;
}

}  // namespace ptv
