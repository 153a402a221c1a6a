// async_copy.cuh -- mbarrier and TMA bulk-copy (cp.async.bulk) helpers shared by the producer / consumer kernels.
// Shared-memory addresses are 32-bit shared-window addresses (__cvta_generic_to_shared).
#pragma once

#include <stdint.h>

namespace bbx {
namespace ac {

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// test_wait never suspends the thread (try_wait may park it for a hardware time limit when the phase is still open):
// the one to use for "is the slot free yet?" polls whose answer is usually no
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: gives up after ~4e6 polls so that a protocol bug ends the kernel with wrong results (caught by the parity
// tests) instead of hanging the device
static __device__ __noinline__ bool mbar_wait_slow(uint32_t bar, uint32_t parity) {
  for (int spin = 0; spin < (1 << 22); spin++)
    if (mbar_try_wait(bar, parity)) return true;
  return false;
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return true;
  return mbar_wait_slow(bar, parity);
}
// wait with a sticky failure word in shared memory: once any wait of the CTA has timed out (~1e6 polls), every later
// wait gives up after 64 polls, so a protocol bug costs milliseconds, not minutes
__device__ __forceinline__ void mbar_wait_or_die(uint32_t bar, uint32_t parity, uint32_t dead_word) {
#pragma unroll 1
  for (int spin = 0; spin < (1 << 20); spin++) {
    if (mbar_try_wait(bar, parity)) return;
    if ((spin & 63) == 63) {
      uint32_t d;
      asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(d) : "r"(dead_word));
      if (d) return;
    }
  }
  asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(dead_word), "r"(1u) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 16-byte asynchronous copy global -> shared (LDGSTS, L2 only) and the arrive that fires on an mbarrier once all of this
// thread's earlier cp.async copies have landed (noinc: the barrier's expected count already includes it)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// 8- and 4-byte forms (LDGSTS through L1): a thread that copies exactly the elements it reads back needs no barrier
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (16-byte aligned addresses and size)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
// volatile + memory clobber: the load must stay behind the mbarrier wait that makes the data visible
__device__ __forceinline__ float2 lds2(uint32_t addr) {
  float2 r;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(addr) : "memory");
  return r;
}

}  // namespace ac
}  // namespace bbx
